"""torchrun -nproc-per-node N tools/gpu_multi_test.py : y-slab run on N GPUs vs the single-GPU run.
Fields must be bit-identical (same per-cell arithmetic, no reduction in the state update)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch, torch.distributed as dist
from swmhd_b200 import abi
from swmhd_b200.context import Context
from swmhd_b200.distributed import SlabModel, split_rows, slab_of_global
from cases import make_case

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
ok_all = True
for kind, Nx, Ny in [("J", 256, 200), ("D", 256, 200), ("BJ", 128, 136), ("BD", 128, 136)]:
    g, cfg, U = make_case(kind, Nx, Ny=Ny, perturb=31)
    cfg.device = local
    ref = Context(cfg); ref.set_state(U); ref.fill_halos(); tr_ref = ref.step_diag(0.002, 4); Uref = ref.get_state(); dref = ref.diagnostics(); ref.close()
    j0, ny = split_rows(Ny, world)[rank]
    extra = lambda k: 1 if (k == abi.V and cfg.topo_y == abi.BOUNDED) else 0
    sm = SlabModel(cfg, rank, world, local)
    sm.set_state([slab_of_global(U[k], j0, ny, extra(k)) for k in range(4)])
    sm.fill_halos()
    tr = sm.step_diag(0.002, 4)
    sm.synchronize()
    out = sm.get_state()
    d = sm.diagnostics()
    ok = all(np.array_equal(out[k][3:3 + ny], Uref[k][3 + j0:3 + j0 + ny]) for k in range(4))
    keys = ("ke", "me", "pe", "sum_h", "max_abs_u", "max_abs_A", "min_h")
    dok = all(abs(d[key] - dref[key]) <= 1e-13 * max(1.0, abs(dref[key])) for key in keys)
    dok = dok and all(abs(a[key] - b[key]) <= 1e-13 * max(1.0, abs(b[key])) for a, b in zip(tr, tr_ref) for key in keys)
    t = torch.tensor([int(ok), int(dok)], device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{kind} {Nx}x{Ny} world={world}: fields bit-identical={bool(t[0])} diagnostics match={bool(t[1])}", flush=True)
    ok_all = ok_all and bool(t[0]) and bool(t[1])
    sm.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
