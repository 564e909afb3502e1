"""world_size-2 (and 3) y-slab runs on CPU over gloo: the decomposition, neighbour map and halo
exchange of swmhd_b200.distributed, with the CPU oracle standing in for the per-slab kernels.
The slab result must be bit-identical to the single-domain oracle run (same per-cell arithmetic)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _worker(rank, world, port, kind, Nx, Ny, nsteps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from swmhd_b200 import abi
        from swmhd_b200 import distributed as D
        from oracle import pyoracle as O
        from cases import make_case

        g, cfg, Ug = make_case(kind, Nx, Ny=Ny, perturb=5)
        O.fill_halos(cfg, Ug)
        j0, ny = D.split_rows(Ny, world)[rank]
        # slab parents: rows j0 .. j0+ny+5 of the global parents (halos included)
        U = [D.slab_of_global(Ug[k], j0, ny) for k in range(4)]
        cs = abi.Config.from_buffer_copy(cfg)
        cs.Ny = ny                                   # the oracle sees the slab as its own (Nx, ny) field
        P = Nx + 6
        rows_t = {}

        def rows_of(f, which):
            r0 = {D.SOUTH_SEND: 3, D.NORTH_SEND: ny, D.SOUTH_HALO: 0, D.NORTH_HALO: ny + 3}[which]
            key = (f, which)
            if key not in rows_t:
                rows_t[key] = torch.from_numpy(U[f])[r0:r0 + 3]
            return rows_t[key]

        gam = [8.0 / 15.0, 5.0 / 12.0, 3.0 / 4.0]
        zet = [0.0, -17.0 / 60.0, -5.0 / 12.0]
        dt = 0.004
        Gm = [np.zeros_like(a) for a in U]
        for n in range(nsteps):
            for s in range(3):
                Gn = O.tendencies(cs, U)
                for k in range(4):
                    it = (slice(3, 3 + ny), slice(3, 3 + Nx))
                    if s == 0:
                        U[k][it] = U[k][it] + dt * gam[0] * Gn[k][it]
                    else:
                        U[k][it] = U[k][it] + dt * (gam[s] * Gn[k][it] + zet[s] * Gm[k][it])
                Gm = Gn
                O.fill_halos(cs, U)                  # x wrap (the local y wrap is overwritten below)
                ops = D.exchange_ops(rows_of, rank, world, periodic_y=True)
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
        O.step(cfg, Ug, dt, nsteps)
        ok = all(np.array_equal(U[k][3:3 + ny], Ug[k][3 + j0:3 + j0 + ny]) for k in range(4))
        halo_ok = all(np.array_equal(U[k], Ug[k][j0:j0 + ny + 6]) for k in range(4))
        q.put((rank, ok, halo_ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "J"), (2, "D"), (3, "J")])
def test_slab_exchange_matches_single_domain(world, kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + world * 7 + (0 if kind == "J" else 3)) % 400
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, 40, 36, 2, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok and hok for _, ok, hok in res), res


def _warm_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import time
        from swmhd_b200 import distributed as D
        # skewed clocks: rank 1 believes much more time has passed than rank 0
        skew = 10.0 * rank
        t = torch.zeros(1)

        def step():                       # a "step" that needs every rank (like the halo exchange)
            dist.all_reduce(t)
            time.sleep(0.01)

        n = D.run_in_step_for(0.2, step, "cpu", clock=lambda: time.perf_counter() + skew * (time.perf_counter() % 1.0))
        q.put((rank, n))
    finally:
        dist.destroy_process_group()


def test_time_boxed_loop_runs_the_same_number_of_steps_on_every_rank():
    """bench.py keeps the GPUs busy for ~1.5 s before timing; the ranks must agree on the step count."""
    world, port = 3, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_warm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(set(res.values())) == 1 and res[0] >= 1, res
