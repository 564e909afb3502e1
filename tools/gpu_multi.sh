#!/bin/bash
# N-GPU checks on one box (usage: tools/gpu_multi.sh N [bench-steps]): in-process n_gpus contexts (pytest), the per-rank
# NCCL ring against the single-GPU run (tools/gpu_multi_test.py), then the bench line at N GPUs.
N=${1:-2}; K=${2:-20}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 150 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "n_gpus or second_device" 2>&1 | tail -8 | tee gpurun_out/pytest_multi_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 150 $TR tools/gpu_multi_test.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -12 | tee gpurun_out/multi_test_n$N.log
timeout 330 $TR bench.py --gpus $N --steps $K --warmup 5 2> gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json; echo "bench rc=$?"
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print("N=%d value %.3e ms/step %.3f"%(d["n_gpus"], d["value"], d["ms_per_step"]), "launches", d["gpu_launches"], "clocks", d["clocks"])
    print("  e2e", d["e2e"] and {k:d["e2e"][k] for k in ("value","ms_per_step")})
    for k in ("config4","config5"):
        print(" ", k, {x:d[k][x] for x in ("value","ms_per_step","clocks")})
    print("  slab_parity", d.get("slab_parity"))
except Exception as ex:
    print("bench parse failed", ex)
PY
