#!/bin/bash
# round 2, first GPU call: parity suite, A/B of the row-blocked kernels against round 1, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_r2a.log
echo "--- r02 kernels"; python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast" | tee gpurun_out/quick_r2a.log
echo "--- r01 kernels"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_r01.so python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast" | tee -a gpurun_out/quick_r2a.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2a.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
    print("value %.3e ms/step %.3f"%(d["value"], d["ms_per_step"]), "launches", d["gpu_launches"], "clocks", d["clocks"])
    print("  per_stage", d["roofline"]["per_stage"], "frac", d["roofline"]["frac"])
    print("  e2e", d["e2e"] and {k:d["e2e"][k] for k in ("value","ms_per_step")})
    for k in ("config4","config5"):
        print(" ", k, {x:d[k][x] for x in ("value","ms_per_step","clocks")})
except Exception as ex:
    print("bench parse failed", ex)
PY
