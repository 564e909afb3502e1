#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_models.py -m gpu -q -x -k "upload or step_seq or models or aligns or flow" 2>&1 | tail -8
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-legs > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2e.json').read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f"%(d["value"], d["ms_per_step"]), "clocks", d["clocks"])
print("  e2e", {k:d["e2e"][k] for k in ("value","ms_per_step","ms_per_step_separate_calls")})
print("  roof", {k:d["roofline"].get(k) for k in ("frac","fp64_pipe_pct","instr_per_cell_substage","traffic","kernel")})
PY
