#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f"%(d["value"], d["ms_per_step"]), "launches", d["gpu_launches"], "clocks", d["clocks"])
print("  per_stage", d["roofline"]["per_stage"]["ms"], "frac", d["roofline"]["frac"])
print("  e2e", d["e2e"] and {k:d["e2e"][k] for k in ("value","ms_per_step")})
PY
