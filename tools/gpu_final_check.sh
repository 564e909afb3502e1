#!/bin/bash
# last check of a round on one GPU: what the driver runs (GPU tests, smoke, the default bench line)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f e2e %.2f ms frac %.3f"%(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"]), d["clocks"], "launches", d["gpu_launches"])
print(sorted(d.keys()))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-200
