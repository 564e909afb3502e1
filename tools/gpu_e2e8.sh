#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m 2>/dev/null | head -14
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extra-legs 2> gpurun_out/bench_e2e_n$N.err > gpurun_out/bench_e2e_n$N.json; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_e2e_n$N.json').read().strip().splitlines()[-1])
print("N=%d value %.3e ms/step %.3f"%(d["n_gpus"], d["value"], d["ms_per_step"]), "clocks", d["clocks"])
print("  e2e", d["e2e"])
PY
