#!/bin/bash
# bench line at N GPUs only (usage: tools/gpu_bench_n.sh N [steps])
N=${1:-2}; K=${2:-20}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535"
timeout 330 $TR bench.py --gpus $N --steps $K --warmup 5 2> gpurun_out/bench_n$N.err > gpurun_out/bench_n$N.json; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print("N=%d value %.3e ms/step %.3f"%(d["n_gpus"], d["value"], d["ms_per_step"]), "clocks", d["clocks"])
print("  e2e ms %.2f agg H2D %.0f GB/s"%(d["e2e"]["ms_per_step"], d["e2e"].get("aggregate_h2d_GBps",0)))
for k in ("config4","config5"): print(" ", k, "ms %.3f value %.3e"%(d[k]["ms_per_step"], d[k]["value"]), d[k]["clocks"]["sm_mhz"])
print("  slab_parity ok:", d["slab_parity"]["ok"])
PY
