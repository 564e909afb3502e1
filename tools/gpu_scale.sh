#!/bin/bash
# weak-scaling bench at N GPUs + parity check; outputs gpurun_out/scale_n$N.json
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
if [ "$N" = "1" ]; then
  python -m pytest tests -m gpu -q 2>&1 | tail -2
  python bench.py --gpus 1 --steps 100 --warmup 10 --no-cpu-baseline 2> gpurun_out/scale_n1.err | tee gpurun_out/scale_n1.json | cut -c1-200
else
  $TR tools/gpu_multi_test.py 2>&1 | grep "world="
  $TR bench.py --gpus $N --steps 100 --warmup 10 2> gpurun_out/scale_n$N.err | tee gpurun_out/scale_n$N.json | cut -c1-200
  $TR bench.py --gpus $N --steps 20 --warmup 5 --size 16384 --scaling strong --form divergence 2> gpurun_out/strong_n$N.err | tee gpurun_out/strong_n$N.json | cut -c1-200
fi
