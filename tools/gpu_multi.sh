#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
$TR tools/gpu_multi_test.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -12
$TR bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | cut -c1-700
tail -5 gpurun_out/bench_n$N.err
