#!/bin/bash
# divergence-form evidence: bench lines (4096^2, 16384^2) + ncu --set full of one step at 2048^2
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --form divergence --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_div_$TAG.json 2> gpurun_out/bench_div_$TAG.err
python bench.py --size 16384 --form divergence --scaling strong --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_div16384_$TAG.json 2>/dev/null
FULLD="python bench.py --form divergence --size 2048 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$FULLD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:substage -s 9 -c 3 -f -o gpurun_out/prof_div_$TAG $FULLD > gpurun_out/ncu3.log 2>&1
echo "ncu full div rc=$?"
cut -c1-200 gpurun_out/bench_div_$TAG.json gpurun_out/bench_div16384_$TAG.json
