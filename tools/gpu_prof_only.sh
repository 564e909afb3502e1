#!/bin/bash
TAG=${1:-x}; shift
mkdir -p gpurun_out
FULL="python bench.py --size 2048 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline $@"
$FULL > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:substage -s 9 -c 3 -f -o gpurun_out/prof_$TAG $FULL > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain_$TAG.log | cut -c1-400
