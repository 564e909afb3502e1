#!/bin/bash
TAG=${1:-r02b}
mkdir -p gpurun_out
python tools/prof_step.py 2048 jacobian && timeout 600 ncu --set full --clock-control none --import-source on -k regex:substage -s 3 -c 3 -f -o gpurun_out/prof_diagstep_$TAG python tools/prof_step.py 2048 jacobian > gpurun_out/ncu_d1.log 2>&1; echo rc=$?
python tools/prof_step.py 2048 divergence && timeout 600 ncu --set full --clock-control none --import-source on -k regex:substage -s 3 -c 3 -f -o gpurun_out/prof_diagstep_div_$TAG python tools/prof_step.py 2048 divergence > gpurun_out/ncu_d2.log 2>&1; echo rc=$?
python tools/quick_bench.py 2048 4096 --fast 2>&1 | grep "N=.*fast"
