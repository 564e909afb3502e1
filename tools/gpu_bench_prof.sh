#!/bin/bash
# bench + ncu evidence for one round; outputs in gpurun_out/   (usage: tools/gpu_bench_prof.sh TAG)
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --form divergence --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_div_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_div_$TAG.json
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
SMALL="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$SMALL > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
# full captures at 2048^2 (400 MB working set >> L2): one step of the plain kernels, the stage-1 DIAG variant, one divergence step
FULL="python bench.py --size 2048 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$FULL > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:substage -s 9 -c 3 -f -o gpurun_out/prof_$TAG $FULL > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:substage_rb_kernel<1, 1>" -s 2 -c 1 -f -o gpurun_out/prof_diag_$TAG $FULL > gpurun_out/ncu2d.log 2>&1
echo "ncu full diag rc=$?"
FULLD="python bench.py --form divergence --size 2048 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$FULLD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:substage -s 9 -c 3 -f -o gpurun_out/prof_div_$TAG $FULLD > gpurun_out/ncu3.log 2>&1
echo "ncu full div rc=$?"
ls -la gpurun_out | tail -12
