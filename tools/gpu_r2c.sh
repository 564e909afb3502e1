#!/bin/bash
# round 2: parity suite, quick bench, bench line, ncu launch list + full captures (TAG = $1)
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_$TAG.log
python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast" | tee gpurun_out/quick_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
SMALL="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extra-legs"
$SMALL > gpurun_out/plain1.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $SMALL > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
# DRAM traffic and FP64 pipe of the three launches of one step at the BENCH size (4096^2)
M=dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:substage -s 3 -c 3 --csv --log-file gpurun_out/dram4096_$TAG.csv python tools/prof_step.py 4096 jacobian > gpurun_out/ncu1b.log 2>&1
echo "ncu dram rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:substage -s 3 -c 3 --csv --log-file gpurun_out/dram4096_div_$TAG.csv python tools/prof_step.py 4096 divergence > gpurun_out/ncu1c.log 2>&1
echo "ncu dram div rc=$?"
# full captures at 2048^2 (400 MB working set >> L2), deterministic launch order (tools/prof_step.py): the three launches of a
# step with the fused diagnostics (stage 1 = DIAG variant), of a plain step, and of a divergence step with diagnostics
P="python tools/prof_step.py 2048"
$P jacobian > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:substage -s 3 -c 3 -f -o gpurun_out/prof_jac_diag_$TAG $P jacobian > gpurun_out/ncu2.log 2>&1
echo "ncu full jacobian (diag step) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:substage -s 3 -c 3 -f -o gpurun_out/prof_jac_plain_$TAG $P jacobian --plain > gpurun_out/ncu2p.log 2>&1
echo "ncu full jacobian (plain step) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:substage -s 3 -c 3 -f -o gpurun_out/prof_div_diag_$TAG $P divergence > gpurun_out/ncu3.log 2>&1
echo "ncu full divergence (diag step) rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
    print("value %.3e ms/step %.3f"%(d["value"], d["ms_per_step"]), "launches", d["gpu_launches"], "clocks", d["clocks"])
    print("  per_stage", d["roofline"]["per_stage"], "frac", d["roofline"]["frac"])
    print("  e2e", d["e2e"] and {k:d["e2e"][k] for k in ("value","ms_per_step")})
    for k in ("config4","config5"):
        print(" ", k, {x:d[k][x] for x in ("value","ms_per_step","clocks")})
except Exception as ex:
    print("bench parse failed", ex)
PY
ls -la gpurun_out | tail -8
