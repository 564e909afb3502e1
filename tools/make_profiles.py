"""Regenerate the tracked round-2 evidence under profiles/ from the scratch files a `tools/gpu_r2c.sh TAG` call left in
gpurun_out/:   python tools/make_profiles.py TAG
(bench line, ncu launch list, ncu summaries, DRAM/pipe metrics at 4096^2 -> traffic.json; the instruction budget and the SASS
excerpt have their own tools: tools/sass_budget.py, cuobjdump)."""
import collections
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1]

shutil.copy(G / f"bench_{tag}.json", P / "r02_bench_jacobian_4096.json")
shutil.copy(G / f"launches_{tag}.csv", P / "r02_ncu_launches_jacobian_4096.csv")
for t in ("jac_diag", "jac_plain", "div_diag"):
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(G / f"prof_{t}_{tag}.ncu-rep")], capture_output=True, text=True).stdout
    (P / f"r02_ncu_full_{t}_2048.txt").write_text(out)


def parse(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    out = collections.OrderedDict()
    for r in rows:
        out.setdefault(r[4], {})[r[12]] = float(r[14])
    return out


res = {}
for form, path in (("jacobian", G / f"dram4096_{tag}.csv"), ("divergence", G / f"dram4096_div_{tag}.csv")):
    k = parse(path)
    ncell = 4096 * 4096
    per, pipe, instr, issue, lsu, dur = [], [], [], [], [], []
    for name, m in k.items():
        per.append((m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / ncell)
        pipe.append(m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"])
        issue.append(m["smsp__issue_active.avg.pct_of_peak_sustained_active"])
        lsu.append(m["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"])
        instr.append(m["smsp__inst_executed.sum"] / (ncell / 32))
        dur.append(m["gpu__time_duration.sum"] / 1e3)
    res[form] = {"bytes_per_cell_per_launch_mean": sum(per) / 3, "per_stage_bytes_per_cell": per,
                 "algorithmic_bytes_per_cell_per_launch_mean": 320 / 3, "algorithmic_per_stage": [96, 128, 96],
                 "fp64_pipe_pct": sum(pipe) / 3, "fp64_pipe_pct_per_stage": pipe, "issue_active_pct_per_stage": issue,
                 "lsu_wavefronts_pct_per_stage": lsu, "instr_per_cell_substage": sum(instr) / 3, "instr_per_cell_per_stage": instr,
                 "ncu_duration_us_per_stage": dur, "kernels": list(k.keys()),
                 "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --clock-control none -k regex:substage -s 3 -c 3 "
                           f"python tools/prof_step.py 4096 {form}  (the three substage launches of one RK3 step with the fused diagnostics, "
                           f"at the bench size 4096^2; tools/gpu_r2c.sh {tag})"}
json.dump(res, open(P / "traffic.json", "w"), indent=1)
for f, v in res.items():
    print(f, {k: ([round(y, 2) for y in x] if isinstance(x, list) and x and isinstance(x[0], float) else (round(x, 2) if isinstance(x, float) else None))
              for k, x in v.items() if k not in ("source", "kernels")})
