"""Summarise an .ncu-rep (raw + source pages) into a small text report for profiles/."""
import csv, io, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__cycles_active.avg"]

def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    res, kern, hdr = {}, None, None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            kern = r[1]; continue
        if r and r[0] == "Address":
            hdr = r; continue
        if not hdr or kern is None or len(r) < len(hdr) - 5:
            continue
        d = dict(zip(hdr, r))
        o = res.setdefault(kern, {})
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                try: o[k] = o.get(k, 0) + float(d[k] or 0)
                except ValueError: pass
    return res

if __name__ == "__main__":
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"{w} [{units[i]}]: " + " | ".join(r[i][:70] for r in rows))
    for k, o in stalls(rep).items():
        s = sum(o.values()) or 1
        print("stalls", k[:90])
        print("   " + ", ".join(f"{kk[6:]} {100*v/s:.1f}%" for kk, v in sorted(o.items(), key=lambda x: -x[1])[:9]))
