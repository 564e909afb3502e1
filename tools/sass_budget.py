"""Per-phase instruction budget of the substage kernels from an `ncu --set full --import-source on` capture.

    python tools/sass_budget.py gpurun_out/prof_r02.ncu-rep [--cells N*N] [--md]

Reads the SASS page of every substage kernel in the report (`ncu -i REP --page source --csv --print-source sass`),
weights every instruction by "Instructions Executed" and prints, per kernel, warp instructions PER CELL-SUBSTAGE by
phase and by class:

    fp64   DFMA DMUL DADD DSETP DMNMX (the 64-lane FP64 pipe)      mufu   MUFU.RCP64H (reciprocal seeds)
    smem   LDS / STS                                              glob   LDG / STG
    shfl   SHFL                                                   unif   uniform-datapath instructions (UMOV, ...)
    int    everything else on the integer / move pipes            ctrl   branches, barriers

Phases are found from the structure of the kernel: the two CTA barriers (`BAR.SYNC`) bracket the tile wait and the
derived-field phase A, the largest backward branch is the row loop; inside the loop an instruction executed R+1 times per
warp belongs to the loop head, R times to the row body, once to the south-face iteration (it = -1).
Also prints the stall-sample share of every phase (where warps spend their time).
"""
import collections
import csv
import re
import subprocess
import sys

CLASSES = ["fp64", "mufu", "smem", "glob", "shfl", "unif", "int", "ctrl"]


def cls(op):
    b = op.split(".")[0]
    if b in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        return "fp64"
    if b == "MUFU":
        return "mufu"
    if b in ("LDS", "STS"):
        return "smem"
    if b in ("LDG", "STG"):
        return "glob"
    if b == "SHFL":
        return "shfl"
    if b.startswith("U") and b not in ("UTMALDG", "UTMAPF"):
        return "unif"
    if b in ("BRA", "BSSY", "BSYNC", "BAR", "EXIT", "WARPSYNC", "NANOSLEEP"):
        return "ctrl"
    return "int"


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernels, hdr = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernels.append([r[1], []])
        elif r and r[0] == "Address":
            hdr = r
        elif kernels and len(r) > 10:
            kernels[-1][1].append(r)
    # ncu prints every kernel twice (two metric tables of the same launch): keep the first of each pair of equal names
    uniq, seen = [], collections.Counter()
    for name, body in kernels:
        seen[name] += 1
        if seen[name] % 2 == 1:
            uniq.append((name, body))
    return hdr, uniq


def analyse(hdr, body, R=4):
    iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    ins = []
    for r in body:
        toks = r[1].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if not toks:
            continue
        n = int(r[iE]) if r[iE].isdigit() else 0
        s = int(r[iS]) if r[iS].isdigit() else 0
        ins.append(dict(addr=int(r[0], 16), op=toks[0], text=r[1].strip(), n=n, samples=s))
    warps = ins[0]["n"] or 1
    bars = [k for k, x in enumerate(ins) if x["op"].startswith("BAR")]
    # row loop = backward branch with the largest span whose body holds FP64 work
    best = None
    for k, x in enumerate(ins):
        if x["op"].startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", x["text"])
            if m:
                tgt = int(m.group(1), 16)
                if tgt < x["addr"] and x["n"] / warps >= 1.5:       # taken more than once per warp (not the mbarrier spin)
                    span = x["addr"] - tgt
                    if best is None or span > best[0]:
                        best = (span, tgt, x["addr"])
    loop_lo, loop_hi = (best[1], best[2]) if best else (1 << 62, 1 << 62)
    b1 = ins[bars[0]]["addr"] if bars else 0
    b2 = ins[bars[1]]["addr"] if len(bars) > 1 else b1
    tab = collections.OrderedDict((p, collections.Counter()) for p in ("P0 (TMA issue)", "wait + phase A", "east pre-pass", "loop head", "south face (it=-1)", "row body", "tail / DIAG"))
    for x in ins:
        a, trips = x["addr"], x["n"] / warps
        if a <= b1:
            ph = "P0 (TMA issue)"
        elif a <= b2:
            ph = "wait + phase A"
        elif a < loop_lo:
            ph = "east pre-pass"
        elif a <= loop_hi:
            ph = "loop head" if trips > R + 0.5 else ("row body" if trips > 1.5 else "south face (it=-1)")
        else:
            ph = "tail / DIAG"
        tab[ph][cls(x["op"])] += x["n"]
        tab[ph]["samples"] += x["samples"]
    return warps, tab


def main():
    rep = sys.argv[1]
    md = "--md" in sys.argv
    hdr, kernels = load(rep)
    for name, body in kernels:
        if "substage" not in name:
            continue
        warps, tab = analyse(hdr, body)
        cells = warps * 32 * 4 / 1.0        # R = 4 rows per warp; cells incl. the idle lane of the divergence kernel
        short = re.sub(r"void swmhd::<unnamed>::|\(swmhd::KParams\)|\(int\)|\(bool\)", "", name)
        tot = collections.Counter()
        total_samples = sum(c["samples"] for c in tab.values()) or 1
        print(f"\n### {short}  ({warps} warps)\n")
        sep = " | " if md else "  "
        print(("| " if md else "") + sep.join(["phase".ljust(20)] + [c.rjust(6) for c in CLASSES] + [" total", "stall samples"]) + (" |" if md else ""))
        if md:
            print("|" + "---|" * (len(CLASSES) + 3))
        for ph, c in tab.items():
            per = {k: c[k] * 32 / cells for k in CLASSES}
            tot.update(per)
            print(("| " if md else "") + sep.join([ph.ljust(20)] + ["%6.1f" % per[k] for k in CLASSES] + ["%6.1f" % sum(per.values()), "%5.1f %%" % (100.0 * c["samples"] / total_samples)]) + (" |" if md else ""))
        print(("| " if md else "") + sep.join(["**per cell-substage**".ljust(20)] + ["%6.1f" % tot[k] for k in CLASSES] + ["%6.1f" % sum(tot.values()), ""]) + (" |" if md else ""))


if __name__ == "__main__":
    main()
