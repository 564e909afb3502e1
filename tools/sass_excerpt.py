"""profiles/r02_sass_excerpt_substage_rb_stage2.txt: trimmed SASS of swmhd::substage_rb_kernel<2,false> from the in-tree build
(cuobjdump -sass): the TMA issue (UTMALDG.2D), the L2 tile prefetch (UTMAPF.L2.2D), the mbarrier wait, a slice of the row walk
(LDS.64 / DFMA / MUFU.RCP64H) and the RK3 update with its stores.   python tools/sass_excerpt.py"""
import collections
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
txt = subprocess.run(["cuobjdump", "-sass", str(ROOT / "swmhd_b200" / "libswmhd_cuda.so")], capture_output=True, text=True).stdout
L, on = [], False
for l in txt.splitlines():
    if "Function :" in l:
        on = "substage_rb_kernelILi2ELb0" in l
    if on and not re.match(r"\s*/\* 0x", l):
        L.append(re.sub(r"\s*/\* 0x[0-9a-f]* \*/$", "", l.rstrip()))
ops = collections.Counter()
for l in L:
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(2).split(".")[0]] += 1


def find(pat, start=0):
    for i in range(start, len(L)):
        if re.search(pat, L[i]):
            return i
    return None


out = ["SASS excerpt of swmhd::substage_rb_kernel<2, false> (sm_100a), from `cuobjdump -sass swmhd_b200/libswmhd_cuda.so` (round-2 build; tools/sass_excerpt.py).",
       "Static opcode counts of the whole kernel: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(16)),
       "No tensor-core instruction (HMMA / UTCMMA / ...) anywhere: nothing on the path is a contraction.", ""]
i = find(r"SYNCS\.EXCH")
out.append("---- P0: mbarrier init + expect_tx, four TMA tile loads (cp.async.bulk.tensor.2d -> UTMALDG.2D), L2 prefetch of the slot's next tile (UTMAPF.L2.2D)")
out += [l for l in L[i - 2:find(r"UTMAPF", i) + 6] if re.search(r"SYNCS|UTMALDG|UTMAPF|ELECT|R2UR|UMOV UR|BRA", l)][:40]
j = find(r"SYNCS\.PHASECHK", i)
out += ["", "---- tile wait: mbarrier.try_wait.parity"] + L[j:j + 2]
k = find(r"MUFU\.RCP64H", find(r"BAR\.SYNC", j + 1) + 400)
out += ["", "---- row walk: shared-memory stencil loads, WENO arithmetic on the FP64 pipe, Newton-refined reciprocals"] + L[k - 60:k + 30]
s = find(r"STG\.E\.64", k)
out += ["", "---- RK3 update and stores (U_new and G^n), request of the next row's G^-"] + L[s - 12:s + 20]
(ROOT / "profiles" / "r02_sass_excerpt_substage_rb_stage2.txt").write_text("\n".join(out) + "\n")
print(len(out), "lines")
