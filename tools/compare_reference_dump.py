"""Compare golden binaries written by julia/dump_reference.jl (real Oceananigans) with the CPU
oracle:  python tools/compare_reference_dump.py <dir>.  Prints rel-L2 per field and step; this is
the check that would turn "parity unpinned" into a pinned statement."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cases import make_case, rel_l2
from oracle import pyoracle as O

d = Path(sys.argv[1])
for form, kind in (("jacobian", "J"), ("divergence", "D")):
    g, cfg, U = make_case(kind, 64)
    O.fill_halos(cfg, U)
    n = 0
    for k in (0, 1, 10, 100, 1000):
        O.step(cfg, U, 0.01, k - n); n = k
        errs = []
        for f, tag in enumerate("uvhA"):
            p = d / f"golden_{form}_64_step{k}_{tag}.f64"
            if not p.exists():
                errs.append(None); continue
            ref = np.fromfile(p, dtype=np.float64).reshape(U[f].shape)
            errs.append(rel_l2(g, U[f], ref, f))
        print(form, "step", k, ["%.3e" % e if e is not None else "missing" for e in errs])
