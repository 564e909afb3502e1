"""Small run of every kernel variant for compute-sanitizer (memcheck / racecheck)."""
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from swmhd_b200 import abi
from swmhd_b200.context import Context
from cases import make_case
for kind, N, Ny in (("J", 70, 50), ("D", 70, 50), ("BJ", 64, 40), ("BD", 65, 43)):
    for arith in (abi.ARITH_FAST, abi.ARITH_STRICT):
        g, cfg, U = make_case(kind, N, Ny=Ny, arith=arith, perturb=3)
        c = Context(cfg); c.set_state(U); c.fill_halos()
        c.step(0.004, 1); c.step_diag(0.004, 1); c.tendencies(); d = c.diagnostics(); c.close()
        assert d["all_finite"] == 1
print("sanitize case ok")
