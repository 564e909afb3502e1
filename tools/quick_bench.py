"""Quick device-side timing of swmhd_step at a few sizes (development aid, not bench.py)."""
import sys, time
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np
from swmhd_b200 import abi
from swmhd_b200.context import Context
from cases import make_case

def run(kind, N, arith, steps=10, warm=3):
    g, cfg, U = make_case(kind, N, arith=arith)
    ctx = Context(cfg)
    ctx.set_state(U); ctx.fill_halos()
    ctx.step(0.01 * 64 / N, warm)
    ctx.step(0.01 * 64 / N, steps)
    ms = ctx.last_step_ms / steps
    ctx.step_diag(0.01 * 64 / N, steps)
    ms_diag = ctx.last_step_ms / steps
    st = ctx.step_profile(0.01 * 64 / N, steps)
    d = ctx.diagnostics()
    ctx.close()
    cu = N * N / (ms * 1e-3)
    print(f"{kind} N={N} arith={'strict' if arith else 'fast'}: {ms:.3f} ms/step  {cu/1e9:.3f} Gcell-updates/s  "
          f"HBM-equiv {cu*320/1e9:.0f} GB/s  with-diag {ms_diag:.3f} ms  stages {[round(float(x), 3) for x in st]}  finite={d['all_finite']}", flush=True)

if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1024, 4096]
    for N in sizes:
        for kind in ("J", "D"):
            for arith in ((abi.ARITH_FAST,) if "--fast" in sys.argv else (abi.ARITH_FAST, abi.ARITH_STRICT)):
                run(kind, N, arith)
