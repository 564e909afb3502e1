"""Deterministic launch sequence for ncu captures: python tools/prof_step.py N {jacobian|divergence} [--plain]
One warm-up step, then two steps: with the fused diagnostics every step = [substage<1,DIAG>, diag_final, halo,
substage<2>, halo, substage<3>, halo]; --plain: without (3 substage kernels + 3 halo fills per step).
`ncu -k regex:substage -s 3 -c 3` therefore captures the three substage launches of the second step."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from swmhd_b200 import abi
from swmhd_b200.context import Context
from cases import make_case

N = int(sys.argv[1]); form = sys.argv[2] if len(sys.argv) > 2 else "jacobian"
g, cfg, U = make_case("J" if form == "jacobian" else "D", N, arith=abi.ARITH_FAST, perturb=3)
ctx = Context(cfg); ctx.set_state(U); ctx.fill_halos()
dt = 0.01 * 64 / N
if "--plain" in sys.argv:
    ctx.step(dt, 3)
else:
    ctx.step_diag(dt, 3)
print("ok", ctx.last_step_ms)
ctx.close()
