"""Appendix-C item C10 (Bounded-y wall semantics of the recalled upstream spec) probed against the four published
Bounded-y figures (energy_plots/*/{64x64,128x128}_low_B_low_U.png, digitised in tests/golden/published_traces.json;
IC: divergence_sw_mhd.jl:17,34-37).  Oracle switches (include/swmhd.h, oracle only):

    (plus, outside C10: JS weights, centred tracer advection, WENO eps = 3e-4)
    D1  SWMHD_FLAG_BC_DEPTH1   no-flux / gradient BCs fill only the first halo row (deeper rows untouched)
    W3  SWMHD_FLAG_WALL_WENO3  WENO3 one cell further from the wall than the centred-2nd-order fallback needs
    VM  SWMHD_FLAG_V_MIRROR    v|vh beyond the wall: odd mirror instead of untouched cells

Prints, per combination, the maximum |oracle - published| of KE, ME, PE over the whole run (t <= 14.5) for each figure.
    python tools/probe_c10.py [--quick]        (CPU only; ~5 min on 8 cores)
"""
import itertools
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from swmhd_b200 import abi
from oracle import pyoracle as O
from cases import make_case

TR = json.loads((ROOT / "tests" / "golden" / "published_traces.json").read_text())
J, D = "jacobian_formulation", "divergence_formulation"
GRAD = (-0.05, -0.05)
FIGS = [(J, 64, None), (D, 64, GRAD), (J, 128, GRAD), (D, 128, GRAD)]      # BC of each published run: tests/test_published_traces.py
FLAGS = [("D1", abi.FLAG_BC_DEPTH1), ("W3", abi.FLAG_WALL_WENO3), ("VM", abi.FLAG_V_MIRROR)]


def run(form, N, grad, flags, T=14.5, dt=0.01, eps=1e-6):
    g, _, U = make_case("BJ" if form == J else "BD", N)
    cfg = abi.make_config(g.Nx, g.Ny, formulation=abi.JACOBIAN if form == J else abi.DIVERGENCE, flags=flags,
                          topo_y=abi.BOUNDED, A_gradient=grad, weno_eps=eps)
    O.fill_halos(cfg, U)
    pub = TR[f"{form}/{N}x{N}_low_B_low_U"]
    worst = {k: 0.0 for k in ("ke", "me", "pe")}
    t = 0.0
    while t < T - 1e-9:
        O.step(cfg, U, dt, 50)
        t += 0.5
        d = O.diagnostics(cfg, U)
        if not np.isfinite(d["ke"]):
            return {k: float("nan") for k in worst}
        for k in worst:
            if k in pub and round(t, 1) in [round(x, 1) for x in pub[k]["t"]]:
                v = pub[k]["v"][[round(x, 1) for x in pub[k]["t"]].index(round(t, 1))]
                worst[k] = max(worst[k], abs(d[k] - v))
    return worst


def main():
    O.set_threads(8)
    rows = []
    for combo in itertools.product((0, 1), repeat=3):
        flags = sum(f for (name, f), on in zip(FLAGS, combo) if on)
        label = "+".join(name for (name, f), on in zip(FLAGS, combo) if on) or "default"
        cells = []
        for form, N, grad in FIGS:
            if "--quick" in sys.argv and N == 128:
                cells.append(None); continue
            w = run(form, N, grad, flags)
            cells.append(w)
        rows.append((label, cells))
        print(label, ["-" if c is None else "KE %.1e ME %.1e PE %.1e" % (c["ke"], c["me"], c["pe"]) for c in cells], flush=True)
    # not wall semantics: what else could keep more magnetic energy at 64^2?
    for label, flags, eps in [("JS weights (C1)", abi.FLAG_WENO_JS, 1e-6), ("A: centred 2nd order", abi.FLAG_TRACER_CEN2, 1e-6),
                              ("A: centred 4th order", abi.FLAG_TRACER_CEN4, 1e-6), ("WENO eps = 3e-4 (C2)", 0, 3e-4)]:
        cells = [None if ("--quick" in sys.argv and N == 128) else run(form, N, grad, flags, eps=eps) for form, N, grad in FIGS]
        rows.append((label, cells))
        print(label, ["-" if c is None else "KE %.1e ME %.1e PE %.1e" % (c["ke"], c["me"], c["pe"]) for c in cells], flush=True)
    print("\n| switches | " + " | ".join(f"{'Jac' if f == J else 'Div'} {N}² (KE / ME / PE)" for f, N, _ in FIGS) + " |")
    print("|---|" + "---|" * len(FIGS))
    for label, cells in rows:
        print(f"| {label} | " + " | ".join("–" if c is None else f"{c['ke']:.1e} / {c['me']:.1e} / {c['pe']:.1e}" for c in cells) + " |")


if __name__ == "__main__":
    main()
