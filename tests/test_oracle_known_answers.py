"""Pins the CPU oracle (oracle/swmhd_oracle.c) against everything the reference offers for this
path: the closed-form answers of the reference's own operator scripts
(test_formulations.jl, MHD_visualize.jl), exact t=0 values and invariants (B.2), the survey's
independent scratch-restatement checksums (B.5); the published energy traces (B.3) are in test_published_traces.py.
No GPU needed."""
import json
from pathlib import Path

import numpy as np
import pytest

from swmhd_b200 import abi
from oracle import pyoracle as O
from cases import make_case, grid_for

GOLD = Path(__file__).parent / "golden"


def analytic_fields(N, fA):
    """A evaluated analytically at centre nodes INCLUDING halos (the scripts pass A as a function of
    the OffsetArray nodes, test_formulations.jl:12, so nothing wraps), h = 1."""
    g = grid_for(N)
    idx = np.arange(-2, N + 4, dtype=np.float64)           # logical i = -2 .. N+3
    xc = g.x0 + (idx - 0.5) * g.dx
    X, Y = np.meshgrid(xc, xc)
    A = np.ascontiguousarray(fA(X, Y))
    h = np.ones_like(A)
    return g, A, h


def nodes_f_c(g):
    i = np.arange(1, g.Nx + 1, dtype=np.float64)
    return g.x0 + (i - 1.0) * g.dx, g.x0 + (i - 0.5) * g.dx     # face, centre (same in y)


# ---- B.1: test_formulations.jl:12-18,188-189 ---------------------------------------------------
B1_JAC = {64: 6.554741e-02, 128: 1.710604e-02, 256: 4.323230e-03, 512: 1.083756e-03}
B1_DIV = {64: 8.998313e-02, 128: 2.297462e-02, 256: 5.778557e-03, 512: 1.447887e-03}


@pytest.mark.parametrize("form,table", [(abi.JACOBIAN, B1_JAC), (abi.DIVERGENCE, B1_DIV)])
def test_lorentz_vs_closed_form_test_formulations(form, table):
    errs = {}
    for N in (64, 128, 256, 512):
        g, A, h = analytic_fields(N, lambda x, y: np.exp(-(x ** 2 + y ** 2)))
        cfg = abi.make_config(N, N, formulation=form)
        Fx, Fy = O.lorentz(cfg, h, A)
        xf, xc = nodes_f_c(g)
        ex = -4 * xf[None, :] * np.exp(-2 * (xf[None, :] ** 2 + xc[:, None] ** 2))   # at (xF, yC)
        ey = -4 * xf[:, None] * np.exp(-2 * (xc[None, :] ** 2 + xf[:, None] ** 2))   # at (xC, yF)
        e_x = np.abs(g.interior(Fx, abi.H) - ex).max()
        e_y = np.abs(g.interior(Fy, abi.H) - ey).max()
        assert abs(e_x - table[N]) <= 2e-6 * table[N] + 1e-9, (N, e_x)
        assert abs(e_y - e_x) <= 1e-12, "x/y errors are identical by symmetry (transpose-bug detector)"
        errs[N] = e_x
    Ns = np.array(sorted(errs))
    order = -np.polyfit(np.log10(Ns), np.log10([errs[n] for n in Ns]), 1)[0]
    assert 1.9 < order < 2.1      # what test_formulations.jl:209-210 prints (1.97 / 1.99)


# ---- B.1: MHD_visualize.jl:8-24,55-77 -------------------------------------------------------------
def test_lorentz_vs_closed_form_mhd_visualize():
    ell, A0 = 2.0, -1.0
    expect = {50: 3.455356e-03, 100: 8.820182e-04, 200: 2.217751e-04, 400: 5.551060e-05}
    for N, ref in expect.items():
        g, A, h = analytic_fields(N, lambda x, y: A0 * np.exp(-(x ** 2 + y ** 2) / ell ** 2))
        cfg = abi.make_config(N, N, formulation=abi.JACOBIAN)
        Fx, Fy = O.lorentz(cfg, h, A)
        xf, xc = nodes_f_c(g)

        def d(x, y):
            e = np.exp(-(x ** 2 + y ** 2) / ell ** 2)
            return dict(x=-A0 * 2 / ell ** 2 * x * e, y=-A0 * 2 / ell ** 2 * y * e,
                        xx=A0 * (4 * x ** 2 - 2 * ell ** 2) / ell ** 4 * e, xy=A0 * 4 * x * y / ell ** 4 * e,
                        yy=A0 * (4 * y ** 2 - 2 * ell ** 2) / ell ** 4 * e)
        a = d(xf[None, :], xc[:, None])
        b = d(xc[None, :], xf[:, None])
        lx = a["x"] * a["yy"] - a["y"] * a["xy"]
        ly = b["y"] * b["xx"] - b["x"] * b["xy"]
        # the script's numerical forms use +∂yA_num / +∂xA_num where the model uses Bx = -∂yA/h, By = ∂xA/h
        ex = np.abs(-g.interior(Fx, abi.H) - lx).max()
        ey = np.abs(-g.interior(Fy, abi.H) - ly).max()
        assert abs(ex - ref) <= 2e-6 * ref, (N, ex)
        assert abs(ey - ref) <= 2e-6 * ref, (N, ey)


# ---- WENO5: fifth order, mirror symmetry, optimal-weight identities (A.3) ---------------------------
def test_weno5_order_and_symmetry():
    cfg = abi.make_config(64, 64)
    errs = []
    for n in (32, 64, 128, 256):
        dx = 2 * np.pi / n
        edges = np.arange(-3, n + 4) * dx                      # faces f .. ; cell k spans [edges[k], edges[k+1]]
        cellavg = (np.cos(edges[:-1]) - np.cos(edges[1:])) / dx    # cell averages of sin
        L, R = O.weno_line(cfg, cellavg)
        f = np.arange(3, cellavg.size - 2)
        exact = np.sin(edges[f])
        errs.append((np.abs(L[f] - exact).max(), np.abs(R[f] - exact).max()))
    errs = np.array(errs)
    oL = -np.polyfit(np.log(np.array([32, 64, 128, 256])), np.log(errs[:, 0]), 1)[0]
    oR = -np.polyfit(np.log(np.array([32, 64, 128, 256])), np.log(errs[:, 1]), 1)[0]
    assert oL > 4.7 and oR > 4.7, (oL, oR)
    assert np.allclose(errs[:, 0], errs[:, 1], rtol=0.15)     # identical error norms on a symmetric test
    # mirror property: right(psi)[f] == left(reversed psi)[mirror f]
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(40)
    L, R = O.weno_line(cfg, psi)
    Lr, Rr = O.weno_line(cfg, psi[::-1].copy())
    f = np.arange(3, 38)
    assert np.array_equal(R[f], Lr[40 - f])
    # linear data: every candidate agrees, any weights give the exact face value
    lin = 0.5 + 0.25 * np.arange(20.0)
    L, R = O.weno_line(cfg, lin)
    assert np.allclose(L[3:18], lin[3:18] - 0.125, atol=1e-14) and np.allclose(R[3:18], lin[3:18] - 0.125, atol=1e-14)


# ---- B.2 exact t=0 values -------------------------------------------------------------------------
def test_initial_energies_exact():
    g, cfg, U = make_case("J", 64)
    O.fill_halos(cfg, U)
    d = O.diagnostics(cfg, U)
    assert abs(d["ke"] - 9.817477042468) < 1e-11
    assert abs(d["me"] - 12.109375) < 1e-13          # 12.5 * (1 - 2/Ny), exactly
    assert d["pe"] == 0.0 and d["min_h"] == 1.0 and d["sum_h"] == 4096.0
    for N, me in ((64, 0.5429248629), (128, 0.5461380972)):
        g, cfg, U = make_case("D", N)
        O.fill_halos(cfg, U)
        d = O.diagnostics(cfg, U)
        assert abs(d["me"] - me) < 1e-10 and d["ke"] == 0.0
        assert d["max_abs_div_hB"] < 5e-15            # div(hB) = 0 identically for hB = z x grad A
    for N, me in ((64, 0.02171699), (128, 0.02184552)):
        g, cfg, U = make_case("G", N)
        O.fill_halos(cfg, U)
        assert abs(O.diagnostics(cfg, U)["me"] - me) < 1e-8


# ---- B.5: survey scratch-restatement checksums (independent transcription of the same spec) --------
B5 = {
    "J": {0: [2.005302619704800e+01, 2.005302619704800e+01, 6.400000000000000e+01, 9.236476600955584e+01],
          1: [2.005207795715480e+01, 2.005173897082251e+01, 6.400000045050143e+01, 9.236474334815557e+01],
          10: [1.992717403520849e+01, 1.989372054377198e+01, 6.400327006382328e+01, 9.236361543454497e+01],
          100: [1.718746869973142e+01, 1.735095581909297e+01, 6.401658365874972e+01, 9.232591398565015e+01]},
    "D": {0: [0.0, 0.0, 6.4e+01, 3.557790435357659e+00],
          1: [1.712111966505029e-02, 3.106791795889113e-02, 6.400000000189905e+01, 3.557746597424233e+00],
          10: [1.383504277339225e-01, 2.824825162243974e-01, 6.400001526161076e+01, 3.553776228609422e+00],
          100: [7.999371597866850e-01, 7.821440900165618e-01, 6.400034136975451e+01, 3.537903100476577e+00]},
}
B5_SAMPLE = {  # value at 1-based (i=10, j=20)
    "J": {1: [-5.188240755509784e-07, 1.226156027064817e-06, 1.000000001439201e+00, 9.765625086747991e-01],
          10: [-3.510360180623907e-07, 1.445974050470896e-06, 9.999999085947795e-01, 9.765625919733665e-01]},
    "D": {1: [2.884302415903964e-13, -1.347126237323913e-13, 9.999999999999684e-01, -1.237122049222736e-06]},
}
B5_ENERGY = {  # centre-averaged squares (FLAG_DIAG_CENTRED), step: (KE, ME, PE)
    "J": {1: (9.817088371875, 12.10864898308, 6.905342208337e-06), 10: (9.733111026534, 12.14372941009, 5.012522757623e-02),
          100: (7.352756498650, 14.31074588128, 2.542293279159e-01)},
    "D": {1: (1.536061871909e-05, 5.429085132700e-01, 2.910895786928e-08), 10: (1.207491007963e-03, 5.413973298886e-01, 2.339319052705e-04),
          100: (1.518419349856e-02, 5.222286514277e-01, 5.232572223375e-03)},
}


@pytest.mark.parametrize("kind", ["J", "D"])
def test_survey_checksums(kind):
    g, cfg, U = make_case(kind, 64, flags=abi.FLAG_DIAG_CENTRED)
    O.fill_halos(cfg, U)
    n = 0
    for step in (0, 1, 10, 100):
        O.step(cfg, U, 0.01, step - n)
        n = step
        l2 = [np.sqrt((g.interior(U[k], k) ** 2).sum()) for k in range(4)]
        for k in range(4):
            ref = B5[kind][step][k]
            assert abs(l2[k] - ref) <= 2e-13 * max(ref, 1e-3), (step, k, l2[k], ref)
        if step in B5_SAMPLE[kind]:
            for k in range(4):
                ref = B5_SAMPLE[kind][step][k]
                v = U[k][20 + 2, 10 + 2]
                assert abs(v - ref) <= 1e-12 * max(abs(ref), 1e-9) + 1e-16, (step, k, v, ref)
        if step in B5_ENERGY[kind]:
            d = O.diagnostics(cfg, U)
            for key, ref in zip(("ke", "me", "pe"), B5_ENERGY[kind][step]):
                assert abs(d[key] - ref) <= 2e-12 * abs(ref) + 1e-18, (step, key, d[key], ref)


# ---- invariants ----------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["J", "D", "BJ", "BD"])
def test_mass_conservation_and_constant_A(kind):
    g, cfg, U = make_case(kind, 48, perturb=4)
    O.fill_halos(cfg, U)
    m0 = g.interior(U[abi.H], abi.H).sum()
    O.step(cfg, U, 0.005, 50)
    m1 = g.interior(U[abi.H], abi.H).sum()
    assert abs(m1 - m0) <= 5e-12 * m0               # flux form: sum(h) conserved to round-off
    assert np.isfinite(np.concatenate([u.ravel() for u in U])).all()
    if kind in ("J", "D"):                          # a constant magnetic potential stays constant (C6)
        g, cfg, U = make_case(kind, 48, perturb=4)
        U[abi.A][...] = 0.75
        O.fill_halos(cfg, U)
        O.step(cfg, U, 0.005, 20)
        assert np.abs(g.interior(U[abi.A], abi.A) - 0.75).max() < 1e-13


def test_bounded_walls_keep_v_zero_and_halos_mirror():
    g, cfg, U = make_case("BJ", 40, perturb=2)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.005, 10)
    v = U[abi.V]
    assert np.all(v[3] == 0.0) and np.all(v[3 + g.Ny] == 0.0)        # v(j=1) = v(j=Ny+1) = 0
    h = U[abi.H]
    for k in (1, 2, 3):
        assert np.array_equal(h[3 - k], h[2 + k]) and np.array_equal(h[2 + g.Ny + k], h[3 + g.Ny - k])
    A = U[abi.A]
    assert np.allclose(A[2] - A[3], 0.05 * g.dy, atol=1e-15)         # A[0] = A[1] - gamma*dy, gamma = -0.05


# B.3 (published energy traces): tests/test_published_traces.py, all twelve figures, machine-digitised.


def test_c10_probe_switches_act_on_the_wall_rows_only():
    """The C10 probe switches (tools/probe_c10.py, profiles/r02_c10_probe.md) change what they say they change:
    D1 leaves the 2nd and 3rd halo rows of centre fields untouched, VM mirrors v oddly beyond the wall, W3 changes
    tendencies only in the rows whose WENO5 footprint crosses a wall — and none of them touches a periodic grid."""
    from cases import make_case
    g, cfg, U = make_case("BJ", 40, Ny=24, perturb=3)
    base = [u.copy() for u in U]
    O.fill_halos(cfg, base)
    G0 = O.tendencies(cfg, base)
    for flag in (abi.FLAG_BC_DEPTH1, abi.FLAG_WALL_WENO3, abi.FLAG_V_MIRROR):
        c = abi.Config.from_buffer_copy(cfg)
        c.flags = flag
        V = [u.copy() for u in U]
        O.fill_halos(c, V)
        if flag == abi.FLAG_BC_DEPTH1:
            assert np.array_equal(V[abi.A][2], base[abi.A][2]) and np.all(V[abi.A][0:2, 3:-3] == 0.0)     # k = 1 filled, k = 2, 3 untouched
        if flag == abi.FLAG_V_MIRROR:
            assert np.array_equal(V[abi.V][2, 3:-3], -base[abi.V][4, 3:-3]) and np.any(V[abi.V][2] != 0.0)
        if flag == abi.FLAG_WALL_WENO3:
            G = O.tendencies(c, base)
            rows = np.unique(np.nonzero(G[abi.H][3:-3, 3:-3] != G0[abi.H][3:-3, 3:-3])[0])
            assert len(rows) and set(rows) <= {1, 2, g.Ny - 3, g.Ny - 2}, rows      # faces 3 and Ny-1: cells 2, 3 and Ny-2, Ny-1 (0-based 1, 2, Ny-3, Ny-2)
    gp, cfgp, Up = make_case("J", 40, Ny=24, perturb=3)
    ref = [u.copy() for u in Up]
    O.fill_halos(cfgp, ref)
    for flag in (abi.FLAG_BC_DEPTH1, abi.FLAG_WALL_WENO3, abi.FLAG_V_MIRROR):
        c = abi.Config.from_buffer_copy(cfgp)
        c.flags = flag
        V = [u.copy() for u in Up]
        O.fill_halos(c, V)
        O.step(c, V, 0.004, 1)
        W = [u.copy() for u in ref]
        O.step(cfgp, W, 0.004, 1)
        assert all(np.array_equal(a, b) for a, b in zip(V, W))
