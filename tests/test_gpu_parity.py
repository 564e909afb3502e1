"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same inputs.

Bars (north_star): STRICT arithmetic bit-identical to the oracle; FAST arithmetic relative
L2 <= 1e-12 per field after one step.
"""
import numpy as np
import pytest

from swmhd_b200 import abi
from swmhd_b200.context import Context
from oracle import pyoracle as O
from cases import make_case, rel_l2

pytestmark = pytest.mark.gpu

TOL_1STEP = 1e-12


def run_gpu(cfg, U, dt, nsteps):
    ctx = Context(cfg)
    ctx.set_state(U)
    ctx.fill_halos()
    ctx.step(dt, nsteps)
    out = ctx.get_state()
    ctx.close()
    return out


@pytest.mark.parametrize("kind", ["J", "D", "G", "GD"])
@pytest.mark.parametrize("N", [64, 100])
def test_strict_bit_identical_one_step(kind, N):
    g, cfg, U = make_case(kind, N, arith=abi.ARITH_STRICT, perturb=7)
    Ug = run_gpu(cfg, U, 0.01 * 64 / N, 1)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.01 * 64 / N, 1)
    for k in range(4):
        assert np.array_equal(Ug[k], U[k]), f"field {k}: max abs diff {np.abs(Ug[k]-U[k]).max():.3e}"


@pytest.mark.parametrize("kind", ["J", "D", "G", "GD"])
@pytest.mark.parametrize("N", [64, 100])
def test_fast_one_step_tolerance(kind, N):
    # J, G perturbed; D, GD as the script sets them (uh = vh = 0: the tight case of SURVEY B.6); the perturbed
    # divergence cases are test_fast_one_step_perturbed_divergence (tests/test_gpu_round2.py)
    g, cfg, U = make_case(kind, N, arith=abi.ARITH_FAST, perturb=7 if kind in ("J", "G") else None)
    Ug = run_gpu(cfg, U, 0.01 * 64 / N, 1)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.01 * 64 / N, 1)
    for k in range(4):
        err = rel_l2(g, Ug[k], U[k], k)
        assert err <= TOL_1STEP, f"field {k}: rel L2 {err:.3e}"


@pytest.mark.parametrize("kind", ["J", "D"])
def test_tendencies_strict(kind):
    g, cfg, U = make_case(kind, 64, arith=abi.ARITH_STRICT, perturb=3)
    O.fill_halos(cfg, U)
    Go = O.tendencies(cfg, U)
    ctx = Context(cfg)
    ctx.set_state(U)
    Gg = ctx.tendencies()
    ctx.close()
    for k in range(4):
        assert np.array_equal(g.interior(Gg[k], k), g.interior(Go[k], k)), f"G[{k}]"


@pytest.mark.parametrize("kind", ["J", "D", "BJ", "BD"])
def test_halo_fill_matches_oracle(kind):
    g, cfg, U = make_case(kind, 72, Ny=40, arith=abi.ARITH_STRICT, perturb=11)
    if kind.startswith("B"):       # set! may leave anything on the wall rows of v: the fill must zero them
        U[abi.V][3] = 0.3
        U[abi.V][3 + g.Ny] = -0.2
    Uo = [u.copy() for u in U]
    O.fill_halos(cfg, Uo)
    for rep in range(3):           # repeated: the fill must be race-free
        ctx = Context(cfg)
        ctx.set_state(U)
        ctx.fill_halos()
        Ug = ctx.get_state()
        ctx.close()
        for k in range(4):
            assert np.array_equal(Ug[k], Uo[k]), (rep, k)


@pytest.mark.parametrize("kind", ["J", "D"])
def test_diagnostics_match_oracle(kind):
    g, cfg, U = make_case(kind, 64, perturb=5)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.01, 3)
    do = O.diagnostics(cfg, U)
    ctx = Context(cfg)
    ctx.set_state(U)
    dg = ctx.diagnostics()
    ctx.close()
    for key in ("ke", "me", "pe", "total", "sum_h"):
        assert abs(dg[key] - do[key]) <= 1e-13 * max(1.0, abs(do[key])), key
    for key in ("max_abs_u", "max_abs_A", "min_h"):
        assert dg[key] == do[key], key
    assert abs(dg["max_abs_div_hB"] - do["max_abs_div_hB"]) < 1e-13
    assert dg["all_finite"] == 1


@pytest.mark.parametrize("kind,form_tol", [("J", 1e-9), ("D", 1e-9)])
def test_fast_1000_steps(kind, form_tol):
    g, cfg, U = make_case(kind, 64, arith=abi.ARITH_FAST)
    Ug = run_gpu(cfg, U, 0.01, 1000)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.01, 1000)
    for k in range(4):
        err = rel_l2(g, Ug[k], U[k], k)
        assert err <= form_tol, f"field {k}: rel L2 {err:.3e}"


@pytest.mark.parametrize("kind", ["J", "D"])
@pytest.mark.parametrize("arith", [abi.ARITH_FAST, abi.ARITH_STRICT])
def test_fused_step_diagnostics(kind, arith):
    """swmhd_step_diag: the diagnostics fused into the stage-1 kernel equal the oracle's
    diagnostics of the state at the start of every step, and do not perturb the step."""
    g, cfg, U = make_case(kind, 100, Ny=72, arith=arith, perturb=9)
    ctx = Context(cfg)
    ctx.set_state(U)
    ctx.fill_halos()
    dg = ctx.step_diag(0.004, 3)
    Ug = ctx.get_state()
    ctx.close()
    ctx2 = Context(cfg)
    ctx2.set_state(U)
    ctx2.fill_halos()
    ctx2.step(0.004, 3)
    Up = ctx2.get_state()
    ctx2.close()
    for k in range(4):
        assert np.array_equal(Ug[k], Up[k])
    O.fill_halos(cfg, U)
    for n in range(3):
        do = O.diagnostics(cfg, U)
        for key in ("ke", "me", "pe", "total", "sum_h"):
            assert abs(dg[n][key] - do[key]) <= 2e-13 * max(1.0, abs(do[key])), (n, key, dg[n][key], do[key])
        for key in ("max_abs_A", "min_h"):
            # extrema are exact reductions of the state: equal whenever the states are (always for STRICT and for
            # the initial state; a FAST state differs from the oracle's by ulps from the first step on)
            if arith == abi.ARITH_STRICT or n == 0:
                assert dg[n][key] == do[key], (n, key)
            else:
                assert abs(dg[n][key] - do[key]) <= 1e-14 * abs(do[key]), (n, key)
        assert abs(dg[n]["max_abs_u"] - do["max_abs_u"]) <= 1e-15 * max(1.0, do["max_abs_u"])
        assert abs(dg[n]["max_abs_div_hB"] - do["max_abs_div_hB"]) < 1e-13
        O.step(cfg, U, 0.004, 1)


@pytest.mark.parametrize("kind", ["J", "D"])
def test_ragged_and_odd_sizes_strict(kind):
    """Odd Nx (row pitch not a multiple of 16 B -> plain-load path instead of TMA) and sizes that
    are not multiples of the 32x8 tile."""
    for (N, Ny) in [(65, 43), (70, 50)]:
        g, cfg, U = make_case(kind, N, Ny=Ny, arith=abi.ARITH_STRICT, perturb=13)
        Ug = run_gpu(cfg, U, 0.004, 2)
        O.fill_halos(cfg, U)
        O.step(cfg, U, 0.004, 2)
        for k in range(4):
            assert np.array_equal(Ug[k], U[k]), (N, Ny, k)


@pytest.mark.parametrize("kind", ["J", "D", "BJ", "BD"])
def test_slab_api_single_rank_equals_step(kind):
    """The multi-GPU entry points (edges -> interior -> finish on two streams) on one slab must give
    exactly what swmhd_step gives."""
    from swmhd_b200.distributed import SlabModel
    g, cfg, U = make_case(kind, 96, Ny=80, arith=abi.ARITH_FAST, perturb=17)
    a = run_gpu(cfg, U, 0.004, 3)
    sm = SlabModel(cfg, rank=0, world=1, device=0, host_exchange=True)
    sm.set_state(U)
    sm.fill_halos()
    tr = sm.step_diag(0.004, 3)
    sm.synchronize()
    b = sm.get_state()
    ref = Context(cfg)
    ref.set_state(U)
    ref.fill_halos()
    tr_ref = ref.step_diag(0.004, 3)
    ref.close()
    for x, y in zip(tr, tr_ref):
        for key in ("ke", "me", "pe", "sum_h", "max_abs_u", "max_abs_A", "min_h", "max_abs_div_hB"):
            assert abs(x[key] - y[key]) <= 1e-13 * max(1.0, abs(y[key])), key
    assert abs(sm.ctx.time - 3 * 0.004) < 1e-15 and sm.ctx.iteration == 3
    sm.close()
    for k in range(4):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("kind", ["BJ", "BD"])
@pytest.mark.parametrize("arith", [abi.ARITH_STRICT, abi.ARITH_FAST])
def test_bounded_y_matches_oracle(kind, arith):
    g, cfg, U = make_case(kind, 72, Ny=56, arith=arith, perturb=19)
    Ug = run_gpu(cfg, U, 0.004, 5)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.004, 5)
    for k in range(4):
        if arith == abi.ARITH_STRICT:
            assert np.array_equal(Ug[k], U[k]), (k, np.abs(Ug[k] - U[k]).max())
        else:
            assert rel_l2(g, Ug[k], U[k], k) <= 5e-12, k


@pytest.mark.parametrize("kind", ["J", "D"])
def test_device_side_writer_outputs(kind):
    """swmhd_get_outputs: u, v, s = sqrt(u^2 + v^2) of the reference's field writer
    (SWMHD_example.jl:67-69,81-84; divergence_sw_mhd.jl:63-66) vs a numpy restatement."""
    g, cfg, U = make_case(kind, 72, Ny=48, perturb=23)
    ctx = Context(cfg)
    ctx.set_state(U)
    ctx.fill_halos()
    ctx.step(0.004, 2)
    St = ctx.get_state()
    u_o, v_o, s_o = ctx.get_outputs()
    St2 = ctx.get_state()
    ctx.step(0.004, 1)                      # the staging buffers are free again: stepping still works
    ctx.close()
    for k in range(4):
        assert np.array_equal(St[k], St2[k])
    u, v, h = St[0], St[1], St[2]
    if kind == "D":
        uvel = u / (0.5 * (np.roll(h, 1, axis=1) + h))
        vvel = v / (0.5 * (np.roll(h, 1, axis=0) + h))
    else:
        uvel, vvel = u, v
    v2 = vvel ** 2
    ixf = 0.5 * (np.roll(v2, 1, axis=1) + v2)
    ixy = 0.5 * (ixf + np.roll(ixf, -1, axis=0))
    s = np.sqrt(uvel ** 2 + ixy)
    it = (slice(3, 3 + g.Ny), slice(3, 3 + g.Nx))
    assert np.allclose(u_o[it], uvel[it], rtol=1e-14, atol=0)
    assert np.allclose(v_o[it], vvel[it], rtol=1e-14, atol=0)
    assert np.allclose(s_o[it], s[it], rtol=1e-14, atol=1e-300)
    # halos of the outputs are the periodic images of their interiors (with_halos = true writers)
    assert np.array_equal(s_o[it][:, :3], s_o[3:3 + g.Ny, 3 + g.Nx:6 + g.Nx])
    assert np.array_equal(s_o[it][:3, :], s_o[3 + g.Ny:6 + g.Ny, 3:3 + g.Nx])


@pytest.mark.parametrize("kind", ["J", "D", "BJ"])
def test_tiny_grids_strict(kind):
    """Smallest legal grids (a tile is larger than the whole domain; TMA boxes hang over every edge)."""
    for (N, Ny) in [(8, 8), (16, 8), (10, 12), (34, 9)]:
        g, cfg, U = make_case(kind, N, Ny=Ny, arith=abi.ARITH_STRICT, perturb=29)
        Ug = run_gpu(cfg, U, 0.002, 3)
        O.fill_halos(cfg, U)
        O.step(cfg, U, 0.002, 3)
        for k in range(4):
            assert np.array_equal(Ug[k], U[k]), (N, Ny, k)


def test_error_paths():
    """Error behaviour of the C ABI: wrong buffer length, bad stage, bad field — no crash, an error code and a message."""
    from swmhd_b200.context import SwmhdError
    g, cfg, U = make_case("J", 32, Ny=16)
    ctx = Context(cfg)
    with pytest.raises(SwmhdError) as e:
        ctx.set_field(abi.U, np.zeros((5, 5)))
    assert e.value.code == abi.ERR_ARG and "length mismatch" in str(e.value)
    with pytest.raises(SwmhdError):
        ctx.substage(0.01, 4)
    with pytest.raises(SwmhdError):
        ctx.substage_interior(0.01, 1)          # edges not called first
    with pytest.raises(SwmhdError):
        ctx.arm_diag(5000)
    U[abi.H][5, 5] = np.nan
    ctx.set_state(U)
    with pytest.raises(SwmhdError) as e:
        ctx.diagnostics()
    assert e.value.code == abi.ERR_NONFINITE
    assert ctx.diagnostics(check_finite=False)["all_finite"] == 0
    ctx.close()


def test_fuzz_strict_bit_identity():
    """Seeded fuzz over sizes, formulations, topologies and data: STRICT arithmetic must stay bit-identical to the
    oracle (whole parent arrays, halos included) — catches value-dependent rounding differences."""
    rng = np.random.default_rng(20261018)
    for trial in range(16):
        kind = ["J", "D", "BJ", "BD"][trial % 4]
        N = int(rng.integers(8, 90))
        Ny = int(rng.integers(8, 70))
        seed = int(rng.integers(1, 10_000))
        nst = int(rng.integers(1, 4))
        g, cfg, U = make_case(kind, N, Ny=Ny, arith=abi.ARITH_STRICT, perturb=seed)
        if kind.startswith("B"):
            cfg.A_grad_south, cfg.A_grad_north = float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.1, 0.1))
        dt = float(rng.uniform(0.001, 0.004))
        Ug = run_gpu(cfg, U, dt, nst)
        O.fill_halos(cfg, U)
        O.step(cfg, U, dt, nst)
        for k in range(4):
            assert np.array_equal(Ug[k], U[k]), (trial, kind, N, Ny, seed, nst, k)


@pytest.mark.parametrize("kind", ["J", "D"])
def test_step_profile_variants_advance_like_step(kind):
    """swmhd_step_profile / swmhd_step_profile_diag (measurement entry points) run the same launches as swmhd_step."""
    g, cfg, U = make_case(kind, 96, arith=abi.ARITH_FAST, perturb=5)
    ref = run_gpu(cfg, [u.copy() for u in U], 0.004, 3)
    for diag in (False, True):
        ctx = Context(cfg)
        ctx.set_state([u.copy() for u in U])
        ctx.fill_halos()
        ms = ctx.step_profile(0.004, 3, diag=diag)
        out = ctx.get_state()
        ctx.close()
        assert len(ms) == 3 and all(m > 0 for m in ms)
        for k in range(4):
            assert np.array_equal(out[k], ref[k])
