"""Round-2 GPU tests: longer FAST parity at a configuration size, perturbed divergence cases, the asynchronous
field writer, ABI argument/state errors, concurrent contexts, and the multi-GPU paths that live inside
libswmhd_cuda.so (single-process n_gpus contexts; skipped on a one-GPU box, the multi-process ring is proven by
bench.py's slab_parity object and tools/gpu_multi_test.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from swmhd_b200 import abi
from swmhd_b200.context import Context, SwmhdError
from oracle import pyoracle as O
from cases import make_case, rel_l2

pytestmark = pytest.mark.gpu
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0


def run_gpu(cfg, U, dt, nsteps):
    ctx = Context(cfg)
    ctx.set_state(U)
    ctx.fill_halos()
    ctx.step(dt, nsteps)
    out = ctx.get_state()
    ctx.close()
    return out


@pytest.mark.parametrize("kind", ["J", "D"])
def test_fast_100_steps_at_1024(kind):
    """north_star: rel-L2 <= 1e-9 after 1000 steps; at BASELINE config-2 size the bar used here is 1e-10 after
    100 steps against the CPU oracle (configs 2-3)."""
    N = 1024
    g, cfg, U = make_case(kind, N, arith=abi.ARITH_FAST)
    dt = 0.01 * 64 / N
    Ug = run_gpu(cfg, U, dt, 100)
    O.set_threads(64)
    O.fill_halos(cfg, U)
    O.step(cfg, U, dt, 100)
    for k in range(4):
        err = rel_l2(g, Ug[k], U[k], k)
        assert err <= 1e-10, f"field {k}: rel L2 {err:.3e}"


@pytest.mark.parametrize("kind", ["D", "GD", "BD"])
@pytest.mark.parametrize("N", [64, 100])
def test_fast_one_step_perturbed_divergence(kind, N):
    """One FAST step of the divergence form with non-zero momentum (smooth 1e-3 perturbation of all four fields)."""
    g, cfg, U = make_case(kind, N, arith=abi.ARITH_FAST, perturb=11)
    Ug = run_gpu(cfg, U, 0.01 * 64 / N, 1)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.01 * 64 / N, 1)
    for k in range(4):
        err = rel_l2(g, Ug[k], U[k], k)
        assert err <= 1e-12, f"field {k}: rel L2 {err:.3e}"


@pytest.mark.parametrize("kind", ["J", "D"])
def test_async_outputs_overlap_stepping(kind):
    """swmhd_get_outputs_async: the TimeInterval(0.1) field writer (SWMHD_example.jl:81-84) overlapped with the
    next steps.  Stepping continues bit-identically while the copy is in flight, and the host arrays hold the
    outputs of the state at the time of the call."""
    g, cfg, U = make_case(kind, 256, Ny=192, perturb=13)
    ref = Context(cfg); ref.set_state(U); ref.fill_halos(); ref.step(0.002, 2)
    u0, v0, s0 = ref.get_outputs(); A0 = ref.get_field(abi.A)
    ref.step(0.002, 3); Uend = ref.get_state(); ref.close()
    ctx = Context(cfg); ctx.set_state(U); ctx.fill_halos(); ctx.step(0.002, 2)
    bufs = [torch.zeros(ctx.field_shape(k), dtype=torch.float64).pin_memory().numpy() for k in (abi.U, abi.V, abi.U, abi.A)]
    ctx.get_outputs_async(*bufs)
    ctx.step(0.002, 3)                      # queued behind nothing: the copy runs on its own stream
    ctx.outputs_wait()
    Uasync = ctx.get_state()
    # a second request while nothing is pending, then wait twice (idempotent)
    ctx.get_outputs_async(*bufs); ctx.outputs_wait(); ctx.outputs_wait()
    ctx.close()
    for k in range(4):
        assert np.array_equal(Uasync[k], Uend[k]), k


@pytest.mark.parametrize("kind", ["J", "D"])
def test_async_outputs_values(kind):
    g, cfg, U = make_case(kind, 128, Ny=96, perturb=13)
    ctx = Context(cfg); ctx.set_state(U); ctx.fill_halos(); ctx.step(0.002, 2)
    u0, v0, s0 = ctx.get_outputs(); A0 = ctx.get_field(abi.A)
    bufs = [np.zeros(ctx.field_shape(k)) for k in (abi.U, abi.V, abi.U, abi.A)]     # pageable host memory works too
    ctx.get_outputs_async(*bufs)
    ctx.step(0.002, 2)
    ctx.outputs_wait()
    ctx.close()
    for got, want in zip(bufs, (u0, v0, s0, A0)):
        assert np.array_equal(got, want)


def test_tendencies_argument_and_state_errors():
    g, cfg, U = make_case("BJ", 40, Ny=24)
    ctx = Context(cfg); ctx.set_state(U); ctx.fill_halos()
    G = [ctx.new_parent(k) for k in range(4)]
    dp = C.POINTER(C.c_double)
    arr = (dp * 4)(*[x.ctypes.data_as(dp) for x in G])
    # v has one row more in a Bounded-y grid: sizing every buffer from field 0 is rejected, not overrun
    assert ctx.lib.swmhd_tendencies(ctx._h, arr, G[0].size) == abi.ERR_ARG
    assert ctx.lib.swmhd_tendencies(ctx._h, arr, G[1].size) == abi.OK
    arr_null = (dp * 4)(G[0].ctypes.data_as(dp), None, G[2].ctypes.data_as(dp), G[3].ctypes.data_as(dp))
    before = ctx.get_state()
    assert ctx.lib.swmhd_tendencies(ctx._h, arr_null, G[1].size) == abi.ERR_ARG
    # between stage 1 and stage 2 the tendency buffers hold G^-: tendencies / outputs are refused
    ctx.substage(0.002, 1)
    with pytest.raises(SwmhdError) as e:
        ctx.tendencies()
    assert e.value.code == abi.ERR_STATE
    with pytest.raises(SwmhdError) as e:
        ctx.get_outputs()
    assert e.value.code == abi.ERR_STATE
    with pytest.raises(SwmhdError) as e:
        ctx.step(0.002, 1)
    assert e.value.code == abi.ERR_STATE
    ctx.substage(0.002, 2); ctx.substage(0.002, 3)
    ctx.tendencies()
    ctx.close()
    ref = Context(cfg); ref.set_state(U); ref.fill_halos(); ref.step(0.002, 1); want = ref.get_state(); ref.close()
    ctx = Context(cfg); ctx.set_state(U); ctx.fill_halos()
    for s in (1, 2, 3):
        ctx.substage(0.002, s)
    got = ctx.get_state(); ctx.close()
    for k in range(4):
        assert np.array_equal(got[k], want[k])


def test_two_contexts_on_one_gpu_reduce_independently():
    """Each context owns the ticket word of its final reduction: interleaved fused-diagnostic steps of two contexts
    on one GPU (own non-blocking streams, so they overlap) give what each gives alone."""
    cases = [make_case("J", 160, Ny=128, perturb=3), make_case("D", 96, Ny=200, perturb=4)]
    alone = []
    for g, cfg, U in cases:
        c = Context(cfg); c.set_state(U); c.fill_halos(); alone.append(c.step_diag(0.002, 6)); c.close()
    ctxs = []
    for g, cfg, U in cases:
        c = Context(cfg); c.set_state(U); c.fill_halos(); ctxs.append(c)
    both = [[], []]
    for n in range(6):
        for q, c in enumerate(ctxs):
            both[q] += c.step_diag(0.002, 1)
    for c in ctxs:
        c.close()
    for q in range(2):
        for a, b in zip(alone[q], both[q]):
            assert a == b


@pytest.mark.skipif(NGPU < 2, reason="needs two GPUs in one process")
@pytest.mark.parametrize("kind", ["J", "D", "BJ", "BD"])
@pytest.mark.parametrize("ng", [2, 4, 8])
def test_single_process_n_gpus_bit_identical(kind, ng):
    """cfg.n_gpus: one context drives ng devices (ncclCommInitAll inside the library); set/get move GLOBAL parent
    arrays; fields after 4 steps are bit-identical to the one-GPU run, halos included."""
    if ng > NGPU:
        pytest.skip("not enough GPUs")
    g, cfg, U = make_case(kind, 128, Ny=200, perturb=31)
    ref = Context(cfg); ref.set_state(U); ref.fill_halos()
    tr_ref = ref.step_diag(0.002, 4); Uref = ref.get_state(); outs_ref = ref.get_outputs(); dref = ref.diagnostics(); ref.close()
    cfg2 = abi.Config.from_buffer_copy(cfg)
    cfg2.n_gpus = ng
    for d in range(ng):
        cfg2.device_ids[d] = d
    mg = Context(cfg2); mg.set_state(U); mg.fill_halos()
    tr = mg.step_diag(0.002, 4); Um = mg.get_state(); outs = mg.get_outputs(); dm = mg.diagnostics()
    assert abs(mg.time - 4 * 0.002) < 1e-15 and mg.iteration == 4
    mg.close()
    for k in range(4):
        assert np.array_equal(Um[k], Uref[k]), k
    for a, b in zip(outs, outs_ref):
        assert np.array_equal(a, b)
    keys = ("ke", "me", "pe", "sum_h", "max_abs_u", "max_abs_A", "min_h", "max_abs_div_hB")
    for x, y in zip(tr + [dm], tr_ref + [dref]):
        for key in keys:
            assert abs(x[key] - y[key]) <= 1e-13 * max(1.0, abs(y[key])), key


@pytest.mark.skipif(NGPU < 2, reason="needs a second GPU")
def test_context_on_second_device():
    """Launch configuration (shared-memory opt-in, L2 prefetch distance) is per device, not per process."""
    g, cfg, U = make_case("J", 128, Ny=96, perturb=5)
    a = run_gpu(cfg, U, 0.002, 2)
    cfg1 = abi.Config.from_buffer_copy(cfg)
    cfg1.device = 1
    b = run_gpu(cfg1, U, 0.002, 2)
    for k in range(4):
        assert np.array_equal(a[k], b[k])


@pytest.mark.parametrize("kind", ["J", "D"])
def test_step_seq_equals_the_same_steps_one_by_one(kind):
    """swmhd_step_seq: a batch with a per-step dt (the aligned sequence up to an output time) is bit-identical to the
    same steps issued one call at a time, clock included; bad sequences are refused before anything runs."""
    g, cfg, U = make_case(kind, 96, Ny=80, perturb=21)
    dts = [0.004, 0.004, 0.004, 0.0013, 0.004, 0.0007]
    a = Context(cfg); a.set_state(U); a.fill_halos()
    for dt in dts:
        a.step(dt, 1)
    Ua, ta, ia = a.get_state(), a.time, a.iteration
    da = a.diagnostics(); a.close()
    b = Context(cfg); b.set_state(U); b.fill_halos()
    tr = b.step_seq(dts, diag=True)
    Ub, tb, ib = b.get_state(), b.time, b.iteration
    with pytest.raises(SwmhdError) as e:
        b.step_seq([0.004, 0.0, 0.004])
    assert e.value.code == abi.ERR_ARG and b.iteration == ib
    b.close()
    assert ta == tb and ia == ib == len(dts) and len(tr) == len(dts)
    for k in range(4):
        assert np.array_equal(Ua[k], Ub[k])


@pytest.mark.parametrize("kind,N,Ny", [("J", 256, 1024), ("D", 248, 776), ("BJ", 128, 520), ("BD", 128, 512), ("J", 96, 80)])
def test_upload_step_equals_set_then_step(kind, N, Ny):
    """swmhd_upload_step (set! + time_step!, the upload pipelined with stage 1 by row bands) is bit-identical to
    set_field x 4 + fill_halos + step_diag(1) — halos of the host arrays need not be filled — and leaves the clock alike."""
    g, cfg, U = make_case(kind, N, Ny=Ny, perturb=41)
    a = Context(cfg); a.set_state(U); a.fill_halos(); da = a.step_diag(0.002, 1)[0]; a.step(0.002, 1)
    Ua, ta = a.get_state(), a.time; a.close()
    pinned = [torch.from_numpy(u.copy()).pin_memory().numpy() for u in U]
    b = Context(cfg)
    db = b.upload_step(pinned, 0.002, diag=True); b.step(0.002, 1)
    Ub, tb = b.get_state(), b.time
    db2 = b.upload_step(pinned, 0.002, diag=True)        # again, from a context that has history
    b.close()
    assert ta == tb
    for k in range(4):
        assert np.array_equal(Ua[k], Ub[k]), k
    for key in ("ke", "me", "pe", "sum_h", "max_abs_u", "max_abs_A", "min_h", "max_abs_div_hB"):
        assert abs(da[key] - db[key]) <= 1e-13 * max(1.0, abs(da[key])), key
        assert db2[key] == db[key], key


@pytest.mark.parametrize("kind", ["J", "D", "BJ"])
def test_fast_1000_steps_at_256(kind):
    """north_star: rel-L2 <= 1e-9 per field after 1000 steps, here at 256^2 (16x the cells of the 64^2 case of
    test_gpu_parity.py::test_fast_1000_steps), periodic and Bounded-y."""
    N = 256
    g, cfg, U = make_case(kind, N, arith=abi.ARITH_FAST)
    dt = 0.01 * 64 / N
    Ug = run_gpu(cfg, U, dt, 1000)
    O.set_threads(64)
    O.fill_halos(cfg, U)
    O.step(cfg, U, dt, 1000)
    for k in range(4):
        err = rel_l2(g, Ug[k], U[k], k)
        assert err <= 1e-9, f"field {k}: rel L2 {err:.3e}"


def test_no_kernel_writes_outside_its_arrays(monkeypatch):
    """compute-sanitizer is closed on this pool: with SWMHD_GUARD=1 every device array sits between two 32 KB guard zones;
    after stepping ragged, odd, tiny, wide and Bounded-y grids through every kernel route (row-blocked, one thread per cell,
    fused diagnostics, tendencies, outputs, upload pipeline) all sentinels must be intact."""
    monkeypatch.setenv("SWMHD_GUARD", "1")
    rng = np.random.default_rng(7)
    cases = [("J", 8, 8), ("D", 8, 9), ("BJ", 10, 12), ("BD", 34, 9), ("J", 70, 50), ("D", 65, 43), ("BJ", 64, 40), ("BD", 96, 80),
             ("J", 130, 33), ("D", 250, 264), ("J", 256, 1024), ("BD", 128, 520)]
    for _ in range(6):
        cases.append((["J", "D", "BJ", "BD"][int(rng.integers(4))], int(rng.integers(8, 200)), int(rng.integers(8, 300))))
    for kind, N, Ny in cases:
        for arith in (abi.ARITH_FAST, abi.ARITH_STRICT):
            g, cfg, U = make_case(kind, N, Ny=Ny, arith=arith, perturb=3)
            c = Context(cfg); c.set_state(U); c.fill_halos()
            c.step(0.002, 2); c.step_diag(0.002, 2); c.tendencies(); c.get_outputs(); c.diagnostics()
            c.upload_step(U, 0.002, diag=True)
            bufs = [np.zeros(c.field_shape(k)) for k in (abi.U, abi.V, abi.U, abi.A)]
            c.get_outputs_async(*bufs); c.outputs_wait()
            c.check_guards()
            c.close()
    monkeypatch.delenv("SWMHD_GUARD")
    g, cfg, U = make_case("J", 32, Ny=16)
    c = Context(cfg)
    with pytest.raises(SwmhdError):
        c.check_guards()                      # created without guards: says so instead of reporting a clean bill
    c.close()
