"""CPU-side checks: the C-ABI library loads and exports every symbol include/swmhd.h declares,
struct layouts agree with the C compiler, there is no CPU fallback, and the host-side helpers
(grid nodes, set!, slab decomposition) behave like the reference's Oceananigans surface."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from swmhd_b200 import abi
from swmhd_b200.grids import RectilinearGrid, Periodic, Bounded, Flat
from swmhd_b200 import distributed as D

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    txt = (ROOT / "include" / "swmhd.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(swmhd_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    declared = header_functions()
    assert len(declared) >= 20
    bound = {name for name, _, _ in abi.SYMBOLS}
    assert set(declared) == bound, set(declared) ^ bound
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.swmhd_abi_version() == abi.ABI_VERSION


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "swmhd.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(swmhd_config), sizeof(swmhd_diag), offsetof(swmhd_config, dx), offsetof(swmhd_config, A_grad_south),'
                   'offsetof(swmhd_config, slab_j0), offsetof(swmhd_config, n_gpus), offsetof(swmhd_config, device_ids));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == C.sizeof(abi.Config)
    assert int(out[1]) == C.sizeof(abi.Diag)
    assert int(out[2]) == abi.Config.dx.offset
    assert int(out[3]) == abi.Config.A_grad_south.offset
    assert int(out[4]) == abi.Config.slab_j0.offset
    assert int(out[5]) == abi.Config.n_gpus.offset
    assert int(out[6]) == abi.Config.device_ids.offset


def test_no_cpu_fallback_and_argument_errors():
    lib = abi.load_library()
    h = C.c_void_p()
    bad = abi.make_config(64, 64)
    bad.Hx = 2
    assert lib.swmhd_create(C.byref(bad), C.byref(h)) == abi.ERR_ARG
    bad = abi.make_config(64, 64, flags=abi.FLAG_WENO_JS)
    assert lib.swmhd_create(C.byref(bad), C.byref(h)) == abi.ERR_ARG
    assert b"oracle-only" in lib.swmhd_last_error(None)
    import torch
    if not torch.cuda.is_available():
        rc = lib.swmhd_create(C.byref(abi.make_config(64, 64)), C.byref(h))
        assert rc == abi.ERR_NODEVICE and not h.value
        assert b"no CPU fallback" in lib.swmhd_last_error(None)
        from swmhd_b200.context import Context, SwmhdError
        with pytest.raises(SwmhdError):
            Context(abi.make_config(64, 64))


def test_product_package_never_imports_the_oracle():
    """Nothing under swmhd_b200/ may import, link, include or dlopen anything under oracle/."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|pyoracle|libswmhd_oracle|swmhd_oracle_[a-z]+\s*\(|#include\s*[\"<].*oracle", re.M)
    for p in (ROOT / "swmhd_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".h", ".cuh"):
            m = pat.search(p.read_text())
            assert m is None, (p, m.group(0))
    out = subprocess.run(["nm", "-D", "--undefined-only", str(abi.library_path())], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_grid_nodes_and_set():
    g = RectilinearGrid(size=(64, 32), x=(-5, 5), y=(-5, 5), topology=(Periodic, Periodic, Flat))
    assert g.dx == 10 / 64 and g.dy == 10 / 32
    xs, ys = g.nodes(abi.U)          # u at (Face, Center)
    assert xs[0] == -5.0 and abs(ys[0] - (-5 + 0.5 * g.dy)) < 1e-15
    xs, ys = g.nodes(abi.V)          # v at (Center, Face)
    assert abs(xs[0] - (-5 + 0.5 * g.dx)) < 1e-15 and ys[0] == -5.0
    p = g.new_parent(abi.H)
    assert p.shape == (32 + 6, 64 + 6)
    g.set_interior(p, abi.H, lambda x, y, z: x + 10 * y)
    xs, ys = g.nodes(abi.H)
    assert p[3, 3] == xs[0] + 10 * ys[0] and p[3 + 31, 3 + 63] == xs[63] + 10 * ys[31]
    assert p[0].sum() == 0 and p[:, 0].sum() == 0          # halos untouched by set!
    gb = RectilinearGrid(size=(64, 32), x=(-5, 5), y=(-5, 5), topology=(Periodic, Bounded, Flat))
    assert gb.parent_shape(abi.V) == (32 + 7, 70) and gb.parent_shape(abi.U) == (38, 70)
    with pytest.raises(ValueError):
        RectilinearGrid(size=(8, 8), x=(0, 1), y=(0, 1), topology=(Bounded, Periodic, Flat))


def test_slab_decomposition_helpers():
    assert D.split_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert sum(n for _, n in D.split_rows(16384, 8)) == 16384
    assert D.neighbours(0, 4, True) == (3, 1) and D.neighbours(3, 4, True) == (2, 0)
    assert D.neighbours(0, 4, False) == (None, 1) and D.neighbours(3, 4, False) == (2, None)
    cfg = abi.make_config(64, 100)
    c2 = D.slab_config(cfg, 2, 3, device=1)
    assert (c2.slab_j0, c2.slab_ny, c2.rank, c2.world, c2.device) == (67, 33, 2, 3, 1)
    assert cfg.world == 1 and cfg.slab_ny == 100        # the global config is not modified


def test_models_validate_reference_options():
    """The host mirror accepts exactly the two model set-ups the reference builds."""
    from swmhd_b200 import models as M
    g = M.RectilinearGrid(size=(64, 64), x=(-5, 5), y=(-5, 5), topology=(M.Periodic, M.Periodic, M.Flat))
    common = dict(grid=g, timestepper=":RungeKutta3", mass_advection=M.WENO5(), tracer_advection=M.WENO5(),
                  gravitational_acceleration=9.81, coriolis=M.FPlane(f=1), tracers=(":A",))
    with pytest.raises(ValueError):     # wrong forcing for the formulation
        M.ShallowWaterModel(momentum_advection=M.WENO5(vector_invariant=M.VelocityStencil()),
                            forcing=dict(uh=M.Forcing(M.div_lorentz_x, discrete_form=True), vh=M.Forcing(M.div_lorentz_y, discrete_form=True)),
                            formulation=M.VectorInvariantFormulation(), **common)
    with pytest.raises(ValueError):     # VorticityStencil is not what the reference selects
        M.ShallowWaterModel(momentum_advection=M.WENO5(vector_invariant=M.VorticityStencil()),
                            forcing=dict(u=M.Forcing(M.lorentz_force_func_x, discrete_form=True), v=M.Forcing(M.lorentz_force_func_y, discrete_form=True)),
                            formulation=M.VectorInvariantFormulation(), **common)
    with pytest.raises(ValueError):
        M.ShallowWaterModel(momentum_advection=M.WENO5(), forcing={}, formulation=M.ConservativeFormulation(), **{**common, "timestepper": "QuasiAdamsBashforth2"})


def test_run_plans_the_aligned_dt_sequence_between_events():
    """run!(simulation) with a TimeInterval(0.1) writer (SWMHD_example.jl:81-84): Simulation._plan_batch replays upstream's
    aligned_time_step on a copy of the clock, so that every batch of steps between two events is ONE swmhd_step_seq call.
    Checked against a stand-in context that ticks the clock like the library (no GPU needed)."""
    from swmhd_b200 import models as M

    class FakeCtx:
        def __init__(self):
            self.time, self.iteration, self.calls = 0.0, 0, []

        def step_seq(self, dts, diag=False):
            self.calls.append(list(dts))
            for dt in dts:          # tick! per RK3 stage: (8/15, 2/15, 1/3) dt
                self.time = ((self.time + (8.0 / 15.0) * dt) + (5.0 / 12.0 - 17.0 / 60.0) * dt) + (3.0 / 4.0 - 5.0 / 12.0) * dt
                self.iteration += 1

    class FakeModel:
        def __init__(self):
            self.ctx = FakeCtx()
            self.clock = self.ctx

        def _mark_stale(self):
            pass

    m = FakeModel()
    sim = M.Simulation(m, dt=0.03, stop_time=0.35)
    fired = []

    class W:
        schedule = M.TimeInterval(0.1)

        def write(self, s):
            fired.append(s.model.clock.time)

    sim.output_writers["fields"] = W()
    M.run_b(sim)
    assert np.allclose(fired, [0.0, 0.1, 0.2, 0.3], atol=1e-12)
    assert abs(m.clock.time - 0.35) < 1e-12
    # three full steps and the remainder up to each output time, one call per output interval
    assert [len(c) for c in m.ctx.calls] == [4, 4, 4, 2]
    for c in m.ctx.calls[:3]:
        assert np.allclose(c[:3], 0.03) and abs(sum(c) - 0.1) < 1e-12 and 0 < c[3] < 0.03
    assert np.allclose(m.ctx.calls[3], [0.03, 0.02], atol=1e-12)
    # iteration schedules batch too: every 5 iterations, stop after 12
    m2 = FakeModel()
    sim2 = M.Simulation(m2, dt=0.01, stop_iteration=12)
    seen = []
    sim2.callbacks["p"] = M.Callback(lambda s: seen.append(s.model.clock.iteration), M.IterationInterval(5))
    M.run_b(sim2)
    assert seen == [0, 5, 10] and [len(c) for c in m2.ctx.calls] == [5, 5, 2]
