"""The Python mirror of the reference's Oceananigans surface, end to end on the GPU."""
import numpy as np
import pytest

from swmhd_b200 import (RectilinearGrid, Periodic, Bounded, Flat, ShallowWaterModel, VectorInvariantFormulation,
                        ConservativeFormulation, WENO5, VelocityStencil, FPlane, Forcing, lorentz_force_func_x,
                        lorentz_force_func_y, div_lorentz_x, div_lorentz_y, set_b, run_b, time_step_b, Simulation,
                        Callback, IterationInterval, TimeInterval, MemoryOutputWriter, abi)
from oracle import pyoracle as O
from cases import make_case, rel_l2

pytestmark = pytest.mark.gpu


def build_jacobian(N=64, arithmetic="strict"):
    grid = RectilinearGrid(size=(N, N), x=(-5, 5), y=(-5, 5), topology=(Periodic, Periodic, Flat))
    model = ShallowWaterModel(grid=grid, timestepper=":RungeKutta3",
                              momentum_advection=WENO5(vector_invariant=VelocityStencil()),
                              mass_advection=WENO5(), tracer_advection=WENO5(), gravitational_acceleration=9.81,
                              coriolis=FPlane(f=1), tracers=(":A",),
                              forcing=dict(u=Forcing(lorentz_force_func_x, discrete_form=True),
                                           v=Forcing(lorentz_force_func_y, discrete_form=True)),
                              formulation=VectorInvariantFormulation(), arithmetic=arithmetic)
    set_b(model, u=lambda x, y, z: 5 * y * np.exp(-(x ** 2 + y ** 2)), v=lambda x, y, z: -5 * x * np.exp(-(x ** 2 + y ** 2)),
          h=lambda x, y, z: 1.0, A=lambda x, y, z: 0.5 * np.abs(y))
    return model


def test_example_script_flow_matches_oracle():
    """SWMHD_example.jl:10-100 in Python: model, set!, Simulation with an every-iteration progress callback
    and an energy writer, run! for 25 iterations; the final state equals the oracle's."""
    model = build_jacobian()
    sim = Simulation(model, dt=0.01, stop_iteration=25)
    seen = []
    sim.callbacks["progress"] = Callback(lambda s: seen.append((s.model.clock.iteration, float(np.abs(s.model.solution.u.interior).max()))), IterationInterval(1))
    sim.output_writers["energies"] = MemoryOutputWriter(model, dict(e=lambda m: m.diagnostics()), IterationInterval(5))
    run_b(sim)
    assert [i for i, _ in seen] == list(range(26))
    assert sim.output_writers["energies"].iterations == [0, 5, 10, 15, 20, 25]
    assert abs(model.clock.time - 0.25) < 1e-12 and model.clock.iteration == 25
    g, cfg, U = make_case("J", 64, arith=abi.ARITH_STRICT)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.01, 25)
    for k, f in enumerate(list(model.solution) + [model.tracers.A]):
        assert np.array_equal(f.data, U[k]), f.name
    e0 = sim.output_writers["energies"].data["e"][0]
    assert abs(e0["ke"] - 9.817477042468) < 1e-10 and abs(e0["me"] - 12.109375) < 1e-12
    model.close()


def test_time_interval_writer_aligns_dt():
    """TimeInterval(0.1) writers clip the step so outputs land on multiples of 0.1 (upstream aligned_time_step)."""
    model = build_jacobian(arithmetic="fast")
    sim = Simulation(model, dt=0.03, stop_time=0.35)
    w = MemoryOutputWriter(model, dict(A=model.tracers.A), TimeInterval(0.1), with_halos=True)
    sim.output_writers["fields"] = w
    run_b(sim)
    assert np.allclose(w.times, [0.0, 0.1, 0.2, 0.3], atol=1e-12)
    assert w.data["A"][0].shape == (70, 70)
    assert abs(model.clock.time - 0.35) < 1e-12
    model.close()


def test_divergence_model_bounded_with_gradient_bc():
    from swmhd_b200 import FieldBoundaryConditions, GradientBoundaryCondition
    grid = RectilinearGrid(size=(48, 40), x=(-5, 5), y=(-5, 5), topology=(Periodic, Bounded, Flat))
    bcs = dict(A=FieldBoundaryConditions(north=GradientBoundaryCondition(-0.05), south=GradientBoundaryCondition(-0.05)))
    model = ShallowWaterModel(grid=grid, timestepper="RungeKutta3", boundary_conditions=bcs,
                              momentum_advection=WENO5(), mass_advection=WENO5(), tracer_advection=WENO5(),
                              gravitational_acceleration=9.81, coriolis=FPlane(f=1), tracers=("A",),
                              forcing=dict(uh=Forcing(div_lorentz_x, discrete_form=True), vh=Forcing(div_lorentz_y, discrete_form=True)),
                              formulation=ConservativeFormulation(), arithmetic="strict")
    set_b(model, uh=lambda x, y, z: y * np.exp(-(x ** 2 + y ** 2)), vh=lambda x, y, z: -x * np.exp(-(x ** 2 + y ** 2)),
          h=1.0, A=lambda x, y, z: -0.05 * y)
    time_step_b(model, 0.005, 8)
    g, cfg, U = make_case("BD", 48, Ny=40, arith=abi.ARITH_STRICT)
    O.fill_halos(cfg, U)
    O.step(cfg, U, 0.005, 8)
    for k, f in enumerate(list(model.solution) + [model.tracers.A]):
        assert np.array_equal(f.data, U[k]), f.name
    model.close()
