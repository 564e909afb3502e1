"""Host-side mirror of the Oceananigans surface the reference scripts drive.

Names, argument meaning and call order follow jacobian_formulation/SWMHD_example.jl:14-100
and divergence_formulation/divergence_sw_mhd.jl:12-99, so a reference script reads almost
line for line in Python (`!` functions become `set_b`, `run_b`, `time_step_b`):

    grid  = RectilinearGrid(size=(Nx, Ny), x=(-Lx/2, Lx/2), y=(-Ly/2, Ly/2),
                            topology=(Periodic, Periodic, Flat))
    model = ShallowWaterModel(grid=grid, timestepper="RungeKutta3",
                              momentum_advection=WENO5(vector_invariant=VelocityStencil()),
                              mass_advection=WENO5(), tracer_advection=WENO5(),
                              gravitational_acceleration=9.81, coriolis=FPlane(f=1),
                              tracers=("A",),
                              forcing=dict(u=Forcing(lorentz_force_func_x, discrete_form=True),
                                           v=Forcing(lorentz_force_func_y, discrete_form=True)),
                              formulation=VectorInvariantFormulation())
    set_b(model, u=u_i, v=v_i, h=h_i, A=A_i)
    simulation = Simulation(model, dt=0.01, stop_time=30.0)
    run_b(simulation)

Everything numerical happens behind the C ABI (`Context`); this module only validates
that the requested model is one of the two the reference builds, moves parent arrays,
and runs the Simulation loop (callbacks, schedules, dt alignment, output writers).
"""
from __future__ import annotations

import math
import time as _time
from dataclasses import dataclass, field as _dc_field
from typing import Callable

import numpy as np

from . import abi
from .context import Context
from .grids import RectilinearGrid, Periodic, Bounded, Flat  # noqa: F401  (re-exported)


# -- option tags ------------------------------------------------------------------------------
class VectorInvariantFormulation:
    pass


class ConservativeFormulation:
    pass


class VelocityStencil:
    pass


class VorticityStencil:
    pass


@dataclass
class WENO5:
    vector_invariant: object | None = None


@dataclass
class FPlane:
    f: float = 0.0


@dataclass
class GradientBoundaryCondition:
    gradient: float


@dataclass
class FieldBoundaryConditions:
    north: GradientBoundaryCondition | None = None
    south: GradientBoundaryCondition | None = None


# The reference's forcing hooks.  They are the names a user passes to `Forcing`; the
# arithmetic they stand for is inlined in the fused CUDA kernel (FORM 0 / FORM 1), so
# here they are only tags that select it.
def lorentz_force_func_x(i, j, k, grid, clock, fields):  # sw_mhd_jacobian_functions.jl:20-22
    raise NotImplementedError("evaluated on the GPU inside the fused substage kernel")


def lorentz_force_func_y(i, j, k, grid, clock, fields):  # sw_mhd_jacobian_functions.jl:24-26
    raise NotImplementedError("evaluated on the GPU inside the fused substage kernel")


def div_lorentz_x(i, j, k, grid, clock, fields):  # sw_mhd_divergence_functions.jl:162-165
    raise NotImplementedError("evaluated on the GPU inside the fused substage kernel")


def div_lorentz_y(i, j, k, grid, clock, fields):  # sw_mhd_divergence_functions.jl:167-170
    raise NotImplementedError("evaluated on the GPU inside the fused substage kernel")


@dataclass
class Forcing:
    func: Callable
    discrete_form: bool = False


_FIELD_NAMES = {
    abi.JACOBIAN: ("u", "v", "h", "A"),
    abi.DIVERGENCE: ("uh", "vh", "h", "A"),
}


class Clock:
    def __init__(self, ctx: Context):
        self._ctx = ctx

    @property
    def time(self):
        return self._ctx.time

    @property
    def iteration(self):
        return self._ctx.iteration


class Field:
    """A haloed prognostic field: host parent array + lazy refresh from the device."""

    def __init__(self, model, index, name):
        self.model, self.index, self.name = model, index, name
        self.parent = model.grid.new_parent(index)
        self._stale = True

    def _refresh(self):
        if self._stale:
            self.model.ctx.get_field(self.index, self.parent)
            self._stale = False

    @property
    def data(self):
        """Parent array including halos (numpy axis 0 = j, axis 1 = i)."""
        self._refresh()
        return self.parent

    @property
    def interior(self):
        self._refresh()
        return self.model.grid.interior(self.parent, self.index)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.interior, dtype=dtype)


class ShallowWaterModel:
    """The two model set-ups of the reference, executed by libswmhd_cuda.so."""

    def __init__(self, grid: RectilinearGrid, timestepper="RungeKutta3", momentum_advection=None,
                 mass_advection=None, tracer_advection=None, gravitational_acceleration=9.81,
                 coriolis: FPlane | None = None, tracers=("A",), forcing=None,
                 formulation=None, boundary_conditions=None, closure=None,
                 arithmetic="fast", device=0):
        if str(timestepper).lstrip(":") != "RungeKutta3":
            raise ValueError("only timestepper = :RungeKutta3 (SWMHD_example.jl:23)")
        if closure is not None:
            raise ValueError("the reference passes no closure")
        if isinstance(tracers, str):
            tracers = (tracers,)
        if tuple(str(t).lstrip(":") for t in tracers) != ("A",):
            raise ValueError("tracers must be (:A) — the magnetic potential")
        for adv in (momentum_advection, mass_advection, tracer_advection):
            if not isinstance(adv, WENO5):
                raise ValueError("advection schemes must be WENO5()")
        forcing = forcing or {}
        funcs = {k: (v.func if isinstance(v, Forcing) else v) for k, v in forcing.items()}
        if isinstance(formulation, VectorInvariantFormulation):
            form = abi.JACOBIAN
            if not isinstance(momentum_advection.vector_invariant, VelocityStencil):
                raise ValueError("VectorInvariantFormulation needs WENO5(vector_invariant=VelocityStencil())")
            if funcs != {"u": lorentz_force_func_x, "v": lorentz_force_func_y}:
                raise ValueError("forcing must be (u=lorentz_force_func_x, v=lorentz_force_func_y)")
        elif isinstance(formulation, ConservativeFormulation):
            form = abi.DIVERGENCE
            if momentum_advection.vector_invariant is not None:
                raise ValueError("ConservativeFormulation uses WENO5() momentum advection")
            if funcs != {"uh": div_lorentz_x, "vh": div_lorentz_y}:
                raise ValueError("forcing must be (uh=div_lorentz_x, vh=div_lorentz_y)")
        else:
            raise ValueError("formulation must be VectorInvariantFormulation() or ConservativeFormulation()")
        if any(isinstance(v, Forcing) and not v.discrete_form for v in forcing.values()):
            raise ValueError("the Lorentz hooks are discrete_form = true forcings")
        A_grad = None
        if boundary_conditions and "A" in boundary_conditions:
            bc = boundary_conditions["A"]
            if not grid.bounded_y:
                raise ValueError("A boundary conditions need a Bounded y topology")
            A_grad = (bc.south.gradient if bc.south else 0.0, bc.north.gradient if bc.north else 0.0)
        self.grid = grid
        self.formulation = formulation
        self.gravitational_acceleration = float(gravitational_acceleration)
        self.coriolis = coriolis or FPlane(0.0)
        arith = {"fast": abi.ARITH_FAST, "strict": abi.ARITH_STRICT}[arithmetic]
        self.cfg = abi.make_config(grid.Nx, grid.Ny, Lx=grid.Lx, Ly=grid.Ly, formulation=form,
                                   topo_y=abi.BOUNDED if grid.bounded_y else abi.PERIODIC,
                                   g=self.gravitational_acceleration, f=float(self.coriolis.f),
                                   arith=arith, A_gradient=A_grad, device=device)
        self.ctx = Context(self.cfg)
        names = _FIELD_NAMES[form]
        self._fields = [Field(self, k, names[k]) for k in range(4)]
        self.solution = _Named(**{names[k]: self._fields[k] for k in range(3)})
        self.tracers = _Named(A=self._fields[3])
        self.clock = Clock(self.ctx)
        self.ctx.fill_halos()

    def fields(self):
        return {f.name: f for f in self._fields}

    def _mark_stale(self):
        for f in self._fields:
            f._stale = True

    def diagnostics(self):
        return self.ctx.diagnostics()

    def close(self):
        self.ctx.close()


class _Named:
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self._order = list(kw)

    def __iter__(self):
        return iter(getattr(self, k) for k in self._order)

    def __getitem__(self, k):
        return getattr(self, k)


def set_b(model: ShallowWaterModel, **kwargs):
    """`set!(model, u=..., v=..., h=..., A=...)`: functions of (x, y, z), arrays or numbers,
    evaluated at each field's own nodes; omitted fields keep their values; halos refilled."""
    by_name = model.fields()
    for name, val in kwargs.items():
        if name not in by_name:
            raise KeyError(f"{name} is not a field of this model ({list(by_name)})")
        f = by_name[name]
        if isinstance(val, np.ndarray) and val.shape == f.parent.shape and val.dtype == np.float64 and val.flags["C_CONTIGUOUS"]:
            model.ctx.set_field(f.index, val)      # a full parent array: straight H2D (fast when pinned)
            f._stale = True
            continue
        f._refresh()
        model.grid.set_interior(f.parent, f.index, val)
        model.ctx.set_field(f.index, f.parent)
    model.ctx.fill_halos()
    model._mark_stale()


def time_step_b(model: ShallowWaterModel, dt, nsteps=1):
    """`time_step!(model, Δt)`: one (or nsteps) RK3 step(s) on the device."""
    model.ctx.step(dt, nsteps)
    model._mark_stale()


def time_step_diag_b(model: ShallowWaterModel, dt, nsteps=1):
    """`time_step!` x nsteps that also returns, per step, the energy / progress diagnostics of the
    state the step started from (SWMHD_example.jl:47-77), evaluated inside the stage-1 kernel."""
    out = model.ctx.step_diag(dt, nsteps)
    model._mark_stale()
    return out


def set_and_step_diag_b(model: ShallowWaterModel, dt, **kwargs):
    """`set!(model, ...)` immediately followed by `time_step!(model, Δt)` with the diagnostics of the uploaded state:
    all four fields must be given as full parent arrays; the upload is pipelined with stage 1 (swmhd_upload_step)."""
    by_name = model.fields()
    if sorted(kwargs) != sorted(by_name):
        raise ValueError(f"set_and_step_diag_b needs all of {list(by_name)}")
    U = [None] * 4
    for name, a in kwargs.items():
        f = by_name[name]
        if not (isinstance(a, np.ndarray) and a.shape == f.parent.shape and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
            raise ValueError(f"{name}: C-contiguous float64 parent array of shape {f.parent.shape} expected")
        U[f.index] = a
    d = model.ctx.upload_step(U, dt, diag=True)
    model._mark_stale()
    return d


# -- Simulation -------------------------------------------------------------------------------
@dataclass
class IterationInterval:
    interval: int

    def due(self, sim):
        return sim.model.clock.iteration % self.interval == 0

    def next_time(self, sim):
        return math.inf


@dataclass
class TimeInterval:
    interval: float
    _next: float = 0.0

    def due(self, sim):
        t = sim.model.clock.time
        if t >= self._next - 1e-12 * max(1.0, abs(self._next)):
            self._next += self.interval
            return True
        return False

    def next_time(self, sim):
        return self._next


@dataclass
class Callback:
    func: Callable
    schedule: object = _dc_field(default_factory=lambda: IterationInterval(1))


class MemoryOutputWriter:
    """Collects named outputs in memory (stand-in for JLD2OutputWriter / NetCDFOutputWriter,
    SWMHD_example.jl:81-92); `outputs` maps names to Fields or callables(model)."""

    def __init__(self, model, outputs: dict, schedule, with_halos=False):
        self.model, self.outputs, self.schedule, self.with_halos = model, outputs, schedule, with_halos
        self.times, self.iterations = [], []
        self.data = {k: [] for k in outputs}

    def write(self, sim):
        self.times.append(self.model.clock.time)
        self.iterations.append(self.model.clock.iteration)
        for k, o in self.outputs.items():
            if isinstance(o, Field):
                self.data[k].append(np.array(o.data if self.with_halos else o.interior))
            else:
                self.data[k].append(o(self.model))

    def save(self, path):
        np.savez_compressed(path, times=np.array(self.times), iterations=np.array(self.iterations),
                            **{k: np.array(v) for k, v in self.data.items()})


class Simulation:
    """`Simulation(model, Δt=, stop_time=)` with callbacks and output_writers dictionaries."""

    def __init__(self, model, dt, stop_time=math.inf, stop_iteration=math.inf):
        self.model, self.dt, self.stop_time, self.stop_iteration = model, float(dt), stop_time, stop_iteration
        self.callbacks: dict[str, Callback] = {}
        self.output_writers: dict[str, MemoryOutputWriter] = {}
        self.run_wall_time = 0.0

    def _aligned_dt(self):
        """upstream aligned_time_step: clip Δt to the next TimeInterval event and to stop_time."""
        t = self.model.clock.time
        dt = min(self.dt, self.stop_time - t)
        for s in [c.schedule for c in self.callbacks.values()] + [w.schedule for w in self.output_writers.values()]:
            nt = s.next_time(self)
            if nt > t + 1e-14 * max(1.0, abs(t)):
                dt = min(dt, nt - t)
        return dt

    def _plan_batch(self, max_steps=256):
        """The Δt sequence up to (and including) the step after which a callback or writer is due, or stop_time /
        stop_iteration is reached: upstream's aligned_time_step replayed on a copy of the clock, which ticks
        (8/15, 2/15, 1/3) Δt per RK3 stage exactly like the library's."""
        g1 = 8.0 / 15.0
        g2z2 = 5.0 / 12.0 + (-17.0 / 60.0)
        g3z3 = 3.0 / 4.0 + (-5.0 / 12.0)
        t, it = self.model.clock.time, self.model.clock.iteration
        scheds = [c.schedule for c in self.callbacks.values()] + [w.schedule for w in self.output_writers.values()]
        nexts = [s.next_time(self) for s in scheds if isinstance(s, TimeInterval)]
        ivals = [s.interval for s in scheds if isinstance(s, IterationInterval)]
        dts = []
        while len(dts) < max_steps:
            dt = min(self.dt, self.stop_time - t)
            for nt in nexts:
                if nt > t + 1e-14 * max(1.0, abs(t)):
                    dt = min(dt, nt - t)
            dts.append(dt)
            t = ((t + g1 * dt) + g2z2 * dt) + g3z3 * dt
            it += 1
            due_time = any(t >= nt - 1e-12 * max(1.0, abs(nt)) for nt in nexts)
            due_iter = any(it % iv == 0 for iv in ivals)
            if due_time or due_iter or t >= self.stop_time - 1e-12 or it >= self.stop_iteration:
                break
        return dts

    def _fire(self):
        for c in self.callbacks.values():
            if c.schedule.due(self):
                c.func(self)
        for w in self.output_writers.values():
            if w.schedule.due(self):
                w.write(self)


def run_b(sim: Simulation):
    """`run!(simulation)` (SWMHD_example.jl:97): between two events (callback, writer, stop) the whole aligned Δt
    sequence runs in ONE library call (swmhd_step_seq), the device never returns to the host in between."""
    m = sim.model
    t0 = _time.perf_counter()
    sim._fire()  # iteration 0
    while m.clock.time < sim.stop_time - 1e-12 and m.clock.iteration < sim.stop_iteration:
        dts = sim._plan_batch()
        m.ctx.step_seq(dts)
        m._mark_stale()
        sim._fire()
    sim.run_wall_time = _time.perf_counter() - t0
    return sim
