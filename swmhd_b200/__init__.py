"""swmhd_b200 — B200-native (sm_100a) implementation of the SWMHD RK3 hot path.

`Context` is the 1:1 wrapper over the C ABI (include/swmhd.h); `models` mirrors the
Oceananigans surface the reference scripts use.  The library has no CPU fallback.
"""
from . import abi  # noqa: F401
from .context import Context, SwmhdError  # noqa: F401
from .grids import RectilinearGrid, Periodic, Bounded, Flat  # noqa: F401
from .models import (  # noqa: F401
    ShallowWaterModel, VectorInvariantFormulation, ConservativeFormulation, WENO5, VelocityStencil,
    VorticityStencil, FPlane, Forcing, FieldBoundaryConditions, GradientBoundaryCondition,
    lorentz_force_func_x, lorentz_force_func_y, div_lorentz_x, div_lorentz_y,
    set_b, time_step_b, time_step_diag_b, set_and_step_diag_b, run_b, Simulation, Callback, IterationInterval, TimeInterval, MemoryOutputWriter,
)

__version__ = "0.1.0"
