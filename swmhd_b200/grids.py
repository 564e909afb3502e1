"""RectilinearGrid and haloed Field storage laid out like Oceananigans' (SURVEY A.1, A.10).

Mirrors what the reference scripts use of `Oceananigans.Grids` / `Oceananigans.Fields`:
`RectilinearGrid(size=, x=, y=, topology=)` (SWMHD_example.jl:14-16, divergence_sw_mhd.jl:12-14),
`nodes`, `interior`, `parent` and the node-wise evaluation done by `set!` (SWMHD_example.jl:41).
Host arrays only; device memory lives behind the C ABI.
"""
from __future__ import annotations

import numpy as np

from . import abi


class Periodic:  # topology tags, as in Oceananigans.Grids
    pass


class Bounded:
    pass


class Flat:
    pass


Face, Center = "Face", "Center"

#: staggered locations of the four prognostic fields (u|uh, v|vh, h, A)
FIELD_LOCATIONS = {abi.U: (Face, Center), abi.V: (Center, Face), abi.H: (Center, Center), abi.A: (Center, Center)}


class RectilinearGrid:
    """Uniform 2-D (x, y, Flat) C-grid with halo 3 (the WENO5 default)."""

    def __init__(self, size, x, y, topology=(Periodic, Periodic, Flat), halo=(3, 3)):
        self.Nx, self.Ny = int(size[0]), int(size[1])
        self.x0, self.x1 = float(x[0]), float(x[1])
        self.y0, self.y1 = float(y[0]), float(y[1])
        self.Lx, self.Ly = self.x1 - self.x0, self.y1 - self.y0
        self.topology = tuple(topology)
        if self.topology[0] is not Periodic:
            raise ValueError("x topology must be Periodic (the reference never runs Bounded-x)")
        if self.topology[1] not in (Periodic, Bounded):
            raise ValueError("y topology must be Periodic or Bounded")
        if tuple(halo) != (3, 3):
            raise ValueError("halo must be (3, 3)")
        self.Hx = self.Hy = abi.HALO
        self.dx, self.dy = self.Lx / self.Nx, self.Ly / self.Ny

    @property
    def bounded_y(self):
        return self.topology[1] is Bounded

    def n_interior(self, field):
        """(Nx_f, Ny_f): Face-located fields have N+1 points in a Bounded direction."""
        lx, ly = FIELD_LOCATIONS[field]
        return self.Nx, self.Ny + (1 if (ly == Face and self.bounded_y) else 0)

    def parent_shape(self, field):
        """numpy shape of the parent array; numpy's LAST axis is Julia's FIRST (i fastest)."""
        nx, ny = self.n_interior(field)
        return (ny + 2 * self.Hy, nx + 2 * self.Hx)

    def nodes(self, field):
        """1-D node coordinates (x[i], y[j]) of the interior points of `field`."""
        lx, ly = FIELD_LOCATIONS[field]
        nx, ny = self.n_interior(field)
        i = np.arange(1, nx + 1, dtype=np.float64)
        j = np.arange(1, ny + 1, dtype=np.float64)
        xs = self.x0 + (i - 0.5) * self.dx if lx == Center else self.x0 + (i - 1.0) * self.dx
        ys = self.y0 + (j - 0.5) * self.dy if ly == Center else self.y0 + (j - 1.0) * self.dy
        return xs, ys

    def new_parent(self, field):
        return np.zeros(self.parent_shape(field), dtype=np.float64)

    def interior(self, parent, field):
        nx, ny = self.n_interior(field)
        return parent[self.Hy:self.Hy + ny, self.Hx:self.Hx + nx]

    def set_interior(self, parent, field, fn):
        """`set!`: evaluate fn(x, y, z) at the field's own nodes on the interior (A.10)."""
        xs, ys = self.nodes(field)
        X, Y = np.meshgrid(xs, ys)  # shape (ny, nx)
        val = fn(X, Y, 0.0) if callable(fn) else fn
        self.interior(parent, field)[...] = np.broadcast_to(np.asarray(val, dtype=np.float64), X.shape)
        return parent
