"""Deterministic initial conditions of SURVEY 8(d), shared by CPU and GPU tests."""
import numpy as np

from swmhd_b200 import abi
from swmhd_b200.grids import RectilinearGrid, Periodic, Bounded, Flat


def grid_for(N, bounded_y=False, Ny=None):
    Ny = N if Ny is None else Ny
    topo = (Periodic, Bounded if bounded_y else Periodic, Flat)
    return RectilinearGrid((N, Ny), (-5, 5), (-5, 5), topology=topo)


def make_case(kind, N, Ny=None, arith=abi.ARITH_FAST, flags=0, perturb=None):
    """kind: 'J' (SWMHD_example.jl:36-40), 'D' (divergence_sw_mhd.jl:33-38, amplitude 0.5),
    'G' (amplitude 0.1, Jacobian form, the published low-B case), 'GD' (same, divergence form),
    'BJ'/'BD' (Bounded-y, A = -0.05 y with gradient BC, unit vortex; divergence_sw_mhd.jl:17,34-37)."""
    bounded = kind in ("BJ", "BD")
    g = grid_for(N, bounded, Ny)
    form = abi.JACOBIAN if kind in ("J", "G", "BJ") else abi.DIVERGENCE
    cfg = abi.make_config(g.Nx, g.Ny, formulation=form, arith=arith, flags=flags,
                          topo_y=abi.BOUNDED if bounded else abi.PERIODIC,
                          A_gradient=(-0.05, -0.05) if bounded else None)
    U = [g.new_parent(k) for k in range(4)]
    gauss2 = lambda amp: (lambda x, y, z: amp * np.exp(-((x - 0.5) ** 2 + y ** 2)) - amp * np.exp(-((x + 0.5) ** 2 + y ** 2)))
    g.set_interior(U[abi.H], abi.H, 1.0)
    if kind == "J":
        g.set_interior(U[abi.U], abi.U, lambda x, y, z: 5 * y * np.exp(-(x ** 2 + y ** 2)))
        g.set_interior(U[abi.V], abi.V, lambda x, y, z: -5 * x * np.exp(-(x ** 2 + y ** 2)))
        g.set_interior(U[abi.A], abi.A, lambda x, y, z: 0.5 * np.abs(y))
    elif kind == "D":
        g.set_interior(U[abi.A], abi.A, gauss2(0.5))
    elif kind in ("G", "GD"):
        g.set_interior(U[abi.A], abi.A, gauss2(0.1))
    elif bounded:
        g.set_interior(U[abi.U], abi.U, lambda x, y, z: y * np.exp(-(x ** 2 + y ** 2)))
        g.set_interior(U[abi.V], abi.V, lambda x, y, z: -x * np.exp(-(x ** 2 + y ** 2)))
        g.set_interior(U[abi.A], abi.A, lambda x, y, z: -0.05 * y)
    else:
        raise ValueError(kind)
    if perturb:
        rng = np.random.default_rng(perturb)
        for k in range(4):
            it = g.interior(U[k], k)
            # smooth random perturbation (a few low Fourier modes), amplitude 1e-3
            xs, ys = g.nodes(k)
            X, Y = np.meshgrid(xs, ys)
            for _ in range(4):
                kx, ky = rng.integers(1, 4, 2)
                ph = rng.uniform(0, 2 * np.pi, 2)
                it += 1e-3 * rng.standard_normal() * np.sin(2 * np.pi * kx * X / 10 + ph[0]) * np.sin(2 * np.pi * ky * Y / 10 + ph[1])
    return g, cfg, U


def rel_l2(g, a, b, field):
    ia, ib = g.interior(a, field), g.interior(b, field)
    den = np.sqrt((ib ** 2).sum())
    num = np.sqrt(((ia - ib) ** 2).sum())
    return num / den if den > 0 else num
