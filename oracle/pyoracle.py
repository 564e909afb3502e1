"""ctypes wrapper around oracle/_build/libswmhd_oracle.so.

TEST INFRASTRUCTURE (see the header of swmhd_oracle.c).  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from swmhd_b200.abi import Config, Diag, BOUNDED, V as FIELD_V

HERE = Path(__file__).resolve().parent
SO = HERE / "_build" / "libswmhd_oracle.so"
_dp = C.POINTER(C.c_double)
_lib = None


def build(force=False):
    src = HERE / "swmhd_oracle.c"
    if force or not SO.exists() or SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return SO


def lib():
    global _lib
    if _lib is None:
        if not SO.exists():
            build()
        _lib = C.CDLL(str(SO))
        _lib.swmhd_oracle_field_len.restype = C.c_size_t
        _lib.swmhd_oracle_field_len.argtypes = [C.POINTER(Config), C.c_int]
    return _lib


def set_threads(n: int) -> int:
    """Use n OpenMP threads for the oracle (returns the number in effect)."""
    return int(lib().swmhd_oracle_set_threads(C.c_int(int(n))))


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def field_shape(cfg: Config, field: int):
    """numpy shape (rows, pitch) of the parent array: i fastest = last numpy axis."""
    rows = cfg.Ny + 6 + (1 if (field == FIELD_V and cfg.topo_y == BOUNDED) else 0)
    return (rows, cfg.Nx + 6)


def alloc(cfg: Config):
    return [np.zeros(field_shape(cfg, k)) for k in range(4)]


def fill_halos(cfg, U):
    lib().swmhd_oracle_fill_halos(C.byref(cfg), *[_p(a) for a in U])


def tendencies(cfg, U):
    G = alloc(cfg)
    lib().swmhd_oracle_tendencies(C.byref(cfg), *[_p(a) for a in U], *[_p(g) for g in G])
    return G


def lorentz(cfg, h, A):
    Fx, Fy = np.zeros_like(h), np.zeros_like(h)
    lib().swmhd_oracle_lorentz(C.byref(cfg), _p(h), _p(A), _p(Fx), _p(Fy))
    return Fx, Fy


def weno_line(cfg, psi):
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    L, R = np.full_like(psi, np.nan), np.full_like(psi, np.nan)
    lib().swmhd_oracle_weno_line(C.byref(cfg), _p(psi), C.c_int(psi.size), _p(L), _p(R))
    return L, R


def substage(cfg, U, Gn, Gm, dt, stage):
    arr = _dp * 4
    lib().swmhd_oracle_substage(C.byref(cfg), *[_p(a) for a in U], arr(*[_p(g) for g in Gn]),
                                arr(*[_p(g) for g in Gm]), C.c_double(dt), C.c_int(stage))


def step(cfg, U, dt, nsteps=1, clock=None):
    """nsteps RK3 steps in place on the four haloed arrays U (halos must be filled)."""
    ck = _p(clock) if clock is not None else None
    rc = lib().swmhd_oracle_step(C.byref(cfg), *[_p(a) for a in U], C.c_double(dt), C.c_int(nsteps), ck)
    assert rc == 0
    return U


def diagnostics(cfg, U) -> dict:
    d = Diag()
    lib().swmhd_oracle_diagnostics(C.byref(cfg), *[_p(a) for a in U], C.byref(d))
    return d.as_dict()
