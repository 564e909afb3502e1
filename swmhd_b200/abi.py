"""ctypes view of include/swmhd.h (the C ABI of libswmhd_cuda.so).

This is the only place the package touches the shared library.  There is no
CPU fallback: if the library is missing, `load_library()` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

ABI_VERSION = 2

OK, ERR_ARG, ERR_CUDA, ERR_NONFINITE, ERR_NODEVICE, ERR_STATE, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
MAX_GPUS = 8
COMM_ID_BYTES = 128
PERIODIC, BOUNDED = 0, 1
JACOBIAN, DIVERGENCE = 0, 1
U, V, H, A = 0, 1, 2, 3
ARITH_FAST, ARITH_STRICT = 0, 1
FLAG_WENO_JS, FLAG_PRESSURE_GHDH, FLAG_CDIVU_OVER_H, FLAG_DIAG_CENTRED = 1, 2, 4, 8
FLAG_BC_DEPTH1, FLAG_WALL_WENO3, FLAG_V_MIRROR = 16, 32, 64      # C10 probe (oracle only)
FLAG_TRACER_CEN2, FLAG_TRACER_CEN4 = 128, 256                     # same probe: centred tracer advection
HALO = 3


class Config(C.Structure):
    """struct swmhd_config (include/swmhd.h)."""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("Nx", C.c_int32), ("Ny", C.c_int32),
        ("Hx", C.c_int32), ("Hy", C.c_int32),
        ("topo_x", C.c_int32), ("topo_y", C.c_int32),
        ("formulation", C.c_int32),
        ("arith", C.c_int32),
        ("flags", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double),
        ("g", C.c_double), ("f", C.c_double),
        ("weno_eps", C.c_double),
        ("h_ref", C.c_double),
        ("A_gradient_bc", C.c_int32),
        ("device", C.c_int32),
        ("A_grad_south", C.c_double), ("A_grad_north", C.c_double),
        ("slab_j0", C.c_int32), ("slab_ny", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32),
        ("n_gpus", C.c_int32), ("device_ids", C.c_int32 * MAX_GPUS), ("reserved0", C.c_int32),
    ]


class Diag(C.Structure):
    """struct swmhd_diag (include/swmhd.h)."""
    _fields_ = [
        ("ke", C.c_double), ("me", C.c_double), ("pe", C.c_double), ("total", C.c_double),
        ("max_abs_u", C.c_double), ("max_abs_A", C.c_double), ("min_h", C.c_double),
        ("max_abs_div_hB", C.c_double), ("sum_h", C.c_double),
        ("all_finite", C.c_int32), ("reserved", C.c_int32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


def make_config(Nx, Ny, Lx=10.0, Ly=10.0, formulation=JACOBIAN, topo_y=PERIODIC, g=9.81, f=1.0,
                arith=ARITH_FAST, flags=0, weno_eps=1e-6, h_ref=1.0, A_gradient=None, device=0,
                slab_j0=0, slab_ny=None, rank=0, world=1, device_ids=None) -> Config:
    c = Config()
    c.abi_version = ABI_VERSION
    c.Nx, c.Ny, c.Hx, c.Hy = Nx, Ny, HALO, HALO
    c.topo_x, c.topo_y = PERIODIC, topo_y
    c.formulation, c.arith, c.flags = formulation, arith, flags
    c.dx, c.dy = Lx / Nx, Ly / Ny
    c.g, c.f, c.weno_eps, c.h_ref = g, f, weno_eps, h_ref
    if A_gradient is not None:
        c.A_gradient_bc = 1
        c.A_grad_south, c.A_grad_north = A_gradient
    c.device = device
    c.slab_j0 = slab_j0
    c.slab_ny = Ny if slab_ny is None else slab_ny
    c.rank, c.world = rank, world
    if device_ids is not None and len(device_ids) > 1:      # single-process multi-GPU: one y-slab per device
        c.n_gpus = len(device_ids)
        for i, d in enumerate(device_ids):
            c.device_ids[i] = d
        c.device = device_ids[0]
    return c


# every symbol include/swmhd.h declares: (name, restype, argtypes)
_dp = C.POINTER(C.c_double)
_ctx = C.c_void_p
SYMBOLS = [
    ("swmhd_create", C.c_int, [C.POINTER(Config), C.POINTER(_ctx)]),
    ("swmhd_destroy", None, [_ctx]),
    ("swmhd_last_error", C.c_char_p, [_ctx]),
    ("swmhd_abi_version", C.c_int, []),
    ("swmhd_set_field", C.c_int, [_ctx, C.c_int, _dp, C.c_size_t]),
    ("swmhd_get_field", C.c_int, [_ctx, C.c_int, _dp, C.c_size_t]),
    ("swmhd_field_len", C.c_size_t, [_ctx, C.c_int]),
    ("swmhd_fill_halos", C.c_int, [_ctx]),
    ("swmhd_step", C.c_int, [_ctx, C.c_double, C.c_int]),
    ("swmhd_step_diag", C.c_int, [_ctx, C.c_double, C.c_int, C.POINTER(Diag)]),
    ("swmhd_upload_step", C.c_int, [_ctx, C.POINTER(_dp), C.c_size_t, C.c_double, C.POINTER(Diag)]),
    ("swmhd_step_seq", C.c_int, [_ctx, C.POINTER(C.c_double), C.c_int, C.POINTER(Diag)]),
    ("swmhd_substage", C.c_int, [_ctx, C.c_double, C.c_int]),
    ("swmhd_tendencies", C.c_int, [_ctx, C.POINTER(_dp), C.c_size_t]),
    ("swmhd_diagnostics", C.c_int, [_ctx, C.POINTER(Diag)]),
    ("swmhd_get_outputs", C.c_int, [_ctx, _dp, _dp, _dp]),
    ("swmhd_get_outputs_async", C.c_int, [_ctx, _dp, _dp, _dp, _dp]),
    ("swmhd_outputs_wait", C.c_int, [_ctx]),
    ("swmhd_pin_host", C.c_int, [C.c_void_p, C.c_size_t]),
    ("swmhd_unpin_host", C.c_int, [C.c_void_p]),
    ("swmhd_comm_unique_id", C.c_int, [C.c_void_p, C.c_size_t]),
    ("swmhd_comm_init", C.c_int, [_ctx, C.c_void_p, C.c_size_t]),
    ("swmhd_split_rows", C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("swmhd_time", C.c_double, [_ctx]),
    ("swmhd_iteration", C.c_int64, [_ctx]),
    ("swmhd_set_clock", C.c_int, [_ctx, C.c_double, C.c_int64]),
    ("swmhd_set_streams", C.c_int, [_ctx, C.c_void_p, C.c_void_p]),
    ("swmhd_substage_edges", C.c_int, [_ctx, C.c_double, C.c_int]),
    ("swmhd_substage_interior", C.c_int, [_ctx, C.c_double, C.c_int]),
    ("swmhd_substage_finish", C.c_int, [_ctx, C.c_int]),
    ("swmhd_exchange_rows", C.c_int, [_ctx, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    ("swmhd_sync", C.c_int, [_ctx]),
    ("swmhd_check_guards", C.c_int, [_ctx]),
    ("swmhd_arm_diag", C.c_int, [_ctx, C.c_int]),
    ("swmhd_get_diag_slots", C.c_int, [_ctx, C.c_int, C.c_int, C.POINTER(Diag)]),
    ("swmhd_step_profile", C.c_int, [_ctx, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    ("swmhd_step_profile_diag", C.c_int, [_ctx, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    ("swmhd_launch_count", C.c_int64, [_ctx]),
    ("swmhd_last_step_ms", C.c_double, [_ctx]),
]

LIB_NAME = "libswmhd_cuda.so"
_lib = None


def library_path() -> Path:
    return Path(__file__).resolve().parent / LIB_NAME


def load_library():
    """dlopen libswmhd_cuda.so (built in-tree by `__graft_entry__.build()` / `swmhd_b200.build`).

    Raises RuntimeError when it is missing: the product path never falls back to a CPU implementation.
    """
    global _lib
    if _lib is not None:
        return _lib
    p = Path(os.environ.get("SWMHD_LIB", library_path()))
    if not p.exists():
        raise RuntimeError(f"{p} not found: build it with `python -m swmhd_b200.build` "
                           "(there is no CPU fallback for the SWMHD hot path)")
    lib = C.CDLL(str(p))
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    if lib.swmhd_abi_version() != ABI_VERSION:
        raise RuntimeError("libswmhd_cuda.so ABI version mismatch")
    _lib = lib
    return lib
