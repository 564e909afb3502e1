"""Thin object wrapper over the C ABI (one `swmhd_ctx` = one y-slab on one GPU)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

_dp = C.POINTER(C.c_double)


class SwmhdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libswmhd_cuda error {code}: {msg}")
        self.code = code


def _ptr(a: np.ndarray):
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("host buffers must be C-contiguous float64 parent arrays")
    return a.ctypes.data_as(_dp)


class Context:
    """Owns a `swmhd_ctx*`.  Every method maps 1:1 onto an include/swmhd.h entry point."""

    def __init__(self, cfg: abi.Config):
        self.lib = abi.load_library()
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = self.lib.swmhd_create(C.byref(cfg), C.byref(self._h))
        if rc != abi.OK:
            raise SwmhdError(rc, (self.lib.swmhd_last_error(None) or b"").decode())
        self.ny = cfg.slab_ny
        self.pitch = cfg.Nx + 2 * abi.HALO

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.swmhd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != abi.OK:
            raise SwmhdError(rc, (self.lib.swmhd_last_error(self._h) or b"").decode())

    # -- fields -----------------------------------------------------------------------------
    def field_shape(self, field):
        """(rows, pitch) of the host parent array: this slab's rows, or the global array for n_gpus > 1."""
        n = self.lib.swmhd_field_len(self._h, field)
        return (n // self.pitch, self.pitch)

    def new_parent(self, field):
        return np.zeros(self.field_shape(field), dtype=np.float64)

    def set_field(self, field, parent: np.ndarray):
        self._ck(self.lib.swmhd_set_field(self._h, field, _ptr(parent), parent.size))

    def get_field(self, field, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = self.new_parent(field)
        self._ck(self.lib.swmhd_get_field(self._h, field, _ptr(out), out.size))
        return out

    def set_state(self, U):
        for k in range(4):
            self.set_field(k, U[k])

    def get_state(self):
        return [self.get_field(k) for k in range(4)]

    def fill_halos(self):
        self._ck(self.lib.swmhd_fill_halos(self._h))

    # -- stepping ---------------------------------------------------------------------------
    def step(self, dt, nsteps=1):
        self._ck(self.lib.swmhd_step(self._h, float(dt), int(nsteps)))

    def step_diag(self, dt, nsteps=1):
        arr = (abi.Diag * nsteps)()
        self._ck(self.lib.swmhd_step_diag(self._h, float(dt), int(nsteps), arr))
        return [d.as_dict() for d in arr]

    def upload_step(self, U, dt, diag=True):
        """set_state(U) + one RK3 step (+ diagnostics of the uploaded state), the upload pipelined with stage 1."""
        arr = (_dp * 4)(*[_ptr(u) for u in U])
        d = abi.Diag() if diag else None
        self._ck(self.lib.swmhd_upload_step(self._h, arr, max(u.size for u in U), float(dt), C.byref(d) if diag else None))
        return d.as_dict() if diag else None

    def step_seq(self, dts, diag=False):
        """One call, len(dts) RK3 steps, step n with dts[n] (the aligned Δt sequence up to the next output time)."""
        n = len(dts)
        arr = (C.c_double * n)(*[float(x) for x in dts])
        dg = (abi.Diag * n)() if diag else None
        self._ck(self.lib.swmhd_step_seq(self._h, arr, n, dg))
        return [d.as_dict() for d in dg] if diag else None

    def step_profile(self, dt, nsteps=1, diag=False):
        """Mean device time (ms) of the three fused substage kernels over nsteps steps
        (diag=True: with the diagnostics fused into stage 1, as step_diag runs them)."""
        out = (C.c_double * 3)()
        fn = self.lib.swmhd_step_profile_diag if diag else self.lib.swmhd_step_profile
        self._ck(fn(self._h, float(dt), int(nsteps), out))
        return list(out)

    def substage(self, dt, stage):
        self._ck(self.lib.swmhd_substage(self._h, float(dt), int(stage)))

    def tendencies(self):
        G = [self.new_parent(k) for k in range(4)]
        arr = (_dp * 4)(*[_ptr(g) for g in G])
        self._ck(self.lib.swmhd_tendencies(self._h, arr, max(g.size for g in G)))
        return G

    def diagnostics(self, check_finite=True) -> dict:
        d = abi.Diag()
        rc = self.lib.swmhd_diagnostics(self._h, C.byref(d))
        if rc == abi.ERR_NONFINITE and not check_finite:
            rc = abi.OK
        self._ck(rc)
        return d.as_dict()

    def get_outputs(self):
        """(u, v, s) parent arrays of the field writer, computed on the device."""
        u, v, s = self.new_parent(abi.U), self.new_parent(abi.V), self.new_parent(abi.U)
        self._ck(self.lib.swmhd_get_outputs(self._h, _ptr(u), _ptr(v), _ptr(s)))
        return u, v, s

    def get_outputs_async(self, u, v, s, A):
        """Queue the field writer's outputs (u, v, s, A of the current state) into four host parent arrays
        (page-locked for a truly asynchronous copy) and return; valid after outputs_wait()."""
        self._ck(self.lib.swmhd_get_outputs_async(self._h, _ptr(u), _ptr(v), _ptr(s), _ptr(A)))

    def outputs_wait(self):
        self._ck(self.lib.swmhd_outputs_wait(self._h))

    # -- NCCL ring (world > 1, one process per GPU) --------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(abi.COMM_ID_BYTES)
        rc = self.lib.swmhd_comm_unique_id(buf, abi.COMM_ID_BYTES)
        if rc != abi.OK:
            raise SwmhdError(rc, (self.lib.swmhd_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, uid: bytes):
        buf = C.create_string_buffer(uid, abi.COMM_ID_BYTES)
        self._ck(self.lib.swmhd_comm_init(self._h, buf, abi.COMM_ID_BYTES))

    # -- clock ------------------------------------------------------------------------------
    @property
    def time(self):
        return self.lib.swmhd_time(self._h)

    @property
    def iteration(self):
        return self.lib.swmhd_iteration(self._h)

    def set_clock(self, time, iteration):
        self._ck(self.lib.swmhd_set_clock(self._h, float(time), int(iteration)))

    # -- slab plumbing ----------------------------------------------------------------------
    def set_streams(self, main_stream: int, edge_stream: int):
        self._ck(self.lib.swmhd_set_streams(self._h, C.c_void_p(main_stream), C.c_void_p(edge_stream)))

    def substage_edges(self, dt, stage):
        self._ck(self.lib.swmhd_substage_edges(self._h, float(dt), int(stage)))

    def substage_interior(self, dt, stage):
        self._ck(self.lib.swmhd_substage_interior(self._h, float(dt), int(stage)))

    def substage_finish(self, stage):
        self._ck(self.lib.swmhd_substage_finish(self._h, int(stage)))

    def exchange_rows(self, field, which):
        """(device pointer, nrows, row_doubles) of a send/recv row block (see swmhd.h)."""
        p, n, w = C.c_void_p(), C.c_int(), C.c_size_t()
        self._ck(self.lib.swmhd_exchange_rows(self._h, field, which, C.byref(p), C.byref(n), C.byref(w)))
        return p.value, n.value, w.value

    def arm_diag(self, slot):
        self._ck(self.lib.swmhd_arm_diag(self._h, int(slot)))

    def get_diag_slots(self, first, count):
        arr = (abi.Diag * count)()
        self._ck(self.lib.swmhd_get_diag_slots(self._h, int(first), int(count), arr))
        return [d.as_dict() for d in arr]

    def check_guards(self):
        """SWMHD_GUARD=1 contexts: raise if any kernel wrote outside its device arrays."""
        self._ck(self.lib.swmhd_check_guards(self._h))

    def sync(self):
        self._ck(self.lib.swmhd_sync(self._h))

    @property
    def launch_count(self):
        return self.lib.swmhd_launch_count(self._h)

    @property
    def last_step_ms(self):
        return self.lib.swmhd_last_step_ms(self._h)
