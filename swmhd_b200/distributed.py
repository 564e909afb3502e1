"""y-slab domain decomposition, host side.  The exchange itself lives in libswmhd_cuda.so (ncclSend/ncclRecv,
include/swmhd.h: swmhd_comm_init, n_gpus); this module holds the decomposition maths, the thin per-rank
caller (`SlabModel`) and the host-driven route over `torch.distributed` point-to-point ops (gloo in CPU tests).

The reference is single-process; the decomposition is the one its layout
suggests: rank p owns global rows [j0+1, j0+ny] plus 3 halo rows on each side, stored
like an (Nx, ny) Field, so each halo message is a contiguous block of 3*(Nx+6) doubles
per field and needs no packing.  Per substage:

    edges    (high-priority stream)  rows within one tile of the slab edges, x-wrapped
    exchange (NCCL stream)           send my edge rows, receive the neighbours' into my halos
    interior (main stream)           everything else, concurrently with the exchange
    finish                           join, swap buffers, tick the clock

There is no data-path collective beyond this neighbour exchange; diagnostics are
combined with one small all_reduce.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import abi

# exchange_rows(which): see include/swmhd.h
SOUTH_SEND, NORTH_SEND, SOUTH_HALO, NORTH_HALO = 0, 1, 2, 3
CURRENT = 4  # add to `which` to address the current state instead of the one being written


def split_rows(Ny: int, world: int):
    """Contiguous row ranges [(j0, ny)] of the `world` slabs (first ranks take the remainder)."""
    base, rem = divmod(Ny, world)
    out, j0 = [], 0
    for r in range(world):
        ny = base + (1 if r < rem else 0)
        out.append((j0, ny))
        j0 += ny
    return out


def neighbours(rank: int, world: int, periodic_y: bool):
    """(south, north) ranks or None at a wall."""
    if world == 1:
        return (rank, rank) if periodic_y else (None, None)
    s = rank - 1 if rank > 0 else (world - 1 if periodic_y else None)
    n = rank + 1 if rank < world - 1 else (0 if periodic_y else None)
    return s, n


def slab_config(cfg: abi.Config, rank: int, world: int, device: int = 0) -> abi.Config:
    c = abi.Config.from_buffer_copy(cfg)
    j0, ny = split_rows(cfg.Ny, world)[rank]
    c.slab_j0, c.slab_ny, c.rank, c.world, c.device = j0, ny, rank, world, device
    return c


def slab_of_global(parent_global: np.ndarray, j0: int, ny: int, extra_rows: int = 0) -> np.ndarray:
    """Rows of a global parent array that form the slab's parent array (3 halo rows each side)."""
    return np.ascontiguousarray(parent_global[j0:j0 + ny + 6 + extra_rows])


class _CudaRows:
    """`__cuda_array_interface__` view of rows inside a libswmhd_cuda buffer."""

    def __init__(self, ptr, nrows, width):
        self.__cuda_array_interface__ = {
            "shape": (nrows, width), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None,
        }


def exchange_ops(rows_of, rank, world, periodic_y, tag_base=0):
    """Build the P2P op list of one halo exchange.

    rows_of(field, which) -> tensor of the 3 rows to send (which = SOUTH_SEND/NORTH_SEND) or to
    receive into (SOUTH_HALO/NORTH_HALO).  My south edge rows go to the south neighbour's north
    halo and vice versa.
    """
    s, n = neighbours(rank, world, periodic_y)
    ops = []
    for f in range(4):
        if n is not None:
            ops.append(dist.P2POp(dist.isend, rows_of(f, NORTH_SEND), n, tag=tag_base + 2 * f))
        if s is not None:
            ops.append(dist.P2POp(dist.irecv, rows_of(f, SOUTH_HALO), s, tag=tag_base + 2 * f))
        if s is not None:
            ops.append(dist.P2POp(dist.isend, rows_of(f, SOUTH_SEND), s, tag=tag_base + 2 * f + 1))
        if n is not None:
            ops.append(dist.P2POp(dist.irecv, rows_of(f, NORTH_HALO), n, tag=tag_base + 2 * f + 1))
    return ops


def self_exchange(rows_of):
    """world == 1 with a slab-style context: periodic wrap by local copies."""
    for f in range(4):
        rows_of(f, SOUTH_HALO).copy_(rows_of(f, NORTH_SEND))
        rows_of(f, NORTH_HALO).copy_(rows_of(f, SOUTH_SEND))


def run_in_step_for(seconds: float, step, device, clock=None) -> int:
    """Call step() repeatedly for about `seconds` of RANK 0's wall clock and return the number of calls,
    the SAME on every rank.  A slab step exchanges halo rows with its neighbours, so ranks that left a
    time-based loop after different numbers of steps would dead-lock: rank 0 decides, the others follow."""
    import time
    clock = clock or time.perf_counter
    t_s = clock()
    go = torch.ones(1, dtype=torch.int32, device=device)
    n = 0
    while int(go.item()):
        step()
        n += 1
        go.fill_(1 if clock() - t_s < seconds else 0)
        dist.broadcast(go, src=0)
    return n


class SlabModel:
    """One y-slab of the global grid on this process's GPU (one process per GPU).

    Thin caller: the halo exchange (ncclSend/ncclRecv), the edge/interior stream overlap and the cross-rank
    reductions of the diagnostics run inside libswmhd_cuda.so once the context holds a communicator.
    torch.distributed only carries the NCCL unique id from rank 0 to the others (plumbing).
    `host_exchange=True` keeps the host-driven path of the ABI instead (swmhd_substage_edges / exchange with
    torch P2P ops / interior / finish): the route a host with its own transport would take.
    """

    def __init__(self, cfg_global: abi.Config, rank=None, world=None, device=None, host_exchange=False):
        from .context import Context
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.device = torch.cuda.current_device() if device is None else device
        self.cfg = slab_config(cfg_global, self.rank, self.world, self.device)
        self.periodic_y = cfg_global.topo_y == abi.PERIODIC
        self.host_exchange = bool(host_exchange)
        self.ctx = Context(self.cfg)
        self._views = {}
        if self.host_exchange:
            self.main = torch.cuda.Stream(device=self.device)
            self.edge = torch.cuda.Stream(device=self.device, priority=-1)
            self.ctx.set_streams(self.main.cuda_stream, self.edge.cuda_stream)
        elif self.world > 1:
            box = [self.ctx.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self.ctx.comm_init(box[0])

    # -- host-driven exchange (legacy route) -------------------------------------------------
    def _rows(self, field, which):
        ptr, n, w = self.ctx.exchange_rows(field, which)
        t = self._views.get(ptr)
        if t is None:
            t = torch.as_tensor(_CudaRows(ptr, n, w), device=f"cuda:{self.device}")
            self._views[ptr] = t
        return t

    def _exchange(self, current: bool, stream):
        if self.world == 1:
            return []   # a single slab owns its own periodic wrap (done by the halo kernel)
        off = CURRENT if current else 0
        rows_of = lambda f, which: self._rows(f, which + off)
        ops = exchange_ops(rows_of, self.rank, self.world, self.periodic_y)
        if not ops:
            return []
        with torch.cuda.stream(stream):
            return dist.batch_isend_irecv(ops)

    def _host_substage(self, dt, stage):
        self.ctx.substage_edges(dt, stage)
        works = self._exchange(current=False, stream=self.edge)
        self.ctx.substage_interior(dt, stage)
        with torch.cuda.stream(self.main):
            for w in works:
                w.wait()
        self.ctx.substage_finish(stage)

    # -- state ------------------------------------------------------------------------------
    def set_state(self, U_slab):
        self.ctx.set_state(U_slab)

    def get_state(self):
        return self.ctx.get_state()

    def fill_halos(self):
        """update_state! after set!: x wrap / walls locally, y halos from the neighbours."""
        self.ctx.fill_halos()
        if self.host_exchange:
            works = self._exchange(current=True, stream=self.main)
            with torch.cuda.stream(self.main):
                for w in works:
                    w.wait()
            self.main.synchronize()

    # -- stepping ---------------------------------------------------------------------------
    def substage(self, dt, stage):
        if self.host_exchange:
            self._host_substage(dt, stage)
        else:
            self.ctx.substage(dt, stage)

    def step(self, dt, nsteps=1):
        if not self.host_exchange:
            return self.ctx.step(dt, nsteps)
        for _ in range(nsteps):
            for stage in (1, 2, 3):
                self._host_substage(dt, stage)

    def step_diag(self, dt, nsteps=1):
        """nsteps steps with the diagnostics of the state at the start of every step, fused into the
        stage-1 kernels; one device->host copy and one small all-reduce for the whole batch."""
        if not self.host_exchange:
            return self.ctx.step_diag(dt, nsteps)
        assert nsteps <= 1024
        for n in range(nsteps):
            self.ctx.arm_diag(n)
            for stage in (1, 2, 3):
                self._host_substage(dt, stage)
        return [self._combine(d) for d in self.ctx.get_diag_slots(0, nsteps)]

    def _combine(self, d):
        """Slab partials are already scaled by the global normalisation: sums add, extrema combine."""
        if self.world == 1:
            return d
        dev = f"cuda:{self.device}"
        sums = torch.tensor([d["ke"], d["me"], d["pe"], d["sum_h"], float(1 - d["all_finite"])], dtype=torch.float64, device=dev)
        maxs = torch.tensor([d["max_abs_u"], d["max_abs_A"], d["max_abs_div_hB"], -d["min_h"]], dtype=torch.float64, device=dev)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
        s, m = sums.tolist(), maxs.tolist()
        return dict(ke=s[0], me=s[1], pe=s[2], total=s[0] + s[1] + s[2], sum_h=s[3], all_finite=int(s[4] == 0),
                    max_abs_u=m[0], max_abs_A=m[1], max_abs_div_hB=m[2], min_h=-m[3])

    def synchronize(self):
        self.ctx.sync()

    @property
    def last_step_ms(self):
        return self.ctx.last_step_ms

    def diagnostics(self) -> dict:
        d = self.ctx.diagnostics(check_finite=False)       # with a communicator: already reduced over the ring
        return self._combine(d) if self.host_exchange else d

    def close(self):
        self._views.clear()
        self.ctx.close()
