#!/usr/bin/env python
"""bench.py — FP64 cell-updates/s of the SWMHD RK3 step on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One "step" = one full RK3 time step (3 fused substage kernels + halo fills) of the
Jacobian-formulation model with the energy / div(hB) diagnostics evaluated every
step (BASELINE config 3: 4096^2 periodic FP64 per GPU).  N > 1 (torchrun, one rank per
GPU) stacks N such slabs in y (weak scaling) with NCCL halo exchange between slabs.

Output: ONE JSON line on rank 0 (contract in the task statement):
  value      device-timed whole-job cell-updates/s, state resident in HBM
  e2e        same metric through the public Python API with host (pinned) buffers:
             set!(model, ...) H2D of the four haloed fields + one step + D2H of the diagnostics
  roofline   fused substage kernel: algorithmic bytes per launch / CUDA-event duration,
             against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the CPU oracle (port of the reference algorithm) on the host cores,
             bounded sample of the same workload
--impl reference times that CPU oracle alone (the reference is Julia-only and cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

# stdout carries exactly one JSON line.  Libraries print there too (NCCL's "NCCL version ..." banner when
# NCCL_DEBUG is set on the box): the process's fd 1 is pointed at stderr and the JSON line goes to a
# duplicate of the original stdout.
try:
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
except OSError:                     # no usable stderr / stdout descriptors: print the JSON line the plain way
    _JSON_OUT = sys.stdout

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "fp64_cell_updates_per_s"
UNIT = "cell-updates/s"
BYTES_PER_CELL_UPDATE = 320.0          # 40 doubles per cell per RK3 step
STAGE_BYTES = (96.0, 128.0, 96.0)      # per cell: stage 1: 4R+8W, stage 2: 8R+8W, stage 3: 8R+4W


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="Nx and rows per GPU")
    ap.add_argument("--form", default="jacobian", choices=["jacobian", "divergence"])
    ap.add_argument("--arith", default="fast", choices=["fast", "strict"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--topology", default="periodic", choices=["periodic", "bounded"], help="y topology (x is periodic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the config4 / config5 / slab_parity objects")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
def initial_state(grid, form, pinned=False, bounded=False):
    """IC-J (SWMHD_example.jl:36-40) or IC-D (divergence_sw_mhd.jl:33-38) on the host; Bounded-y:
    IC-B, A = -0.05 y with its gradient BC and a unit vortex (divergence_sw_mhd.jl:17,34-37)."""
    from swmhd_b200 import abi
    U = []
    for k in range(4):
        shape = grid.parent_shape(k)
        if pinned:
            import torch
            a = torch.zeros(shape, dtype=torch.float64).pin_memory().numpy()
        else:
            a = np.zeros(shape)
        U.append(a)
    grid.set_interior(U[abi.H], abi.H, 1.0)
    if bounded:
        grid.set_interior(U[abi.U], abi.U, lambda x, y, z: y * np.exp(-(x ** 2 + y ** 2)))
        grid.set_interior(U[abi.V], abi.V, lambda x, y, z: -x * np.exp(-(x ** 2 + y ** 2)))
        grid.set_interior(U[abi.A], abi.A, lambda x, y, z: -0.05 * y)
    elif form == abi.JACOBIAN:
        grid.set_interior(U[abi.U], abi.U, lambda x, y, z: 5 * y * np.exp(-(x ** 2 + y ** 2)))
        grid.set_interior(U[abi.V], abi.V, lambda x, y, z: -5 * x * np.exp(-(x ** 2 + y ** 2)))
        grid.set_interior(U[abi.A], abi.A, lambda x, y, z: 0.5 * np.abs(y))
    else:
        grid.set_interior(U[abi.A], abi.A, lambda x, y, z: 0.5 * np.exp(-((x - 0.5) ** 2 + y ** 2)) - 0.5 * np.exp(-((x + 0.5) ** 2 + y ** 2)))
    return U


class ClockSampler:
    """nvidia-smi -lms sampler of SM clocks / throttle reasons running DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.strip()]
            except Exception:
                pass
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(form, ncells):
    """DRAM bytes per substage-kernel launch (mean of the three stages) from the committed
    ncu --set full capture (profiles/traffic.json holds bytes per cell measured at 2048^2), or None."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())[form]["bytes_per_cell_per_launch_mean"] * ncells
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------
def cpu_reference_rate(form, Nx, rows, steps, warmup=0):
    """cell-updates/s of the CPU oracle on an Nx x rows periodic sample of the workload."""
    from swmhd_b200 import abi
    from swmhd_b200.grids import RectilinearGrid
    from oracle import pyoracle as O
    O.set_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    grid = RectilinearGrid((Nx, rows), (-5, 5), (-5 * rows / Nx, 5 * rows / Nx))
    cfg = abi.make_config(Nx, rows, Lx=grid.Lx, Ly=grid.Ly, formulation=form)
    U = initial_state(grid, form)
    O.fill_halos(cfg, U)
    dt = 0.01 * 64 / Nx
    if warmup:
        O.step(cfg, U, dt, warmup)
    t0 = time.perf_counter()
    O.step(cfg, U, dt, steps)
    el = time.perf_counter() - t0
    return Nx * rows * steps / el, el


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The reference itself is
    Julia + Oceananigans (not installable here), so this is the CPU oracle port, all threads."""
    from swmhd_b200 import abi
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    form = abi.JACOBIAN if args.form == "jacobian" else abi.DIVERGENCE
    cores = os.cpu_count() or 1
    Nx, rows = args.size, 256
    rate, el = cpu_reference_rate(form, Nx, rows, args.steps, args.warmup)
    ms = el / args.steps * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.form} formulation {args.size}^2 periodic FP64 RK3 step (BASELINE config 3)",
                   "sample": f"each step = one RK3 step of a {Nx}x{rows} periodic band of the workload grid",
                   "note": "reference is Julia/Oceananigans (no Julia in this image): CPU oracle port, OpenMP"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} RK3 steps of a {Nx}x{rows} band, {cores} OpenMP threads"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ---------------------------------------------------------------------------------------------
class Env:
    """Launch environment of one rank (torchrun exports RANK / LOCAL_RANK / WORLD_SIZE)."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the SWMHD hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = f"cuda:{self.local}"
        self.numa = None
        if self.world > 1:
            # N ranks push their slabs through the host at once in the e2e leg: keep this rank (and the pinned pages it
            # first-touches) on the CPUs / NUMA node next to its GPU.  Not at N = 1: the CPU baseline wants every core.
            try:
                import pynvml
                pynvml.nvmlInit()
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.local))
                self.numa = f"cpu affinity of GPU {self.local}: {len(os.sched_getaffinity(0))} cpus"
            except Exception as ex:
                self.numa = f"unset ({type(ex).__name__})"
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device(self.dev))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if not self.dist:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if not self.dist:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t)
        return float(t.item())


class Leg:
    """One workload on all ranks: global (Nx, NyG) grid, this rank's y-slab resident in HBM."""

    def __init__(self, env, form_name, Nx, NyG, bounded, arith_name="fast", pinned=True):
        from swmhd_b200 import abi
        from swmhd_b200.grids import RectilinearGrid, Periodic, Bounded, Flat
        from swmhd_b200.context import Context
        from swmhd_b200.distributed import SlabModel, split_rows
        self.env, self.Nx, self.NyG, self.bounded = env, Nx, NyG, bounded
        self.form = abi.JACOBIAN if form_name == "jacobian" else abi.DIVERGENCE
        self.form_name = form_name
        self.arith = abi.ARITH_FAST if arith_name == "fast" else abi.ARITH_STRICT
        self.Lx, self.Ly = 10.0, 10.0 * NyG / Nx
        self.dt = 0.01 * 64 / Nx
        topo = (Periodic, Bounded if bounded else Periodic, Flat)
        self.cfg = abi.make_config(Nx, NyG, Lx=self.Lx, Ly=self.Ly, formulation=self.form, arith=self.arith, device=env.local,
                                   topo_y=abi.BOUNDED if bounded else abi.PERIODIC, A_gradient=(-0.05, -0.05) if bounded else None)
        j0, ny = split_rows(NyG, env.world)[env.rank]
        self.ny = ny
        dy = self.Ly / NyG
        # slab ICs straight from the closed forms at this slab's nodes (halo rows come from fill_halos)
        grid = RectilinearGrid((Nx, ny), (-self.Lx / 2, self.Lx / 2), (-self.Ly / 2 + j0 * dy, -self.Ly / 2 + (j0 + ny) * dy), topology=topo)
        self.U0 = initial_state(grid, self.form, pinned=pinned, bounded=bounded)
        self.m = Context(self.cfg) if env.world == 1 else SlabModel(self.cfg, env.rank, env.world, env.local)
        self.ctx = self.m if env.world == 1 else self.m.ctx
        self.reset()

    def reset(self):
        self.m.set_state(self.U0)
        self.m.fill_halos()

    def keep_busy(self, seconds):
        """Untimed load while nvidia-smi starts sampling; every rank runs the same number of steps."""
        from swmhd_b200.distributed import run_in_step_for
        if self.env.world == 1:
            t_s = time.perf_counter()
            while time.perf_counter() - t_s < seconds:
                self.m.step(self.dt, 3)
        else:
            run_in_step_for(seconds, lambda: self.m.step(self.dt, 2), self.env.dev)

    def timed(self, K, W, sampler=None, busy=1.5):
        """W warm-up steps, then K steps (diagnostics every step, fused into stage 1) between CUDA events on
        the launching stream (inside swmhd_step_diag), barrier + synchronize on both sides, max over ranks."""
        env = self.env
        if sampler:                 # rank 0 only
            sampler.start()
        if busy > 0:                # EVERY rank: a slab step exchanges halo rows with its neighbours
            self.keep_busy(busy)
            self.reset()
        self.m.step_diag(self.dt, max(W, 1))
        env.barrier()
        l0 = self.ctx.launch_count
        diags = self.m.step_diag(self.dt, K)
        ms = env.max_over_ranks(self.ctx.last_step_ms)
        env.barrier()
        launches = int(env.sum_over_ranks(self.ctx.launch_count - l0))
        clocks = sampler.stop() if sampler else None
        finite = all(d["all_finite"] for d in diags)
        ncell = self.Nx * self.NyG
        return {"value": ncell * K / (ms * 1e-3), "ms_per_step": ms / K, "steps": K, "warmup": W, "gpu_launches": launches,
                "all_finite": bool(finite), "clocks": clocks,
                "hbm_frac_per_gpu": ncell * BYTES_PER_CELL_UPDATE * K / (ms * 1e-3) / 1e9 / env.world / hbm_peak()[0]}

    def workload(self):
        w = self.env.world
        return (f"{self.form_name} formulation {self.Nx}x{self.NyG} {'Bounded-y' if self.bounded else 'periodic'} FP64 RK3 step, "
                f"energy/div(hB) diagnostics every step" + (f"; {w} y-slabs of {self.Nx}x{self.NyG // w}, NCCL halo exchange inside libswmhd_cuda.so" if w > 1 else ""))

    def close(self):
        self.m.close()


def roofline_single(leg, K):
    """Dominant kernel, live: CUDA-event pair around every fused substage launch (single GPU)."""
    ctx = leg.ctx
    st_ms = ctx.step_profile(leg.dt, min(K, 20), diag=True)      # the launches of the timed region (stage 1 with diagnostics)
    st_ms_plain = ctx.step_profile(leg.dt, min(K, 20))
    ncell = leg.Nx * leg.NyG
    peak, peak_src = hbm_peak()
    ach = [ncell * b / (t * 1e-3) / 1e9 for b, t in zip(STAGE_BYTES, st_ms)]
    mean_bytes = ncell * sum(STAGE_BYTES) / 3.0
    mean_ms = sum(st_ms) / 3.0
    achieved = mean_bytes / (mean_ms * 1e-3) / 1e9
    kern = "substage_rb_kernel" if leg.form_name == "jacobian" else "substage_rbd_kernel"
    kernel = (f"swmhd::{kern}<1,DIAG> (stage 1), swmhd::{kern}<2,0>, swmhd::{kern}<3,0> (csrc/substage_rb.cu)"
              if leg.arith == 0 and leg.Nx % 2 == 0 else "swmhd::substage_kernel<FORM,STAGE,DIAG,TMA,NSTG> (csrc/substage_kernel.cu)")
    prof = fp64_profile(leg.form_name)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": ncu_traffic(leg.form_name, ncell), "peak_source": peak_src, "kernel": kernel,
            "bytes_per_launch": mean_bytes, "ms_per_launch": mean_ms,
            "per_stage": {"ms": st_ms, "GB/s": ach, "ms_without_diagnostics": st_ms_plain},
            "second_ceiling": "FP64 issue: 148 SM x 64 lanes; see DESIGN.md (the kernel is FP64-pipe bound, not HBM bound)"}
    roof.update(prof)
    return roof


def fp64_profile(form_name):
    """FP64-pipe utilisation and instructions per cell-substage of the committed ncu capture (profiles/traffic.json)."""
    p = ROOT / "profiles" / "traffic.json"
    out = {"fp64_pipe_pct": None, "instr_per_cell_substage": None}
    try:
        d = json.loads(p.read_text())[form_name]
        out["fp64_pipe_pct"] = d.get("fp64_pipe_pct")
        out["instr_per_cell_substage"] = d.get("instr_per_cell_substage")
    except Exception:
        pass
    return out


def e2e_single(env, leg, K, arith_name):
    """set!(model, 4 haloed fields from pinned host memory) + time_step! + diagnostics to the host, per step,
    through the Python mirror of the reference's model API."""
    from swmhd_b200 import abi
    from swmhd_b200 import models as M
    Nx, NyG, U0, dt, form = leg.Nx, leg.NyG, leg.U0, leg.dt, leg.form
    mgrid = M.RectilinearGrid(size=(Nx, NyG), x=(-leg.Lx / 2, leg.Lx / 2), y=(-leg.Ly / 2, leg.Ly / 2), topology=(M.Periodic, M.Periodic, M.Flat))
    if form == abi.JACOBIAN:
        model = M.ShallowWaterModel(grid=mgrid, timestepper="RungeKutta3", momentum_advection=M.WENO5(vector_invariant=M.VelocityStencil()),
                                    mass_advection=M.WENO5(), tracer_advection=M.WENO5(), gravitational_acceleration=9.81,
                                    coriolis=M.FPlane(f=1), tracers=("A",),
                                    forcing=dict(u=M.Forcing(M.lorentz_force_func_x, discrete_form=True), v=M.Forcing(M.lorentz_force_func_y, discrete_form=True)),
                                    formulation=M.VectorInvariantFormulation(), arithmetic=arith_name, device=env.local)
        names = ("u", "v", "h", "A")
    else:
        model = M.ShallowWaterModel(grid=mgrid, timestepper="RungeKutta3", momentum_advection=M.WENO5(), mass_advection=M.WENO5(),
                                    tracer_advection=M.WENO5(), gravitational_acceleration=9.81, coriolis=M.FPlane(f=1), tracers=("A",),
                                    forcing=dict(uh=M.Forcing(M.div_lorentz_x, discrete_form=True), vh=M.Forcing(M.div_lorentz_y, discrete_form=True)),
                                    formulation=M.ConservativeFormulation(), arithmetic=arith_name, device=env.local)
        names = ("uh", "vh", "h", "A")
    ke = min(K, 10)
    h2d = sum(a.nbytes for a in U0)
    fields = {n: U0[k] for k, n in enumerate(names)}
    for _ in range(2):
        M.set_b(model, **fields)
        M.time_step_diag_b(model, dt)
        M.set_and_step_diag_b(model, dt, **fields)
    env.torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(ke):                                              # two separate calls: set!, then time_step!
        M.set_b(model, **fields)                                     # H2D of the four haloed fields (pinned)
        M.time_step_diag_b(model, dt)                                # one RK3 step + D2H of its diagnostics
    env.torch.cuda.synchronize()
    el_sep = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(ke):                                              # one call: the upload pipelined with stage 1 by row bands
        M.set_and_step_diag_b(model, dt, **fields)
    env.torch.cuda.synchronize()
    el = time.perf_counter() - t0
    model.close()
    return {"value": Nx * NyG * ke / el, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 9 * 8,
            "ms_per_step": el / ke * 1e3, "steps": ke, "ms_per_step_separate_calls": el_sep / ke * 1e3,
            "what": "set_and_step!(model, 4 haloed fields from pinned host memory, Δt): upload pipelined with stage 1 (swmhd_upload_step) + stages 2, 3 + "
                    "energy/div(hB) diagnostics to host, per step; ms_per_step_separate_calls = set!(...) then time_step!(...) as two calls",
            "bound": "PCIe: the upload of the four fields is %.1f ms of the step at the measured rate" % (h2d / 55e9 * 1e3)}


def e2e_slabs(env, leg, K):
    """Every rank uploads its slab from pinned host memory, halos are exchanged, one step runs and the ring-reduced
    diagnostics come back to the host, per step."""
    ke = min(K, 10)
    h2d = sum(a.nbytes for a in leg.U0)
    for _ in range(2):
        leg.reset(); leg.m.step_diag(leg.dt, 1)
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        leg.reset()
        leg.m.step_diag(leg.dt, 1)
    env.barrier()
    el = env.max_over_ranks(time.perf_counter() - t0)
    return {"value": leg.Nx * leg.NyG * ke / el, "unit": UNIT, "h2d_bytes_per_step": h2d * env.world, "d2h_bytes_per_step": 9 * 8 * env.world,
            "ms_per_step": el / ke * 1e3, "steps": ke,
            "what": "per rank: slab upload from pinned host memory (4 haloed fields) + NCCL halo exchange + one RK3 step + ring-reduced diagnostics to host, per step",
            "aggregate_h2d_GBps": h2d * env.world * ke / el / 1e9,
            "bound": "host side: every rank pushes its slab through the host at once; the aggregate host->device rate of the box saturates "
                     "(~170 GB/s measured at 8 ranks against 55 GB/s for one), the device part of the step is unchanged",
            "numa": env.numa}


def slab_parity(env):
    """N-GPU y-slab runs against the single-GPU run of the same problem, untimed: fields must be BIT-IDENTICAL
    (same per-cell arithmetic, no reduction in the state update), diagnostics equal to 1e-13.  All four kinds:
    both formulations, periodic and Bounded-y.  Rank 0 also drives a single-process n_gpus context."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    from cases import make_case
    from swmhd_b200 import abi
    from swmhd_b200.context import Context
    from swmhd_b200.distributed import SlabModel, split_rows, slab_of_global
    torch, dist = env.torch, env.dist
    out = {"world": env.world, "cases": {}, "steps": 4}
    keys = ("ke", "me", "pe", "sum_h", "max_abs_u", "max_abs_A", "min_h")
    ok_all = True
    for kind, Nx, Ny in [("J", 256, 264), ("D", 256, 264), ("BJ", 128, 200), ("BD", 128, 200)]:
        g, cfg, U = make_case(kind, Nx, Ny=Ny, perturb=31)
        cfg.device = env.local
        ref = Context(cfg); ref.set_state(U); ref.fill_halos()
        tr_ref = ref.step_diag(0.002, 4); Uref = ref.get_state(); ref.close()
        j0, ny = split_rows(Ny, env.world)[env.rank]
        extra = lambda k: 1 if (k == abi.V and cfg.topo_y == abi.BOUNDED) else 0
        sm = SlabModel(cfg, env.rank, env.world, env.local)
        sm.set_state([slab_of_global(U[k], j0, ny, extra(k)) for k in range(4)])
        sm.fill_halos()
        tr = sm.step_diag(0.002, 4)
        outU = sm.get_state()
        sm.close()
        ok = all(np.array_equal(outU[k][3:3 + ny], Uref[k][3 + j0:3 + j0 + ny]) for k in range(4))
        dok = all(abs(a[key] - b[key]) <= 1e-13 * max(1.0, abs(b[key])) for a, b in zip(tr, tr_ref) for key in keys)
        t = torch.tensor([int(ok), int(dok)], device=env.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out["cases"][kind] = {"grid": [Nx, Ny], "fields_bit_identical": bool(t[0]), "diagnostics_1e-13": bool(t[1])}
        ok_all = ok_all and bool(t[0]) and bool(t[1])
    # single-process multi-GPU context (cfg.n_gpus): rank 0 drives every visible device of the job, the others wait
    inproc = None
    env.barrier()
    if env.rank == 0 and torch.cuda.device_count() >= env.world:
        try:
            g, cfg, U = make_case("J", 256, Ny=264, perturb=31)
            ref = Context(cfg); ref.set_state(U); ref.fill_halos(); ref.step(0.002, 4); Uref = ref.get_state(); ref.close()
            cfg2 = abi.Config.from_buffer_copy(cfg)
            cfg2.n_gpus = env.world
            for d in range(env.world):
                cfg2.device_ids[d] = d
            mg = Context(cfg2); mg.set_state(U); mg.fill_halos(); mg.step(0.002, 4); Um = mg.get_state(); mg.close()
            inproc = all(np.array_equal(Um[k], Uref[k]) for k in range(4))
        except Exception as ex:       # reported, never fatal for the bench line
            inproc = f"failed: {ex}"
    env.barrier()
    out["single_process_n_gpus_bit_identical"] = inproc
    out["ok"] = bool(ok_all and (inproc is None or inproc is True))
    return out


def run_native(args):
    env = Env(args)
    torch = env.torch
    world, rank = env.world, env.rank
    Nx = args.size
    NyG = args.size * world if args.scaling == "weak" else args.size
    K, W = args.steps, args.warmup
    bounded = args.topology == "bounded"
    sampler = ClockSampler(env.local) if rank == 0 else None

    # ---- primary leg: the headline workload (BASELINE config 3 per GPU unless flags say otherwise) -------
    leg = Leg(env, args.form, Nx, NyG, bounded, args.arith)
    res = leg.timed(K, W, sampler)
    workload = leg.workload()
    roof = e2e = None
    if world == 1:
        roof = roofline_single(leg, K)
        if not args.no_e2e and not bounded:
            e2e = e2e_single(env, leg, K, args.arith)
    else:
        peak, peak_src = hbm_peak()
        ach = res["hbm_frac_per_gpu"] * peak
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src, "note": "per GPU, whole step (kernels + exchange), not a single-kernel figure"}
        if not args.no_e2e:
            e2e = e2e_slabs(env, leg, K)
    ic = "IC-B (divergence_sw_mhd.jl:17,34-37)" if bounded else ("IC-J (SWMHD_example.jl:36-40)" if args.form == "jacobian" else "IC-D (divergence_sw_mhd.jl:33-38)")
    dt = leg.dt
    leg.close()
    del leg

    # ---- the other BASELINE configurations and the N-GPU parity proof, in the same line -----------------
    extra = {}
    if not args.no_extra_legs:
        Kx, Wx = min(K, 20), 3
        # config 4: divergence formulation 16384^2, strong scaling (the whole grid on 1, 2, 4, 8 GPUs)
        lg = Leg(env, "divergence", 16384, 16384, False, args.arith, pinned=False)
        r4 = lg.timed(Kx, Wx, ClockSampler(env.local) if rank == 0 else None, busy=1.0)
        r4.update({"workload": lg.workload() + " (BASELINE config 4)", "scaling": "strong", "unit": UNIT})
        lg.close(); del lg
        extra["config4"] = r4
        # config 5: Jacobian formulation, Bounded-y, 8192^2 per GPU, weak scaling
        lg = Leg(env, "jacobian", 8192, 8192 * world, True, args.arith, pinned=False)
        r5 = lg.timed(Kx, Wx, ClockSampler(env.local) if rank == 0 else None, busy=1.0)
        r5.update({"workload": lg.workload() + " (BASELINE config 5)", "scaling": "weak", "unit": UNIT})
        lg.close(); del lg
        extra["config5"] = r5
        if world > 1:
            extra["slab_parity"] = slab_parity(env)

    if rank != 0:
        if env.dist:
            env.dist.destroy_process_group()
        return
    cpu = None
    form = 0 if args.form == "jacobian" else 1
    if world == 1 and not args.no_cpu_baseline:
        try:
            rows = 256
            r1, el1 = cpu_reference_rate(form, Nx, rows, 1)
            n = int(max(2, min(40, 15.0 / max(el1, 1e-3))))
            rate, el = cpu_reference_rate(form, Nx, rows, n)
            cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{n} RK3 steps of a {Nx}x{rows} periodic band of the workload ({el:.1f} s), CPU oracle (OpenMP, all cores); "
                             "the reference is Julia/Oceananigans and cannot be built here"}
        except Exception as ex:  # the oracle is a reported baseline, never a dependency of the product path
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload + (" (BASELINE config 3 per GPU)" if args.size == 4096 and args.form == "jacobian" and not bounded else ""),
                   "arith": args.arith, "dt": dt, "ic": ic,
                   "l2": "working set 12 fields x %.0f MB >> 126 MB L2 (inputs larger than L2, no flush needed)" % (Nx * (NyG // world) * 8 / 1e6),
                   "timing": "CUDA events on the launching stream around K steps (inside swmhd_step_diag), barrier + synchronize on both sides, max over ranks",
                   "all_finite": res["all_finite"]},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": res["gpu_launches"], "clocks": res["clocks"],
    }
    line.update(extra)
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if env.dist:
        env.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
