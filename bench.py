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
def run_native(args):
    import torch
    from swmhd_b200 import abi
    from swmhd_b200.grids import RectilinearGrid
    from swmhd_b200.context import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SWMHD hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    form = abi.JACOBIAN if args.form == "jacobian" else abi.DIVERGENCE
    arith = abi.ARITH_FAST if args.arith == "fast" else abi.ARITH_STRICT
    Nx = args.size
    if args.scaling == "weak":
        NyG = args.size * world
    else:
        NyG = args.size
    Lx, Ly = 10.0, 10.0 * NyG / Nx
    dt = 0.01 * 64 / Nx
    K, W = args.steps, args.warmup
    sampler = ClockSampler(local) if rank == 0 else None

    bounded = args.topology == "bounded"
    from swmhd_b200.grids import Periodic, Bounded, Flat
    topo = (Periodic, Bounded if bounded else Periodic, Flat)
    cfg_g = abi.make_config(Nx, NyG, Lx=Lx, Ly=Ly, formulation=form, arith=arith, device=local,
                            topo_y=abi.BOUNDED if bounded else abi.PERIODIC, A_gradient=(-0.05, -0.05) if bounded else None)
    e2e = None
    roof = None
    launches = 0
    if world == 1:
        grid = RectilinearGrid((Nx, NyG), (-Lx / 2, Lx / 2), (-Ly / 2, Ly / 2), topology=topo)
        U0 = initial_state(grid, form, pinned=True, bounded=bounded)
        ctx = Context(cfg_g)
        ctx.set_state(U0)
        ctx.fill_halos()
        if sampler:
            sampler.start()
            t_s = time.perf_counter()
            while time.perf_counter() - t_s < 1.5:   # nvidia-smi needs ~1 s before its first sample:
                ctx.step(dt, 5)                      # keep the GPU under the same load meanwhile (untimed)
            ctx.set_state(U0)
            ctx.fill_halos()
        ctx.step_diag(dt, W)
        torch.cuda.synchronize()
        l0 = ctx.launch_count
        diags = ctx.step_diag(dt, K)            # CUDA events on the launching stream inside
        ms_total = ctx.last_step_ms
        launches = ctx.launch_count - l0
        torch.cuda.synchronize()
        # dominant kernel, live: event pair around every fused substage launch
        st_ms = ctx.step_profile(dt, min(K, 20), diag=True)      # the launches of the timed region (stage 1 with diagnostics)
        st_ms_plain = ctx.step_profile(dt, min(K, 20))
        clocks = sampler.stop() if sampler else None
        ncell = Nx * NyG
        peak, peak_src = hbm_peak()
        ach = [ncell * b / (t * 1e-3) / 1e9 for b, t in zip(STAGE_BYTES, st_ms)]
        mean_bytes = ncell * sum(STAGE_BYTES) / 3.0
        mean_ms = sum(st_ms) / 3.0
        achieved = mean_bytes / (mean_ms * 1e-3) / 1e9
        traffic = ncu_traffic(args.form, ncell)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": ("swmhd::substage_rb_kernel<1,DIAG> (stage 1) + swmhd::substage_kernel<0,STAGE> (stages 2, 3)"
                           if form == abi.JACOBIAN and arith == abi.ARITH_FAST else "swmhd::substage_kernel<FORM,STAGE>"),
                "bytes_per_launch": mean_bytes, "ms_per_launch": mean_ms,
                "per_stage": {"ms": st_ms, "GB/s": ach, "ms_without_diagnostics": st_ms_plain},
                "second_ceiling": "FP64 issue: 148 SM x 64 lanes; see DESIGN.md (the kernel is FP64-pipe bound, not HBM bound)"}
        finite = all(d["all_finite"] for d in diags)
        # ---- end to end through the public API with host buffers -------------------------
        if not args.no_e2e and not bounded:
            from swmhd_b200 import models as M
            ctx.close()
            mgrid = M.RectilinearGrid(size=(Nx, NyG), x=(-Lx / 2, Lx / 2), y=(-Ly / 2, Ly / 2), topology=(M.Periodic, M.Periodic, M.Flat))
            if form == abi.JACOBIAN:
                model = M.ShallowWaterModel(grid=mgrid, timestepper="RungeKutta3", momentum_advection=M.WENO5(vector_invariant=M.VelocityStencil()),
                                            mass_advection=M.WENO5(), tracer_advection=M.WENO5(), gravitational_acceleration=9.81,
                                            coriolis=M.FPlane(f=1), tracers=("A",),
                                            forcing=dict(u=M.Forcing(M.lorentz_force_func_x, discrete_form=True), v=M.Forcing(M.lorentz_force_func_y, discrete_form=True)),
                                            formulation=M.VectorInvariantFormulation(), arithmetic=args.arith, device=local)
                names = ("u", "v", "h", "A")
            else:
                model = M.ShallowWaterModel(grid=mgrid, timestepper="RungeKutta3", momentum_advection=M.WENO5(), mass_advection=M.WENO5(),
                                            tracer_advection=M.WENO5(), gravitational_acceleration=9.81, coriolis=M.FPlane(f=1), tracers=("A",),
                                            forcing=dict(uh=M.Forcing(M.div_lorentz_x, discrete_form=True), vh=M.Forcing(M.div_lorentz_y, discrete_form=True)),
                                            formulation=M.ConservativeFormulation(), arithmetic=args.arith, device=local)
                names = ("uh", "vh", "h", "A")
            ke = min(K, 10)
            h2d = sum(a.nbytes for a in U0)
            for _ in range(2):
                M.set_b(model, **{n: U0[k] for k, n in enumerate(names)})
                M.time_step_diag_b(model, dt)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(ke):
                M.set_b(model, **{n: U0[k] for k, n in enumerate(names)})   # H2D of the four haloed fields (pinned)
                d = M.time_step_diag_b(model, dt)                            # one RK3 step + D2H of its diagnostics
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            e2e = {"value": ncell * ke / el, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 9 * 8,
                   "ms_per_step": el / ke * 1e3, "steps": ke,
                   "what": "set!(model, 4 haloed fields from pinned host memory) + time_step! + energy/div(hB) diagnostics to host, per step"}
            model.close()
        else:
            ctx.close()
            e2e = None
    else:
        from swmhd_b200.distributed import SlabModel, split_rows, run_in_step_for
        j0, ny = split_rows(NyG, world)[rank]
        # slab ICs straight from the closed forms at this slab's nodes (halo rows come from the exchange)
        gridl = RectilinearGrid((Nx, ny), (-Lx / 2, Lx / 2), (-Ly / 2 + j0 * (Ly / NyG), -Ly / 2 + (j0 + ny) * (Ly / NyG)), topology=topo)
        U0 = initial_state(gridl, form, pinned=True, bounded=bounded)
        sm = SlabModel(cfg_g, rank, world, local)
        sm.set_state(U0)
        sm.fill_halos()
        sm.step_diag(dt, W)
        sm.synchronize()
        if sampler:
            sampler.start()
        # same untimed load on every rank while nvidia-smi starts; every rank runs the same number of steps
        # (rank 0's clock decides: a slab step exchanges halo rows, uneven counts would dead-lock)
        def _warm():
            sm.step(dt, 2)
            sm.synchronize()
        run_in_step_for(1.5, _warm, f"cuda:{local}")
        sm.synchronize()
        dist.barrier(); torch.cuda.synchronize()
        l0 = sm.ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sm.main):
            e0.record(sm.main)
        sm.ctx.sync()
        for n in range(K):                       # diagnostics every step, fused in the stage-1 kernels
            sm.ctx.arm_diag(n % 1024)
            for stage in (1, 2, 3):
                sm.substage(dt, stage)
        with torch.cuda.stream(sm.main):
            e1.record(sm.main)
        sm.synchronize(); torch.cuda.synchronize()
        finite = all(d["all_finite"] for d in sm.ctx.get_diag_slots(0, min(K, 1024)))
        ms_local = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
        dist.barrier()
        dist.all_reduce(ms_local, op=dist.ReduceOp.MAX)
        ms_total = float(ms_local.item())
        nl = torch.tensor([sm.ctx.launch_count - l0], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(nl)
        launches = int(nl.item())
        clocks = sampler.stop() if sampler else None
        ncell = Nx * NyG
        peak, peak_src = hbm_peak()
        achieved = ncell * BYTES_PER_CELL_UPDATE * K / (ms_total * 1e-3) / 1e9 / world
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "note": "per GPU, whole step (kernels + exchange), not a single-kernel figure"}
        # ---- end to end at N GPUs: every rank uploads its slab from pinned host memory, halos are exchanged,
        # one step runs and the (all-reduced) diagnostics come back to the host, per step
        if not args.no_e2e:
            ke = min(K, 10)
            h2d = sum(a.nbytes for a in U0)
            for _ in range(2):
                sm.set_state(U0); sm.fill_halos(); sm.step_diag(dt, 1)
            sm.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for _ in range(ke):
                sm.set_state(U0)
                sm.fill_halos()
                d = sm.step_diag(dt, 1)
            sm.synchronize(); dist.barrier()
            el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
            el = float(el.item())
            e2e = {"value": ncell * ke / el, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 9 * 8 * world,
                   "ms_per_step": el / ke * 1e3, "steps": ke,
                   "what": "per rank: slab upload from pinned host memory (4 haloed fields) + NCCL halo exchange + one RK3 step + diagnostics (D2H + all-reduce), per step"}
        sm.close()

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    value = ncell * K / (ms_total * 1e-3)
    cpu = None
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
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.form} formulation {Nx}x{NyG} {'Bounded-y' if bounded else 'periodic'} FP64 RK3 step, energy/div(hB) diagnostics every step "
                               f"(BASELINE config 3 per GPU; y-slabs of {Nx}x{NyG // world})",
                   "arith": args.arith, "dt": dt, "ic": "IC-B (divergence_sw_mhd.jl:17,34-37)" if bounded else ("IC-J (SWMHD_example.jl:36-40)" if form == abi.JACOBIAN else "IC-D (divergence_sw_mhd.jl:33-38)"),
                   "l2": "working set 12 fields x %.0f MB >> 126 MB L2 (inputs larger than L2, no flush needed)" % (Nx * (NyG // world) * 8 / 1e6),
                   "timing": "CUDA events on the launching stream around K steps, max over ranks",
                   "all_finite": bool(finite)},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if dist:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
