"""Compare golden binaries written by julia/dump_reference.jl (real Oceananigans) with the CPU
oracle:  python tools/compare_reference_dump.py <dir>.  Prints rel-L2 per field and step; this is
the check that would turn "parity unpinned" into a pinned statement."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from cases import make_case, rel_l2
from oracle import pyoracle as O

d = Path(sys.argv[1])
for form, kind in (("jacobian", "J"), ("divergence", "D")):
    g, cfg, U = make_case(kind, 64)
    O.fill_halos(cfg, U)
    n = 0
    for k in (0, 1, 10, 100, 1000):
        O.step(cfg, U, 0.01, k - n); n = k
        errs = []
        for f, tag in enumerate("uvhA"):
            p = d / f"golden_{form}_64_step{k}_{tag}.f64"
            if not p.exists():
                errs.append(None); continue
            ref = np.fromfile(p, dtype=np.float64).reshape(U[f].shape)
            errs.append(rel_l2(g, U[f], ref, f))
        print(form, "step", k, ["%.3e" % e if e is not None else "missing" for e in errs])


# ---- Bounded-y with activity at the walls: which C10 combination does upstream implement? -------------------------
from swmhd_b200 import abi
from swmhd_b200.grids import RectilinearGrid, Periodic, Bounded, Flat


def wall_case(form, flags):
    g = RectilinearGrid((64, 64), (-5, 5), (-5, 5), topology=(Periodic, Bounded, Flat))
    cfg = abi.make_config(64, 64, formulation=abi.JACOBIAN if form == "jacobian_wall" else abi.DIVERGENCE, flags=flags,
                          topo_y=abi.BOUNDED, A_gradient=(-0.05, -0.05))
    U = [g.new_parent(k) for k in range(4)]
    g.set_interior(U[abi.H], abi.H, 1.0)
    g.set_interior(U[abi.A], abi.A, lambda x, y, z: -0.05 * y + 0.02 * np.exp(-(x ** 2 + (y - 4) ** 2)))
    g.set_interior(U[abi.U], abi.U, lambda x, y, z: (y - 3.5) * np.exp(-(x ** 2 + (y - 3.5) ** 2)))
    g.set_interior(U[abi.V], abi.V, lambda x, y, z: -x * np.exp(-(x ** 2 + (y - 3.5) ** 2)))
    return g, cfg, U


if (d / "golden_jacobian_wall_64_step0_A.f64").exists():
    names = [("D1", abi.FLAG_BC_DEPTH1), ("W3", abi.FLAG_WALL_WENO3), ("VM", abi.FLAG_V_MIRROR)]
    for form in ("jacobian_wall", "divergence_wall"):
        for combo in range(8):
            flags = sum(f for i, (_, f) in enumerate(names) if combo >> i & 1)
            label = "+".join(n for i, (n, _) in enumerate(names) if combo >> i & 1) or "default"
            g, cfg, U = wall_case(form, flags)
            O.fill_halos(cfg, U)
            n, line = 0, []
            for k in (0, 1, 10, 100):
                O.step(cfg, U, 0.01, k - n); n = k
                worst = 0.0
                for f, tag in enumerate("uvhA"):
                    ref = np.fromfile(d / f"golden_{form}_64_step{k}_{tag}.f64", dtype=np.float64).reshape(U[f].shape)
                    worst = max(worst, float(np.abs(U[f] - ref).max()))     # whole parents: the halo rows are the point
                line.append("%.2e" % worst)
            print(form, label.ljust(9), "max |oracle - upstream| over whole parents at steps 0/1/10/100:", line)
