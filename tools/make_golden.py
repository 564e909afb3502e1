"""Generate tests/golden/*.npz: small input/output vectors of the CPU oracle.

The reference (Julia + Oceananigans) cannot run in this environment, so these fixtures are
produced by oracle/swmhd_oracle.c (itself pinned by tests/test_oracle_known_answers.py).
They pin the oracle against regressions and give the CUDA path fixed vectors that do not need the
oracle at test time.  julia/dump_reference.jl writes the same layout from real Oceananigans.

    python tools/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from cases import make_case  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

CASES = [  # name, kind, Nx, Ny, dt, steps, perturb
    ("jacobian_40x24", "J", 40, 24, 0.004, (1, 10), 21),
    ("divergence_40x24", "D", 40, 24, 0.004, (1, 10), 22),
    ("jacobian_bounded_36x20", "BJ", 36, 20, 0.004, (1, 10), 23),
    ("divergence_bounded_36x20", "BD", 36, 20, 0.004, (1, 10), 24),
]

if __name__ == "__main__":
    out = ROOT / "tests" / "golden"
    out.mkdir(exist_ok=True)
    for name, kind, Nx, Ny, dt, steps, seed in CASES:
        g, cfg, U = make_case(kind, Nx, Ny=Ny, perturb=seed)
        O.fill_halos(cfg, U)
        data = {f"in_{k}": U[k].copy() for k in range(4)}
        G = O.tendencies(cfg, U)
        data.update({f"G_{k}": G[k] for k in range(4)})
        n = 0
        for s in steps:
            O.step(cfg, U, dt, s - n)
            n = s
            data.update({f"step{s}_{k}": U[k].copy() for k in range(4)})
            d = O.diagnostics(cfg, U)
            data[f"diag{s}"] = np.array([d[k] for k in ("ke", "me", "pe", "max_abs_u", "max_abs_A", "min_h", "sum_h")])
        np.savez_compressed(out / f"{name}.npz", kind=kind, Nx=Nx, Ny=Ny, dt=dt, steps=np.array(steps), seed=seed, **data)
        print("wrote", name, {k: v.shape for k, v in data.items() if k.startswith("in_")})
