"""Golden fixtures (tests/golden, made by tools/make_golden.py): CPU leg pins the oracle,
GPU leg checks the CUDA path against the stored vectors without calling the oracle."""
from pathlib import Path

import numpy as np
import pytest

from swmhd_b200 import abi
from cases import make_case, rel_l2

GOLD = sorted((Path(__file__).parent / "golden").glob("*.npz"))


def load(p):
    z = np.load(p)
    g, cfg, U = make_case(str(z["kind"]), int(z["Nx"]), Ny=int(z["Ny"]), perturb=int(z["seed"]))
    return z, g, cfg, U


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_oracle_reproduces_golden(path):
    from oracle import pyoracle as O
    z, g, cfg, U = load(path)
    O.fill_halos(cfg, U)
    for k in range(4):
        assert np.array_equal(U[k], z[f"in_{k}"]), "input construction drifted"
    G = O.tendencies(cfg, U)
    for k in range(4):
        assert np.array_equal(G[k], z[f"G_{k}"])
    n = 0
    for s in z["steps"]:
        O.step(cfg, U, float(z["dt"]), int(s) - n)
        n = int(s)
        for k in range(4):
            assert np.array_equal(U[k], z[f"step{s}_{k}"]), (s, k)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
@pytest.mark.parametrize("arith", [abi.ARITH_STRICT, abi.ARITH_FAST])
def test_cuda_matches_golden(path, arith):
    from swmhd_b200.context import Context
    z, g, cfg, U = load(path)
    cfg.arith = arith
    ctx = Context(cfg)
    ctx.set_state([np.ascontiguousarray(z[f"in_{k}"]) for k in range(4)])
    n = 0
    for s in z["steps"]:
        ctx.step(float(z["dt"]), int(s) - n)
        n = int(s)
        out = ctx.get_state()
        for k in range(4):
            ref = z[f"step{s}_{k}"]
            if arith == abi.ARITH_STRICT:
                assert np.array_equal(out[k], ref), (s, k, np.abs(out[k] - ref).max())
            else:
                assert rel_l2(g, out[k], ref, k) <= (1e-12 if s == 1 else 1e-11), (s, k)
        d = ctx.diagnostics()
        ref = z[f"diag{s}"]
        got = np.array([d[k] for k in ("ke", "me", "pe", "max_abs_u", "max_abs_A", "min_h", "sum_h")])
        assert np.allclose(got, ref, rtol=1e-11, atol=1e-15)
    ctx.close()
