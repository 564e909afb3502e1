"""Occupancy assumptions of the row-blocked kernels, checked from the ptxas log of the in-tree build and from
the shared-memory layout constants of the source: 4 CTAs of 128 threads per SM need <= 128 registers per
thread, no spills, and <= (228 KB - 4 x 1 KB) / 4 of dynamic shared memory (profiles/README.md: the fourth
CTA per SM was worth 5-10 % in both formulations)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LOG = ROOT / "swmhd_b200" / "csrc" / "_obj" / "substage_rb.o.log"
SRC = (ROOT / "swmhd_b200" / "csrc" / "substage_rb.cu").read_text()


def _kernels():
    from swmhd_b200 import build
    build.build(force=not LOG.exists())             # no-op when the library is up to date and its log is there
    text = LOG.read_text()
    out = {}
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers",
                         text, re.S):
        name = re.search(r"substage_(rbd?)_kernelILi(\d)ELb([01])", m.group(1))
        out[(name.group(1), int(name.group(2)), int(name.group(3)))] = dict(stack=int(m.group(2)), spill=int(m.group(3)) + int(m.group(4)), regs=int(m.group(5)))
    return out


def test_row_blocked_kernels_fit_four_ctas_per_sm():
    ks = _kernels()
    assert set(ks) == {(f, s, 0) for f in ("rb", "rbd") for s in (1, 2, 3)} | {("rb", 1, 1), ("rbd", 1, 1)}
    for key, k in ks.items():
        assert k["regs"] <= 128, (key, k)           # 65536 registers / (4 CTAs x 128 threads)
        assert k["spill"] == 0 and k["stack"] == 0, (key, k)


def test_shared_memory_layout_fits_four_ctas_per_sm():
    def macro(name):
        return int(re.search(rf"#define {name} (\d+)", SRC).group(1))
    TX, TYB, R = 32, macro("RB_TY"), macro("RB_R")
    NW, NDIAG = TYB // R, 9
    SZP = ((TX + 6) * (TYB + 6) * 8 + 127) // 128 * 16
    jac = 4 * SZP + 3 * (TX + 6) * (TYB + 5) + 2 * (TX + 2) * (TYB + 2) + 2 + NW * (4 * R + 2)    # ffc arrays at the raw pitch; per-warp east-column scratch
    div = 4 * SZP + 4 * (TX + 6) * (TYB + 4) + (TX + 6) * (TYB + 2) + 2 + NW * NDIAG          # derived arrays at the raw pitch
    assert "constexpr int DERIVED = o_By + NC;" in SRC and "constexpr int DERIVED_D = 4 * NB + NRH;" in SRC
    budget = (228 * 1024 - 4 * 1024) // 4           # 228 KB per SM, 1 KB reserved per resident CTA
    assert jac * 8 <= budget and div * 8 <= budget, (jac * 8, div * 8, budget)
    assert macro("RB_MINB") == 4 and macro("RB_MINB_D") == 4
