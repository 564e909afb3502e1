"""The row-blocked Jacobian kernel (csrc/substage_rb.cu) runs every stage by default; SWMHD_RB_STAGES selects
the stages (bit s-1 = stage s).  The library reads the mask once per process, so the parity suite is
re-run in child processes with stage 1 only on the row-blocked kernel and with none."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.parametrize("mask", ["1", "0"])
def test_parity_suite_with_stage_mask(mask):
    env = dict(os.environ, SWMHD_RB_STAGES=mask)
    r = subprocess.run([sys.executable, "-m", "pytest", str(ROOT / "tests" / "test_gpu_parity.py"), "-m", "gpu", "-q", "-x"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
