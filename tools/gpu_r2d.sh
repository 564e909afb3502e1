#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "diag or fused or slab_api or fullsize or models or error or profile" 2>&1 | tail -6
python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast"
python - <<'PY'
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from swmhd_b200 import abi
from swmhd_b200.context import Context
from cases import make_case
for kind in ("J","D"):
    g,cfg,U=make_case(kind,4096,arith=abi.ARITH_FAST)
    c=Context(cfg); c.set_state(U); c.fill_halos(); c.step_diag(0.01*64/4096,5)
    a=c.step_profile(0.01*64/4096,20,diag=True); b=c.step_profile(0.01*64/4096,20)
    print(kind,"stage ms with diag",[round(x,4) for x in a],"plain",[round(x,4) for x in b]); c.close()
PY
