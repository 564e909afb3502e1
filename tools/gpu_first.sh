#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "--- persistent TMA"; python tools/quick_bench.py 4096 2>&1 | grep fast
echo "--- nonpersistent TMA"; SWMHD_PERSISTENT=0 python tools/quick_bench.py 4096 2>&1 | grep fast
echo "--- nonpersistent NO_TMA"; SWMHD_PERSISTENT=0 SWMHD_NO_TMA=1 python tools/quick_bench.py 4096 2>&1 | grep fast
