#!/bin/bash
# A/B on one box: parity tests, then kernel variants (SWMHD_LIB) vs the one-thread-per-cell kernel (SWMHD_NO_RB=1)
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
echo "--- rb default";  python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
for v in "$@"; do
  echo "--- $v"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_$v.so python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
done
echo "--- old"; SWMHD_NO_RB=1 python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
