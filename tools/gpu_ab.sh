#!/bin/bash
# A/B on one box: parity tests, then the stage masks of the row-blocked kernel and optional variants (SWMHD_LIB)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_rb_stages.py -m gpu -q -x 2>&1 | tail -3
echo "--- default (rb stage 1)";  python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
echo "--- rb all stages";  SWMHD_RB_STAGES=7 python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
for v in "$@"; do
  echo "--- $v"; SWMHD_RB_STAGES=7 SWMHD_LIB=$PWD/swmhd_b200/libswmhd_$v.so python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
done
echo "--- one thread per cell"; SWMHD_RB_STAGES=0 python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
