#!/bin/bash
echo "--- default"; python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast"
for v in "$@"; do
  echo "--- $v"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_$v.so python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast"
  SWMHD_LIB=$PWD/swmhd_b200/libswmhd_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast_one_step or fused or bounded" 2>&1 | tail -1
done
echo "--- default, L2_AHEAD=1184"; SWMHD_L2_AHEAD=1184 python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast"
echo "--- default, L2_AHEAD=296"; SWMHD_L2_AHEAD=296 python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast"
