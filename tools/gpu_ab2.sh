#!/bin/bash
# A/B on one box: the in-tree build, then variants built with tools/build_variant.sh (arguments: variant tags)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -2
echo "--- default"; python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast"
for v in "$@"; do
  echo "--- $v"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_$v.so python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast"
  SWMHD_LIB=$PWD/swmhd_b200/libswmhd_$v.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fast_one_step or fused or bounded" 2>&1 | tail -1
done
