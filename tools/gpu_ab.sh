#!/bin/bash
echo "--- 16x16 tests"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_t16x16.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -3
echo "--- current 32x8"; python tools/quick_bench.py 4096 2>&1 | grep fast
echo "--- 16x16"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_t16x16.so python tools/quick_bench.py 4096 2>&1 | grep fast
