#!/bin/bash
echo "--- rcp1 tests"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_r1.so python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
echo "--- current"; python tools/quick_bench.py 4096 2>&1 | grep fast
echo "--- rcp1"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_r1.so python tools/quick_bench.py 4096 2>&1 | grep fast
