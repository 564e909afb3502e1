#!/bin/bash
echo "--- current (minb 3) nonpersistent TMA"; python tools/quick_bench.py 4096 2>&1 | grep fast
echo "--- minb 2"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_mb2.so python tools/quick_bench.py 4096 2>&1 | grep fast
echo "--- minb 4"; SWMHD_LIB=$PWD/swmhd_b200/libswmhd_mb4.so python tools/quick_bench.py 4096 2>&1 | grep fast
