#!/bin/bash
# A/B of the L2 tile prefetch (SWMHD_L2_AHEAD=0 disables it)
echo "--- default"
python tools/quick_bench.py 4096 --fast 2>&1 | grep "fast"
echo "--- SWMHD_L2_AHEAD=0"
SWMHD_L2_AHEAD=0 python tools/quick_bench.py 4096 --fast 2>&1 | grep "fast"
echo "--- SWMHD_RB_STAGES=0"
SWMHD_RB_STAGES=0 python tools/quick_bench.py 4096 --fast 2>&1 | grep "J N.*fast"
