#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for t in 1 2 4 8 16; do echo "--- TPC=$t"; SWMHD_TPC=$t python tools/quick_bench.py 4096 2>&1 | grep "J.*fast"; done
