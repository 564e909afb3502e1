#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_r2b.log
echo "--- r02 kernels"; python tools/quick_bench.py 4096 --fast 2>&1 | grep "N=.*fast" | tee gpurun_out/quick_r2b.log
