#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_full.log
timeout 600 python tools/bench_kernels.py --batch 16 --bwd 2>&1 | tee gpurun_out/kernels_bwd.log
