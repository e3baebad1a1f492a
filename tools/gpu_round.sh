#!/bin/bash
# full GPU check: tests, smoke, kernel sweep, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python tools/bench_kernels.py --batch 16 2>&1 | tee gpurun_out/kernels.log
timeout 900 python bench.py --steps 2 --warmup 1 --batch-per-gpu ${BPG:-32} --no-cpu-baseline 2>&1 | tail -5 | tee gpurun_out/bench.log
