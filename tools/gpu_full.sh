#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_full.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout 1500 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
