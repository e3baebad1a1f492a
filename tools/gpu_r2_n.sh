#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s -k "dq_bf16_accumulation" 2>&1 | grep -i "relative error\|passed\|failed"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -1 gpurun_out/n_bench.json | cut -c1-1500
