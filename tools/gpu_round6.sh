#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/pytest_full.log
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_nocpu.json | cut -c1-300
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r6.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench_r6.log 2>&1
python tools/step_breakdown.py gpurun_out/launches_r6.csv 30 | tee gpurun_out/step_breakdown_r6.txt
