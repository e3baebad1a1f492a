#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 --batch-per-gpu 256 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-330 | tee gpurun_out/bench_cl.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cl4.csv python bench.py --steps 1 --warmup 1 --batch-per-gpu 256 --no-cpu-baseline > gpurun_out/ncu_bench_cl.log 2>&1
