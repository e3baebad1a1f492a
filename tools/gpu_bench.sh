#!/bin/bash
mkdir -p gpurun_out
BPG=${BPG:-256}
timeout 900 python bench.py --steps 3 --warmup 2 --batch-per-gpu $BPG --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench.log
nvidia-smi --query-gpu=memory.used --format=csv | tee -a gpurun_out/bench.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch-per-gpu $BPG --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
