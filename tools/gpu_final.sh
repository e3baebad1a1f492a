#!/bin/bash
# Round-end measurement pass: tests, smoke, both bench arms, ncu launch list + step breakdown, ncu --set full of the
# attention kernels inside the benchmark step (DRAM traffic for roofline.traffic), other configs, kernel sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_full.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout 1500 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-400
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-300
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r8.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench_r8.log 2>&1
python tools/step_breakdown.py gpurun_out/launches_r8.csv 30 | tee gpurun_out/step_breakdown_r8.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:attn_.wd_sm100 -c 12 -f -o gpurun_out/prof_bench_attn_r8 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench_attn_r8.log 2>&1
tail -2 gpurun_out/ncu_bench_attn_r8.log
timeout 900 python tools/bench_configs.py coco_panoptic city_instance 2>&1 | tail -2 | tee gpurun_out/bench_configs.jsonl
timeout 900 python tools/bench_kernels.py --batch 16 --bwd --sweep --out gpurun_out/kernel_sweep.json 2>&1 | tail -20 > gpurun_out/kernels_sweep.log
