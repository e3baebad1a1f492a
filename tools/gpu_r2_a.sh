#!/bin/bash
# round 2, pass A (1 GPU): full GPU test-suite, default bench, the other BASELINE configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt
free -g >> gpurun_out/a_gpu.txt; nproc >> gpurun_out/a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err; echo "bench rc $?"
timeout 600 python bench.py --workload coco_panoptic --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_bench_coco.json 2> gpurun_out/a_bench_coco.err; echo "coco rc $?"
timeout 600 python bench.py --workload city_instance_infer --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_city.json 2> gpurun_out/a_bench_city.err; echo "city rc $?"
cut -c1-600 gpurun_out/a_bench_default.json; tail -3 gpurun_out/a_bench_default.err
cut -c1-400 gpurun_out/a_bench_coco.json; tail -3 gpurun_out/a_bench_coco.err
cut -c1-400 gpurun_out/a_bench_city.json; tail -3 gpurun_out/a_bench_city.err
