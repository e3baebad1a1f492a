#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_input.py tests/test_gpu_unet.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/input_test.log
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_nocpu.json | python -c "import json,sys; d=json.load(sys.stdin); print({k: d[k] for k in ('value','ms_per_step','host_enqueue_ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['achieved'])"
