#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_kernels.py --generalised --out gpurun_out/kernel_sweep_generalised.json 2>&1 | tail -20 | tee gpurun_out/kernels_generalised.log
timeout 600 python tools/bench_kernels.py --eager --out gpurun_out/kernel_sweep_eager.json 2>&1 | tail -10 | tee gpurun_out/kernels_eager.log
