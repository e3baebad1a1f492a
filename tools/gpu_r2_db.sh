#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_network_parity.py tests/test_gpu_ln_view.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/bench_qkv.py 2>&1 | tail -5
