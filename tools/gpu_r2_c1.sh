#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network_parity.py tests/test_gpu_unet.py -m gpu -q -x 2>&1 | tail -4
