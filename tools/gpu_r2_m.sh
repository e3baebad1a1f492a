#!/bin/bash
mkdir -p gpurun_out
for l in 4 8; do echo "lanes per row $l"; MU_CE_LPR=$l timeout 300 python tools/bench_ce.py 2>&1 | tail -3; done
MU_CE_LPR=4 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x -k "cross_entropy" 2>&1 | tail -3
MU_CE_LPR=8 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py -m gpu -q -x -k "cross_entropy" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_network_parity.py tests/test_gpu_instance_loss.py -m gpu -q -x 2>&1 | tail -3
