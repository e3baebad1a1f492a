#!/bin/bash
echo "tanh"; timeout 300 python tools/bench_bn.py 2>&1 | tail -4 | cut -c1-330
echo "ex2+rcp"; MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_gelu_ex2.so timeout 300 python tools/bench_bn.py 2>&1 | tail -4 | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_network_parity.py tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -3
cat gpurun_out/network_parity.json | head -60
