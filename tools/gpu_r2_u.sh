#!/bin/bash
# staged epilogue of the projection data gradient (P2) + the 64-register forward projection: tests, kernels alone, bench A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_attention.py tests/test_gpu_network_parity.py tests/test_gpu_query_attention.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
echo staged; timeout 300 python tools/bench_qkv.py 2>&1 | tail -5
echo p2direct; MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_p2direct.so timeout 300 python tools/bench_qkv.py 2>&1 | tail -5
done
bash tools/gpu_ab.sh staged=default p2direct=maskunet_b200/variant_p2direct.so
