#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_query_attention.py -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/query_attn_test.log
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -5 | tee -a gpurun_out/query_attn_test.log
