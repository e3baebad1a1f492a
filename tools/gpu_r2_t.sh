#!/bin/bash
mkdir -p gpurun_out
echo "tests, forced pairs"; MU_CONV_PAIR=2 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -2
echo "tests, no pairs"; MU_CONV_PAIR=0 timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x 2>&1 | tail -2
echo "tests, default"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_unet.py tests/test_gpu_network_parity.py -m gpu -q -x 2>&1 | tail -2
echo "pairs + resident"; timeout 600 python tools/bench_conv_ours.py > gpurun_out/t_conv_pair_res.jsonl 2>&1; python tools/conv_totals.py gpurun_out/t_conv_pair_res.jsonl | tail -1
echo "resident only"; MU_CONV_PAIR=0 timeout 600 python tools/bench_conv_ours.py > gpurun_out/t_conv_res.jsonl 2>&1; python tools/conv_totals.py gpurun_out/t_conv_res.jsonl | tail -1
