#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_instance_loss.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/instance_loss_test.log
timeout 600 python tools/bench_configs.py coco_instance_term 2>&1 | tail -3 | tee gpurun_out/coco_instance_term.jsonl
