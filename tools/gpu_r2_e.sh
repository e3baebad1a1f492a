#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/graph_bisect.py --modules > gpurun_out/e_graph_bisect.log 2>&1; grep -v "ok after fwd" gpurun_out/e_graph_bisect.log | tail -40
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_unet.py::test_trainer_cuda_graph_step_matches_eager_steps > gpurun_out/e_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/e_pytest.log
