#!/bin/bash
# Run the GPU test groups in separate processes (a CUDA trap poisons its process, not the next group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest -k expr
  echo "=== $1" | tee -a gpurun_out/check.log
  timeout "$2" python -m pytest tests -m gpu -q -x -k "$3" 2>&1 | tail -40 | tee -a gpurun_out/check.log
}
: > gpurun_out/check.log
run mask 300 "mask_binarize or mask_bits"
run stage_fp32 300 "stagewise and float32"
run module_fp32 300 "module_matches and float32"
run sdpa 300 "sdpa_oracle"
run tcgen05_small 300 "stagewise and bfloat16"
run tcgen05_cross 300 "tcgen05_forward"
run module_bf16 300 "module_matches and bfloat16"
run rest 300 "view_not_permute or resample"
