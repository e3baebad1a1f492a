#!/bin/bash
# refresh of the HBM table for the kernels changed after the first capture (fused re-view LayerNorm, P3 with the db column,
# 1x1 head kernels): ncu --set full of their launches inside one benchmark step
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none -k regex:'residual_ln|qkv_dw|conv_wgrad_sm100_kernel<160>|conv_fprop_sm100_kernel<160' -c 34 -f -o gpurun_out/h2_prof_hbm python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/h2_ncu_hbm.log 2>&1
tail -1 gpurun_out/h2_ncu_hbm.log
python tools/ncu_hbm_table.py gpurun_out/h2_prof_hbm.ncu-rep gpurun_out/h2_hbm_table.txt
rm -f gpurun_out/h2_prof_hbm.ncu-rep
