#!/bin/bash
# ncu --set full of the HBM-bound kernel families inside one benchmark step (one step, B = 256)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:'bn_|residual_ln|qkv_|upcat|maxpool|ce_fused|sample_ln|mask_binarize|transpose|column_sums' -c 330 -f -o gpurun_out/h_prof_hbm python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/h_ncu_hbm.log 2>&1
tail -2 gpurun_out/h_ncu_hbm.log
python tools/ncu_hbm_table.py gpurun_out/h_prof_hbm.ncu-rep gpurun_out/h_hbm_table.txt | head -50
ls -la gpurun_out/h_prof_hbm.ncu-rep
