#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/kernels_fwd_poly.log
for v in poly8 poly4; do
  echo "== $v" | tee -a gpurun_out/kernels_fwd_poly.log
  MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_$v.so timeout 300 python tools/bench_kernels.py --batch 16 2>&1 | head -3 | tee -a gpurun_out/kernels_fwd_poly.log
done
