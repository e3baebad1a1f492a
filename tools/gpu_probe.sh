#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/bwd_probes3.log
for v in s3 p1 p5; do
  echo "== $v" | tee -a gpurun_out/bwd_probes3.log
  MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_$v.so timeout 300 python tools/bench_kernels.py --batch 16 --bwd --site 0 2>&1 | tail -1 | tee -a gpurun_out/bwd_probes3.log
done
