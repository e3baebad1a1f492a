#!/bin/bash
echo base; timeout 300 python tools/bench_kernels.py --bwd --batch 64 --site 0 2>&1 | tail -1
for v in probe5 probe1 probe2; do echo $v; MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_$v.so timeout 300 python tools/bench_kernels.py --bwd --batch 64 --site 0 2>&1 | tail -1; done
