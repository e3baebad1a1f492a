#!/bin/bash
for v in pipe4 pipe8 pipe16; do echo $v; MASKUNET_B200_LIB=$PWD/maskunet_b200/variant_$v.so timeout 300 python tools/bench_kernels.py --batch 64 2>&1 | tail -6 | cut -c1-110; done
