#!/bin/bash
mkdir -p gpurun_out
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_trace.so timeout 300 python tools/fwd_trace.py 2>&1 | tail -24 | tee gpurun_out/fwd_trace.txt
