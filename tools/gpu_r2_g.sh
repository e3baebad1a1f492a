#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network_parity.py tests/test_gpu_attention.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc $?"; tail -6 gpurun_out/g_pytest.log
echo "== sa6 B=128: free-running / deterministic (publish at once) / deterministic (lagged publish)"
timeout 300 python tools/bench_kernels.py --bwd --batch 128 --site 0 2>&1 | grep "^{"
timeout 300 python tools/bench_kernels.py --bwd --batch 128 --site 0 --deterministic 2>&1 | grep "^{"
MASKUNET_B200_LIB=$PWD/maskunet_b200/build_variant_detlag.so timeout 300 python tools/bench_kernels.py --bwd --batch 128 --site 0 --deterministic 2>&1 | grep "^{"
echo "== all sites deterministic"
timeout 300 python tools/bench_kernels.py --bwd --batch 16 --deterministic 2>&1 | grep "^{"
cat gpurun_out/network_parity.json | python -c "
import json,sys
r=json.load(sys.stdin)
for k,v in r.items():
    if k.startswith('deterministic'): print(k, v)
"
