#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ln_view.py tests/test_gpu_network_parity.py -m gpu -q -x 2>&1 | tail -2
python tools/with_clocks.py gpurun_out/ln_view_kernels.json -- python tools/bench_ln_view.py
python - <<'PY'
import json
for r in json.load(open('gpurun_out/ln_view_kernels.json'))['records']: print(r)
PY
