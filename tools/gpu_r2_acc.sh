#!/bin/bash
mkdir -p gpurun_out
python tools/with_clocks.py gpurun_out/acc_conv_dgrad.json -- python tools/bench_conv_acc.py
python - <<'PY'
import json
for r in json.load(open('gpurun_out/acc_conv_dgrad.json'))['records']: print(r)
PY
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q -x 2>&1 | tail -2
