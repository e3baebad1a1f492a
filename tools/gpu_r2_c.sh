#!/bin/bash
# scaling attribution at N=2: independent replicas / hooks+packing only / 4 buckets / one bucket / 12 buckets
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
i=0
for v in "--no-ddp" "--ddp-dryrun" "" "--bucket-mb 1000" "--bucket-mb 8" "--no-ddp" ""; do
  i=$((i+1))
  timeout 600 $TR --master-port $((29600+i)) bench.py --gpus 2 --steps 20 --warmup 3 $v > gpurun_out/c_n2_$i.json 2> gpurun_out/c_n2_$i.err
  python - <<PY
import json
try:
    r=[json.loads(l) for l in open("gpurun_out/c_n2_$i.json") if l.startswith("{")][-1]
    print("$i [$v]", round(r["ms_per_step"],2), [round(p["ms_per_step"],2) for p in r["per_rank"]], [p["sm_mhz"] for p in r["per_rank"]], r["config"].get("grad_buckets"))
except Exception as e: print("$i [$v] failed", e)
PY
done
