#!/bin/bash
# 2-GPU pass: gradient equivalence on hardware + the default bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddp_equivalence.py -m gpu -q -x 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo "n2 rc $?"
timeout 600 $TR --master-port 29712 bench.py --gpus 2 --workload coco_panoptic --steps 2 --warmup 1 > gpurun_out/n2_bench_coco.json 2> gpurun_out/n2_bench_coco.err; echo "n2 coco rc $?"
for f in n2_bench n2_bench_coco; do python - <<PY
import json
try:
    r=[json.loads(l) for l in open("gpurun_out/$f.json") if l.startswith("{")][-1]; print("$f", round(r["ms_per_step"],2), round(r["value"],1), "e2e", round(r["e2e"]["value"],1), [round(p["ms_per_step"],1) for p in r["per_rank"]], [p["sm_mhz"] for p in r["per_rank"]], r["clocks"]["reasons"])
except Exception as e: print("$f failed", e)
PY
done
tail -3 gpurun_out/n2_bench.err
