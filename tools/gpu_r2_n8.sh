#!/bin/bash
# round 2, 8-GPU pass: weak-scaling default workload with and without the exchange, COCO-panoptic global batch 2048
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29701 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "n8 rc $?"
timeout 600 $TR --master-port 29702 bench.py --gpus 8 --steps 10 --warmup 3 --no-ddp > gpurun_out/n8_bench_noddp.json 2> gpurun_out/n8_bench_noddp.err; echo "n8 noddp rc $?"
timeout 600 $TR --master-port 29703 bench.py --gpus 8 --workload coco_panoptic --steps 3 --warmup 2 > gpurun_out/n8_bench_coco.json 2> gpurun_out/n8_bench_coco.err; echo "n8 coco rc $?"
timeout 600 $TR --master-port 29704 bench.py --gpus 8 --workload city_instance_infer --steps 5 --warmup 3 > gpurun_out/n8_bench_city.json 2> gpurun_out/n8_bench_city.err; echo "n8 city rc $?"
for f in n8_bench n8_bench_noddp n8_bench_coco n8_bench_city; do python - <<PY
import json
try:
    r=[json.loads(l) for l in open("gpurun_out/$f.json") if l.startswith("{")][-1]; print("$f", round(r["ms_per_step"],2), round(r["value"],1), "e2e", round(r["e2e"]["value"],1), [round(p["ms_per_step"],1) for p in r["per_rank"]], [p["sm_mhz"] for p in r["per_rank"]], r["clocks"]["reasons"])
except Exception as e: print("$f failed", e)
PY
done
