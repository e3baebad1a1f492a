#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err; echo "n2 rc $?"
timeout 600 $TR --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-ddp > gpurun_out/b_bench_n2_noddp.json 2> gpurun_out/b_bench_n2_noddp.err; echo "n2 noddp rc $?"
timeout 600 $TR --master-port 29545 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/b_bench_n2_again.json 2> gpurun_out/b_bench_n2_again.err; echo "n2 again rc $?"
timeout 600 $TR --master-port 29543 tools/trace_step.py > gpurun_out/b_trace.log 2>&1; echo "trace rc $?"
timeout 600 $TR --master-port 29544 bench.py --gpus 2 --workload coco_panoptic --steps 2 --warmup 1 > gpurun_out/b_bench_coco_n2.json 2> gpurun_out/b_bench_coco_n2.err; echo "coco n2 rc $?"
timeout 600 python bench.py --workload coco_panoptic --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_bench_coco_n1.json 2> gpurun_out/b_bench_coco_n1.err; echo "coco n1 rc $?"
timeout 300 python -m pytest tests/test_gpu_input.py -m gpu -x -q > gpurun_out/b_input.log 2>&1; tail -3 gpurun_out/b_input.log
for f in b_bench_n2 b_bench_n2_noddp b_bench_n2_again b_bench_coco_n2 b_bench_coco_n1; do python - <<PY
import json
try:
    r=[json.loads(l) for l in open("gpurun_out/$f.json") if l.startswith("{")][-1]; print("$f", r["ms_per_step"], r["value"], r.get("per_rank"), r["roofline"]["kernel_ms"], r["clocks"], r["host_enqueue_ms_per_step"])
except Exception as e: print("$f failed", e)
PY
done
tail -3 gpurun_out/b_trace.log | cut -c1-2500
