#!/bin/bash
# round 2, pass D (1 GPU): full GPU suite (new: CUDA graph, masked-row clearing, resize), default bench, host-bound small
# batch eager vs CUDA graph, descriptor-cache counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/d_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/d_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench_default.json 2> gpurun_out/d_bench_default.err; echo "bench rc $?"
for B in 16 64; do
timeout 600 python bench.py --steps 20 --warmup 4 --batch-per-gpu $B --no-cpu-baseline > gpurun_out/d_bench_b${B}_eager.json 2> gpurun_out/d_bench_b${B}_eager.err; echo "b$B eager rc $?"
timeout 600 python bench.py --steps 20 --warmup 4 --batch-per-gpu $B --no-cpu-baseline --cuda-graph > gpurun_out/d_bench_b${B}_graph.json 2> gpurun_out/d_bench_b${B}_graph.err; echo "b$B graph rc $?"
done
timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --cuda-graph > gpurun_out/d_bench_b256_graph.json 2> gpurun_out/d_bench_b256_graph.err; echo "b256 graph rc $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/d_bench_*.json")):
    try:
        r=[json.loads(l) for l in open(f) if l.startswith("{")][-1]
        print(f, "ms", round(r["ms_per_step"],2), "img/s", round(r["value"],1), "e2e", round(r["e2e"]["value"],1), "host_ms", round(r["host_enqueue_ms_per_step"],2), "launches", r["gpu_launches"], r["config"].get("cuda_graph"), r["clocks"]["sm_mhz"], (r.get("roofline") or {}).get("frac"))
    except Exception as e: print(f, "failed", e)
PY
python - <<'PY'
import torch, maskunet_b200
from maskunet_b200 import _lib
from maskunet_b200.train import Trainer
dev=torch.device("cuda",0)
m=maskunet_b200.UNet(3,19,compute_dtype=torch.bfloat16,channels_last=True).to(dev).to(memory_format=torch.channels_last)
tr=Trainer(m); x=torch.rand(4,3,128,128,device=dev); y=torch.randint(0,19,(4,128,128),device=dev)
for i in range(6):
    tr.step(x,y); torch.cuda.synchronize(); print("step",i,_lib.tmap_cache_stats())
PY
