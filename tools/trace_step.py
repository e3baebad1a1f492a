#!/usr/bin/env python
"""Timeline of the data-parallel train step (nsys is not in this image: torch.profiler / CUPTI instead).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/trace_step.py

Rank 0 profiles 2 steps after warm-up and writes gpurun_out/trace_step_n{N}.json (per-stream kernel summary: NCCL
kernels, our kernels, gaps) and the chrome trace (gzipped) beside it."""
import gzip
import json
import os
import shutil
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import maskunet_b200
    from maskunet_b200.train import Trainer
    B = int(os.environ.get("TRACE_BATCH", "256"))
    torch.manual_seed(42)
    model = maskunet_b200.UNet(3, 150, compute_dtype=torch.bfloat16, channels_last=True).to(dev).to(memory_format=torch.channels_last)
    tr = Trainer(model, data_parallel=world > 1)
    torch.manual_seed(42 + rank)
    x = torch.rand(B, 3, 128, 128, device=dev)
    y = torch.randint(0, 150, (B, 128, 128), device=dev)
    for _ in range(3):
        tr.step(x, y)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    acts = [ProfilerActivity.CUDA, ProfilerActivity.CPU]
    if rank == 0:
        with profile(activities=acts) as prof:
            for _ in range(2):
                tr.step(x, y)
            torch.cuda.synchronize()
    else:
        for _ in range(2):
            tr.step(x, y)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        path = f"gpurun_out/trace_step_n{world}.chrome.json"
        prof.export_chrome_trace(path)
        ev = json.load(open(path))["traceEvents"]
        kern = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
        kern.sort(key=lambda e: e["ts"])
        streams = {}
        for e in kern:
            s = streams.setdefault(e["args"].get("stream", e.get("tid")), {"n": 0, "busy_us": 0.0, "nccl_us": 0.0,
                                                                         "first": e["ts"], "last": 0.0})
            s["n"] += 1
            s["busy_us"] += e["dur"]
            if "nccl" in e["name"].lower():
                s["nccl_us"] += e["dur"]
            s["last"] = max(s["last"], e["ts"] + e["dur"])
        for s in streams.values():
            s["span_us"] = s["last"] - s["first"]
        nccl = [{"name": e["name"][:60], "ts_us": e["ts"] - kern[0]["ts"], "dur_us": e["dur"],
                 "stream": e["args"].get("stream")} for e in kern if "nccl" in e["name"].lower()]
        # what runs on the main stream while an NCCL kernel is in flight, and how long it takes there
        by_name = {}
        for e in kern:
            if "nccl" in e["name"].lower():
                continue
            overl = any(n["ts_us"] + kern[0]["ts"] < e["ts"] + e["dur"] and e["ts"] < n["ts_us"] + kern[0]["ts"] + n["dur_us"]
                        for n in nccl)
            d = by_name.setdefault(e["name"][:50], {"alone_us": [], "with_nccl_us": []})
            d["with_nccl_us" if overl else "alone_us"].append(e["dur"])
        top = sorted(by_name.items(), key=lambda kv: -(sum(kv[1]["alone_us"]) + sum(kv[1]["with_nccl_us"])))[:14]
        summ = {"world": world, "batch_per_gpu": B, "steps_profiled": 2,
                "streams": {str(k): v for k, v in streams.items()}, "nccl_kernels": nccl,
                "top_kernels": {k: {"alone_n": len(v["alone_us"]), "alone_total_us": sum(v["alone_us"]),
                                    "with_nccl_n": len(v["with_nccl_us"]), "with_nccl_total_us": sum(v["with_nccl_us"])}
                                for k, v in top}}
        json.dump(summ, open(f"gpurun_out/trace_step_n{world}.json", "w"), indent=1)
        with open(path, "rb") as fi, gzip.open(path + ".gz", "wb") as fo:
            shutil.copyfileobj(fi, fo)
        os.unlink(path)
        print(json.dumps({k: v for k, v in summ.items() if k != "top_kernels"})[:3000])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
