#!/usr/bin/env python
"""Benchmark of the MaskAttn-UNet hot path on B200 (BASELINE.json metric: train images/s @128x128).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is the reference's loop body (ade_semantic.py:394-401): zero_grad, forward, CrossEntropyLoss,
backward, [gradient all-reduce], AdamW.  Workload at every N: UNet(3, 150) ADE20K-shape semantic training,
synthetic 128x128 images, per-GPU batch 256, bf16 activations with fp32 master parameters (BASELINE.json
configs[1]; weak scaling: global batch = 256 N, so N = 8 reaches configs[2]'s global batch of 2048).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` includes the per-step
host->device copy of images/labels from pinned memory and the loss read-back.  `--impl reference` times
the CPU oracle port of the reference step (oracle/unet_oracle.py; the Python reference itself cannot travel
to the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_images_per_s_128x128"
UNIT = "images/s"
WORKLOAD = "MaskAttn-UNet semantic ADE20K-shape train step (UNet(3,150), CE + AdamW), synthetic 128x128"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(", ") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step_rate(steps: int, warmup: int, batch: int = 2):
    """images/s of the reference step restated on CPU (oracle port), all host threads."""
    import torch
    from oracle.unet_oracle import OracleTrainer
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = OracleTrainer(3, 150, lr=5e-5, weight_decay=1e-1, seed=42, dropout_p=0.3)
    images = torch.rand(batch, 3, 128, 128, generator=torch.Generator().manual_seed(0))
    labels = torch.randint(0, 150, (batch, 128, 128), generator=torch.Generator().manual_seed(1))
    for _ in range(warmup):
        tr.step(images, labels)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.step(images, labels)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, cores, batch


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bounded: each CPU step is ~6 s at batch 2 on 8 cores; cap the whole arm near three minutes
    budget_steps = max(1, int(os.environ.get("MASKUNET_REF_MAX_STEPS", "24")))
    if steps + warmup > budget_steps:
        warmup = min(warmup, 1)
        steps = min(steps, budget_steps - warmup)
    rate, sec, cores, batch = cpu_reference_step_rate(steps, warmup, batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": batch, "device": "cpu",
                   "note": "oracle port of the reference step (reference is Python and cannot travel); "
                           "images/s is batch-insensitive on CPU (dense 16384^2 scores, 1 GiB per image per tensor)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} steps of batch {batch} after {warmup} warm-up"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import maskunet_b200
    from maskunet_b200 import ops
    from maskunet_b200.train import Trainer

    B, c_out = args.batch_per_gpu, 150
    torch.manual_seed(42)                                   # identical weights on every rank
    cl = os.environ.get("MASKUNET_CHANNELS_LAST", "1") == "1"   # NHWC activations: the production layout
    model = maskunet_b200.UNet(3, c_out, compute_dtype=torch.bfloat16, channels_last=cl).to(dev)
    if cl:
        model = model.to(memory_format=torch.channels_last)
    trainer = Trainer(model, lr=5e-5, weight_decay=1e-1, data_parallel=world > 1)
    torch.manual_seed(42 + rank)                            # masks / dropout differ per rank
    img_host = torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(rank)).pin_memory()
    lab_host = torch.randint(0, c_out, (B, 128, 128), generator=torch.Generator().manual_seed(1000 + rank)).pin_memory()
    img_dev, lab_dev = img_host.to(dev), lab_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = []

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(args.warmup):
        trainer.step(img_dev, lab_dev)
    # host time to ENQUEUE one step into an empty queue (no sync inside a step): the launch-bound floor of the step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    trainer.step(img_dev, lab_dev)
    host_ms.append((time.perf_counter() - t0) * 1e3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.KERNEL_TIMING["events"].clear()
    ops.KERNEL_TIMING["enabled"] = True
    launches0 = ops.LAUNCHES["count"]
    ms_dev = timed(lambda: trainer.step(img_dev, lab_dev), args.steps)
    launches = ops.LAUNCHES["count"] - launches0
    ops.KERNEL_TIMING["enabled"] = False
    clocks = sampler.stop() if rank == 0 else None

    # end to end: host buffers in, loss out, every step
    losses = []
    ms_e2e = timed(lambda: losses.append(trainer.step(img_host, lab_host).item()), args.steps)

    # roofline of the dominant kernel of ours: attention at the 16384-token site
    roof = None
    peaks = load_peaks()
    ev = ops.KERNEL_TIMING["events"]
    site = model.self_attention6
    n_keep_sum = float(site._compaction[2].sum().item())
    best = None
    for name, mult in (("mu_attn_bwd", 8.0), ("mu_attn_fwd", 4.0)):
        recs = [(a, b) for a, b, meta in ev.get(name, []) if meta == (B, 16384, 64)]
        if not recs:
            continue
        avg_ms = sum(a.elapsed_time(b) for a, b in recs) / len(recs)
        flops = mult * 16384 * n_keep_sum * 64
        cand = dict(kernel=name, ms=avg_ms, achieved=flops / (avg_ms * 1e-3) / 1e12)
        if best is None or cand["ms"] > best["ms"]:
            best = cand
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "attn_b256_dram_traffic.json")   # from tools/ncu_traffic.py (ncu --set full)
    if best is not None and os.path.isfile(tpath):
        try:
            with open(tpath) as fh:
                traffic = json.load(fh).get(best["kernel"], {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    if best is not None:
        peak = peaks["tf_sustained"]
        roof = {"bound": "tensor", "achieved": best["achieved"], "peak": peak, "unit": "TFLOP/s",
                "frac": best["achieved"] / peak, "traffic": traffic, "kernel": best["kernel"],
                "kernel_ms": best["ms"], "peak_source": peaks["source"] + ", sustained (kernel timed inside the step)",
                "flops_counted": "useful: kept keys only (4 N n_keep C fwd, 8 N n_keep C bwd), site N=16384 C=64"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, cores, cb = cpu_reference_step_rate(steps=2, warmup=1, batch=2)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"2 steps of batch {cb} after 1 warm-up (fp32, oracle port of the reference step)"}
    gb = B * world
    h2d = img_host.numel() * img_host.element_size() + lab_host.numel() * lab_host.element_size()
    line = {
        "metric": METRIC, "value": gb / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": gb, "c_out": c_out,
                   "parallelism": f"dp{world}", "optimizer": "AdamW(lr=5e-5, wd=1e-1)",
                   "l2": "per-step working set (tens of GB of activations) exceeds the 126 MB L2; no flush needed",
                   "precision": "bf16 activations, fp32 master parameters / statistics / accumulation",
                   "layout": "channels_last" if cl else "nchw"},
        "clocks": clocks,
        "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e, "last_loss": losses[-1] if losses else None},
        "gpu_launches": launches,
        "host_enqueue_ms_per_step": host_ms[0] if host_ms else None,   # one step into an empty queue; below ms_per_step = GPU-bound
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch-per-gpu", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: relaunch ourselves under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
