#!/usr/bin/env python
"""Benchmark of the MaskAttn-UNet hot path on B200 (BASELINE.json metric: train images/s @128x128).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs):
  ade20k_semantic (default, configs[1])  UNet(3, 150) train step, per-GPU batch 256, bf16; weak scaling
                                         (global batch 256 N; N = 8 reaches configs[2]'s global batch of 2048).
  coco_panoptic (configs[2])             UNet(3, 133), loss = 0.9 CE + 0.1 InstanceContrastiveLoss
                                         (coco_panoptic.py:464, :547-554), GLOBAL batch 2048 split over the ranks
                                         (strong scaling), gradient accumulation in micro-batches of 128 (the
                                         reference's instance loss is defined for B <= H = 128).
  city_instance_infer (configs[3])       InstanceUNet(3, 19, 16).eval(), batch 1024, bf16, no_grad: a step is one
                                         forward; ms_per_step is the latency.
  kernel_sweep (configs[4])              attention forward / backward TFLOP/s over the token grids (one JSON line
                                         holding every record + the clocks they were measured at).

A train step is the reference's loop body (ade_semantic.py:394-401): zero_grad, forward, loss, backward, [gradient
all-reduce], AdamW.  Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` includes the
per-step host->device copy of images/labels from pinned memory and the read-back of the result.

`--impl reference` times the reference's own op sequence on the host cores: oracle/unet_oracle.OracleTrainer with
reference_ops=True (nn.Linear / matmul / `/` / `+ mask` / F.softmax / F.layer_norm under autograd, CE, AdamW) --
tests/test_oracle_vs_reference.py holds it to the reference's own classes in value (same losses) and speed (1.00x).
The Python reference itself cannot travel to the GPU box.  It keeps the requested K and W; each step is a bounded
sample: the largest batch <= 8 (configs[0] is batch 8) for which K + W steps fit ~2.5 minutes (images/s is
batch-insensitive on CPU: dense 16384^2 scores, 1 GiB per image per tensor).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tools.clocks import ClockSampler  # noqa: E402

UNIT = "images/s"
WORKLOADS = {
    "ade20k_semantic": dict(
        metric="train_images_per_s_128x128", c_out=150, batch_per_gpu=256, scaling="weak",
        text="MaskAttn-UNet semantic ADE20K-shape train step (UNet(3,150), CE + AdamW), synthetic 128x128"),
    "coco_panoptic": dict(
        metric="train_images_per_s_128x128", c_out=133, global_batch=2048, micro_batch=128, scaling="strong",
        text="MaskAttn-UNet COCO-panoptic train step (UNet(3,133), 0.9 CE + 0.1 InstanceContrastiveLoss, AdamW), "
             "synthetic 128x128, global batch 2048, micro-batches of 128"),
    "city_instance_infer": dict(
        metric="inference_images_per_s_128x128", c_out=19, batch_per_gpu=1024, scaling="weak",
        text="MaskAttn-UNet Cityscapes-instance inference (InstanceUNet(3,19,16).eval(), no_grad), synthetic 128x128"),
    "kernel_sweep": dict(metric="mask_attention_tflops", text="mask-attention kernel sweep"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def synthetic_instances(B: int, n_per_image: int = 6, seed: int = 2):
    """Panoptic-style instance labels: 0 background, rectangles with large segment ids (coco_panoptic.py:76-85)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    h0 = torch.randint(0, 100, (B, n_per_image), generator=g)
    w0 = torch.randint(0, 90, (B, n_per_image), generator=g)
    rows = torch.arange(128).view(1, 128, 1)
    cols = torch.arange(128).view(1, 1, 128)
    im = torch.zeros(B, 128, 128, dtype=torch.int64)
    ids = 1_000_003 * (torch.arange(B) + 1)
    for j in range(n_per_image):                    # later rectangles overwrite earlier ones
        inside = ((rows >= h0[:, j].view(B, 1, 1)) & (rows < h0[:, j].view(B, 1, 1) + 24)
                  & (cols >= w0[:, j].view(B, 1, 1)) & (cols < w0[:, j].view(B, 1, 1) + 36))
        im = torch.where(inside, (ids + j).view(B, 1, 1), im)
    return im


# ------------------------------------------------------------------------------------------- CPU arm
def _mem_available_gb() -> float:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1048576.0
    except Exception:
        pass
    return 0.0


def cpu_reference_rate(workload: str, steps: int, warmup: int, budget_s: float, max_batch: int = 8):
    """images/s of the reference step in the reference's own op sequence on the host cores, all threads.
    Returns dict(rate, sec_per_step, cores, batch, steps, warmup)."""
    import torch
    import torch.nn.functional as F
    from oracle import unet_oracle as uo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    w = WORKLOADS[workload]
    c_out = w["c_out"]
    infer = workload == "city_instance_infer"
    gen = torch.Generator().manual_seed(0)

    def data(b):
        return (torch.rand(b, 3, 128, 128, generator=gen), torch.randint(0, c_out, (b, 128, 128), generator=gen))

    if infer:
        torch.manual_seed(42)
        sd = uo.init_state(3, c_out, "instance")

        def make_step(b):
            keeps = uo.LazyKeeps(b)
            img, _ = data(b)

            def run():
                with torch.no_grad():
                    uo.unet_forward(sd, img, keeps, variant="instance", attention=uo.mask_attention_reference_ops)
            return run
    else:
        def make_step(b):
            tr = uo.OracleTrainer(3, c_out, lr=5e-5, weight_decay=1e-1, seed=42, dropout_p=0.3, reference_ops=True)
            img, lab = data(b)
            if workload != "coco_panoptic":
                return lambda: tr.step(img, lab)
            from oracle.instance_loss_oracle import instance_contrastive_loss
            inst = synthetic_instances(b)

            def run():      # coco_panoptic.py:547-554
                if tr.keeps is None:
                    tr.keeps = uo.LazyKeeps(b)
                tr.opt.zero_grad(set_to_none=True)
                logits = uo.unet_forward(tr.sd, img, tr.keeps, training=True, dropout_p=0.3, update_stats=True,
                                         attention=tr.attention)
                loss = 0.9 * F.cross_entropy(logits, lab) + 0.1 * instance_contrastive_loss(logits, inst)[0]
                loss.backward()
                tr.opt.step()
            return run

    # calibrate on one batch-1 step, then the largest batch (<= 8 = configs[0]) that keeps K + W steps in the budget
    one = make_step(1)
    t0 = time.perf_counter()
    one()
    t1 = time.perf_counter() - t0
    mem = _mem_available_gb()
    batch = max_batch
    while batch > 1 and ((steps + warmup) * batch * t1 > budget_s or mem < 4.0 * batch + 6.0):
        batch //= 2        # ~3.7 GB of RSS per image in the train step (29.6 GB at batch 8)
    fn = make_step(batch)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    return dict(rate=batch * steps / dt, sec=dt / steps, cores=cores, batch=batch, steps=steps, warmup=warmup,
                mem_gb=round(mem, 1))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "kernel_sweep":
        print(json.dumps({"impl": "reference", "unavailable": "the kernel sweep has no CPU arm (see its eager records)"}))
        return
    w = WORKLOADS[args.workload]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_reference_rate(args.workload, steps, warmup, budget_s=float(os.environ.get("MASKUNET_REF_BUDGET_S", "150")))
    sample = (f"{r['steps']} steps of batch {r['batch']} after {r['warmup']} warm-up, fp32, the reference's own op "
              f"sequence (oracle port, speed 1.00x the reference's classes: tests/test_oracle_vs_reference.py); "
              f"host MemAvailable {r['mem_gb']} GB")
    line = {
        "impl": "reference", "metric": w["metric"], "value": r["rate"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["sec"] * 1e3, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["text"], "batch_per_step": r["batch"], "device": "cpu",
                   "note": "images/s is batch-insensitive on CPU (dense 16384^2 scores, 1 GiB per image per tensor); "
                           "the batch is the largest <= 8 (BASELINE configs[0]) that fits the time and memory budget"},
        "cpu_baseline": {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["rate"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- kernel sweep
def run_kernel_sweep(args):
    import torch
    from tools import bench_kernels as bk
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --workload kernel_sweep needs a CUDA device")
    peaks = load_peaks()
    sampler = ClockSampler(0)
    sampler.start()
    recs = []
    for N, C in [(16384, 64), (4096, 64), (4096, 128), (1024, 128), (1024, 256), (256, 256)]:
        recs.append(bk.run(16 if N >= 4096 else 128, N, C, True, peaks["tf_burst"], verbose=False))
    if args.full_sweep:
        for side in (8, 16, 32, 64):
            for C in (64, 128, 256):
                recs.append(bk.run(max(16, 148 * 2 * 128 // (side * side) + 1), side * side, C, True, peaks["tf_burst"], verbose=False))
        for heads in (4, 8):
            for Q in (50, 100, 200):
                recs.append(bk.run_generalised(64, Q, 4096, heads, peaks["tf_burst"], verbose=False))
    clocks = sampler.stop()
    best = max(recs, key=lambda r: r.get("bwd_tflops", 0))
    print(json.dumps({"metric": "mask_attention_tflops", "value": best.get("bwd_tflops"), "unit": "TFLOP/s",
                      "n_gpus": 1, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": "mask-attention kernel sweep (BASELINE configs[4]); value = best backward, "
                                             "useful FLOPs (kept keys)", "peak_tflops_burst": peaks["tf_burst"],
                                 "l2": "256 MB flush buffer written between timed launches"},
                      "clocks": clocks, "records": recs}), flush=True)


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

    if args.deterministic:
        maskunet_b200.set_deterministic(True)
    wl = args.workload
    w = WORKLOADS[wl]
    c_out = w["c_out"]
    infer = wl == "city_instance_infer"
    micro = None
    if wl == "coco_panoptic":
        if w["global_batch"] % world:
            raise SystemExit("coco_panoptic: the global batch of 2048 must divide over the ranks")
        B = args.batch_per_gpu or w["global_batch"] // world
        micro = min(w["micro_batch"], B)
    else:
        B = args.batch_per_gpu or w["batch_per_gpu"]
    torch.manual_seed(42)                                   # identical weights on every rank
    cl = os.environ.get("MASKUNET_CHANNELS_LAST", "1") == "1"   # NHWC activations: the production layout
    cls = maskunet_b200.InstanceUNet if infer else maskunet_b200.UNet
    model = cls(3, c_out, compute_dtype=torch.bfloat16, channels_last=cl).to(dev)
    if cl:
        model = model.to(memory_format=torch.channels_last)
    trainer = None
    if infer:
        model.eval()
    else:
        crit = maskunet_b200.InstanceContrastiveLoss() if wl == "coco_panoptic" else None
        trainer = Trainer(model, lr=5e-5, weight_decay=1e-1, data_parallel=world > 1 and not args.no_ddp,
                          instance_loss=crit, bucket_bytes=int(args.bucket_mb * 1024 * 1024),
                          cuda_graph=args.cuda_graph)
        if trainer.reducer is not None and args.ddp_dryrun:
            trainer.reducer.dry_run = True
    torch.manual_seed(42 + rank)                            # masks / dropout differ per rank
    img_host = torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(rank)).pin_memory()
    lab_host = torch.randint(0, c_out, (B, 128, 128), generator=torch.Generator().manual_seed(1000 + rank)).pin_memory()
    inst_host = synthetic_instances(B, seed=2 + rank).pin_memory() if wl == "coco_panoptic" else None
    img_dev, lab_dev = img_host.to(dev), lab_host.to(dev)
    inst_dev = inst_host.to(dev) if inst_host is not None else None

    result_host = torch.empty((B, 128, 128), dtype=torch.int64).pin_memory() if infer else None

    def step_dev():
        if infer:
            with torch.no_grad():
                return model(img_dev)
        return trainer.step(img_dev, lab_dev, inst_dev, micro_batch=micro)

    # end to end: every step copies ITS batch host -> device from pinned memory and reads its result back.  The copy of
    # batch k + 1 is issued on a copy stream before step k is enqueued (double buffering, what maskunet_b200.data.
    # DevicePrefetcher does for a real loader), so it overlaps step k's kernels; the first copy and every result
    # read-back are exposed.  K steps = K copies + K read-backs inside the timed region.
    copy_stream = torch.cuda.Stream(dev)
    host_batch = [img_host] if infer else [t for t in (img_host, lab_host, inst_host) if t is not None]

    def stage():
        with torch.cuda.stream(copy_stream):
            tensors = [t.to(dev, non_blocking=True) for t in host_batch]
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return tensors, ev

    staged = []

    def step_e2e():
        if not staged:
            staged.append(stage())
        tensors, ev = staged.pop()
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        for t in tensors:
            t.record_stream(cur)
        staged.append(stage())                  # batch k + 1 travels while step k computes
        if infer:        # host images in, class map (city_instance.py:461-462: argmax of the semantic logits) out
            with torch.no_grad():
                sem = model(tensors[0])[0]
                result_host.copy_(maskunet_b200.segmentation_argmax(sem), non_blocking=True)
            cur.synchronize()
            return None
        inst = tensors[2] if len(tensors) > 2 else None
        return trainer.step(tensors[0], tensors[1], inst, micro_batch=micro).item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        mine = e0.elapsed_time(e1) / steps
        ms = torch.tensor([mine], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), mine

    for _ in range(args.warmup):
        step_dev()
    # host time to ENQUEUE one step into an empty queue (no sync inside a step): the launch-bound floor of the step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_dev()
    host_ms = (time.perf_counter() - t0) * 1e3
    sampler = ClockSampler(local)
    sampler.start()
    ops.KERNEL_TIMING["events"].clear()
    ops.KERNEL_TIMING["enabled"] = True
    launches0 = ops.LAUNCHES["count"]
    ms_dev, ms_mine = timed(step_dev, args.steps)
    launches = ops.LAUNCHES["count"] - launches0
    if trainer is not None and trainer._graph is not None:
        launches += trainer.graph_launches * args.steps        # replays launch the captured kernels; Python does not see them
    ops.KERNEL_TIMING["enabled"] = False
    clocks = sampler.stop()

    # end to end: host buffers in, result out, every step
    results, e2e_wall = [], []

    def e2e_once():
        t0 = time.perf_counter()
        results.append(step_e2e())
        e2e_wall.append(round((time.perf_counter() - t0) * 1e3, 2))      # host wall time of the step (it ends in a sync)

    # one untimed end-to-end step first: the staging tensors of this path are new to the caching allocator, and their
    # first allocation (a cudaMalloc beside ~45 GB of cached activations) took up to 150 ms in one run out of four
    step_e2e()
    sampler_e2e = ClockSampler(local)
    sampler_e2e.start()
    ms_e2e, _ = timed(e2e_once, args.steps)
    clocks_e2e = sampler_e2e.stop()
    staged.clear()

    # roofline of the dominant kernel of ours: attention at the 16384-token site
    roof = None
    peaks = load_peaks()
    ev = ops.KERNEL_TIMING["events"]
    site = model.self_attention6
    n_keep_sum = float(site._compaction[2].sum().item())     # kept keys of one launch (one micro-batch)
    best = None
    for name, mult in (("mu_attn_bwd", 8.0), ("mu_attn_fwd", 4.0)):
        recs = [(a, b) for a, b, meta in ev.get(name, []) if meta[1:] == (16384, 64)]
        if not recs:
            continue
        avg_ms = sum(a.elapsed_time(b) for a, b in recs) / len(recs)
        flops = mult * 16384 * n_keep_sum * 64
        cand = dict(kernel=name, ms=avg_ms, achieved=flops / (avg_ms * 1e-3) / 1e12, launches=len(recs))
        if best is None or cand["ms"] > best["ms"]:
            best = cand
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "attn_b256_dram_traffic.json")   # from tools/ncu_traffic.py (ncu --set full)
    if best is not None and wl == "ade20k_semantic" and B == 256 and os.path.isfile(tpath):
        try:
            with open(tpath) as fh:
                traffic = json.load(fh).get(best["kernel"], {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    if best is not None:
        peak = peaks["tf_sustained"]
        roof = {"bound": "tensor", "achieved": best["achieved"], "peak": peak, "unit": "TFLOP/s",
                "frac": best["achieved"] / peak, "traffic": traffic, "kernel": best["kernel"],
                "kernel_ms": best["ms"], "launches_timed": best["launches"],
                "peak_source": peaks["source"] + ", sustained (kernel timed inside the step)",
                "frac_of_burst_peak": best["achieved"] / peaks["tf_burst"],
                "flops_counted": "useful: kept keys only (4 N n_keep C fwd, 8 N n_keep C bwd), site N=16384 C=64"}

    # per-rank view (scaling attribution): every rank's own device time and clocks, gathered on rank 0
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "ms_per_step": ms_mine, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_rate(wl, steps=1, warmup=1, budget_s=75.0)
        cpu = {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"1 step of batch {r['batch']} after 1 warm-up (fp32, the reference's own op sequence: "
                         f"oracle.unet_oracle reference_ops, 1.00x the speed of the reference's classes)"}
    gb = B * world
    h2d = sum(t.numel() * t.element_size() for t in (img_host, lab_host, inst_host) if t is not None)
    if infer:
        h2d = img_host.numel() * img_host.element_size()
    d2h = result_host.numel() * result_host.element_size() if infer else 4
    cfg = {"workload": w["text"], "batch_per_gpu": B, "global_batch": gb, "c_out": c_out,
           "parallelism": f"dp{world}",
           "l2": "per-step working set (tens of GB of activations) exceeds the 126 MB L2; no flush needed",
           "precision": "bf16 activations, fp32 master parameters / statistics / accumulation",
           "layout": "channels_last" if cl else "nchw",
           "e2e_input": "pinned host batch per step; batch k+1 is copied on a side stream while step k runs "
                        "(K steps = K + 1 copies issued, K consumed), result read back every step"}
    if not infer:
        cfg["optimizer"] = "AdamW(lr=5e-5, wd=1e-1)"
    if micro:
        cfg["micro_batch"] = micro
    if args.cuda_graph:
        cfg["cuda_graph"] = trainer is not None and trainer._graph is not None
    if args.deterministic:
        cfg["deterministic"] = "fixed-order reductions (maskunet_b200.set_deterministic): bit-reproducible step"
    if args.no_ddp and world > 1:
        cfg["note"] = "--no-ddp: independent replicas, no gradient exchange (scaling attribution run)"
    if world > 1 and not args.no_ddp:
        cfg["grad_buckets"] = len(trainer.reducer.buckets) if trainer is not None and trainer.reducer else None
        cfg["bucket_mb"] = args.bucket_mb
        if args.ddp_dryrun:
            cfg["note"] = "--ddp-dryrun: hooks + bucket packing without the collective (scaling attribution run)"
    line = {
        "metric": w["metric"], "value": gb / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": w["scaling"],
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": cfg,
        "clocks": clocks,
        "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e, "ms_steps_host_wall": e2e_wall,
                "clocks": clocks_e2e,
                "last_loss": results[-1] if results and results[-1] is not None else None},
        "gpu_launches": launches,
        "host_enqueue_ms_per_step": host_ms,   # one step into an empty queue; below ms_per_step = GPU-bound
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    if infer:
        line["latency_ms"] = ms_dev
    if per_rank is not None:
        line["per_rank"] = per_rank
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="ade20k_semantic")
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="override the workload's per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ddp", action="store_true", help="N independent replicas without the gradient exchange")
    ap.add_argument("--deterministic", action="store_true", help="fixed-order reductions: bit-reproducible step")
    ap.add_argument("--cuda-graph", action="store_true", help="capture the train step in a CUDA graph after warm-up")
    ap.add_argument("--ddp-dryrun", action="store_true", help="hooks and bucket packing, but no collective")
    ap.add_argument("--bucket-mb", type=float, default=25.0, help="gradient bucket size of the data-parallel exchange")
    ap.add_argument("--full-sweep", action="store_true", help="kernel_sweep: add the token-grid and generalised sweeps")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: relaunch ourselves under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "kernel_sweep":
        run_kernel_sweep(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
