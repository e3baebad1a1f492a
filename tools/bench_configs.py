"""BASELINE.json configs[2] and configs[3] on one GPU (bench.py carries configs[1], the headline):
  coco_panoptic : UNet(3, 133) train step, per-GPU batch 256 (global 2048 at 8 GPUs), bf16, CE + AdamW
  city_instance : InstanceUNet(3, 19, embed_dim=16).eval(), batch 1024, bf16, no_grad: latency + images/s
Prints one JSON line per config."""
import json, statistics, sys, torch
sys.path.insert(0, ".")
import maskunet_b200
from maskunet_b200.train import Trainer

dev = torch.device("cuda", 0)


def ev_time(fn, n):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return ts


def coco_panoptic(B=256):
    torch.manual_seed(42)
    model = maskunet_b200.UNet(3, 133, compute_dtype=torch.bfloat16, channels_last=True).to(dev).to(memory_format=torch.channels_last)
    tr = Trainer(model, lr=5e-5, weight_decay=1e-1)
    x = torch.rand(B, 3, 128, 128, device=dev)
    y = torch.randint(0, 133, (B, 128, 128), device=dev)
    for _ in range(3):
        tr.step(x, y)
    ts = ev_time(lambda: tr.step(x, y), 5)
    ms = statistics.median(ts)
    print(json.dumps({"config": "coco_panoptic train (UNet(3,133), CE + AdamW), bf16, batch/GPU %d" % B, "ms_per_step": round(ms, 2),
                      "images_per_s": round(B / ms * 1e3, 1)}), flush=True)


def synthetic_instances(B, n_per_image=6, seed=2):
    """Panoptic-style instance labels: 0 background, rectangles with large segment ids (coco_panoptic.py:76-85)."""
    g = torch.Generator().manual_seed(seed)
    im = torch.zeros(B, 128, 128, dtype=torch.int64)
    for b in range(B):
        for j in range(n_per_image):
            h0, w0 = int(torch.randint(0, 100, (1,), generator=g)), int(torch.randint(0, 90, (1,), generator=g))
            im[b, h0:h0 + 24, w0:w0 + 36] = 1_000_003 * (b + 1) + j
    return im


def coco_panoptic_instance_term(B=128):
    """The real COCO-panoptic step (coco_panoptic.py:544-553): loss = 0.9 CE + 0.1 InstanceContrastiveLoss.  Batch 128:
    the reference's loss uses the batch index of a pixel as a row index (:502), so it is defined for B <= H = 128."""
    torch.manual_seed(42)
    model = maskunet_b200.UNet(3, 133, compute_dtype=torch.bfloat16, channels_last=True).to(dev).to(memory_format=torch.channels_last)
    x = torch.rand(B, 3, 128, 128, device=dev)
    y = torch.randint(0, 133, (B, 128, 128), device=dev)
    inst = synthetic_instances(B).to(dev)
    rec = {"config": "coco_panoptic train with the instance term (0.9 CE + 0.1 triplet), bf16, batch/GPU %d" % B,
           "instances_per_step": int(inst.unique().numel()) - 1}
    for name, crit in (("ce_only", None), ("with_instance_term", maskunet_b200.InstanceContrastiveLoss())):
        tr = Trainer(model, lr=1e-5, weight_decay=1e-4, instance_loss=crit)
        for _ in range(3):
            loss = tr.step(x, y, inst)
        ms = statistics.median(ev_time(lambda: tr.step(x, y, inst), 5))
        rec[name] = {"ms_per_step": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1), "loss": round(float(loss), 4)}
    logits = model(x).detach()
    crit = maskunet_b200.InstanceContrastiveLoss()
    for _ in range(2):
        crit(logits, inst)
    rec["loss_forward_only_ms"] = round(statistics.median(ev_time(lambda: crit(logits, inst), 5)), 3)
    print(json.dumps(rec), flush=True)


def city_instance(B=1024):
    torch.manual_seed(42)
    model = maskunet_b200.InstanceUNet(3, 19, embed_dim=16, compute_dtype=torch.bfloat16, channels_last=True).to(dev)
    model = model.to(memory_format=torch.channels_last).eval()
    x = torch.rand(B, 3, 128, 128, device=dev)
    with torch.no_grad():
        for _ in range(3):
            out = model(x)
        ts = ev_time(lambda: model(x), 20)
    ms = statistics.median(ts)
    print(json.dumps({"config": "city_instance inference (InstanceUNet(3,19,16).eval()), bf16, batch %d" % B,
                      "latency_ms_median_of_20": round(ms, 2), "images_per_s": round(B / ms * 1e3, 1),
                      "outputs": [list(o.shape) for o in out]}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["coco_panoptic", "city_instance"]
    if "coco_panoptic" in which:
        coco_panoptic()
    torch.cuda.empty_cache()
    if "coco_instance_term" in which:
        coco_panoptic_instance_term()
    torch.cuda.empty_cache()
    if "city_instance" in which:
        city_instance()
