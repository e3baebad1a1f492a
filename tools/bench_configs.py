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
    if "city_instance" in which:
        city_instance()
