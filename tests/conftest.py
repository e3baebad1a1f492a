import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

ATTN_CASES = ["attn_b2_c64_8x8", "attn_b2_c128_16x16", "attn_b1_c256_16x16", "attn_b3_c64_16x8", "attn_b2_c64_20x20"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_attn_golden(name):
    import torch
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: torch.from_numpy(z[k]) for k in z.files}
    g["keep"] = g["keep"].bool()
    g["params"] = {k[len("param."):]: v for k, v in g.items() if k.startswith("param.")}
    g["grads"] = {k[len("grad."):]: v for k, v in g.items() if k.startswith("grad.")}
    return g


def rel_err(a, b):
    """norm-wise relative error ||a - b|| / ||b|| in float64."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
