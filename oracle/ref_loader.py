"""Load the reference's OWN model classes without running its scripts.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference scripts run dataset construction, plotting and a CUDA training
loop at import time, so they cannot be imported.  Their model classes depend
only on torch, so we ``ast.parse`` the file and ``exec`` just the ClassDef
nodes we need.  Nothing is copied into this repo; the reference tree is read
at call time and is only present in the build container (never on the GPU
box), which is why the goldens under tests/golden/ exist.
"""
from __future__ import annotations

import ast
import os
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("MASKUNET_REFERENCE_ROOT", "/root/reference")

MODEL_CLASSES = ("Mask2FormerAttention", "ConvBlock", "DownSample", "UpSample", "UNet")

# canonical copies, SURVEY.md section 8(c)
SCRIPTS = {
    "ade_semantic": "code/ade20k/ade_semantic.py",      # :152-314
    "coco_panoptic": "code/coco/coco_panoptic.py",      # :173-335
    "city_instance": "code/cityscapes/city_instance.py",  # :127-276 (3-output variant)
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, SCRIPTS["ade_semantic"]))


def load_reference_classes(script: str = "ade_semantic", names=MODEL_CLASSES) -> SimpleNamespace:
    """Return a namespace holding the reference's classes ``names`` from ``script``."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    path = os.path.join(REFERENCE_ROOT, SCRIPTS.get(script, script))
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    wanted = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in names]
    found = {n.name for n in wanted}
    missing = set(names) - found
    if missing:
        raise RuntimeError(f"{path}: classes not found: {sorted(missing)}")
    ns = {"torch": torch, "nn": nn, "F": F, "__name__": f"reference_{script}"}
    module = ast.Module(body=wanted, type_ignores=[])
    exec(compile(module, path, "exec"), ns)
    return SimpleNamespace(**{k: ns[k] for k in names})


def load_reference_functions(script: str = "ade_semantic", names=("mean_iou",)) -> SimpleNamespace:
    """The reference's own module-level FUNCTIONS ``names`` (e.g. mean_iou, ade_semantic.py:128-146), same mechanism."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    path = os.path.join(REFERENCE_ROOT, SCRIPTS.get(script, script))
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in wanted}
    if missing:
        raise RuntimeError(f"{path}: functions not found: {sorted(missing)}")
    ns = {"torch": torch, "nn": nn, "F": F, "__name__": f"reference_{script}"}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), ns)
    return SimpleNamespace(**{k: ns[k] for k in names})
