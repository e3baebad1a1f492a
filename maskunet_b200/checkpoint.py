"""Reference checkpoint compatibility.

The reference saves ``model.state_dict()`` with ``torch.save`` (ade_semantic.py:341-344, :412, :426) -- under
``nn.DataParallel`` every key carries a ``module.`` prefix that its loaders strip (ade_panoptic.py:432-435) -- and
transfers weights between tasks by dropping ``final_layer.*`` and loading with ``strict=False``
(cityscapes/city_semantic.py:335-338).  Our modules keep the reference's state_dict keys (including the dead
``emb_layer.*``), so a reference checkpoint loads unchanged; these helpers are the three idioms in one place.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch


def strip_data_parallel_prefix(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """ade_panoptic.py:432-435: ``{k.replace('module.', ''): v}``."""
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}


def load_reference_checkpoint(model: torch.nn.Module, path_or_state, drop_prefixes: Iterable[str] = (),
                              strict: bool = True, map_location="cpu") -> Tuple[list, list]:
    """Load a checkpoint written by any of the reference scripts into ``model`` (ours or the reference's).

    ``drop_prefixes=("final_layer.",), strict=False`` reproduces the cross-task transfer of city_semantic.py:335-338.
    Returns (missing_keys, unexpected_keys) like ``load_state_dict``.
    """
    state = path_or_state if isinstance(path_or_state, dict) else torch.load(path_or_state, map_location=map_location)
    state = strip_data_parallel_prefix(state)
    drop = tuple(drop_prefixes)
    if drop:
        state = {k: v for k, v in state.items() if not k.startswith(drop)}
    result = model.load_state_dict(state, strict=strict)
    return list(result.missing_keys), list(result.unexpected_keys)


def export_reference_checkpoint(model: torch.nn.Module, path: str, data_parallel_prefix: bool = False) -> None:
    """Write ``model.state_dict()`` the way the reference does (:341-344), optionally with the ``module.`` prefix a
    DataParallel-wrapped reference model would have produced, so that the reference scripts can load it."""
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    if data_parallel_prefix:
        state = {"module." + k: v for k, v in state.items()}
    torch.save(state, path)
