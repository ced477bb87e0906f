"""Released-checkpoint compatibility: read the reference's Lightning `last.ckpt` files or the stripped
`.pth` state dicts (scripts/strip_checkpoints.py:52-61) and hand them to the fused plans / drop-in modules under the
reference's own key names (`model.conv1.linear.weight`, `model.bn1.running_var`, ...).

A Lightning checkpoint stores the trainer module's state under "state_dict" with the prefixes "model." (the network) and
"ema.module." (the EMA copy, bcos/training/ema.py); the stripped files hold the network's state dict directly.  Both end up
as the same dictionary here; shapes are validated against the architecture before any weight is packed."""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import torch
from torch import Tensor

MODEL_STATE_DICT_KEY = "state_dict"
MODEL_PREFIX = "model."
EMA_PREFIX = "ema.module."

__all__ = ["strip_state_dict", "load_state_dict_file", "check_state_dict", "resnet_plan_from_checkpoint"]


def strip_state_dict(pl_checkpoint: Mapping, ema: bool = False) -> Dict[str, Tensor]:
    """scripts/strip_checkpoints.py:52-61: keep the entries of `checkpoint["state_dict"]` that start with the model (or
    EMA) prefix and drop that prefix."""
    prefix = EMA_PREFIX if ema else MODEL_PREFIX
    state = {k[len(prefix):]: v for k, v in pl_checkpoint[MODEL_STATE_DICT_KEY].items() if k.startswith(prefix)}
    if not state:
        raise KeyError(f"no '{prefix}*' entries in the checkpoint" + (" (was it trained without EMA?)" if ema else ""))
    return state


def load_state_dict_file(path, ema: bool = False) -> Dict[str, Tensor]:
    """`last.ckpt` (Lightning) or stripped `.pth` -> state dict with the network's key names, on the CPU."""
    obj = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(obj, Mapping) and MODEL_STATE_DICT_KEY in obj and isinstance(obj[MODEL_STATE_DICT_KEY], Mapping):
        return strip_state_dict(obj, ema)
    if ema:
        raise ValueError("`ema=True` needs a Lightning checkpoint; a stripped file holds one set of weights only")
    if not isinstance(obj, Mapping) or not all(isinstance(v, Tensor) for v in obj.values()):
        raise ValueError(f"{path}: neither a Lightning checkpoint nor a state dict")
    return dict(obj)


def check_state_dict(state: Mapping[str, Tensor], shapes: Mapping[str, Tuple[int, ...]], ignore_missing=("num_batches_tracked",
                                                                                                       "running_mean")) -> None:
    """Raise with the full list of problems unless `state` provides every tensor the architecture needs, in its shape.
    Extra entries are an error too (a checkpoint of another architecture must not load silently)."""
    missing = [k for k in shapes if k not in state and not k.endswith(tuple(ignore_missing))]
    unexpected = [k for k in state if k not in shapes]
    wrong = [f"{k}: {tuple(state[k].shape)} != {tuple(shapes[k])}" for k in shapes if k in state and tuple(state[k].shape) != tuple(shapes[k])]
    if missing or unexpected or wrong:
        raise ValueError("checkpoint does not match the architecture: "
                         f"missing {missing[:8]}{'...' if len(missing) > 8 else ''}, "
                         f"unexpected {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}, wrong shape {wrong[:8]}")


def resnet_plan_from_checkpoint(arch: str, path, batch: int, ema: bool = False, **plan_kwargs):
    """Fused forward+explain plan (engine.resnet.ResNetPlan) over a released B-cosified ResNet checkpoint.  The weights
    are packed once, here; load another file -> build another plan (the module-level path re-packs by itself when a
    parameter's version counter changes, modules/_runtime.py `_PlanCache`)."""
    from .engine.resnet import ResNetPlan
    from .models import resnet_state_shapes
    state = load_state_dict_file(path, ema)
    check_state_dict(state, resnet_state_shapes(arch))
    return ResNetPlan(arch, state, batch, **plan_kwargs)
