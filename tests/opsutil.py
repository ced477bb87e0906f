"""Helpers for the GPU parity tests: move launch records between devices and compare their outputs."""
from __future__ import annotations

import dataclasses
from typing import Dict, Iterable

import torch
from torch import Tensor

from bcos_b200.engine import ops as O

OUTPUT_FIELDS = {
    O.IgemmOp: ["y", "gain", "maskbits", "sq_out", "out2", "amax"],
    O.InputPrepOp: ["out", "sq"],
    O.PatchNormOp: ["inv_norm"],
    O.AvgPoolFwdOp: ["y", "sq"],
    O.AvgPoolBwdMulOp: ["gx"],
    O.GapLogitsOp: ["logits", "pred"],
    O.FcSeedOp: ["out1", "out2"],
    O.ContribMapOp: ["cmap", "grad6"],
    O.ExplanationImageOp: ["out"],
    O.TrunkOutOp: ["out"],
    O.SeedFromNchwOp: ["out1", "out2"],
    O.VitPatchifyOp: ["out", "sq"],
    O.VitContribMapOp: ["cmap", "grad6"],
    O.VitLnFwdOp: ["y", "rstd", "sq"],
    O.VitLnBwdOp: ["G_out", "ghat"],
    O.VitGeluFwdOp: ["a", "sq", "gain"],
    O.VitAttentionOp: ["out"],
    O.PixelSqsumOp: ["sq"],
    O.DenseBnReluFwdOp: ["y", "sq", "maskbits"],
    O.DenseBnReluBwdOp: ["G"],
    O.DenseSliceCastOp: ["out"],
    O.CopyChannelsOp: ["dst"],
    O.StemIm2colOp: ["out", "inv_norm"],
    O.SgemmOp: ["c"],
    O.HeadTokensOp: ["tokens"],
    O.RowSoftmaxOp: ["s"],
    O.SeedFromTokensOp: ["out1", "out2"],
}


def to_device(op, device, memo: Dict[int, Tensor] | None = None):
    """Copy a launch record; tensors are cloned to `device` (aliasing preserved through `memo`)."""
    memo = {} if memo is None else memo
    changes = {}
    for f in dataclasses.fields(op):
        v = getattr(op, f.name)
        if isinstance(v, Tensor):
            if id(v) not in memo:
                base = v._base
                if base is not None:
                    # a view (interior of a zero-bordered buffer, or a reshaped alias of another operand): move the whole
                    # buffer once and keep the view, so that aliasing between launch records survives the copy
                    if id(base) not in memo:
                        memo[id(base)] = base.detach().clone().to(device)
                    memo[id(v)] = memo[id(base)].as_strided(v.shape, v.stride(), v.storage_offset())
                else:
                    memo[id(v)] = v.detach().clone().to(device)
            changes[f.name] = memo[id(v)]
    return dataclasses.replace(op, **changes)


def max_rel_err(a: Tensor, b: Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return ((a - b).abs().max() / denom).item()


def _planes_of(op, name: str) -> int:
    if isinstance(op, O.IgemmOp):
        return {"y": 1 if op.y_f32 else op.y_planes, "out2": op.out2_planes}.get(name, 1)
    if name in ("out", "y", "gx", "out1", "out2") and hasattr(op, "planes"):
        return op.planes
    return 1


def _join(t: Tensor, planes: int) -> Tensor:
    c = t.shape[-1] // planes
    return sum(t[..., p * c:(p + 1) * c].double() for p in range(planes))


def compare(op_ref, op_dev, tol: float, fields: Iterable[str] | None = None) -> Dict[str, float]:
    errs = {}
    for name in (fields or OUTPUT_FIELDS[type(op_ref)]):
        r, d = getattr(op_ref, name), getattr(op_dev, name)
        if r is None:
            continue
        planes = _planes_of(op_ref, name)
        if planes > 1 and r.dtype not in (torch.int32, torch.int64, torch.float32):
            r, d = _join(r.cpu(), planes), _join(d.cpu(), planes)
        if r.dtype in (torch.int32, torch.int64, torch.uint8):
            mism = (r.cpu() != d.cpu()).float().mean().item()
            errs[name] = mism
            # ReLU decisions / kept MaxOut units of values that tie to the last bit may fall the other way
            assert mism <= (2e-3 if name in ("maskbits", "amax") else 0.0), (op_ref.name, name, mism)
        else:
            e = max_rel_err(d, r)
            errs[name] = e
            assert e <= tol, (op_ref.name, name, e)
    return errs
