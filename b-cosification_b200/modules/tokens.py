"""Token-model modules of the B-cosified ViT on the CUDA kernels: DetachableLayerNorm (reference
bcos/modules/norms/centered_norms.py:187-245), MyGELU (bcosify_vit.py:27-32), plain linear (`to_qkv`, which both
converters leave as nn.Linear: vit.py:140, bcosify_vit.py:138) and the frozen-probability attention core
(bcos/models/vit.py:143-158)."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib as L
from . import _runtime as R
from .common import DetachableModule

__all__ = ["DetachableLayerNorm", "MyGELU", "PlainLinear", "frozen_attention", "l2_normalize_rows"]

_NOT_BUILT = ("bcos_b200: only the explanation-mode backward (detached dynamic weights) is built; "
              "the full training backward is outside this round's scope")


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, detach):
        d = x.shape[-1]
        rows = x.numel() // d
        y = torch.empty_like(x)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        L.layernorm_fwd(x, rows, d, weight, bias, eps, y, rstd)
        ctx.detach, ctx.weight = detach, weight
        ctx.save_for_backward(rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        if not ctx.detach:
            raise NotImplementedError(_NOT_BUILT)
        (rstd,) = ctx.saved_tensors
        gy = gy.contiguous()
        d = gy.shape[-1]
        gx = torch.empty_like(gy)
        L.layernorm_explain_bwd(gy, gy.numel() // d, d, ctx.weight, rstd, gx)
        return gx, None, None, None, None


class DetachableLayerNorm(nn.LayerNorm, DetachableModule):
    """LayerNorm whose variance is detached in explanation mode (the mean stays in the graph)."""

    def __init__(self, *args, **kwargs):
        DetachableModule.__init__(self)
        super().__init__(*args, **kwargs)

    def forward(self, input: Tensor) -> Tensor:
        assert len(self.normalized_shape) == 1, "bcos_b200: LayerNorm over the last dimension only"
        R._require_cuda(input, "DetachableLayerNorm")
        x = input.float().contiguous()
        w = None if self.weight is None else self.weight.detach().float().contiguous()
        b = None if self.bias is None else self.bias.detach().float().contiguous()
        return _LayerNormFn.apply(x, w, b, float(self.eps), self.detach).type(input.dtype)

    @classmethod
    def from_standard_module(cls, standard_module: nn.LayerNorm, model_config: dict):
        new_mod = cls(normalized_shape=standard_module.normalized_shape, eps=standard_module.eps,
                      elementwise_affine=standard_module.elementwise_affine)
        if model_config.get("weights", None) is not None:
            new_mod.weight.data = standard_module.weight.data
            if standard_module.bias is not None:
                new_mod.bias.data = standard_module.bias.data
        return new_mod


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, detach):
        d = x.shape[-1]
        rows = x.numel() // d
        y = torch.empty_like(x)
        inv = torch.empty(rows, dtype=torch.float32, device=x.device)
        L.l2norm_rows(x, rows, d, y, inv)
        ctx.detach = detach
        ctx.save_for_backward(inv)
        return y

    @staticmethod
    def backward(ctx, gy):
        if not ctx.detach:
            raise NotImplementedError(_NOT_BUILT)
        (inv,) = ctx.saved_tensors
        gy = gy.contiguous()
        d = gy.shape[-1]
        gx = torch.empty_like(gy)
        L.row_scale(gy, gy.numel() // d, d, inv, gx)
        return gx, None


def l2_normalize_rows(x: Tensor, detach: bool) -> Tensor:
    """x / x.norm(dim=-1, keepdim=True) with the norm detached in explanation mode (bcosattnpool.py:29-32)."""
    R._require_cuda(x, "l2_normalize_rows")
    return _L2NormFn.apply(x.float().contiguous(), detach).type(x.dtype)


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, detach):
        y = torch.empty_like(x)
        L.gelu_gate(x, None, x.numel(), y)
        ctx.detach = detach
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        if not ctx.detach:
            raise NotImplementedError(_NOT_BUILT)
        (x,) = ctx.saved_tensors
        gx = torch.empty_like(x)
        L.gelu_gate(x, gy.contiguous(), x.numel(), gx)
        return gx, None


class MyGELU(DetachableModule):
    """x * Phi(x) with the gate Phi(x) detached in explanation mode."""

    def forward(self, x):
        R._require_cuda(x, "MyGELU")
        return _GeluFn.apply(x.float().contiguous(), self.detach).type(x.dtype)


class PlainLinear(nn.Linear, DetachableModule):
    """nn.Linear (no B-cos transform) evaluated by the same tcgen05 kernel with the scale switched off; its
    explanation backward is the ordinary data gradient."""

    def __init__(self, in_features, out_features, bias=False, device=None, dtype=None):
        DetachableModule.__init__(self)
        super().__init__(in_features, out_features, bias=bias, device=device, dtype=dtype)
        self._cache = R._PlanCache()

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, R._PlanCache() if k == "_cache" else copy.deepcopy(v, memo))
        return new

    def forward(self, input: Tensor) -> Tensor:
        lead = input.shape[:-1]
        x = input.reshape(-1, self.in_features, 1, 1)
        y = R.bcos_map(x, self._cache, self.weight, self.bias, lambda: self.weight[:, :, None, None], 1, 0, 1.0, True)
        return y.reshape(*lead, self.out_features)


class _AttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads, scale, detach):
        b, n, three_hd = qkv.shape
        dh = three_hd // (3 * heads)
        out = torch.empty(b, n, heads * dh, dtype=torch.float32, device=qkv.device)
        L.attention(qkv, None, b, n, heads, dh, scale, False, out)
        ctx.heads, ctx.scale, ctx.detach = heads, scale, detach
        ctx.save_for_backward(qkv)
        return out

    @staticmethod
    def backward(ctx, g):
        if not ctx.detach:
            raise NotImplementedError(_NOT_BUILT)
        (qkv,) = ctx.saved_tensors
        b, n, three_hd = qkv.shape
        gqkv = torch.zeros_like(qkv)          # q, k are detached (vit.py:148-150): only the v block receives gradient
        L.attention(qkv, g.contiguous(), b, n, ctx.heads, three_hd // (3 * ctx.heads), ctx.scale, True, gqkv)
        return gqkv, None, None, None


def frozen_attention(qkv: Tensor, heads: int, scale: float, detach: bool) -> Tensor:
    """softmax(q k^T * scale) v on a packed qkv tensor [B, N, 3*heads*64] -> [B, N, heads*64]."""
    R._require_cuda(qkv, "attention")
    return _AttnFn.apply(qkv.float().contiguous(), heads, float(scale), detach)
