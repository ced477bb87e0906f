"""`BcosAttentionPool2d` -- mirror of reference bcos/modules/bcosattnpool.py:10-77 (CLIP attention pooling with the query
and keys detached in explanation mode, no positional embedding in pooled mode) on the CUDA kernels: one packed q|k|v
projection through the tcgen05 linear kernel, the frozen-probability attention kernel, and the output projection."""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import _runtime as R
from .common import DetachableModule
from .tokens import frozen_attention, l2_normalize_rows

__all__ = ["BcosAttentionPool2d"]


class BcosAttentionPool2d(DetachableModule):
    def __init__(self, spacial_dim: int, embed_dim: int, num_heads: int, output_dim: int = None, attn_unpool: bool = False):
        super().__init__()
        self.positional_embedding = nn.Parameter(torch.randn(spacial_dim ** 2 + 1, embed_dim) / embed_dim ** 0.5)
        if not attn_unpool:
            self.k_proj = nn.Linear(embed_dim, embed_dim)
            self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.c_proj = nn.Linear(embed_dim, output_dim or embed_dim)
        self.num_heads = num_heads
        self.attn_unpool = attn_unpool
        self._qkv_cache = R._PlanCache()
        self._out_cache = R._PlanCache()

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, R._PlanCache() if k in ("_qkv_cache", "_out_cache") else copy.deepcopy(v, memo))
        return new

    @staticmethod
    def _plain(x: Tensor, cache, weight: Tensor, key_tensor: Tensor, extra: tuple = ()) -> Tensor:
        lead = x.shape[:-1]
        y = R.bcos_map(x.reshape(-1, x.shape[-1], 1, 1), cache, key_tensor, None, lambda: weight[:, :, None, None], 1, 0, 1.0, True,
                       extra=extra)
        return y.reshape(*lead, weight.shape[0])

    def forward(self, x):
        if self.attn_unpool:
            return self._forward_unpool(x)
        # like the reference's pooled path (bcosattnpool.py:40-58: in_proj_bias=None, out_proj_bias=None) the projections' biases,
        # if any, are NOT used here
        n, c = x.shape[0], x.shape[1]
        t = x.flatten(start_dim=2).permute(0, 2, 1)                        # N (HW) C
        t = torch.cat([t.mean(dim=1, keepdim=True), t], dim=1)             # N (HW+1) C ; no positional embedding (:33-34)
        wqkv = torch.cat([self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], 0)
        qkv = self._plain(t, self._qkv_cache, wqkv, self.v_proj.weight, extra=(self.q_proj.weight, self.k_proj.weight))
        dh = c // self.num_heads
        # every token attends, only the query of the mean token (index 0) is used; q, k frozen in explanation mode
        o = frozen_attention(qkv, self.num_heads, dh ** -0.5, self.detach)[:, 0]      # backward outside explanation mode raises
        cw = self.c_proj.weight if isinstance(self.c_proj, nn.Linear) else self.c_proj.linear.weight
        return self._plain(o, self._out_cache, cw, cw)                     # c_proj acts as a plain linear (:56)

    def _forward_unpool(self, x):
        """bcosattnpool.py:23-33: every spatial token goes through v_proj (plain) and c_proj (a B-cos linear once
        bcosified, bcosify.py:81-97) and is L2-normalised with a detachable norm; output (HW) x N x D'."""
        for lin in (self.v_proj, self.c_proj):
            if getattr(lin, "bias", None) is not None:
                raise NotImplementedError("bcos_b200: attention pooling is built bias-free (all registered configs strip biases)")
        t = x.flatten(start_dim=2).permute(2, 0, 1).contiguous()           # NCHW -> (HW) N C
        t = self._plain(t, self._qkv_cache, self.v_proj.weight, self.v_proj.weight)
        if isinstance(self.c_proj, nn.Linear):
            t = self._plain(t, self._out_cache, self.c_proj.weight, self.c_proj.weight)
        else:
            t = self.c_proj(t)
        return l2_normalize_rows(t, self.detach)

    @classmethod
    def from_standard_module(cls, model, module, model_config):
        new_module = cls(model.input_resolution // 32, model.conv1.out_channels * 64, module.num_heads, model.output_dim,
                         model_config.get("attn_unpool", False))
        if model_config.get("weights", None) is not None:
            for name, param in module.named_parameters():
                if new_module.attn_unpool and ("k_proj" not in name) and ("q_proj" not in name):
                    obj = new_module
                    parts = name.split(".")
                    for p_ in parts[:-1]:
                        obj = getattr(obj, p_)
                    getattr(obj, parts[-1]).data = param.data
        return new_module
