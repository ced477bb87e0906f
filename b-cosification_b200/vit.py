"""B-cosified SimpleViT on the CUDA-backed modules -- mirror of reference bcos/models/vit.py:64-339 as converted by
bcosify_vit.py:45-153 (patch-embedding weights doubled for the 6-channel input, Linear -> BcosifyLinear except `to_qkv`,
GELU -> MyGELU, LayerNorm -> DetachableLayerNorm, all biases removed, `gap_reorder`: head per token, then mean).
State-dict keys equal the reference's (`model.transformer.encoder_3.attn.to_out.linear.weight`, ...)."""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .bcosify import IMAGENET_MEAN_ADDINVERSE, IMAGENET_STD_ADDINVERSE, Normalize6
from .explain import BcosUtilMixin
from .modules import BcosifyLinear, LogitLayer
from .modules.common import DetachableModule
from .modules.tokens import DetachableLayerNorm, MyGELU, PlainLinear, frozen_attention

VIT_ARCH = {  # name: (dim, depth, heads, mlp_dim)   reference vit.py:441-467
    "simple_vit_ti_patch16_224": (192, 12, 3, 768),
    "simple_vit_s_patch16_224": (384, 12, 6, 1536),
    "simple_vit_b_patch16_224": (768, 12, 12, 3072),
}


def _ln(dim):
    m = DetachableLayerNorm(dim)
    m.bias = None            # the factories strip every bias (vit_bcosification/model.py:20-25)
    return m


def posemb_sincos_2d(h: int, w: int, dim: int, temperature: float = 10000.0) -> Tensor:
    """reference vit.py:64-86"""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden, b):
        super().__init__()
        self.net = nn.Sequential(OrderedDict(norm=_ln(dim), linear1=BcosifyLinear(dim, hidden, b=b), act=MyGELU(),
                                             linear2=BcosifyLinear(hidden, dim, b=b)))

    def forward(self, x):
        return self.net(x)


class Attention(DetachableModule):
    def __init__(self, dim, heads, dim_head, b):
        super().__init__()
        self.heads, self.scale = heads, dim_head ** -0.5
        self.norm = _ln(dim)
        self.to_qkv = PlainLinear(dim, dim_head * heads * 3, bias=False)
        self.to_out = BcosifyLinear(dim_head * heads, dim, bias=False, b=b)

    def forward(self, x):
        qkv = self.to_qkv(self.norm(x))
        return self.to_out(frozen_attention(qkv, self.heads, self.scale, self.detach))


class Encoder(nn.Module):
    def __init__(self, dim, heads, dim_head, mlp_dim, b):
        super().__init__()
        self.attn = Attention(dim, heads, dim_head, b)
        self.ff = FeedForward(dim, mlp_dim, b)

    def forward(self, x):
        x = self.attn(x) + x
        return self.ff(x) + x


class SimpleViT(nn.Module):
    def __init__(self, image_size=224, patch_size=16, num_classes=1000, dim=192, depth=12, heads=3, mlp_dim=768, b=2,
                 gap_reorder=True):
        super().__init__()
        self.patch, self.dim, self.gap_reorder = patch_size, dim, gap_reorder
        self.to_patch_embedding = nn.Sequential(OrderedDict(linear=BcosifyLinear(patch_size * patch_size * 6, dim, b=b)))
        self.transformer = nn.Sequential(OrderedDict(
            (f"encoder_{i}", Encoder(dim, heads, dim // heads, mlp_dim, b)) for i in range(depth)))
        self.linear_head = nn.Sequential(OrderedDict(norm=_ln(dim), linear=BcosifyLinear(dim, num_classes, b=b)))
        g = image_size // patch_size
        self.register_buffer("pos_emb", posemb_sincos_2d(g, g, dim), persistent=False)

    def forward(self, img):
        B, C, H, W = img.shape
        p = self.patch
        # Rearrange "b c (h p1) (w p2) -> b h w (p1 p2 c)"  (reference vit.py:290-294)
        x = img.view(B, C, H // p, p, W // p, p).permute(0, 2, 4, 3, 5, 1).reshape(B, (H // p) * (W // p), p * p * C)
        x = self.to_patch_embedding(x) + self.pos_emb
        x = self.transformer(x)
        if self.gap_reorder:
            return self.linear_head(x).mean(dim=1)
        return self.linear_head(x.mean(dim=1))


class BcosifiedViT(BcosUtilMixin, nn.Module):
    """`bcosify_vit.BcosifyNetwork.forward` (bcosify_vit.py:79-82): normalise the 6-channel input, model, LogitLayer."""

    def __init__(self, model: SimpleViT, logit_bias: Optional[float] = -math.log(1000 - 1), logit_temperature=None):
        super().__init__()
        self.model = model
        self.logit_layer = LogitLayer(logit_temperature=logit_temperature, logit_bias=logit_bias)
        self.bcosifynormalize = Normalize6(IMAGENET_MEAN_ADDINVERSE, IMAGENET_STD_ADDINVERSE)

    def forward(self, x):
        return self.logit_layer(self.model(self.bcosifynormalize(x)))


def add_channels_patch_linear(w3: Tensor) -> Tensor:
    """bcosify_vit.py:84-125: [out, p*p*3] -> [out, p*p*6] with per-pixel [W/2, -W/2]."""
    out_f = w3.shape[0]
    wr = w3.view(out_f, -1, 3) / 2
    return torch.cat([wr, -wr], dim=2).reshape(out_f, -1)


def bcosified_simple_vit(arch: str = "simple_vit_ti_patch16_224", b: float = 2) -> BcosifiedViT:
    dim, depth, heads, mlp = VIT_ARCH[arch]
    return BcosifiedViT(SimpleViT(dim=dim, depth=depth, heads=heads, mlp_dim=mlp, b=b))
