"""B-cos CLIP ViT image encoder on the CUDA-backed modules -- mirror of CLIP/clip/model.py:157-241 (`LayerNorm`, `QuickGELU`,
`ResidualAttentionBlock`, `Transformer`, `VisionTransformer`) as converted by the reference's bcosify.py with `clip_kd`.

What bcosify.py:74-113 does to this encoder (and what therefore runs on libbcosk.so here): the patch-embedding `conv1` becomes
a `BcosifyConv2d` over 6 input channels, `mlp.c_fc` / `mlp.c_proj` become `BcosifyLinear`, `attn.out_proj` becomes a
`BcosifyLinear` OBJECT whose weight `nn.MultiheadAttention` keeps using as a plain linear map, every `nn.Sequential` becomes a
`BcosSequential`.  `LayerNorm`, `QuickGELU` and the attention itself are not touched by the reference (stock torch modules,
not detachable - SURVEY.md 8c notes that the reference registers no experiment config for this encoder); the factory strips
all `.bias` attributes and the positional embedding (clip_bcosification/model.py:17-25).  State-dict keys equal CLIP's.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from .bcosify import BcosifyNetwork


class LayerNorm(nn.LayerNorm):
    """CLIP/clip/model.py:157-163: LayerNorm evaluated in fp32."""

    def forward(self, x: torch.Tensor):
        orig_type = x.dtype
        return super().forward(x.type(torch.float32)).type(orig_type)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x: torch.Tensor):
        self.attn_mask = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        return self.attn(x, x, x, need_weights=False, attn_mask=self.attn_mask)[0]

    def forward(self, x: torch.Tensor):
        x = x + self.attention(self.ln_1(x))
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.conv1 = nn.Conv2d(in_channels=3, out_channels=width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor):
        x = self.conv1(x)                                          # [N, width, grid, grid]: B-cos patch embedding
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        x = torch.cat([self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device), x], dim=1)
        if self.positional_embedding is not None:
            x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.permute(1, 0, 2)).permute(1, 0, 2)  # NLD -> LND -> NLD
        x = self.ln_post(x[:, 0, :])
        if self.proj is not None:
            x = x @ self.proj
        return x


def bcosified_clip_vit(input_resolution: int = 224, patch_size: int = 32, width: int = 768, layers: int = 12, heads: int = 12,
                       output_dim: int = 512) -> BcosifyNetwork:
    """Offline equivalent of clip_bcosification/model.py:8-25 applied to CLIP's ViT image encoder (defaults: ViT-B/32)."""
    cfg = dict(is_bcos=True, name="vitclip", bcos_args=dict(b=2, max_out=1),
               bcosify_args=dict(clip_kd=True, fix_b=True, norm_layer="BnUncV2", use_bias=False))
    m = BcosifyNetwork(VisionTransformer(input_resolution, patch_size, width, layers, heads, output_dim), cfg, add_channels=True,
                       logit_layer=False)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
        if hasattr(mod, "positional_embedding") and mod.positional_embedding is not None:
            mod.positional_embedding = None
    return m
