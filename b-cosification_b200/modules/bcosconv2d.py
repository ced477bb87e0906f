"""B-cos 2-D convolutions -- drop-in mirror of reference bcos/modules/bcosconv2d.py and bcosifyconv2d.py.

Same constructors, attributes (`.linear`, `.weight`, `.b`, `.max_out`, `.detach`, ...) and state-dict keys
(`*.linear.weight`); the arithmetic runs on libbcosk.so (tcgen05 implicit GEMM with the B-cos epilogue).
"""
from __future__ import annotations

import warnings
from typing import Tuple, Union

import torch
import torch.linalg as LA
import torch.nn as nn
from torch import Tensor
from torch.nn.modules.utils import _pair

from . import _runtime as R
from .common import DetachableModule

__all__ = ["NormedConv2d", "BcosConv2d", "BcosConv2dWithScale", "BcosifyConv2d"]


class NormedConv2d(nn.Conv2d):
    """nn.Conv2d whose weights are used with unit L2 norm per output unit (bcosconv2d.py:17-41).  Only a parameter
    container here: the normalisation is folded into the packed weights by the owning BcosConv2d."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.scale = None
        self.use_weight_norm = True

    def effective_weight(self) -> Tensor:
        w = self.weight
        if self.use_weight_norm:
            w = w / LA.vector_norm(w, dim=(1, 2, 3), keepdim=True)
            if self.scale is not None:
                w = self.scale * w
        return w

    def forward(self, in_tensor: Tensor) -> Tensor:  # plain linear map (B = 1)
        raise RuntimeError("NormedConv2d is evaluated through its BcosConv2d")

    def set_scale(self, weight: Tensor, trainable=False):
        self.scale = nn.Parameter(weight.norm(p=2, dim=(1, 2, 3), keepdim=True), requires_grad=trainable)

    def toggle_weight_norm(self, use_weight_norm):
        self.use_weight_norm = use_weight_norm


def _single(v) -> int:
    a, b = _pair(v)
    if a != b:
        raise NotImplementedError("bcos_b200: only square kernels / strides / paddings are built")
    return int(a)


def _patch_norms(in_tensor: Tensor, k: int, s: int, p: int) -> Tensor:
    """sqrt(sum-pool_{k,s,p}(sum_c x^2) + 1e-6), [N,1,Ho,Wo]: exact fp32 per-pixel sums of squares, then bcosk_patch_inv_norm."""
    from .. import _lib as L
    x = in_tensor.float().contiguous()
    nb, c, h, w = x.shape
    sq = torch.empty(nb * h * w, dtype=torch.float32, device=x.device)
    L.pixel_sqsum_nchw_f32(x, sq)
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    inv = torch.empty(nb * oh * ow, dtype=torch.float32, device=x.device)
    L.patch_inv_norm(sq, 1, nb, h, w, k, k, s, p, 1e-6, 0.0, inv, oh, ow)
    return (1.0 / inv).view(nb, 1, oh, ow)


class BcosConv2d(DetachableModule):
    """`out = (w_hat . x) * |cos(x, w_hat)|^(B-1)` (bcosconv2d.py:43-262)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[int, Tuple[int, ...]] = 1,
                 stride: Union[int, Tuple[int, ...]] = 1, padding: Union[int, Tuple[int, ...]] = 0,
                 dilation: Union[int, Tuple[int, ...]] = 1, groups: int = 1, padding_mode: str = "zeros", device=None,
                 dtype=None, bias: bool = False, b: Union[int, float] = 2, max_out: int = 1, **kwargs):
        assert max_out > 0, f"max_out should be greater than 0, was {max_out}"
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = _pair(dilation)
        self.groups = groups
        self.padding_mode = padding_mode
        self.device = device
        self.dtype = dtype
        self.bias = None
        self.b = b
        self.max_out = max_out
        if any(d > 1 for d in self.dilation):
            warnings.warn("dilation > 1 is not built in bcos_b200")
        self.linear = self._make_linear()
        self._cache = R._PlanCache()

    def _make_linear(self) -> nn.Module:
        return NormedConv2d(in_channels=self.in_channels, out_channels=self.out_channels * self.max_out,
                            kernel_size=self.kernel_size, stride=self.stride, padding=self.padding, dilation=self.dilation,
                            groups=self.groups, bias=False, padding_mode=self.padding_mode, device=self.device,
                            dtype=self.dtype)

    # deep copies (EMA, ExplanationsLogger) must not share launch plans
    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_group_caches", "_group_bias"):
                continue
            setattr(new, k, R._PlanCache() if k == "_cache" else copy.deepcopy(v, memo))
        return new

    def _effective_weight(self) -> Tensor:
        return self.linear.effective_weight()

    def forward(self, in_tensor: Tensor) -> Tensor:
        return self.forward_impl(in_tensor)

    def forward_impl(self, in_tensor: Tensor) -> Tensor:
        if any(d > 1 for d in self.dilation) or self.padding_mode != "zeros":
            # the reference's own patch norm (calc_patch_norms, bcosconv2d.py:196-231: a zero-padded, undilated sum pool) does not
            # cover these either: its output shape / border values disagree with the convolution's
            raise NotImplementedError("bcos_b200: dilation > 1 and non-zero padding modes are not built (the reference's "
                                      "calc_patch_norms does not support them; no registered config uses them)")
        if self.max_out > 1 and (self.out_channels // self.groups) % 8 != 0:
            raise NotImplementedError("bcos_b200: MaxOut needs (out_channels / groups) % 8 == 0")
        b = float(self.b.detach()) if isinstance(self.b, torch.Tensor) else float(self.b)
        lin = self.linear
        if self.groups > 1:
            return self._forward_groups(in_tensor, b)
        extra = (getattr(lin, "scale", None), getattr(lin, "use_weight_norm", None))     # what else the effective weight depends on
        k, st, pd = _single(self.kernel_size), _single(self.stride), _single(self.padding)
        if k == st and pd == 0 and k * k > 64:
            # patch-embedding convolution (CLIP ViT conv1: 16x16 / 32x32 non-overlapping patches): the same B-cos transform is
            # a 1x1 map over the unfolded patches, channel order (c, i, j) = the weight's own flattening; the patch norm sums
            # the whole patch either way.  (The implicit GEMM walks at most 64 filter taps.)
            n, c, h, w = in_tensor.shape
            assert h % k == 0 and w % k == 0, "patch-embedding conv needs an input that is a whole number of patches"
            xp = in_tensor.reshape(n, c, h // k, k, w // k, k).permute(0, 1, 3, 5, 2, 4).reshape(n, c * k * k, h // k, w // k)
            return R.bcos_map(xp, self._cache, lin.weight, getattr(lin, "bias", None),
                              lambda: self._effective_weight().reshape(lin.weight.shape[0], c * k * k, 1, 1), 1, 0, b, self.detach,
                              max_out=self.max_out, extra=extra)
        return R.bcos_map(in_tensor, self._cache, lin.weight, getattr(lin, "bias", None), self._effective_weight,
                          st, pd, b, self.detach, max_out=self.max_out, extra=extra)

    def _forward_groups(self, in_tensor: Tensor, b: float) -> Tensor:
        """groups > 1 (bcosconv2d.py:201-209, 224-229): every group is its own B-cos map -- its filters see only the group's input
        channels and the patch norm is the group's -- so each runs as one launch plan over a channel slice (own plan cache)."""
        G, lin = self.groups, self.linear
        assert self.in_channels % G == 0 and self.out_channels % G == 0
        if not hasattr(self, "_group_caches") or len(self._group_caches) != G:
            self._group_caches = [R._PlanCache() for _ in range(G)]
        cg, og = self.in_channels // G, self.out_channels * self.max_out // G
        st, pd = _single(self.stride), _single(self.padding)
        bias = getattr(lin, "bias", None)
        pad_rows = (-og) % 8                      # launches write a multiple of 8 output columns: zero filters, sliced off again
        assert pad_rows == 0 or self.max_out == 1
        outs = []
        for g in range(G):
            rows = slice(g * og, (g + 1) * og)
            extra = (getattr(lin, "scale", None), getattr(lin, "use_weight_norm", None), g)

            def eff(rows=rows):
                w = self._effective_weight()[rows]
                return w if pad_rows == 0 else torch.cat([w, w.new_zeros((pad_rows,) + tuple(w.shape[1:]))], 0)

            bg = None
            if bias is not None and pad_rows == 0:
                bg = bias[rows]
            elif bias is not None:              # padded copy kept per (group, bias version): a fresh tensor per call would defeat the plan cache
                key = (g, bias.data_ptr(), bias._version)
                store = self.__dict__.setdefault("_group_bias", {})
                if key not in store:
                    for old in [k_ for k_ in store if k_[0] == g]:
                        del store[old]
                    store[key] = torch.cat([bias.detach()[rows], bias.new_zeros(pad_rows)], 0)
                bg = store[key]
            y = R.bcos_map(in_tensor[:, g * cg:(g + 1) * cg], self._group_caches[g], lin.weight, bg, eff, st, pd, b, self.detach,
                           max_out=self.max_out, extra=extra + ((bias.data_ptr(), bias._version) if bias is not None else ()))
            outs.append(y if pad_rows == 0 else y[:, :og])
        return torch.cat(outs, 1)

    def calc_patch_norms(self, in_tensor: Tensor) -> Tensor:
        """||patch|| per output position, [N,1,Ho,Wo] (bcosconv2d.py:196-231) through bcosk_patch_inv_norm; groups > 1: [N,O,Ho,Wo] with
        the norm of every output channel's own group, like the reference's repeat_interleave."""
        R._require_cuda(in_tensor, "calc_patch_norms")
        k, s, p = _single(self.kernel_size), _single(self.stride), _single(self.padding)
        if self.groups > 1:
            G, cg = self.groups, self.in_channels // self.groups
            per = [_patch_norms(in_tensor[:, g * cg:(g + 1) * cg], k, s, p) for g in range(G)]
            return torch.repeat_interleave(torch.cat(per, 1), self.out_channels // G, dim=1)
        return _patch_norms(in_tensor, k, s, p)

    def extra_repr(self) -> str:
        s = "B={b}"
        if self.max_out > 1:
            s += ", max_out={max_out}"
        s += ","
        extra = dict(b=self.b.data.item()) if isinstance(self.b, nn.Parameter) else {}
        return s.format(**{**self.__dict__, **extra})


class BcosConv2dWithScale(BcosConv2d):
    """Deprecated reference class (bcosconv2d.py:265-326): output divided by a constant scale."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, groups=1,
                 padding_mode="zeros", device=None, dtype=None, b=2, max_out=1, scale=None, scale_factor=100.0, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, padding_mode,
                         device=device, dtype=dtype, b=b, max_out=max_out, **kwargs)
        if scale is None:
            ks = kernel_size if not isinstance(kernel_size, tuple) else (kernel_size[0] * kernel_size[1]) ** 0.5
            self.scale = (ks * (in_channels / groups) ** 0.5) / scale_factor
        else:
            assert scale != 1.0, "For scale=1.0, use the normal BcosConv2d instead!"
            self.scale = scale

    def forward(self, in_tensor: Tensor) -> Tensor:
        return self.forward_impl(in_tensor) / self.scale


class BcosifyConv2d(BcosConv2d):
    """B-cosified convolution: plain (un-normalised) nn.Conv2d weights, optional bias (bcosifyconv2d.py:7-187)."""

    def __init__(self, *args, clamping: bool = False, b_loss: bool = False, **kwargs):
        # the reference builds its nn.Conv2d with bias=self.bias, which BcosConv2d.__init__ has set to None: no `linear.bias`
        # exists unless from_standard_module copies one in together with the weights (bcosifyconv2d.py:14-30, 143-147)
        self._want_bias = False
        super().__init__(*args, **kwargs)
        self.clamping = clamping
        self.b_loss = b_loss

    def _make_linear(self) -> nn.Module:
        return nn.Conv2d(in_channels=self.in_channels, out_channels=self.out_channels * self.max_out,
                         kernel_size=self.kernel_size, stride=self.stride, padding=self.padding, dilation=self.dilation,
                         groups=self.groups, bias=self._want_bias, padding_mode=self.padding_mode, device=self.device,
                         dtype=self.dtype)

    @property
    def weight(self) -> Tensor:  # CLIP reads conv1.weight.dtype (bcosifyconv2d.py:35-37)
        return self.linear.weight

    def _effective_weight(self) -> Tensor:
        return self.linear.weight

    def forward_impl(self, in_tensor: Tensor) -> Tensor:
        if self.clamping or self.b_loss:
            raise NotImplementedError("bcos_b200: learnable-B variants (clamping / b_loss) are not built (dead in all configs)")
        return super().forward_impl(in_tensor)

    @classmethod
    def from_standard_module(cls, mod, model_config):
        """bcosifyconv2d.py:116-148."""
        new_mod = cls(in_channels=mod.in_channels, out_channels=mod.out_channels, kernel_size=mod.kernel_size,
                      stride=mod.stride, padding=mod.padding, dilation=mod.dilation, groups=mod.groups,
                      bias=mod.bias is not None, padding_mode=mod.padding_mode,
                      clamping=model_config["bcosify_args"].get("clamping", False),
                      b_loss=model_config["bcosify_args"].get("learn_b", False), b=model_config["bcos_args"].get("b", 1))
        if model_config.get("weights", None) is not None:
            new_mod.linear.weight.data = mod.weight.data
            if mod.bias is not None:
                new_mod.linear.bias = nn.Parameter(mod.bias.data)
        return new_mod

    @classmethod
    def from_standard_module_linear(cls, mod, model_config):
        """nn.Linear -> 1x1 B-cos conv for the classifier applied before GAP (bcosifyconv2d.py:151-182)."""
        new_mod = cls(in_channels=mod.in_features, out_channels=mod.out_features, kernel_size=1, stride=1, padding=0,
                      dilation=1, groups=1, bias=mod.bias is not None, padding_mode="zeros",
                      clamping=model_config["bcosify_args"].get("clamping", False),
                      b_loss=model_config["bcosify_args"].get("learn_b", False), b=model_config["bcos_args"].get("b", 1))
        if model_config.get("weights", None) is not None:
            new_mod.linear.weight.data = mod.weight.data.view_as(new_mod.linear.weight.data)
            if mod.bias is not None:
                new_mod.linear.bias = nn.Parameter(mod.bias.data)
        return new_mod
