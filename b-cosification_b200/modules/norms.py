"""Uncentered batch norm and helpers -- mirror of reference bcos/modules/norms/uncentered_norms/batchnorm_uncentered.py
and bcos/modules/norms/utils.py."""
from __future__ import annotations

from functools import wraps

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib as L
from . import _runtime as R
from .common import DetachableModule

__all__ = ["BatchNormUncentered2d", "NoBias", "Unaffine", "batch_norm_uncentered_2d"]


def batch_norm_uncentered_2d(input: Tensor, running_var, weight=None, bias=None, training: bool = False,
                             momentum: float = 0.1, eps: float = 1e-5, detach: bool = False) -> Tensor:
    """batchnorm_uncentered.py:21-60 on the CUDA kernels (bcosk_channel_stats_nchw + bcosk_scale_bias_nchw)."""
    assert input.dim() == 4, "input should be a 4d tensor!"
    R._require_cuda(input, "batch_norm_uncentered_2d")
    x = input.float().contiguous()
    nb, c = x.shape[0], x.shape[1]
    hw = x.shape[2] * x.shape[3]
    if training:
        mean = torch.empty(c, dtype=torch.float32, device=x.device)
        var = torch.empty(c, dtype=torch.float32, device=x.device)
        L.channel_stats_nchw(x.detach(), nb, c, hw, mean, var)      # centred, biased variance (:39)
        if running_var is not None:
            running_var.copy_((1 - momentum) * running_var + momentum * var)
    else:
        assert running_var is not None, "running_var must be defined in eval mode"
        var = running_var.float()
    alpha = 1.0 / (var + eps).sqrt()
    if weight is not None:
        alpha = alpha * weight.detach().float()
    const_ok = (not training) or detach       # statistics are constants w.r.t. x in eval and in explanation mode
    y = R.ChannelAffineFn.apply(x, alpha.contiguous(), None if bias is None else bias.detach().float().contiguous(), 1.0, 0.0,
                                const_ok)
    return y.type(input.dtype)


class BatchNormUncentered2d(nn.BatchNorm2d, DetachableModule):
    """batchnorm_uncentered.py:63-141."""

    def __init__(self, *args, **kwargs):
        self.bias = kwargs.pop("bias", None)
        DetachableModule.__init__(self)
        super().__init__(*args, **kwargs)

    def forward(self, input):
        if self.momentum is None:
            exponential_average_factor = 0.0
        else:
            exponential_average_factor = self.momentum
        if self.training and self.track_running_stats:
            if self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
                if self.momentum is None:
                    exponential_average_factor = 1.0 / float(self.num_batches_tracked)
                else:
                    exponential_average_factor = self.momentum
        if self.training:
            bn_training = True
        else:
            bn_training = (self.running_mean is None) and (self.running_var is None)
        return batch_norm_uncentered_2d(
            input=input,
            running_var=self.running_var if not self.training or self.track_running_stats else None,
            weight=self.weight, bias=self.bias, training=bn_training, momentum=exponential_average_factor, eps=self.eps,
            detach=self.detach)

    @classmethod
    def from_standard_module(cls, mod, model_config):
        """'BnUncV2' fold so that the eval output equals the standard BatchNorm2d (batchnorm_uncentered.py:118-141)."""
        new_mod = cls(num_features=mod.num_features, eps=mod.eps, momentum=mod.momentum, affine=mod.affine,
                      track_running_stats=mod.track_running_stats, bias=mod.bias is not None)
        new_mod.weight.data = mod.weight.data
        norm_layer = model_config["bcosify_args"].get("norm_layer", "BnUncV2")
        if mod.bias is not None and norm_layer == "BnUncV2":
            std = (mod.running_var.data + mod.eps).sqrt()
            new_mod.bias.data = mod.bias.data - ((mod.running_mean.data / std) * mod.weight.data)
        else:
            new_mod.bias.data = mod.bias.data
        if mod.running_var is not None:
            new_mod.running_var.data = mod.running_var.data
        if mod.running_mean is not None:
            new_mod.running_mean.data = mod.running_mean.data
        return new_mod


def _append_to_name(mod, suffix):
    old_name = mod.__class__.__name__
    mod._get_name = lambda: old_name + suffix


def NoBias(make_layer):
    """norms/utils.py:18-51: build the layer, then remove its bias."""
    @wraps(make_layer)
    def init(*args, **kwargs):
        norm = make_layer(*args, **kwargs)
        assert norm.bias is not None, "It makes no sense to use this wrapper if you set affine=False!"
        norm.bias = None
        _append_to_name(norm, "NoBias")
        return norm
    if hasattr(make_layer, "__name__"):
        init.__name__ = make_layer.__name__ + "NoBias"
    if hasattr(make_layer, "__qualname__"):
        init.__qualname__ = make_layer.__qualname__ + "NoBias"
    return init


def Unaffine(make_layer):
    """norms/utils.py:54-88: build the layer, then remove bias and weight."""
    @wraps(make_layer)
    def init(*args, **kwargs):
        norm = make_layer(*args, **kwargs)
        assert norm.bias is not None, "It makes no sense to use this wrapper if you set affine=False!"
        norm.bias = None
        norm.weight = None
        _append_to_name(norm, "Unaffine")
        return norm
    if hasattr(make_layer, "__name__"):
        init.__name__ = make_layer.__name__ + "Unaffine"
    if hasattr(make_layer, "__qualname__"):
        init.__qualname__ = make_layer.__qualname__ + "Unaffine"
    return init
