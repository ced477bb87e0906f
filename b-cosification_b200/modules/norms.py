"""Norm layers with detachable statistics -- mirror of the reference's bcos/modules/norms: uncentered_norms/
{batchnorm,groupnorm,posnorm,allnorm}_uncentered.py, the 2-D classes of centered_norms.py and utils.py."""
from __future__ import annotations

from functools import wraps

import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib as L
from . import _runtime as R
from .common import DetachableModule

__all__ = ["BatchNormUncentered2d", "NoBias", "Unaffine", "batch_norm_uncentered_2d", "group_norm_uncentered",
           "GroupNormUncentered2d", "GNInstanceNormUncentered2d", "GNLayerNormUncentered2d", "DetachableGroupNorm2d",
           "DetachableGNInstanceNorm2d", "DetachableGNLayerNorm2d", "PositionNormUncentered2d", "DetachablePositionNorm2d",
           "AllNormUncentered2d", "all_norm_uncentered_2d"]

_NOT_BUILT = ("bcos_b200: only the explanation-mode backward (detached statistics) is built; "
              "the full training backward is outside this round's scope")


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


# ------------------------------------------------------------------------------------------------
# group / position norms (bcosk_groupnorm_* / bcosk_positionnorm_*)
# ------------------------------------------------------------------------------------------------
def _f32(t):
    return None if t is None else t.detach().float().contiguous()


class _GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, groups, weight, bias, eps, centred, detach):
        nb, c = x.shape[0], x.shape[1]
        hw = x.numel() // max(nb * c, 1)
        y = torch.empty_like(x)
        rstd = torch.empty(nb * groups, dtype=torch.float32, device=x.device)
        L.groupnorm_fwd(x, nb, c, hw, groups, weight, bias, eps, centred, y, rstd)
        ctx.args = (nb, c, hw, groups, weight, centred, detach)
        ctx.save_for_backward(rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        nb, c, hw, groups, weight, centred, detach = ctx.args
        if not detach:
            raise NotImplementedError(_NOT_BUILT)
        (rstd,) = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.empty_like(gy)
        L.groupnorm_explain_bwd(gy, nb, c, hw, groups, weight, rstd, centred, gx)
        return gx, None, None, None, None, None, None


class _PositionNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, centred, detach):
        nb, c = x.shape[0], x.shape[1]
        hw = x.shape[2] * x.shape[3]
        y = torch.empty_like(x)
        rstd = torch.empty(nb * hw, dtype=torch.float32, device=x.device)
        L.positionnorm_fwd(x, nb, c, hw, weight, bias, eps, centred, y, rstd)
        ctx.args = (nb, c, hw, weight, centred, detach)
        ctx.save_for_backward(rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        nb, c, hw, weight, centred, detach = ctx.args
        if not detach:
            raise NotImplementedError(_NOT_BUILT)
        (rstd,) = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.empty_like(gy)
        L.positionnorm_explain_bwd(gy, nb, c, hw, weight, rstd, centred, gx)
        return gx, None, None, None, None, None


def _group_norm(input: Tensor, num_groups: int, weight, bias, eps: float, centred: bool, detach: bool, who: str) -> Tensor:
    assert input.shape[1] % num_groups == 0, (
        "Number of channels in input should be divisible by num_groups, "
        f"but got input of shape {input.shape} and num_groups={num_groups}")
    R._require_cuda(input, who)
    x = input.float().contiguous()
    return _GroupNormFn.apply(x, int(num_groups), _f32(weight), _f32(bias), float(eps), centred, detach).type(input.dtype)


def group_norm_uncentered(input: Tensor, num_groups: int, weight=None, bias=None, eps: float = 1e-5, detach: bool = False):
    """groupnorm_uncentered.py:21-61: x / sqrt(var + eps) per (image, group), var = centred biased variance."""
    return _group_norm(input, num_groups, weight, bias, eps, False, detach, "group_norm_uncentered")


class GroupNormUncentered2d(nn.GroupNorm, DetachableModule):
    """groupnorm_uncentered.py:64-80."""

    def __init__(self, num_groups: int, num_channels: int, eps: float = 1e-5, affine: bool = True) -> None:
        DetachableModule.__init__(self)
        super().__init__(num_groups, num_channels, eps, affine)

    def forward(self, input: Tensor) -> Tensor:
        return group_norm_uncentered(input, self.num_groups, self.weight, self.bias, self.eps, detach=self.detach)


class GNInstanceNormUncentered2d(GroupNormUncentered2d):
    def __init__(self, num_channels: int, *args, **kwargs):
        super().__init__(num_groups=num_channels, num_channels=num_channels, *args, **kwargs)


class GNLayerNormUncentered2d(GroupNormUncentered2d):
    def __init__(self, num_channels: int, *args, **kwargs):
        super().__init__(num_groups=1, num_channels=num_channels, *args, **kwargs)


class DetachableGroupNorm2d(nn.GroupNorm, DetachableModule):
    """centered_norms.py:93-160: (x - mean) / sqrt(var + eps); only the variance is detached in explanation mode."""

    def __init__(self, *args, **kwargs):
        DetachableModule.__init__(self)
        super().__init__(*args, **kwargs)

    def forward(self, input: Tensor) -> Tensor:
        assert input.dim() == 4, f"Expected 4D input got {input.dim()}D instead!"
        return _group_norm(input, self.num_groups, self.weight, self.bias, self.eps, True, self.detach, "DetachableGroupNorm2d")

    @classmethod
    def from_standard_module(cls, standard_module: nn.GroupNorm, model_config: dict):
        new_mod = cls(num_groups=standard_module.num_groups, num_channels=standard_module.num_channels,
                      eps=standard_module.eps, affine=standard_module.affine)
        if model_config.get("weights", None) is not None:
            new_mod.weight.data = standard_module.weight.data
            if standard_module.bias is not None:
                new_mod.bias.data = standard_module.bias.data
        return new_mod


class DetachableGNInstanceNorm2d(DetachableGroupNorm2d):
    def __init__(self, num_channels: int, *args, **kwargs):
        super().__init__(num_groups=num_channels, num_channels=num_channels, *args, **kwargs)


class DetachableGNLayerNorm2d(DetachableGroupNorm2d):
    """A CNN detachable layer norm (centered_norms.py:172-184)."""

    def __init__(self, num_channels: int, *args, **kwargs):
        super().__init__(num_groups=1, num_channels=num_channels, *args, **kwargs)


class _PositionNormBase(nn.LayerNorm, DetachableModule):
    _centred = False

    def __init__(self, features: int, eps: float = 1e-5, affine: bool = True, device=None, dtype=None) -> None:
        assert isinstance(features, int), f"Provide #features as an int not {type(features)=}"
        DetachableModule.__init__(self)
        super().__init__(normalized_shape=features, eps=eps, elementwise_affine=affine, device=device, dtype=dtype)
        self.features = features

    def forward(self, x: Tensor) -> Tensor:
        assert x.dim() == 4, f"input should be 4D not {x.dim()}D"
        R._require_cuda(x, type(self).__name__)
        x32 = x.float().contiguous()
        return _PositionNormFn.apply(x32, _f32(self.weight), _f32(self.bias), float(self.eps), self._centred,
                                     self.detach).type(x.dtype)


class PositionNormUncentered2d(_PositionNormBase):
    """posnorm_uncentered.py:18-58: x / sqrt(var_c + eps) per pixel."""
    _centred = False


class DetachablePositionNorm2d(_PositionNormBase):
    """centered_norms.py:251-297: channel-wise layer norm per pixel, variance detached in explanation mode."""
    _centred = True


def all_norm_uncentered_2d(input: Tensor, running_var, weight=None, bias=None, training: bool = False, momentum: float = 0.1,
                           eps: float = 1e-5, detach: bool = False) -> Tensor:
    """allnorm_uncentered.py:21-61: one variance for the whole batch tensor (bcosk_channel_stats_nchw with the batch
    tensor as ONE channel), then the per-channel affine kernel with a broadcast multiplier."""
    assert input.dim() == 4, "input should be a 4d tensor!"
    R._require_cuda(input, "all_norm_uncentered_2d")
    x = input.float().contiguous()
    nb, c = x.shape[0], x.shape[1]
    if training:
        mean = torch.empty(1, dtype=torch.float32, device=x.device)
        var = torch.empty(1, dtype=torch.float32, device=x.device)
        L.channel_stats_nchw(x.detach(), 1, 1, x.numel(), mean, var)
        if running_var is not None:
            running_var.copy_((1 - momentum) * running_var + momentum * var)
    else:
        assert running_var is not None, "running_var must be defined in eval mode"
        var = running_var.float()
    alpha = 1.0 / (var + eps).sqrt()
    if weight is not None:
        alpha = alpha * weight.detach().float()
    alpha = alpha.reshape(-1).expand(c).contiguous()
    beta = None if bias is None else bias.detach().float().reshape(-1).expand(c).contiguous()
    y = R.ChannelAffineFn.apply(x, alpha, beta, 1.0, 0.0, (not training) or detach)
    return y.type(input.dtype)


class AllNormUncentered2d(nn.BatchNorm2d, DetachableModule):
    """allnorm_uncentered.py:64-128."""

    def __init__(self, num_features: int, *args, **kwargs) -> None:
        DetachableModule.__init__(self)
        super().__init__(1, *args, **kwargs)

    def forward(self, input):
        self._check_input_dim(input)
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
        return all_norm_uncentered_2d(
            input=input, running_var=self.running_var if not self.training or self.track_running_stats else None,
            weight=self.weight, bias=self.bias, training=bn_training, momentum=exponential_average_factor, eps=self.eps,
            detach=self.detach)


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
