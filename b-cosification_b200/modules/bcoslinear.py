"""B-cos linear layers -- drop-in mirror of reference bcos/modules/bcoslinear.py and bcosifylinear.py.
Evaluated as a 1x1 B-cos "convolution" over the flattened token axis (same tcgen05 kernel; norm = ||x|| + 1e-12)."""
from __future__ import annotations

from typing import Union

import torch
import torch.linalg as LA
import torch.nn as nn
from torch import Tensor

from . import _runtime as R
from .common import DetachableModule

__all__ = ["NormedLinear", "BcosLinear", "BcosifyLinear"]


class NormedLinear(nn.Linear):
    """nn.Linear used with unit-norm rows (bcoslinear.py:20-27); parameter container."""

    def effective_weight(self) -> Tensor:
        return self.weight / LA.vector_norm(self.weight, dim=1, keepdim=True)

    def forward(self, input: Tensor) -> Tensor:
        raise RuntimeError("NormedLinear is evaluated through its BcosLinear")


class BcosLinear(DetachableModule):
    """bcoslinear.py:30-141."""

    def __init__(self, in_features: int, out_features: int, bias: bool = False, device=None, dtype=None,
                 b: Union[int, float] = 2, max_out: int = 1) -> None:
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.bias = False
        self.device = device
        self.dtype = dtype
        self.b = b
        self.max_out = max_out
        self.linear = self._make_linear(bias)
        self._cache = R._PlanCache()

    def _make_linear(self, bias: bool) -> nn.Module:
        return NormedLinear(self.in_features, self.out_features * self.max_out, bias=False, device=self.device, dtype=self.dtype)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, R._PlanCache() if k == "_cache" else copy.deepcopy(v, memo))
        return new

    def _effective_weight(self) -> Tensor:
        return self.linear.effective_weight()[:, :, None, None]

    def forward(self, in_tensor: Tensor) -> Tensor:
        if self.max_out > 1 and self.out_features % 8 != 0:
            raise NotImplementedError("bcos_b200: MaxOut needs out_features % 8 == 0")
        b = float(self.b.detach()) if isinstance(self.b, torch.Tensor) else float(self.b)
        lead = in_tensor.shape[:-1]
        x = in_tensor.reshape(-1, self.in_features, 1, 1)
        lin = self.linear
        y = R.bcos_map(x, self._cache, lin.weight, getattr(lin, "bias", None), self._effective_weight, 1, 0, b, self.detach,
                       linear_eps=True, max_out=self.max_out)
        return y.reshape(*lead, self.out_features)

    def extra_repr(self) -> str:
        s = "B={b}"
        if self.max_out > 1:
            s += ", max_out={max_out}"
        s += ","
        extra = dict(b=self.b.data.item()) if isinstance(self.b, nn.Parameter) else {}
        return s.format(**{**self.__dict__, **extra})


class BcosifyLinear(BcosLinear):
    """B-cosified linear: plain nn.Linear weights, optional bias (bcosifylinear.py:17-133)."""

    def __init__(self, *args, clamping: bool = False, b_loss: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.clamping = clamping
        self.b_loss = b_loss

    def _make_linear(self, bias: bool) -> nn.Module:
        # the reference passes bias=self.bias, which BcosLinear.__init__ leaves False/None: no `linear.bias` exists unless
        # from_standard_module copies one in together with the weights (bcosifylinear.py:28-34, 128-132)
        self.bias = None
        return nn.Linear(in_features=self.in_features, out_features=self.out_features * self.max_out, bias=False,
                         device=self.device, dtype=self.dtype)

    @property
    def weight(self) -> Tensor:
        return self.linear.weight

    def _effective_weight(self) -> Tensor:
        return self.linear.weight[:, :, None, None]

    def forward(self, in_tensor: Tensor) -> Tensor:
        if self.clamping or self.b_loss:
            raise NotImplementedError("bcos_b200: learnable-B variants (clamping / b_loss) are not built (dead in all configs)")
        return super().forward(in_tensor)

    @classmethod
    def from_standard_module(cls, mod, model_config):
        """bcosifylinear.py:109-133."""
        new_mod = cls(mod.in_features, mod.out_features, bias=mod.bias is not None, device=mod.weight.device,
                      dtype=mod.weight.dtype, max_out=1, clamping=model_config["bcosify_args"].get("clamping", False),
                      b_loss=model_config["bcosify_args"].get("learn_b", False), b=model_config["bcos_args"].get("b", 1))
        if model_config.get("weights", None) is not None:
            new_mod.linear.weight.data = mod.weight.data
            if mod.bias is not None:
                new_mod.linear.bias = nn.Parameter(mod.bias.data)
        return new_mod
