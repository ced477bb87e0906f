"""`LogitLayer` -- mirror of reference bcos/modules/logitlayer.py:11-36: out / T + b."""
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _runtime as R

__all__ = ["LogitLayer"]


class LogitLayer(nn.Module):
    def __init__(self, logit_temperature: Optional[float] = None, logit_bias: Optional[float] = None):
        super().__init__()
        self.logit_bias = logit_bias
        self.logit_temperature = logit_temperature

    def forward(self, in_tensor: Tensor) -> Tensor:
        if self.logit_temperature is None and self.logit_bias is None:
            return in_tensor
        R._require_cuda(in_tensor, "LogitLayer")
        x = in_tensor.float().contiguous()
        x4 = x.reshape(x.shape[0], -1, 1, 1)
        smul = 1.0 if self.logit_temperature is None else 1.0 / float(self.logit_temperature)
        sadd = 0.0 if self.logit_bias is None else float(self.logit_bias)
        y = R.ChannelAffineFn.apply(x4, None, None, smul, sadd, True)
        return y.reshape(x.shape).type(in_tensor.dtype)

    def extra_repr(self) -> str:
        ret = ""
        if self.logit_temperature is not None:
            ret += f"logit_temperature={self.logit_temperature}, "
        if self.logit_bias is not None:
            ret += f"logit_bias={self.logit_bias}, "
        return ret[:-2]
