"""Drop-in mirror of the reference's `bcos.modules` surface, executed on libbcosk.so (CUDA only)."""
from . import norms
from ._runtime import config, set_precision
from .bcosattnpool import BcosAttentionPool2d
from .bcosconv2d import BcosConv2d, BcosConv2dWithScale, BcosifyConv2d, NormedConv2d
from .bcoslinear import BcosifyLinear, BcosLinear, NormedLinear
from .common import BcosSequential, DetachableModule
from .logitlayer import LogitLayer
from .norms import BatchNormUncentered2d, NoBias, Unaffine, batch_norm_uncentered_2d
from .tokens import DetachableLayerNorm, MyGELU, PlainLinear, frozen_attention

__all__ = ["BcosAttentionPool2d", "BcosConv2d", "BcosConv2dWithScale", "BcosifyConv2d", "NormedConv2d", "BcosLinear", "BcosifyLinear",
           "NormedLinear", "BcosSequential", "DetachableModule", "LogitLayer", "BatchNormUncentered2d", "NoBias", "Unaffine",
           "batch_norm_uncentered_2d", "DetachableLayerNorm", "MyGELU", "PlainLinear", "frozen_attention", "norms", "config",
           "set_precision"]
