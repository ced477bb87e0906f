"""Drop-in mirror of the reference's `bcos.modules` surface, executed on libbcosk.so (CUDA only)."""
from . import norms
from ._runtime import config, set_precision
from .bcosattnpool import BcosAttentionPool2d
from .bcosconv2d import BcosConv2d, BcosConv2dWithScale, BcosifyConv2d, NormedConv2d
from .bcoslinear import BcosifyLinear, BcosLinear, NormedLinear
from .common import BcosSequential, DetachableModule
from .logitlayer import LogitLayer
from .norms import (AllNormUncentered2d, BatchNormUncentered2d, DetachableGNInstanceNorm2d, DetachableGNLayerNorm2d,
                    DetachableGroupNorm2d, DetachablePositionNorm2d, GNInstanceNormUncentered2d, GNLayerNormUncentered2d,
                    GroupNormUncentered2d, NoBias, PositionNormUncentered2d, Unaffine, batch_norm_uncentered_2d,
                    group_norm_uncentered)
from .tokens import DetachableLayerNorm, MyGELU, PlainLinear, frozen_attention

__all__ = ["BcosAttentionPool2d", "BcosConv2d", "BcosConv2dWithScale", "BcosifyConv2d", "NormedConv2d", "BcosLinear", "BcosifyLinear",
           "NormedLinear", "BcosSequential", "DetachableModule", "LogitLayer", "BatchNormUncentered2d", "NoBias", "Unaffine",
           "batch_norm_uncentered_2d", "group_norm_uncentered", "GroupNormUncentered2d", "GNInstanceNormUncentered2d",
           "GNLayerNormUncentered2d", "DetachableGroupNorm2d", "DetachableGNInstanceNorm2d", "DetachableGNLayerNorm2d",
           "PositionNormUncentered2d", "DetachablePositionNorm2d", "AllNormUncentered2d", "DetachableLayerNorm", "MyGELU", "PlainLinear", "frozen_attention", "norms", "config",
           "set_precision"]
