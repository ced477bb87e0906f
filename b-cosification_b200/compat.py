"""Let the reference's own code (bcosify.py, bcosify_vit.py, bcos/models/*, evaluate.py) run on top of bcos_b200:
`install_as_bcos()` registers our modules under the reference's import paths (`bcos.modules`, `bcos.common`, ...)."""
from __future__ import annotations

import sys
import types


def install_as_bcos() -> None:
    from . import explain, modules
    from .modules import bcosconv2d, bcoslinear, common, logitlayer, norms

    def alias(name, **attrs):
        m = sys.modules.get(name)
        if m is None or not getattr(m, "__bcos_b200__", False):
            m = types.ModuleType(name)
            m.__bcos_b200__ = True
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    pub = {k: getattr(modules, k) for k in modules.__all__ if k not in ("norms", "config", "set_precision")}
    alias("bcos")
    alias("bcos.common", BcosUtilMixin=explain.BcosUtilMixin, explanation_mode=explain.explanation_mode,
          gradient_to_image=explain.gradient_to_image)
    alias("bcos.modules", **pub)
    alias("bcos.modules.bcosattnpool", BcosAttentionPool2d=modules.BcosAttentionPool2d)
    alias("bcos.modules.norms.centered_norms", DetachableLayerNorm=modules.DetachableLayerNorm,
          DetachableGroupNorm2d=norms.DetachableGroupNorm2d, DetachableGNInstanceNorm2d=norms.DetachableGNInstanceNorm2d,
          DetachableGNLayerNorm2d=norms.DetachableGNLayerNorm2d, DetachablePositionNorm2d=norms.DetachablePositionNorm2d)
    alias("bcos.modules.common", DetachableModule=common.DetachableModule, BcosSequential=common.BcosSequential)
    alias("bcos.modules.bcosconv2d", BcosConv2d=bcosconv2d.BcosConv2d, NormedConv2d=bcosconv2d.NormedConv2d,
          BcosConv2dWithScale=bcosconv2d.BcosConv2dWithScale)
    alias("bcos.modules.bcosifyconv2d", BcosifyConv2d=bcosconv2d.BcosifyConv2d)
    alias("bcos.modules.bcoslinear", BcosLinear=bcoslinear.BcosLinear, NormedLinear=bcoslinear.NormedLinear)
    alias("bcos.modules.bcosifylinear", BcosifyLinear=bcoslinear.BcosifyLinear)
    alias("bcos.modules.logitlayer", LogitLayer=logitlayer.LogitLayer)
    unc = {k: getattr(norms, k) for k in ("BatchNormUncentered2d", "GroupNormUncentered2d", "GNInstanceNormUncentered2d",
                                          "GNLayerNormUncentered2d", "PositionNormUncentered2d", "AllNormUncentered2d",
                                          "group_norm_uncentered", "batch_norm_uncentered_2d")}
    cen = {k: getattr(norms, k) for k in ("DetachableGroupNorm2d", "DetachableGNInstanceNorm2d", "DetachableGNLayerNorm2d",
                                          "DetachablePositionNorm2d")}
    alias("bcos.modules.norms", NoBias=norms.NoBias, Unaffine=norms.Unaffine, DetachableLayerNorm=modules.DetachableLayerNorm,
          **unc, **cen)
    alias("bcos.modules.norms.uncentered_norms", **unc)
    alias("bcos.modules.norms.utils", NoBias=norms.NoBias, Unaffine=norms.Unaffine)
