"""Let the reference's own code (bcosify.py, bcosify_vit.py, bcos/models/*, evaluate.py) run on top of bcos_b200:
`install_as_bcos()` registers our modules under the reference's import paths (`bcos.modules`, `bcos.common`, ...)."""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types
from typing import Optional

# packages of the reference whose __init__ pulls in the training stack (torchmetrics, pytorch_lightning, ftfy ...): when a
# reference checkout is known they are registered as path-only packages so that their hot-path submodules
# (bcos.models.resnet / standard_models / vit, CLIP.clip.model) import without it
_PATH_ONLY = (("bcos.models", "bcos/models"), ("CLIP", "CLIP"), ("CLIP.clip", "CLIP/clip"))


def _find_reference_root(explicit: Optional[str]) -> Optional[str]:
    cands = [explicit, os.environ.get("BCOS_REFERENCE_ROOT")]
    real = sys.modules.get("bcos")
    if real is not None and not getattr(real, "__bcos_b200__", False) and getattr(real, "__path__", None):
        cands.append(os.path.dirname(list(real.__path__)[0]))
    cands += list(sys.path)
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "bcos", "modules", "bcosconv2d.py")):
            return os.path.abspath(c)
        if c and c.endswith(".zip") and os.path.isfile(c):        # an archive of the checkout (zipimport resolves archive/sub/dir paths)
            import zipfile
            with zipfile.ZipFile(c) as z:
                if "bcos/modules/bcosconv2d.py" in z.namelist():
                    return os.path.abspath(c)
    return None


def _path_package(name: str, path: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    m.__spec__.submodule_search_locations = [path]
    return m


def install_as_bcos(reference_root: Optional[str] = None) -> Optional[str]:
    """Register the bcos_b200 modules under the reference's import paths.

    Only `bcos.modules[.*]` and `bcos.common` are replaced.  A real `bcos` package that is already imported stays in place
    (its other subpackages - bcos.models, bcos.data, bcos.experiments, bcos.training - keep working); when none is imported,
    `bcos` becomes a path-only package over the reference checkout (argument, $BCOS_REFERENCE_ROOT, or a sys.path entry that
    holds bcos/modules/bcosconv2d.py; a .zip archive of the checkout works too), so `import bcos.models.resnet` etc. resolve to
    the reference's own files.
    Returns the reference root in use (None: no checkout found, only the replaced modules are importable)."""
    from . import explain, modules
    from .modules import bcosconv2d, bcoslinear, common, logitlayer, norms

    ref = _find_reference_root(reference_root)
    top = sys.modules.get("bcos")
    if top is None or getattr(top, "__bcos_b200__", False):
        top = _path_package("bcos", os.path.join(ref, "bcos")) if ref else types.ModuleType("bcos")
        if not ref:
            top.__path__ = []
        top.__bcos_b200__ = True
        sys.modules["bcos"] = top
    if ref:
        for name, rel in _PATH_ONLY:
            if name not in sys.modules:
                sys.modules[name] = _path_package(name, os.path.join(ref, rel))
                parent, _, leaf = name.rpartition(".")
                if parent in sys.modules:
                    setattr(sys.modules[parent], leaf, sys.modules[name])

    def alias(name, **attrs):
        m = sys.modules.get(name)
        if m is None or not getattr(m, "__bcos_b200__", False):
            m = types.ModuleType(name)
            m.__bcos_b200__ = True
            m.__path__ = []
            sys.modules[name] = m
            parent, _, leaf = name.rpartition(".")
            if parent in sys.modules:
                setattr(sys.modules[parent], leaf, m)       # `bcos.modules` attribute access, also on a real parent package
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    pub = {k: getattr(modules, k) for k in modules.__all__ if k not in ("norms", "config", "set_precision")}
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
    return ref
