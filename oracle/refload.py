"""Import shim for the read-only reference checkout (THIS CONTAINER ONLY).

TEST INFRASTRUCTURE -- never imported by the product package.

`import bcos` on the reference fails here (torchmetrics / pytorch_lightning / ftfy are
absent), so we pre-seed `sys.modules` with empty package objects whose `__path__` points
into /root/reference; their `__init__.py` never run, while the hot-path sub-modules
(bcos.modules.*, bcos.models.{resnet,densenet,vit,standard_models}, bcosify, bcosify_vit,
CLIP.clip.model) import unchanged.  Used only by oracle/make_golden.py and the
reference-vs-oracle pin tests, which skip when /root/reference is absent (GPU box).
"""
import importlib.machinery
import os
import sys
import types

REF = os.environ.get("BCOS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "bcos", "modules"))


def _ns(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    m.__spec__.submodule_search_locations = [path]
    sys.modules[name] = m
    return m


_loaded = False


def load():
    """Make `bcos.modules`, `bcos.models.*`, `bcosify`, `bcosify_vit`, `CLIP.clip.model` importable."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF}")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _ns("bcos", REF + "/bcos")
    _ns("bcos.models", REF + "/bcos/models")
    _ns("CLIP", REF + "/CLIP")
    _ns("CLIP.clip", REF + "/CLIP/clip")
    import bcos.modules  # noqa: F401
    import bcos.common  # noqa: F401
    import bcos.models.standard_models  # noqa: F401
    _loaded = True
