"""Import shim for the read-only reference checkout (THIS CONTAINER ONLY).

TEST INFRASTRUCTURE -- never imported by the product package.

`import bcos` on the reference fails here (torchmetrics / pytorch_lightning / ftfy are
absent), so we pre-seed `sys.modules` with empty package objects whose `__path__` points
into /root/reference; their `__init__.py` never run, while the hot-path sub-modules
(bcos.modules.*, bcos.models.{resnet,densenet,vit,standard_models}, bcosify, bcosify_vit,
CLIP.clip.model) import unchanged.  Used by oracle/make_golden.py and the reference-vs-oracle
pin tests (which skip when /root/reference is absent), and by bench.py's CPU legs.

Where /root/reference does not exist (the GPU box) the same unmodified files are imported from the
archive `oracle/_ref/bcos_reference.zip` that `oracle/stage_ref.py` packs in this container.
"""
import importlib.machinery
import os
import sys
import types

LIVE = os.environ.get("BCOS_REFERENCE_ROOT", "/root/reference")
ARCHIVE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "bcos_reference.zip")


def live() -> bool:
    """The reference checkout itself is mounted (this container)."""
    return os.path.isdir(os.path.join(LIVE, "bcos", "modules"))


def staged() -> bool:
    return os.path.isfile(ARCHIVE)


def available() -> bool:
    return live() or staged()


def source() -> str:
    return "checkout" if live() else ("archive" if staged() else "none")


REF = LIVE if live() or not staged() else ARCHIVE


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
        raise RuntimeError(f"reference not found: neither {LIVE} nor {ARCHIVE}")
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
