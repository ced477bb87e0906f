"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/bcosk.h declares.
No compute call is made here (there is no GPU); product entry points must refuse to run without one."""
import ctypes
import re

import pytest
import torch

from bcos_b200 import _lib as L


def test_library_exports_every_declared_symbol(bcosk_lib):
    with open(L.HEADER) as fh:
        src = fh.read()
    declared = set(re.findall(r"\b(bcosk_[a-z0-9_]+)\s*\(", src))
    declared.discard("bcosk_igemm_params")
    assert len(declared) >= 16
    for name in sorted(declared):
        assert hasattr(bcosk_lib, name), f"libbcosk.so lacks {name}"
    assert bcosk_lib.bcosk_version() >= 100


def test_struct_layout_matches_header(bcosk_lib):
    assert bcosk_lib.bcosk_sizeof_igemm_params() == ctypes.sizeof(L.IgemmParams)
    names = [f[0] for f in L.IgemmParams._fields_]
    for must in ("a", "b", "y", "inv_norm", "tap_off_w", "seg_a_choff", "mask2", "os_q"):
        assert must in names


def test_bad_arguments_are_rejected_without_touching_a_device(bcosk_lib):
    p = L.IgemmParams()
    rc = bcosk_lib.bcosk_igemm(ctypes.byref(p), None)
    assert rc == L.BCOSK_EINVAL
    assert b"null" in bcosk_lib.bcosk_last_error()
    assert bcosk_lib.bcosk_igemm(None, None) == L.BCOSK_EINVAL


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_product_fails_loudly_without_gpu(bcosk_lib):
    from bcos_b200.engine import ResNetPlan
    from bcos_b200.models import resnet_state_shapes
    from bcos_b200.utils import synth
    with pytest.raises(L.BcoskError):
        L.require_device()
    plan = ResNetPlan("resnet18", synth.synth_state_dict(resnet_state_shapes("resnet18")), 1, device="cpu", image_size=32)
    with pytest.raises(L.BcoskError):
        plan.explain(torch.zeros(1, 6, 32, 32))
