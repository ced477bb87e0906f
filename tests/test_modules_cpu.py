"""CPU: the drop-in module surface (constructors, attributes, state-dict keys, deepcopy, CPU construction) and - where
the reference checkout exists - that the reference's OWN bcosify.py runs unchanged on top of our modules."""
import copy
import subprocess
import sys
import os

import pytest
import torch
import torch.nn as nn

import refload
from bcos_b200 import _lib as L
import bcos_b200.modules as M
from bcos_b200.bcosify import bcosified_resnet
from bcos_b200.models import resnet_state_shapes
from bcos_b200.utils import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constructor_signatures_and_attributes():
    c = M.BcosConv2d(8, 16, 3, 2, 1, 1, 1, "zeros", None, None, False, 2, 1)       # positional order of the reference
    assert (c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, c.b, c.max_out) == (8, 16, 3, 2, 1, 2, 1)
    assert c.linear.weight.shape == (16, 8, 3, 3) and c.bias is None and c.detach is False
    c.set_explanation_mode(True)
    assert c.is_in_explanation_mode
    # like the reference (checked against the live classes): the constructor never creates `linear.bias` (BcosConv2d sets
    # self.bias = None before the nn.Conv2d is built); only from_standard_module copies one in together with the weights
    bc = M.BcosifyConv2d(8, 16, kernel_size=1, bias=True, b=2)
    assert bc.weight is bc.linear.weight and bc.linear.bias is None
    assert list(bc.state_dict().keys()) == ["linear.weight"]
    import torch.nn as nn
    cfg = dict(weights=None, bcos_args=dict(b=2), bcosify_args={})
    assert list(M.BcosifyConv2d.from_standard_module(nn.Conv2d(3, 4, 3, bias=True), cfg).state_dict()) == ["linear.weight"]
    assert list(M.BcosifyLinear.from_standard_module(nn.Linear(8, 4, bias=True), cfg).state_dict()) == ["linear.weight"]
    cfg["weights"] = "IMAGENET1K_V1"
    assert list(M.BcosifyConv2d.from_standard_module(nn.Conv2d(3, 4, 3, bias=True), cfg).state_dict()) == ["linear.weight", "linear.bias"]
    assert list(M.BcosifyLinear.from_standard_module(nn.Linear(8, 4, bias=True), cfg).state_dict()) == ["linear.weight", "linear.bias"]
    lin = M.BcosLinear(32, 10, b=2, max_out=2)
    assert lin.linear.weight.shape == (20, 32) and lin.bias is False
    bl = M.BcosifyLinear(32, 10, bias=False, b=2)
    assert bl.weight.shape == (10, 32)
    bn = M.BatchNormUncentered2d(16)
    assert set(bn.state_dict()) == {"weight", "bias", "running_mean", "running_var", "num_batches_tracked"}
    nb = M.NoBias(M.BatchNormUncentered2d)(16)
    assert nb.bias is None and "NoBias" in nb._get_name()
    ll = M.LogitLayer(logit_temperature=None, logit_bias=-1.0)
    assert "logit_bias" in ll.extra_repr()


def test_from_standard_module_and_bn_fold():
    cfg = dict(weights="x", bcos_args=dict(b=2), bcosify_args=dict(norm_layer="BnUncV2"))
    conv = nn.Conv2d(3, 8, 3, 2, 1, bias=False)
    bc = M.BcosifyConv2d.from_standard_module(conv, cfg)
    assert torch.equal(bc.linear.weight, conv.weight) and bc.stride == (2, 2) and bc.b == 2
    fc = nn.Linear(16, 10)
    f1 = M.BcosifyConv2d.from_standard_module_linear(fc, cfg)
    assert f1.linear.weight.shape == (10, 16, 1, 1) and torch.equal(f1.linear.weight.flatten(1), fc.weight)
    bn = nn.BatchNorm2d(8)
    bn.running_mean.uniform_(-1, 1); bn.running_var.uniform_(0.5, 2); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.uniform_(-1, 1)
    u = M.BatchNormUncentered2d.from_standard_module(bn, cfg)
    std = (bn.running_var + bn.eps).sqrt()
    assert torch.allclose(u.bias.data, bn.bias.data - bn.running_mean / std * bn.weight.data)


def test_resnet_builder_matches_reference_state_dict_layout_and_deepcopies():
    m = bcosified_resnet("resnet18")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == resnet_state_shapes("resnet18")
    m.load_state_dict(synth.synth_state_dict(shapes), strict=True)
    m2 = copy.deepcopy(m).to("cpu")                      # ExplanationsLogger / EMA deep-copy models
    assert m2.model.conv1._cache is not m.model.conv1._cache
    with pytest.raises(L.BcoskError):                    # no CPU execution path
        m2(torch.zeros(1, 6, 32, 32))
    with m.explanation_mode():
        assert all(mod.detach for mod in m.modules() if hasattr(mod, "set_explanation_mode"))
    assert not any(mod.detach for mod in m.modules() if hasattr(mod, "set_explanation_mode"))


def test_clip_vit_mirror_has_the_reference_state_dict_layout():
    """CLIP ViT image encoder (CLIP/clip/model.py:166-241) after bcosify.py with clip_kd: conv1 / mlp -> B-cos modules with
    `.linear.weight` keys, out_proj a BcosifyLinear object inside nn.MultiheadAttention, Sequentials -> BcosSequential with
    positional keys; biases and positional embedding stripped by the factory (clip_bcosification/model.py:17-25)."""
    from bcos_b200.clip_vit import bcosified_clip_vit
    from bcos_oracle import clip_vit_state_shapes
    m = bcosified_clip_vit(64, 32, 64, 2, 2, 32)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == clip_vit_state_shapes(64, 32, 64, 2, 32)
    blk = m.model.transformer.resblocks[0]
    assert isinstance(m.model.conv1, M.BcosifyConv2d) and isinstance(blk.mlp, M.BcosSequential)
    assert isinstance(blk.mlp[0], M.BcosifyLinear) and isinstance(blk.attn.out_proj, M.BcosifyLinear)
    assert m.model.positional_embedding is None and blk.mlp[0].linear.bias is None


@pytest.mark.skipif(not refload.live(), reason="reference checkout not present (GPU box)")
def test_reference_bcosify_runs_unchanged_on_our_modules():
    """The reference's bcosify.py + bcos/models/standard_models.py, imported unmodified, build the network out of OUR
    modules (registered under the reference's import paths) with the reference's state-dict layout."""
    code = r"""
import sys, importlib.machinery, types
sys.dont_write_bytecode = True
sys.path.insert(0, %r); sys.path.insert(0, %r)
from bcos_b200.compat import install_as_bcos
REF = %r
sys.path.insert(0, REF)
assert install_as_bcos() == REF                  # finds the checkout on sys.path; no hand-made package stubs
import bcos_b200.modules as M
import bcos.models.resnet, bcos.models.vit       # untouched reference subpackages resolve to the reference's files
assert bcos.models.resnet.__file__.startswith(REF)
import torch.nn as nn
import bcosify                                   # the reference file, unmodified
from bcos.models.standard_models import ResNetBcos
from torchvision.models.resnet import Bottleneck
cfg = dict(is_bcos=True, name='resnet50', last_layer_name='fc', weights=None, bcos_args=dict(b=2, max_out=1),
           bcosify_args=dict(fix_b=True, use_bias=False, norm_layer='BnUncV2', manual_optim=False, gap=True, act_layer=True))
net = bcosify.BcosifyNetwork(ResNetBcos(Bottleneck, [3, 4, 6, 3]), cfg, add_channels=True, logit_layer=True)
assert isinstance(net.model.conv1, M.BcosifyConv2d) and isinstance(net.model.bn1, M.BatchNormUncentered2d)
assert isinstance(net.model.fc, M.BcosifyConv2d) and isinstance(net.logit_layer, M.LogitLayer)
assert isinstance(net.model.layer1, M.BcosSequential)
net.model.maxpool = nn.AvgPool2d(3, 2, 1)
for mod in net.modules():
    if hasattr(mod, 'bias') and mod.bias is not None: mod.bias = None
from bcos_b200.models import resnet_state_shapes
assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == resnet_state_shapes('resnet50')
print('REFERENCE_BCOSIFY_OK')
""" % (ROOT, os.path.join(ROOT, "oracle"), refload.REF)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "REFERENCE_BCOSIFY_OK" in out.stdout, out.stdout + out.stderr
