"""B-cosification of standard CNNs on top of `bcos_b200.modules` -- mirror of reference bcosify.py:22-113 and the
FC-before-GAP torchvision subclasses of bcos/models/standard_models.py:36-63.

This is the module-level (drop-in) route: every Conv2d / Linear / BatchNorm2d of a torchvision model is swapped for the
CUDA-backed B-cos module, the first conv gets 6 input channels (`cat(W, -W) / 2`), the input is normalised and a
LogitLayer is appended.  The fused route for the same networks is `bcos_b200.engine.ResNetPlan`.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F
from torchvision.models import DenseNet, ResNet
from torchvision.models.resnet import BasicBlock, Bottleneck

from .explain import BcosUtilMixin
from .modules import BatchNormUncentered2d, BcosifyConv2d, BcosifyLinear, BcosSequential, LogitLayer
from .modules import _runtime as R
from .modules._runtime import ChannelAffineFn

IMAGENET_MEAN_ADDINVERSE = (0.485, 0.456, 0.406, 0.515, 0.544, 0.594)
IMAGENET_STD_ADDINVERSE = (0.229, 0.224, 0.225, 0.229, 0.224, 0.225)
CLIP_MEAN_ADDINVERSE = (0.48145466, 0.4578275, 0.40821073, 0.51854534, 0.5421725, 0.59178927)
CLIP_MEAN_ZERO = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
CLIP_STD_ADDINVERSE = (0.26862954, 0.26130258, 0.27577711, 0.26862954, 0.26130258, 0.27577711)


class ResNetBcos(ResNet):
    """standard_models.py:36-54: classifier applied per position before global average pooling."""

    def _forward_impl(self, x):
        x = self.conv1(x)
        x = self.bn1(x)
        x = self.relu(x)
        x = self.maxpool(x)
        x = self.layer1(x)
        x = self.layer2(x)
        x = self.layer3(x)
        x = self.layer4(x)
        x = self.fc(x)
        x = self.avgpool(x)
        return x.flatten(1)


class DenseNetBcos(DenseNet):
    """standard_models.py:56-63."""

    def forward(self, x):
        out = F.relu(self.features(x), inplace=True)
        out = self.classifier(out)
        out = F.adaptive_avg_pool2d(out, (1, 1))
        return torch.flatten(out, 1)


class Normalize6(nn.Module):
    """torchvision `transforms.Normalize(mean6, std6)` (bcosify.py:39-43) on the bcosk_scale_bias_nchw kernel."""

    def __init__(self, mean, std):
        super().__init__()
        self.register_buffer("inv_std", 1.0 / torch.tensor(std, dtype=torch.float32), persistent=False)
        self.register_buffer("shift", -torch.tensor(mean, dtype=torch.float32) / torch.tensor(std, dtype=torch.float32), persistent=False)

    def forward(self, x):
        R._require_cuda(x, "Normalize6")
        return ChannelAffineFn.apply(x.float().contiguous(), self.inv_std, self.shift, 1.0, 0.0, True).type(x.dtype)


class BcosifyNetwork(BcosUtilMixin, nn.Module):
    """bcosify.py:22-53."""

    def __init__(self, model, model_config, add_channels=True, logit_layer=False):
        super().__init__()
        self.model = model
        self.model_config = model_config
        self.logit_layer = LogitLayer(logit_temperature=None, logit_bias=-math.log(1000 - 1)) if logit_layer else None
        args = model_config["bcosify_args"]
        self.clip_kd = args.get("clip_kd", None)
        self.bfy_mean_zero = model_config.get("bfy_mean_zero", False)
        self.linearprobe_clip = args.get("linearprobe_clip", False)
        if self.clip_kd and self.bfy_mean_zero:
            self.bcosifynormalize = Normalize6(CLIP_MEAN_ZERO, CLIP_STD_ADDINVERSE)
        elif (self.clip_kd or self.linearprobe_clip) and not self.bfy_mean_zero:
            self.bcosifynormalize = Normalize6(CLIP_MEAN_ADDINVERSE, CLIP_STD_ADDINVERSE)
        else:
            self.bcosifynormalize = Normalize6(IMAGENET_MEAN_ADDINVERSE, IMAGENET_STD_ADDINVERSE)
        if add_channels:
            BcosifyNetwork.add_channels(self.model)
        BcosifyNetwork.bcosify(self.model, self.model_config)

    def forward(self, x):
        out = self.model(self.bcosifynormalize(x))
        return self.logit_layer(out) if self.logit_layer else out

    @classmethod
    def add_channels(cls, model):
        """bcosify.py:55-72: 3 -> 6 input channels with cat(W, -W) / 2."""
        found = False
        for module in model.modules():
            if isinstance(module, nn.Conv2d) and module.in_channels == 3:
                if found:
                    warnings.warn("Found multiple layers with 3 input channels. Bcosification might thus not work as intended.")
                found = True
                module.in_channels = 6
                module.weight.data = torch.cat((module.weight.data, -module.weight.data), dim=1) / 2
        if not found:
            warnings.warn("No conv layer with 3 input channels was found although 'add_channels' was set.")

    @classmethod
    def bcosify(cls, model, model_config):
        """bcosify.py:74-113 (CLIP attention pooling is handled by the CLIP plan, not here)."""
        args = model_config.get("bcosify_args", {})
        clip_kd = args.get("clip_kd", False)
        norm_layer = args.get("norm_layer", "BnUncV2")
        gap = args.get("gap", True)
        last = model_config.get("last_layer_name", "NoLastLayerName")
        for n, module in model.named_children():
            if len(list(module.children())) > 0:
                cls.bcosify(module, model_config)
            if isinstance(module, nn.Conv2d):
                setattr(model, n, BcosifyConv2d.from_standard_module(module, model_config))
            elif isinstance(module, nn.Linear) and (n != last or clip_kd or (not gap)):
                if n not in ("k_proj", "v_proj", "q_proj"):
                    setattr(model, n, BcosifyLinear.from_standard_module(module, model_config))
            elif isinstance(module, nn.Linear) and n == last and gap:
                setattr(model, n, BcosifyConv2d.from_standard_module_linear(module, model_config))
            elif isinstance(module, nn.Sequential):
                setattr(model, n, BcosSequential.from_standard_module(module))
            elif isinstance(module, nn.BatchNorm2d) and norm_layer in ("BnUnc2d", "BnUncV2"):
                setattr(model, n, BatchNormUncentered2d.from_standard_module(module, model_config))
            if isinstance(module, nn.ReLU) and not args.get("act_layer", True):
                setattr(model, n, nn.Identity())


def default_config(name: str, last_layer_name: str = "fc", b: float = 2):
    return dict(is_bcos=True, name=name, last_layer_name=last_layer_name, weights=None, bcos_args=dict(b=b, max_out=1),
                bcosify_args=dict(fix_b=True, use_bias=False, norm_layer="BnUncV2", manual_optim=False, gap=True, act_layer=True))


def bcosified_resnet(arch: str = "resnet50") -> BcosifyNetwork:
    """Offline equivalent of bcos/experiments/ImageNet/bcosification/model.py:15-57 (weights=None): B-cosify the
    torchvision ResNet, AvgPool2d(3,2,1) instead of the max pool, strip every bias."""
    kind, layers = {"resnet18": (BasicBlock, [2, 2, 2, 2]), "resnet34": (BasicBlock, [3, 4, 6, 3]),
                    "resnet50": (Bottleneck, [3, 4, 6, 3]), "resnet101": (Bottleneck, [3, 4, 23, 3])}[arch]
    m = BcosifyNetwork(ResNetBcos(kind, layers), default_config(arch), add_channels=True, logit_layer=True)
    m.model.maxpool = nn.AvgPool2d(3, 2, 1)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
    return m


def bcosified_densenet(arch: str = "densenet121") -> BcosifyNetwork:
    """Offline equivalent of the reference's densenet_121 config (experiment_parameters.py:108-129)."""
    growth, blocks, init = {"densenet121": (32, (6, 12, 24, 16), 64)}[arch]
    m = BcosifyNetwork(DenseNetBcos(growth, blocks, init), default_config(arch, last_layer_name="classifier"),
                       add_channels=True, logit_layer=True)
    m.model.features[3] = nn.AvgPool2d(kernel_size=3, stride=2, padding=1)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
    return m
