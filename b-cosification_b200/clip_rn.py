"""B-cos CLIP ResNet image encoder on the CUDA-backed modules -- mirror of CLIP/clip/model.py:10-154 (`Bottleneck`,
`ModifiedResNet`: 3-conv stem, anti-aliasing average pools, attention pooling) as converted by the reference's
bcosify.py with `clip_kd` (CLIP normalisation constants, no LogitLayer; clip_bcosification/model.py:8-25)."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from .bcosify import BcosifyNetwork
from .modules import BcosifyLinear
from .modules.bcosattnpool import BcosAttentionPool2d


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu2 = nn.ReLU(inplace=True)
        self.avgpool = nn.AvgPool2d(stride) if stride > 1 else nn.Identity()
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion)
        self.relu3 = nn.ReLU(inplace=True)
        self.downsample = None
        self.stride = stride
        if stride > 1 or inplanes != planes * Bottleneck.expansion:
            self.downsample = nn.Sequential(OrderedDict([
                ("-1", nn.AvgPool2d(stride)),
                ("0", nn.Conv2d(inplanes, planes * self.expansion, 1, stride=1, bias=False)),
                ("1", nn.BatchNorm2d(planes * self.expansion))]))

    def forward(self, x):
        identity = x
        out = self.relu1(self.bn1(self.conv1(x)))
        out = self.relu2(self.bn2(self.conv2(out)))
        out = self.avgpool(out)
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu3(out + identity)


class ModifiedResNet(nn.Module):
    def __init__(self, layers, output_dim, heads, input_resolution=224, width=64, attn_unpool=False):
        super().__init__()
        self.output_dim, self.input_resolution = output_dim, input_resolution
        self.conv1 = nn.Conv2d(3, width // 2, kernel_size=3, stride=2, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(width // 2)
        self.relu1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(width // 2, width // 2, kernel_size=3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(width // 2)
        self.relu2 = nn.ReLU(inplace=True)
        self.conv3 = nn.Conv2d(width // 2, width, kernel_size=3, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(width)
        self.relu3 = nn.ReLU(inplace=True)
        self.avgpool = nn.AvgPool2d(2)
        self._inplanes = width
        self.layer1 = self._make_layer(width, layers[0])
        self.layer2 = self._make_layer(width * 2, layers[1], stride=2)
        self.layer3 = self._make_layer(width * 4, layers[2], stride=2)
        self.layer4 = self._make_layer(width * 8, layers[3], stride=2)
        self.attnpool = BcosAttentionPool2d(input_resolution // 32, width * 32, heads, output_dim, attn_unpool)

    def _make_layer(self, planes, blocks, stride=1):
        layers = [Bottleneck(self._inplanes, planes, stride)]
        self._inplanes = planes * Bottleneck.expansion
        for _ in range(1, blocks):
            layers.append(Bottleneck(self._inplanes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.relu1(self.bn1(self.conv1(x)))
        x = self.relu2(self.bn2(self.conv2(x)))
        x = self.relu3(self.bn3(self.conv3(x)))
        x = self.avgpool(x)
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.attnpool(x)


def bcosified_clip_rn50(attn_unpool: bool = False) -> BcosifyNetwork:
    """Offline equivalent of clip_bcosification/model.py:8-25 for `resnet_50_clip_b2_noBias...` (random init).
    `attn_unpool=True` is the `model_config["attn_unpool"]` variant (bcosattnpool.py:65): per-token unit embeddings
    (HW) x N x D' instead of the pooled one, used by the text-localisation analysis."""
    cfg = dict(is_bcos=True, name="resnet50clip", bcos_args=dict(b=2, max_out=1),
               bcosify_args=dict(clip_kd=True, fix_b=True, norm_layer="BnUncV2", use_bias=False))
    m = BcosifyNetwork(ModifiedResNet((3, 4, 6, 3), 1024, 32, 224, 64, attn_unpool), cfg, add_channels=True, logit_layer=False)
    # bcosify.py:81-83,97: inside the attention pool only c_proj becomes a BcosifyLinear object (its weight is what is used)
    ap = m.model.attnpool
    if isinstance(ap.c_proj, nn.Linear):
        ap.c_proj = BcosifyLinear.from_standard_module(ap.c_proj, cfg)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
        if hasattr(mod, "positional_embedding") and mod.positional_embedding is not None:
            mod.positional_embedding = None
    return m
