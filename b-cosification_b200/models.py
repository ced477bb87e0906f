"""Offline descriptions of the B-cosified networks (reference state-dict key names).

The reference's factories download weights (bcos/experiments/ImageNet/bcosification/model.py:15-57); here the
networks are described by their state-dict layout so that released checkpoints (`*.linear.weight`, BN buffers;
reference scripts/strip_checkpoints.py:52-61) or the synthetic checkpoint load directly into the fused plans.
"""
from __future__ import annotations

from typing import Dict, Tuple

from .engine.clip_rn import CLIPResNetPlan
from .engine.densenet import DENSENET_ARCH, DenseNetPlan, dn_names
from .engine.resnet import RESNET_ARCH, ResNetPlan
from .engine.vit import VIT_ARCH, ViTPlan
from .utils import synth


def resnet_state_shapes(arch: str, num_classes: int = 1000) -> Dict[str, Tuple[int, ...]]:
    """Keys/shapes of `BcosifyNetwork(ResNetBcos(...))` (bcosify.py:22-53 over the torchvision ResNet skeleton,
    bcos/models/standard_models.py:36-54) after the factories removed every bias."""
    kind, layers = RESNET_ARCH[arch]
    exp = 1 if kind == "basic" else 4
    shapes: Dict[str, Tuple[int, ...]] = {}

    def bn(prefix, c):
        shapes[prefix + ".weight"] = (c,)
        shapes[prefix + ".running_mean"] = (c,)
        shapes[prefix + ".running_var"] = (c,)
        shapes[prefix + ".num_batches_tracked"] = ()

    shapes["model.conv1.linear.weight"] = (64, 6, 7, 7)
    bn("model.bn1", 64)
    inplanes = 64
    for li, (planes, nblocks) in enumerate(zip([64, 128, 256, 512], layers), start=1):
        for bi in range(nblocks):
            stride = 2 if (li > 1 and bi == 0) else 1
            p = f"model.layer{li}.{bi}"
            if kind == "basic":
                shapes[p + ".conv1.linear.weight"] = (planes, inplanes, 3, 3)
                bn(p + ".bn1", planes)
                shapes[p + ".conv2.linear.weight"] = (planes, planes, 3, 3)
                bn(p + ".bn2", planes)
            else:
                shapes[p + ".conv1.linear.weight"] = (planes, inplanes, 1, 1)
                bn(p + ".bn1", planes)
                shapes[p + ".conv2.linear.weight"] = (planes, planes, 3, 3)
                bn(p + ".bn2", planes)
                shapes[p + ".conv3.linear.weight"] = (planes * 4, planes, 1, 1)
                bn(p + ".bn3", planes * 4)
            if stride != 1 or inplanes != planes * exp:
                shapes[p + ".downsample.0.linear.weight"] = (planes * exp, inplanes, 1, 1)
                bn(p + ".downsample.1", planes * exp)
            inplanes = planes * exp
    shapes["model.fc.linear.weight"] = (num_classes, 512 * exp, 1, 1)
    return shapes


def synthetic_resnet_plan(arch: str, batch: int, **plan_kwargs) -> ResNetPlan:
    """Fused plan over the synthetic (random-init, BN-calibrated) checkpoint - the benchmark workload."""
    sd = synth.synthetic_checkpoint(arch, resnet_state_shapes(arch))
    return ResNetPlan(arch, sd, batch, **plan_kwargs)


def clip_rn_state_shapes(layers=(3, 4, 6, 3), output_dim: int = 1024, width: int = 64) -> Dict[str, Tuple[int, ...]]:
    """Keys/shapes of the B-cos CLIP ResNet image encoder: CLIP/clip/model.py:94-154 `ModifiedResNet` converted by
    bcosify.py:74-113 (`clip_kd`), biases and positional embedding removed (clip_bcosification/model.py:15-23).  The named
    ("-1", "0", "1") downsample Sequential becomes a positional BcosSequential: avg pool 0, conv 1, BN 2."""
    shapes: Dict[str, Tuple[int, ...]] = {}

    def bn(prefix, c):
        shapes[prefix + ".weight"] = (c,)
        shapes[prefix + ".running_mean"] = (c,)
        shapes[prefix + ".running_var"] = (c,)
        shapes[prefix + ".num_batches_tracked"] = ()

    shapes["model.conv1.linear.weight"] = (width // 2, 6, 3, 3); bn("model.bn1", width // 2)
    shapes["model.conv2.linear.weight"] = (width // 2, width // 2, 3, 3); bn("model.bn2", width // 2)
    shapes["model.conv3.linear.weight"] = (width, width // 2, 3, 3); bn("model.bn3", width)
    inpl = width
    for li, (planes, nblocks) in enumerate(zip([width, width * 2, width * 4, width * 8], layers), start=1):
        for bi in range(nblocks):
            stride = 2 if (li > 1 and bi == 0) else 1
            p = f"model.layer{li}.{bi}"
            shapes[p + ".conv1.linear.weight"] = (planes, inpl, 1, 1); bn(p + ".bn1", planes)
            shapes[p + ".conv2.linear.weight"] = (planes, planes, 3, 3); bn(p + ".bn2", planes)
            shapes[p + ".conv3.linear.weight"] = (planes * 4, planes, 1, 1); bn(p + ".bn3", planes * 4)
            if stride > 1 or inpl != planes * 4:
                shapes[p + ".downsample.1.linear.weight"] = (planes * 4, inpl, 1, 1); bn(p + ".downsample.2", planes * 4)
            inpl = planes * 4
    e = width * 32
    for nme in ("k_proj", "q_proj", "v_proj"):
        shapes[f"model.attnpool.{nme}.weight"] = (e, e)
    shapes["model.attnpool.c_proj.linear.weight"] = (output_dim, e)
    return shapes


def synthetic_clip_rn50_plan(batch: int, **plan_kwargs) -> CLIPResNetPlan:
    """Fused CLIP RN50 plan over the synthetic (random-init, BN-calibrated) checkpoint - BASELINE config 4."""
    sd = synth.synthetic_checkpoint("clip_rn50", clip_rn_state_shapes())
    return CLIPResNetPlan(sd, batch, **plan_kwargs)


def vit_state_shapes(arch: str, num_classes: int = 1000, patch: int = 16) -> Dict[str, Tuple[int, ...]]:
    """Keys/shapes of the B-cosified SimpleViT (bcos/models/vit.py:253-339 converted by bcosify_vit.py:45-153, biases stripped by
    the factories vit_bcosification/model.py:20-25)."""
    dim, depth, heads, mlp = VIT_ARCH[arch]
    s: Dict[str, Tuple[int, ...]] = {"model.to_patch_embedding.linear.linear.weight": (dim, patch * patch * 6)}
    for i in range(depth):
        p = f"model.transformer.encoder_{i}"
        s[p + ".attn.norm.weight"] = (dim,)
        s[p + ".attn.to_qkv.weight"] = (3 * dim, dim)
        s[p + ".attn.to_out.linear.weight"] = (dim, dim)
        s[p + ".ff.net.norm.weight"] = (dim,)
        s[p + ".ff.net.linear1.linear.weight"] = (mlp, dim)
        s[p + ".ff.net.linear2.linear.weight"] = (dim, mlp)
    s["model.linear_head.norm.weight"] = (dim,)
    s["model.linear_head.linear.linear.weight"] = (num_classes, dim)
    return s


def synthetic_vit_plan(arch: str, batch: int, **plan_kwargs) -> ViTPlan:
    """Fused SimpleViT plan over the synthetic (random-init) checkpoint - BASELINE config 3."""
    return ViTPlan(arch, synth.synth_state_dict(vit_state_shapes(arch), 0), batch, **plan_kwargs)


def densenet_state_shapes(arch: str, num_classes: int = 1000, bn_size: int = 4) -> Dict[str, Tuple[int, ...]]:
    """Keys/shapes of `BcosifyNetwork(DenseNetBcos(...))` (bcosify.py:22-113 over torchvision's DenseNet, classifier as a 1x1 B-cos
    conv before the global average: bcos/models/standard_models.py:56-63), biases removed by the factories."""
    growth, blocks, init = DENSENET_ARCH[arch]
    shapes: Dict[str, Tuple[int, ...]] = {}

    def bn(prefix, c):
        shapes[prefix + ".weight"] = (c,)
        shapes[prefix + ".running_mean"] = (c,)
        shapes[prefix + ".running_var"] = (c,)
        shapes[prefix + ".num_batches_tracked"] = ()

    nm = dn_names(len(blocks))
    shapes[nm["conv0"] + ".linear.weight"] = (init, 6, 7, 7)
    bn(nm["norm0"], init)
    c = init
    for bi, nlayers in enumerate(blocks, start=1):
        for li in range(1, nlayers + 1):
            p = nm[f"denseblock{bi}"] + f".denselayer{li}"
            bn(p + ".norm1", c)
            shapes[p + ".conv1.linear.weight"] = (bn_size * growth, c, 1, 1)
            bn(p + ".norm2", bn_size * growth)
            shapes[p + ".conv2.linear.weight"] = (growth, bn_size * growth, 3, 3)
            c += growth
        if bi != len(blocks):
            bn(nm[f"transition{bi}.norm"], c)
            shapes[nm[f"transition{bi}.conv"] + ".linear.weight"] = (c // 2, c, 1, 1)
            c //= 2
    bn(nm["norm5"], c)
    shapes["model.classifier.linear.weight"] = (num_classes, c, 1, 1)
    return shapes


def synthetic_densenet_plan(arch: str, batch: int, **plan_kwargs) -> DenseNetPlan:
    """Fused DenseNet plan over the synthetic (random-init, BN-calibrated) checkpoint."""
    sd = synth.synthetic_checkpoint(arch, densenet_state_shapes(arch))
    return DenseNetPlan(arch, sd, batch, **plan_kwargs)


def clip_vit_state_shapes(input_resolution: int = 224, patch: int = 32, width: int = 768, layers: int = 12,
                          output_dim: int = 512) -> Dict[str, Tuple[int, ...]]:
    """Keys/shapes of the B-cos CLIP ViT image encoder: CLIP/clip/model.py:206-241 `VisionTransformer` converted by bcosify.py:74-113
    (`clip_kd`), `.bias` attributes and the positional embedding stripped (clip_bcosification/model.py:17-25; `in_proj_bias` is not a
    `.bias` attribute and survives).  The named mlp Sequential becomes a positional BcosSequential: c_fc 0, gelu 1, c_proj 2."""
    s: Dict[str, Tuple[int, ...]] = {"model.class_embedding": (width,), "model.proj": (width, output_dim),
                                     "model.conv1.linear.weight": (width, 6, patch, patch), "model.ln_pre.weight": (width,),
                                     "model.ln_post.weight": (width,)}
    for i in range(layers):
        p = f"model.transformer.resblocks.{i}"
        s[p + ".attn.in_proj_weight"] = (3 * width, width)
        s[p + ".attn.in_proj_bias"] = (3 * width,)
        s[p + ".attn.out_proj.linear.weight"] = (width, width)
        s[p + ".ln_1.weight"] = (width,)
        s[p + ".ln_2.weight"] = (width,)
        s[p + ".mlp.0.linear.weight"] = (4 * width, width)
        s[p + ".mlp.2.linear.weight"] = (width, 4 * width)
    return s
