"""CPU oracle for the B-cos forward + dynamic-linear explanation path.

TEST INFRASTRUCTURE.  This file restates, in plain fp32 PyTorch on the CPU, the arithmetic of the
reference's hot path (shrebox/B-cosification).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it - it is the checker, never
the product: nothing under `b-cosification_b200/` imports this module and the product path raises
when its CUDA library is missing.

Parity status: the reference ships NO tests or golden vectors for this path, so
the oracle is pinned against the reference *itself*: `oracle/make_golden.py` imports the reference
modules read-only in the build container (oracle/refload.py), feeds both the same synthetic
state-dict and images, asserts agreement and writes `tests/golden/*.npz`.  `tests/test_oracle_*.py`
re-check the oracle against those committed vectors everywhere and against the live reference
where /root/reference exists.

Every function cites the reference lines it follows (paths relative to the reference root).
The arithmetic lives in ATen (torch==2.2.1 pinned by the reference, requirements.txt:83); the
container's torch 2.11 runs the same calls.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

IMAGENET_MEAN_ADDINVERSE = (0.485, 0.456, 0.406, 0.515, 0.544, 0.594)  # bcosify.py:14
IMAGENET_STD_ADDINVERSE = (0.229, 0.224, 0.225, 0.229, 0.224, 0.225)  # bcosify.py:15
CLIP_MEAN_ADDINVERSE = (0.48145466, 0.4578275, 0.40821073, 0.51854534, 0.5421725, 0.59178927)  # bcosify.py:17
CLIP_STD_ADDINVERSE = (0.26862954, 0.26130258, 0.27577711, 0.26862954, 0.26130258, 0.27577711)  # bcosify.py:19
LOGIT_BIAS_1000 = -math.log(1000 - 1)  # bcosify.py:31


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


# ----------------------------------------------------------------------------------------------
# B-cos conv / linear  (bcos/modules/bcosconv2d.py, bcosifyconv2d.py, bcoslinear.py, bcosifylinear.py)
# ----------------------------------------------------------------------------------------------
def patch_norms(x: Tensor, kernel_size, stride, padding, groups: int = 1, out_channels: Optional[int] = None) -> Tensor:
    """bcosconv2d.py:196-231 `calc_patch_norms`: sqrt(sumpool_k,s,p(sum_c x^2) + 1e-6)."""
    sq = x * x
    if groups == 1:
        sq = sq.sum(1, keepdim=True)
    else:
        sq = sq.unflatten(1, (groups, x.shape[1] // groups)).sum(2)
    n = (F.avg_pool2d(sq, _pair(kernel_size), padding=_pair(padding), stride=_pair(stride), divisor_override=1) + 1e-6).sqrt()
    if groups > 1:
        n = torch.repeat_interleave(n, repeats=out_channels // groups, dim=1)
    return n


def patch_norms_slow(x: Tensor, weight: Tensor, stride, padding, dilation, groups) -> Tensor:
    """bcosconv2d.py:233-250 `_calc_patch_norms_slow` (ones-kernel conv; the in-code cross-check)."""
    return (F.conv2d(x * x, torch.ones_like(weight), None, stride, padding, dilation, groups) + 1e-6).sqrt()


def normed_weight(weight: Tensor) -> Tensor:
    """bcosconv2d.py:26-35 `NormedConv2d` / bcoslinear.py:20-27 `NormedLinear`: unit L2 norm per output unit."""
    dims = tuple(range(1, weight.dim()))
    return weight / torch.linalg.vector_norm(weight, dim=dims, keepdim=True)


def _maxout(out: Tensor, max_out: int, dim: int) -> Tensor:
    """bcosconv2d.py:166-170: channel c = o*M + m, max over m."""
    if max_out > 1:
        out = out.unflatten(dim, (out.shape[dim] // max_out, max_out)).max(dim=dim + 1 if dim >= 0 else dim).values
    return out


def _bcos_scale(lin: Tensor, norm: Tensor, b: float, detach: bool) -> Tensor:
    """bcosconv2d.py:181-193: b==2 -> |lin|/norm ; else (|lin/norm| + 1e-6)^(b-1); detached in explanation mode."""
    l, n = (lin.detach(), norm.detach()) if detach else (lin, norm)
    if b == 2:
        return l.abs() / n
    return ((l / n).abs() + 1e-6).pow(b - 1)


def bcos_conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride=1, padding=0, dilation=1,
                groups: int = 1, b: float = 2, max_out: int = 1, detach: bool = False,
                normalize_weight: bool = False) -> Tensor:
    """`BcosConv2d.forward_impl` bcosconv2d.py:153-194 (normalize_weight=True, NormedConv2d) and
    `BcosifyConv2d.forward_impl` bcosifyconv2d.py:50-102 (normalize_weight=False, plain nn.Conv2d)."""
    w = normed_weight(weight) if normalize_weight else weight
    lin = F.conv2d(x, w, bias, _pair(stride), _pair(padding), _pair(dilation), groups)
    lin = _maxout(lin, max_out, 1)
    if b == 1:
        return lin
    if _pair(dilation) != (1, 1):
        assert max_out == 1
        norm = patch_norms_slow(x, weight, stride, padding, dilation, groups)
    else:
        norm = patch_norms(x, weight.shape[2:], stride, padding, groups, lin.shape[1])
    return _bcos_scale(lin, norm, b, detach) * lin


def bcos_linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, b: float = 2, max_out: int = 1,
                detach: bool = False, normalize_weight: bool = False) -> Tensor:
    """`BcosLinear.forward` bcoslinear.py:88-130 / `BcosifyLinear.forward` bcosifylinear.py:42-95.
    NB the eps sits OUTSIDE the sqrt here: norm = ||x||_2 + 1e-12 (bcoslinear.py:113)."""
    w = normed_weight(weight) if normalize_weight else weight
    lin = F.linear(x, w, bias)
    lin = _maxout(lin, max_out, -1) if max_out > 1 else lin
    if b == 1:
        return lin
    norm = torch.linalg.vector_norm(x, dim=-1, keepdim=True) + 1e-12
    return _bcos_scale(lin, norm, b, detach) * lin


# ----------------------------------------------------------------------------------------------
# norms / small layers
# ----------------------------------------------------------------------------------------------
def batch_norm_uncentered_2d(x: Tensor, running_var: Optional[Tensor], weight: Optional[Tensor] = None,
                             bias: Optional[Tensor] = None, training: bool = False, momentum: float = 0.1,
                             eps: float = 1e-5, detach: bool = False) -> Tensor:
    """bcos/modules/norms/uncentered_norms/batchnorm_uncentered.py:21-60."""
    if training:
        xs = x.detach() if detach else x
        var = xs.var(dim=(0, 2, 3), unbiased=False)  # centred, biased variance of an uncentred signal (:39)
        if running_var is not None:
            running_var.copy_((1 - momentum) * running_var + momentum * var.detach())  # (:43)
    else:
        var = running_var
    y = x / (var + eps).sqrt()[None, :, None, None]
    if weight is not None:
        y = weight[None, :, None, None] * y
    if bias is not None:
        y = y + bias[None, :, None, None]
    return y.type(x.dtype)


def bn_uncentered_from_standard(weight, bias, running_mean, running_var, eps) -> Tuple[Tensor, Optional[Tensor]]:
    """`BatchNormUncentered2d.from_standard_module` batchnorm_uncentered.py:118-141 ("BnUncV2" fold)."""
    if bias is None:
        return weight, None
    std = (running_var + eps).sqrt()
    return weight, bias - (running_mean / std) * weight


def layer_norm_detachable(x: Tensor, weight: Optional[Tensor], bias: Optional[Tensor], eps: float = 1e-5,
                          detach: bool = False) -> Tensor:
    """`DetachableLayerNorm.forward` bcos/modules/norms/centered_norms.py:187-224: mean stays in-graph,
    only the variance is detached in explanation mode; otherwise stock F.layer_norm."""
    if not detach:
        return F.layer_norm(x, x.shape[-1:], weight, bias, eps)
    var, mean = torch.var_mean(x, dim=-1, unbiased=False, keepdim=True)
    std = (var + eps).sqrt().detach()
    y = (x - mean) / std
    if weight is not None:
        y = weight * y
    if bias is not None:
        y = y + bias
    return y


def group_norm_detachable(x: Tensor, num_groups: int, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None,
                          eps: float = 1e-5, detach: bool = False, centred: bool = False) -> Tensor:
    """centred=False: `group_norm_uncentered` bcos/modules/norms/uncentered_norms/groupnorm_uncentered.py:21-61 (x / std,
    std from the CENTRED biased variance, taken from x.detach() in explanation mode).  centred=True:
    `DetachableGroupNorm2d.forward` bcos/modules/norms/centered_norms.py:93-138 ((x - mean) / std, only the variance
    detached; stock F.group_norm when not detaching).  num_groups = 1 / C are the GN-LayerNorm / GN-InstanceNorm wrappers."""
    n, c = x.shape[:2]
    assert c % num_groups == 0
    if centred and not detach:
        return F.group_norm(x, num_groups, weight, bias, eps)
    xg = x.reshape(n, num_groups, c // num_groups, *x.shape[2:])
    dims = tuple(range(2, xg.dim()))
    if centred:
        var, mean = torch.var_mean(xg, dim=dims, unbiased=False, keepdim=True)
        y = (xg - mean) / (var.detach() + eps).sqrt()
    else:
        var = (xg.detach() if detach else xg).var(dim=dims, unbiased=False, keepdim=True)
        y = xg / (var + eps).sqrt()
    y = y.reshape(x.shape)
    if weight is not None:
        y = weight[None, :, None, None] * y
    if bias is not None:
        y = y + bias[None, :, None, None]
    return y


def position_norm_detachable(x: Tensor, weight: Optional[Tensor] = None, bias: Optional[Tensor] = None, eps: float = 1e-5,
                             detach: bool = False, centred: bool = False) -> Tensor:
    """centred=False: `PositionNormUncentered2d.forward` bcos/modules/norms/uncentered_norms/posnorm_uncentered.py:39-58;
    centred=True: `DetachablePositionNorm2d.forward` bcos/modules/norms/centered_norms.py:251-297 (channel-wise layer norm
    per pixel; the variance is detached in explanation mode, the mean is not)."""
    assert x.dim() == 4
    if centred and not detach:
        return F.layer_norm(x.permute(0, 2, 3, 1), x.shape[1:2], weight, bias, eps).permute(0, 3, 1, 2)
    var, mean = torch.var_mean(x, dim=1, unbiased=False, keepdim=True)
    if detach:
        var = var.detach()
    y = ((x - mean) if centred else x) / (var + eps).sqrt()
    if weight is not None:
        y = weight[None, :, None, None] * y
    if bias is not None:
        y = y + bias[None, :, None, None]
    return y


def all_norm_uncentered_2d(x: Tensor, running_var: Optional[Tensor], weight: Optional[Tensor] = None,
                           bias: Optional[Tensor] = None, training: bool = False, momentum: float = 0.1, eps: float = 1e-5,
                           detach: bool = False) -> Tensor:
    """bcos/modules/norms/uncentered_norms/allnorm_uncentered.py:21-61: one variance for the whole batch tensor."""
    if training:
        var = (x.detach() if detach else x).var(unbiased=False)
        if running_var is not None:
            running_var.copy_((1 - momentum) * running_var + momentum * var.detach())
    else:
        var = running_var
    y = x / (var + eps).sqrt()[None, ..., None, None]
    if weight is not None:
        y = weight[None, ..., None, None] * y
    if bias is not None:
        y = y + bias[None, ..., None, None]
    return y


def gelu_detachable(x: Tensor, detach: bool = False) -> Tensor:
    """`MyGELU` bcosify_vit.py:27-32: gate = 0.5*(1+erf(x/sqrt2)) (detached in explanation mode) * x."""
    gate = 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    if detach:
        gate = gate.detach()
    return gate * x


def logit_layer(x: Tensor, logit_temperature: Optional[float] = None, logit_bias: Optional[float] = None) -> Tensor:
    """bcos/modules/logitlayer.py:22-27."""
    if logit_temperature is not None:
        x = x / logit_temperature
    if logit_bias is not None:
        x = x + logit_bias
    return x


def normalize6(x: Tensor, mean: Sequence[float] = IMAGENET_MEAN_ADDINVERSE,
               std: Sequence[float] = IMAGENET_STD_ADDINVERSE) -> Tensor:
    """`BcosifyNetwork.forward` bcosify.py:50-53 -> torchvision Normalize on the 6-channel [x, 1-x] input."""
    m = torch.tensor(mean, dtype=x.dtype).view(1, -1, 1, 1)
    s = torch.tensor(std, dtype=x.dtype).view(1, -1, 1, 1)
    return (x - m) / s


def add_channels_conv(weight3: Tensor) -> Tensor:
    """`BcosifyNetwork.add_channels` bcosify.py:55-72: stem weight cat(W, -W)/2 along C_in."""
    return torch.cat((weight3, -weight3), dim=1) / 2


# ----------------------------------------------------------------------------------------------
# B-cosified ResNet (torchvision skeleton; bcos/models/standard_models.py:36-54 ResNetBcos:
# classifier applied per position BEFORE global average pooling; maxpool -> AvgPool2d(3,2,1),
# bcos/experiments/ImageNet/bcosification/model.py:47-49; all biases None :53-55)
# ----------------------------------------------------------------------------------------------
RESNET_ARCH = {
    "resnet18": ("basic", [2, 2, 2, 2]),
    "resnet34": ("basic", [3, 4, 6, 3]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]),
    "resnet101": ("bottleneck", [3, 4, 23, 3]),
}


def resnet_state_shapes(arch: str, num_classes: int = 1000) -> Dict[str, Tuple[int, ...]]:
    """State-dict keys/shapes of `BcosifyNetwork(ResNetBcos(...))` after biases are stripped (probe: SURVEY.md section 5)."""
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


class OracleResNet:
    """Functional B-cosified ResNet over a reference-keyed state dict."""

    def __init__(self, arch: str, sd: Dict[str, Tensor], b: float = 2, eps: float = 1e-5,
                 mean=IMAGENET_MEAN_ADDINVERSE, std=IMAGENET_STD_ADDINVERSE, logit_bias: Optional[float] = LOGIT_BIAS_1000):
        self.arch, self.sd, self.b, self.eps = arch, sd, b, eps
        self.kind, self.layers = RESNET_ARCH[arch]
        self.mean, self.std, self.logit_bias = mean, std, logit_bias
        self.training = False  # True => BN uses batch statistics and updates running_var
        self.momentum = 0.1
        self.taps: Optional[Dict[str, Tensor]] = None  # set to {} to record intermediate activations

    # --- building blocks -------------------------------------------------------------------
    def _conv(self, name, x, stride, padding, detach):
        return bcos_conv2d(x, self.sd[name + ".linear.weight"], self.sd.get(name + ".linear.bias"), stride, padding,
                           b=self.b, detach=detach)

    def _bn(self, name, x, detach):
        return batch_norm_uncentered_2d(x, self.sd[name + ".running_var"], self.sd.get(name + ".weight"),
                                        self.sd.get(name + ".bias"), self.training, self.momentum, self.eps, detach)

    def _tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t.detach()

    def _block(self, p, x, stride, detach):
        identity = x
        if self.kind == "basic":  # torchvision BasicBlock: stride on conv1
            out = F.relu(self._bn(p + ".bn1", self._conv(p + ".conv1", x, stride, 1, detach), detach))
            out = self._bn(p + ".bn2", self._conv(p + ".conv2", out, 1, 1, detach), detach)
        else:  # torchvision Bottleneck v1.5: stride on the 3x3
            out = F.relu(self._bn(p + ".bn1", self._conv(p + ".conv1", x, 1, 0, detach), detach))
            out = F.relu(self._bn(p + ".bn2", self._conv(p + ".conv2", out, stride, 1, detach), detach))
            out = self._bn(p + ".bn3", self._conv(p + ".conv3", out, 1, 0, detach), detach)
        if (p + ".downsample.0.linear.weight") in self.sd:
            identity = self._bn(p + ".downsample.1", self._conv(p + ".downsample.0", x, stride, 0, detach), detach)
        out = F.relu(out + identity)
        self._tap(p, out)
        return out

    def forward(self, x6: Tensor, detach: bool = False) -> Tensor:
        """x6: [B,6,H,W] un-normalised `[x, 1-x]`.  Returns logits [B, classes]."""
        x = normalize6(x6, self.mean, self.std)
        x = F.relu(self._bn("model.bn1", self._conv("model.conv1", x, 2, 3, detach), detach))
        self._tap("stem", x)
        x = F.avg_pool2d(x, 3, 2, 1)
        for li, nblocks in enumerate(self.layers, start=1):
            for bi in range(nblocks):
                x = self._block(f"model.layer{li}.{bi}", x, 2 if (li > 1 and bi == 0) else 1, detach)
        x = self._conv("model.fc", x, 1, 0, detach)  # classifier before GAP (standard_models.py:50-52)
        self._tap("fc", x)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return logit_layer(x, None, self.logit_bias)

    __call__ = forward

    def calibrate_bn(self, x6: Tensor) -> None:
        """one train-mode forward with momentum=1.0 so running_var := batch variance."""
        self.training, self.momentum = True, 1.0
        with torch.no_grad():
            self.forward(x6)
        self.training, self.momentum = False, 0.1


# ----------------------------------------------------------------------------------------------
# B-cosified DenseNet (torchvision skeleton; bcos/models/standard_models.py:56-63 DenseNetBcos: classifier (1x1 B-cos
# conv) before global average pooling; features[3] (pool0) -> AvgPool2d(3,2,1), experiment_parameters.py:108-129)
# ----------------------------------------------------------------------------------------------
DENSENET_ARCH = {"densenet121": (32, (6, 12, 24, 16), 64), "densenet169": (32, (6, 12, 32, 32), 64),
                 "densenet201": (32, (6, 12, 48, 32), 64)}     # torchvision densenet.py: growth rate, block config, initial features


def _dn_names(nblocks: int):
    """State-dict prefixes: `BcosSequential.from_standard_module` (bcos/modules/common.py:46-51) rebuilds every
    nn.Sequential positionally, so `features` and the transitions lose their child names (features.conv0 ->
    features.0, transitionK.conv -> features.N.2); dense blocks are ModuleDicts and keep `denselayerL`."""
    f = "model.features"
    names = {"conv0": f + ".0", "norm0": f + ".1", "norm5": f + f".{3 + 2 * nblocks}"}
    for bi in range(1, nblocks + 1):
        names[f"denseblock{bi}"] = f + f".{2 + 2 * bi}"
        names[f"transition{bi}.norm"] = f + f".{3 + 2 * bi}.0"
        names[f"transition{bi}.conv"] = f + f".{3 + 2 * bi}.2"
    return names


def densenet_state_shapes(arch: str, num_classes: int = 1000, bn_size: int = 4) -> Dict[str, Tuple[int, ...]]:
    growth, blocks, init = DENSENET_ARCH[arch]
    shapes: Dict[str, Tuple[int, ...]] = {}

    def bn(prefix, c):
        shapes[prefix + ".weight"] = (c,)
        shapes[prefix + ".running_mean"] = (c,)
        shapes[prefix + ".running_var"] = (c,)
        shapes[prefix + ".num_batches_tracked"] = ()

    nm = _dn_names(len(blocks))
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


class OracleDenseNet:
    """Functional B-cosified DenseNet over a reference-keyed state dict (torchvision densenet.py layer order:
    norm -> relu -> conv inside dense layers and transitions)."""

    def __init__(self, arch: str, sd: Dict[str, Tensor], b: float = 2, eps: float = 1e-5,
                 mean=IMAGENET_MEAN_ADDINVERSE, std=IMAGENET_STD_ADDINVERSE, logit_bias: Optional[float] = LOGIT_BIAS_1000):
        self.arch, self.sd, self.b, self.eps = arch, sd, b, eps
        self.growth, self.blocks, self.init = DENSENET_ARCH[arch]
        self.mean, self.std, self.logit_bias = mean, std, logit_bias
        self.training, self.momentum = False, 0.1

    def _conv(self, name, x, stride, padding, detach):
        return bcos_conv2d(x, self.sd[name + ".linear.weight"], None, stride, padding, b=self.b, detach=detach)

    def _bn(self, name, x, detach):
        return batch_norm_uncentered_2d(x, self.sd[name + ".running_var"], self.sd.get(name + ".weight"),
                                        self.sd.get(name + ".bias"), self.training, self.momentum, self.eps, detach)

    def forward(self, x6: Tensor, detach: bool = False) -> Tensor:
        nm = _dn_names(len(self.blocks))
        x = normalize6(x6, self.mean, self.std)
        x = F.relu(self._bn(nm["norm0"], self._conv(nm["conv0"], x, 2, 3, detach), detach))
        x = F.avg_pool2d(x, 3, 2, 1)
        for bi, nlayers in enumerate(self.blocks, start=1):
            feats = [x]
            for li in range(1, nlayers + 1):
                p = nm[f"denseblock{bi}"] + f".denselayer{li}"
                cat = torch.cat(feats, 1)
                h = self._conv(p + ".conv1", F.relu(self._bn(p + ".norm1", cat, detach)), 1, 0, detach)
                h = self._conv(p + ".conv2", F.relu(self._bn(p + ".norm2", h, detach)), 1, 1, detach)
                feats.append(h)
            x = torch.cat(feats, 1)
            if bi != len(self.blocks):
                x = self._conv(nm[f"transition{bi}.conv"], F.relu(self._bn(nm[f"transition{bi}.norm"], x, detach)), 1, 0,
                               detach)
                x = F.avg_pool2d(x, 2, 2)
        x = F.relu(self._bn(nm["norm5"], x, detach))
        x = self._conv("model.classifier", x, 1, 0, detach)       # classifier before GAP (standard_models.py:59-62)
        x = F.adaptive_avg_pool2d(x, 1).flatten(1)
        return logit_layer(x, None, self.logit_bias)

    __call__ = forward

    def calibrate_bn(self, x6: Tensor) -> None:
        self.training, self.momentum = True, 1.0
        with torch.no_grad():
            self.forward(x6)
        self.training, self.momentum = False, 0.1


# ----------------------------------------------------------------------------------------------
# B-cosified SimpleViT (bcos/models/vit.py:232-339 converted by bcosify_vit.py:45-153; gap_reorder = True:
# the B-cos head is applied per token before the mean, vit.py:331-334)
# ----------------------------------------------------------------------------------------------
VIT_ARCH = {  # name: (dim, depth, heads, mlp_dim)   vit.py:441-467
    "simple_vit_ti_patch16_224": (192, 12, 3, 768),
    "simple_vit_s_patch16_224": (384, 12, 6, 1536),
    "simple_vit_b_patch16_224": (768, 12, 12, 3072),
}


def vit_state_shapes(arch: str, num_classes: int = 1000, patch: int = 16) -> Dict[str, Tuple[int, ...]]:
    dim, depth, heads, mlp = VIT_ARCH[arch]
    sh: Dict[str, Tuple[int, ...]] = {"model.to_patch_embedding.linear.linear.weight": (dim, patch * patch * 6)}
    for i in range(depth):
        p = f"model.transformer.encoder_{i}"
        sh[p + ".attn.norm.weight"] = (dim,)
        sh[p + ".attn.to_qkv.weight"] = (3 * dim, dim)
        sh[p + ".attn.to_out.linear.weight"] = (dim, dim)
        sh[p + ".ff.net.norm.weight"] = (dim,)
        sh[p + ".ff.net.linear1.linear.weight"] = (mlp, dim)
        sh[p + ".ff.net.linear2.linear.weight"] = (dim, mlp)
    sh["model.linear_head.norm.weight"] = (dim,)
    sh["model.linear_head.linear.linear.weight"] = (num_classes, dim)
    return sh


def posemb_sincos_2d(h: int, w: int, dim: int, temperature: float = 10000.0) -> Tensor:
    """`PosEmbSinCos2d.forward` vit.py:64-86."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1)


def vit_add_channels_linear(w3: Tensor) -> Tensor:
    """`BcosifyNetwork.add_channels` bcosify_vit.py:84-125: per pixel [W/2, -W/2] interleaved on the (p1 p2 c) axis."""
    out_f = w3.shape[0]
    wr = w3.view(out_f, -1, 3) / 2
    return torch.cat([wr, -wr], dim=2).reshape(out_f, -1)


class OracleViT:
    def __init__(self, arch: str, sd: Dict[str, Tensor], b: float = 2, patch: int = 16, eps: float = 1e-5,
                 mean=IMAGENET_MEAN_ADDINVERSE, std=IMAGENET_STD_ADDINVERSE, logit_bias: Optional[float] = LOGIT_BIAS_1000,
                 logit_temperature: Optional[float] = None):
        self.arch, self.sd, self.b, self.patch, self.eps = arch, sd, b, patch, eps
        self.dim, self.depth, self.heads, self.mlp = VIT_ARCH[arch]
        self.mean, self.std, self.logit_bias, self.logit_temperature = mean, std, logit_bias, logit_temperature

    def _lin(self, name, x, detach):
        return bcos_linear(x, self.sd[name + ".linear.weight"], None, b=self.b, detach=detach)

    def _ln(self, name, x, detach):
        return layer_norm_detachable(x, self.sd[name + ".weight"], self.sd.get(name + ".bias"), self.eps, detach)

    def forward(self, x6: Tensor, detach: bool = False) -> Tensor:
        x = normalize6(x6, self.mean, self.std)
        B, C, H, W = x.shape
        p = self.patch
        hh, ww = H // p, W // p
        # Rearrange "b c (h p1) (w p2) -> b h w (p1 p2 c)"  (vit.py:290-294)
        x = x.view(B, C, hh, p, ww, p).permute(0, 2, 4, 3, 5, 1).reshape(B, hh, ww, p * p * C)
        x = self._lin("model.to_patch_embedding.linear", x, detach)
        x = x.reshape(B, hh * ww, self.dim) + posemb_sincos_2d(hh, ww, self.dim).to(x.dtype)       # vit.py:325-326
        dh = self.dim // self.heads
        for i in range(self.depth):
            pfx = f"model.transformer.encoder_{i}"
            h = self._ln(pfx + ".attn.norm", x, detach)
            qkv = F.linear(h, self.sd[pfx + ".attn.to_qkv.weight"]).chunk(3, dim=-1)               # plain linear, vit.py:140
            q, k, v = (t.view(B, -1, self.heads, dh).transpose(1, 2) for t in qkv)
            if detach:
                q, k = q.detach(), k.detach()                                                      # vit.py:148-150
            attn = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * dh ** -0.5, dim=-1)
            o = torch.matmul(attn, v).transpose(1, 2).reshape(B, -1, self.dim)
            x = self._lin(pfx + ".attn.to_out", o, detach) + x
            h = self._ln(pfx + ".ff.net.norm", x, detach)
            h = self._lin(pfx + ".ff.net.linear1", h, detach)
            h = gelu_detachable(h, detach)
            x = self._lin(pfx + ".ff.net.linear2", h, detach) + x
        x = self._lin("model.linear_head.linear", self._ln("model.linear_head.norm", x, detach), detach)   # gap_reorder
        x = x.mean(dim=1)
        return logit_layer(x, self.logit_temperature, self.logit_bias)

    __call__ = forward


# ----------------------------------------------------------------------------------------------
# B-cos CLIP RN50 image encoder: CLIP/clip/model.py:94-154 `ModifiedResNet` (3-conv stem, anti-aliasing avg pools,
# CLIP/clip/model.py:10-55 `Bottleneck`) converted by bcosify.py with clip_kd (CLIP mean/std, no LogitLayer) and
# `BcosAttentionPool2d` (bcos/modules/bcosattnpool.py:22-59) without positional embedding or biases
# (bcos/experiments/ImageNet/clip_bcosification/model.py:15-23)
# ----------------------------------------------------------------------------------------------
def clip_rn_state_shapes(layers=(3, 4, 6, 3), output_dim: int = 1024, width: int = 64,
                         attn_unpool: bool = False) -> Dict[str, Tuple[int, ...]]:
    sh: Dict[str, Tuple[int, ...]] = {}

    def bn(prefix, c):
        sh[prefix + ".weight"] = (c,)
        sh[prefix + ".running_mean"] = (c,)
        sh[prefix + ".running_var"] = (c,)
        sh[prefix + ".num_batches_tracked"] = ()

    sh["model.conv1.linear.weight"] = (width // 2, 6, 3, 3); bn("model.bn1", width // 2)
    sh["model.conv2.linear.weight"] = (width // 2, width // 2, 3, 3); bn("model.bn2", width // 2)
    sh["model.conv3.linear.weight"] = (width, width // 2, 3, 3); bn("model.bn3", width)
    inpl = width
    for li, (planes, nb) in enumerate(zip([width, width * 2, width * 4, width * 8], layers), start=1):
        for bi in range(nb):
            stride = 2 if (li > 1 and bi == 0) else 1
            p = f"model.layer{li}.{bi}"
            sh[p + ".conv1.linear.weight"] = (planes, inpl, 1, 1); bn(p + ".bn1", planes)
            sh[p + ".conv2.linear.weight"] = (planes, planes, 3, 3); bn(p + ".bn2", planes)
            sh[p + ".conv3.linear.weight"] = (planes * 4, planes, 1, 1); bn(p + ".bn3", planes * 4)
            if stride > 1 or inpl != planes * 4:
                # BcosSequential rebuilds the ("-1","0","1") Sequential positionally: avgpool=0, conv=1, bn=2
                sh[p + ".downsample.1.linear.weight"] = (planes * 4, inpl, 1, 1); bn(p + ".downsample.2", planes * 4)
            inpl = planes * 4
    e = width * 32
    for nme in (("v_proj",) if attn_unpool else ("k_proj", "q_proj", "v_proj")):
        sh[f"model.attnpool.{nme}.weight"] = (e, e)
    sh["model.attnpool.c_proj.linear.weight"] = (output_dim, e)
    return sh


class OracleCLIPResNet:
    def __init__(self, sd: Dict[str, Tensor], layers=(3, 4, 6, 3), heads: int = 32, b: float = 2, eps: float = 1e-5,
                 mean=CLIP_MEAN_ADDINVERSE, std=CLIP_STD_ADDINVERSE):
        self.sd, self.layers, self.heads, self.b, self.eps, self.mean, self.std = sd, layers, heads, b, eps, mean, std
        self.training, self.momentum = False, 0.1
        self.unpool = "model.attnpool.q_proj.weight" not in sd        # attn_unpool variant has no q/k projections (:15-17)

    def _conv(self, name, x, stride, padding, detach):
        return bcos_conv2d(x, self.sd[name + ".linear.weight"], None, stride, padding, b=self.b, detach=detach)

    def _bn(self, name, x, detach):
        return batch_norm_uncentered_2d(x, self.sd[name + ".running_var"], self.sd.get(name + ".weight"),
                                        self.sd.get(name + ".bias"), self.training, self.momentum, self.eps, detach)

    def _block(self, p, x, stride, detach):
        out = F.relu(self._bn(p + ".bn1", self._conv(p + ".conv1", x, 1, 0, detach), detach))
        out = F.relu(self._bn(p + ".bn2", self._conv(p + ".conv2", out, 1, 1, detach), detach))
        if stride > 1:
            out = F.avg_pool2d(out, stride)
        out = self._bn(p + ".bn3", self._conv(p + ".conv3", out, 1, 0, detach), detach)
        idn = x
        if (p + ".downsample.1.linear.weight") in self.sd:
            idn = F.avg_pool2d(x, stride) if stride > 1 else x
            idn = self._bn(p + ".downsample.2", self._conv(p + ".downsample.1", idn, 1, 0, detach), detach)
        return F.relu(out + idn)

    def attnpool(self, x: Tensor, detach: bool) -> Tensor:
        """bcosattnpool.py:34-59 (pooled mode): query = mean token, q/k detached in explanation mode, bias-free
        projections, `c_proj` used as a plain linear (only its .weight is read)."""
        N, C = x.shape[0], x.shape[1]
        t = x.flatten(2).permute(2, 0, 1)
        t = torch.cat([t.mean(dim=0, keepdim=True), t], dim=0)          # (HW+1) N C
        q, k = t[:1], t
        if detach:
            q, k = q.detach(), k.detach()
        H, dh = self.heads, C // self.heads
        qp = F.linear(q, self.sd["model.attnpool.q_proj.weight"]) * dh ** -0.5
        kp = F.linear(k, self.sd["model.attnpool.k_proj.weight"])
        vp = F.linear(t, self.sd["model.attnpool.v_proj.weight"])
        qh = qp.reshape(1, N * H, dh).transpose(0, 1)                   # [N*H, 1, dh]
        kh = kp.reshape(-1, N * H, dh).transpose(0, 1)
        vh = vp.reshape(-1, N * H, dh).transpose(0, 1)
        attn = torch.softmax(torch.bmm(qh, kh.transpose(1, 2)), dim=-1)
        o = torch.bmm(attn, vh).transpose(0, 1).reshape(1, N, C)
        return F.linear(o, self.sd["model.attnpool.c_proj.linear.weight"]).squeeze(0)

    def attn_unpool(self, x: Tensor, detach: bool) -> Tensor:
        """bcosattnpool.py:23-33 (`attn_unpool=True`): every spatial token -> v_proj (plain, stays nn.Linear
        bcosify.py:95) -> c_proj (BcosifyLinear, b=2) -> divided by its L2 norm, the norm detached in explanation mode.
        Returns (HW) x N x D'."""
        t = x.flatten(2).permute(2, 0, 1)
        t = F.linear(t, self.sd["model.attnpool.v_proj.weight"])
        t = bcos_linear(t, self.sd["model.attnpool.c_proj.linear.weight"], None, b=self.b, detach=detach, normalize_weight=False)
        norm = t.norm(dim=-1, keepdim=True)
        if detach:
            norm = norm.detach()
        return t / norm

    def trunk(self, x6: Tensor, detach: bool = False) -> Tensor:
        """CLIP/clip/model.py:139-152: stem, avg pool, four stages (everything in front of the attention pool)."""
        x = normalize6(x6, self.mean, self.std)
        x = F.relu(self._bn("model.bn1", self._conv("model.conv1", x, 2, 1, detach), detach))
        x = F.relu(self._bn("model.bn2", self._conv("model.conv2", x, 1, 1, detach), detach))
        x = F.relu(self._bn("model.bn3", self._conv("model.conv3", x, 1, 1, detach), detach))
        x = F.avg_pool2d(x, 2)
        for li, nb in enumerate(self.layers, start=1):
            for bi in range(nb):
                x = self._block(f"model.layer{li}.{bi}", x, 2 if (li > 1 and bi == 0) else 1, detach)
        return x

    def forward(self, x6: Tensor, detach: bool = False) -> Tensor:
        x = self.trunk(x6, detach)
        return self.attn_unpool(x, detach) if self.unpool else self.attnpool(x, detach)

    __call__ = forward

    def calibrate_bn(self, x6: Tensor) -> None:
        self.training, self.momentum = True, 1.0
        with torch.no_grad():
            self.forward(x6)
        self.training, self.momentum = False, 0.1


# ----------------------------------------------------------------------------------------------
# B-cos CLIP ViT image encoder (CLIP/clip/model.py:157-241 converted by bcosify.py:74-113 with clip_kd; biases and the
# positional embedding stripped by clip_bcosification/model.py:17-25).  Only conv1, mlp.c_fc and mlp.c_proj are B-cos
# transforms; LayerNorm, QuickGELU and the attention (incl. its out_proj weight) are the stock, non-detachable torch ops.
# ----------------------------------------------------------------------------------------------
def clip_vit_state_shapes(input_resolution: int = 224, patch: int = 32, width: int = 768, layers: int = 12,
                          output_dim: int = 512) -> Dict[str, Tuple[int, ...]]:
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
        # bcosify.py:101-103 turns the named nn.Sequential into a BcosSequential built from its values: positional keys
        s[p + ".mlp.0.linear.weight"] = (4 * width, width)
        s[p + ".mlp.2.linear.weight"] = (width, 4 * width)
    return s


class OracleCLIPViT:
    def __init__(self, sd: Dict[str, Tensor], heads: int = 12, b: float = 2, mean=CLIP_MEAN_ADDINVERSE, std=CLIP_STD_ADDINVERSE):
        self.sd, self.heads, self.b, self.mean, self.std = sd, heads, b, mean, std
        self.layers = len([k for k in sd if k.endswith(".attn.in_proj_weight")])

    def forward(self, x6: Tensor, detach: bool = False) -> Tensor:
        sd = self.sd
        x = normalize6(x6, self.mean, self.std)
        w = sd["model.conv1.linear.weight"]
        x = bcos_conv2d(x, w, None, w.shape[-1], 0, b=self.b, detach=detach)         # patch embedding: stride = kernel
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        cls = sd["model.class_embedding"].to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype)
        x = torch.cat([cls, x], dim=1)
        d = x.shape[-1]
        x = F.layer_norm(x, (d,), sd["model.ln_pre.weight"], None, 1e-5)
        x = x.permute(1, 0, 2)
        for i in range(self.layers):
            p = f"model.transformer.resblocks.{i}"
            y = F.layer_norm(x, (d,), sd[p + ".ln_1.weight"], None, 1e-5)
            a, _ = F.multi_head_attention_forward(y, y, y, d, self.heads, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"],
                                                  None, None, False, 0.0, sd[p + ".attn.out_proj.linear.weight"], None,
                                                  training=False, need_weights=False)
            x = x + a
            y = F.layer_norm(x, (d,), sd[p + ".ln_2.weight"], None, 1e-5)
            y = bcos_linear(y, sd[p + ".mlp.0.linear.weight"], None, b=self.b, detach=detach)
            y = y * torch.sigmoid(1.702 * y)
            y = bcos_linear(y, sd[p + ".mlp.2.linear.weight"], None, b=self.b, detach=detach)
            x = x + y
        x = x.permute(1, 0, 2)
        x = F.layer_norm(x[:, 0, :], (d,), sd["model.ln_post.weight"], None, 1e-5)
        return x @ sd["model.proj"]

    __call__ = forward


def clip_seed_direction(dim: int = 1024, seed: int = 0) -> Tensor:
    """Fixed unit 'text embedding' t for the CLIP explanation target cos(emb, t)
    (interpretability/analyses/text_localisation.py:77-100 back-propagates from the image-text cosine)."""
    g = torch.Generator().manual_seed(4242 + seed)
    t = torch.randn(dim, generator=g)
    return t / t.norm()


def explain_cosine(forward: Callable[..., Tensor], x6: Tensor, t: Tensor) -> Dict[str, Tensor]:
    xb = x6.clone().requires_grad_(True)
    with torch.enable_grad():
        emb = forward(xb, detach=True)
        target = F.cosine_similarity(emb, t.to(emb.dtype)[None], dim=1).sum()
        (grad,) = torch.autograd.grad(target, [xb])
    return {"embedding": emb.detach(), "dynamic_linear_weights": grad, "contribution_map": (x6 * grad).sum(1)}


# ----------------------------------------------------------------------------------------------
# explanation  (bcos/common.py:92-188 `BcosUtilMixin.explain`, batched form SURVEY.md A.4)
# ----------------------------------------------------------------------------------------------
def explain_batched(forward: Callable[..., Tensor], x6: Tensor, idx: Optional[Tensor] = None,
                    seed_grad: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """fwd under explanation mode (detach=True) -> out.max(1) (or `idx`, or an arbitrary output
    gradient `seed_grad`) -> backward(inputs=[x]) -> dynamic linear weights = x.grad,
    contribution_map = (x * x.grad).sum(1)   (bcos/common.py:163-181).
    Images are independent in eval mode, so one batched backward of the summed logits equals the
    per-sample `model.explain` loop."""
    xb = x6.clone().requires_grad_(True)
    with torch.enable_grad():
        out = forward(xb, detach=True)
        if seed_grad is not None:
            target = (out * seed_grad).sum()
            pred = out.argmax(1)
        else:
            pred = out.argmax(1)
            sel = pred if idx is None else idx
            target = out.gather(1, sel.view(-1, 1)).sum()
        (grad,) = torch.autograd.grad(target, [xb])
    return {
        "logits": out.detach(),
        "prediction": pred,
        "dynamic_linear_weights": grad,
        "contribution_map": (x6 * grad).sum(1),
    }


def gradient_to_image(image: Tensor, linear_mapping: Tensor, smooth: int = 15, alpha_percentile: float = 99.5) -> Tensor:
    """`gradient_to_image` bcos/common.py:387-436 -> RGBA [H,W,4] (torch, no numpy/matplotlib)."""
    contribs = (image * linear_mapping).sum(0, keepdim=True)[0]
    rgb_grad = linear_mapping / (linear_mapping.abs().max(0, keepdim=True).values + 1e-12)
    rgb_grad = rgb_grad.clamp(min=0)
    rgb_grad = rgb_grad[:3] / (rgb_grad[:3] + rgb_grad[3:] + 1e-12)
    alpha = linear_mapping.norm(p=2, dim=0, keepdim=True)
    alpha = torch.where(contribs[None] < 0, torch.zeros_like(alpha) + 1e-12, alpha)
    if smooth:
        alpha = F.avg_pool2d(alpha, smooth, stride=1, padding=(smooth - 1) // 2)
    alpha = alpha / torch.quantile(alpha.flatten(), q=alpha_percentile / 100)
    alpha = alpha.clamp(0, 1)
    rgb_grad = torch.cat([rgb_grad, alpha], dim=0)
    return rgb_grad.permute(1, 2, 0)


# ----------------------------------------------------------------------------------------------
# parity metrics (BASELINE.json north_star tolerances)
# ----------------------------------------------------------------------------------------------
def parity_metrics(logits: Tensor, maps: Tensor, ref_logits: Tensor, ref_maps: Tensor,
                   logit_bias: float = LOGIT_BIAS_1000) -> Dict[str, float]:
    """argmax equality, logit relative error, per-image map cosine and max-abs / range."""
    logits, maps, ref_logits, ref_maps = (t.detach().double().cpu() for t in (logits, maps, ref_logits, ref_maps))
    rel = ((logits - ref_logits).abs().max() / ref_logits.abs().max()).item()
    rel_nb = ((logits - ref_logits).abs().max() / (ref_logits - logit_bias).abs().max()).item()
    a, r = maps.flatten(1), ref_maps.flatten(1)
    cos = F.cosine_similarity(a, r, dim=1)
    rng = (r.max(1).values - r.min(1).values)
    mar = ((a - r).abs().max(1).values / rng)
    return {
        "argmax_equal": bool((logits.argmax(1) == ref_logits.argmax(1)).all()),
        "logit_rel_err": rel,
        "logit_rel_err_vs_unbiased": rel_nb,
        "map_cos_min": cos.min().item(),
        "map_maxabs_over_range": mar.max().item(),
    }


def gradient_to_image_batched(x6: Tensor, grad6: Tensor, smooth: int = 15, alpha_percentile: float = 99.5) -> Tensor:
    """RGBA explanations [nb, H, W, 4] - the reference's gradient_to_image (bcos/common.py:387-436) applied per image:
    colour = clamp(w / (max|w| + 1e-12), 0), rgb = c[:3] / (c[:3] + c[3:] + 1e-12); alpha = ||w||_2, 1e-12 where the
    contribution (x * w).sum(0) is negative; avg_pool2d(alpha, smooth, 1, (smooth-1)//2); / torch.quantile; clip."""
    import torch.nn.functional as F
    outs = []
    for image, lin in zip(x6, grad6):
        contribs = (image * lin).sum(0, keepdim=True)
        rgb = lin / (lin.abs().max(0, keepdim=True).values + 1e-12)
        rgb = rgb.clamp(min=0)
        rgb = rgb[:3] / (rgb[:3] + rgb[3:] + 1e-12)
        alpha = lin.norm(p=2, dim=0, keepdim=True)
        alpha = torch.where(contribs < 0, torch.full_like(alpha, 1e-12), alpha)
        if smooth:
            alpha = F.avg_pool2d(alpha, smooth, stride=1, padding=(smooth - 1) // 2)
        alpha = (alpha / torch.quantile(alpha, q=alpha_percentile / 100)).clip(0, 1)
        outs.append(torch.cat([rgb, alpha], 0).permute(1, 2, 0))
    return torch.stack(outs)


def localisation_scores(attributions: Tensor, cell: int, smooth: int = 0, neg: bool = False) -> Tensor:
    """interpretability/analyses/localisation.py:306-388 (the post-processing inside `LocalisationAnalyser.analysis`; it
    is inline in a method that needs the experiment/data stack, so it is restated with the same ATen calls):
    channel sum (:309-311), smoothing (:314-317), sign (:319-320), clamp (:322), region average pooling, transposition to
    column-major regions and the fraction of the total (:378-388).  Returns [T, regions]; metric_t = out[t, t] (:389)."""
    a = attributions.sum(1, keepdim=True)
    if smooth:
        a = F.avg_pool2d(a, smooth, stride=1, padding=(smooth - 1) // 2)
    if neg:
        a = -a
    a = a.clamp(min=0)
    contribs = F.avg_pool2d(a, cell, stride=cell).permute(0, 1, 3, 2).reshape(a.shape[0], -1)
    total = contribs.sum(1, keepdim=True)
    return torch.where(total * contribs > 0, contribs / total, torch.zeros_like(contribs))


def text_localisation_target(out: Tensor, zeroshot_weight: Tensor, attn_unpool: bool, pool_cosine: float = 1,
                             norm_max_cosine: bool = False) -> Tensor:
    """interpretability/analyses/text_localisation.py:80-105 (`compute_attributions`): the scalar whose gradient is the
    explanation -- cosine of the image embedding (per token with attn_unpool) with the text embedding [D, 1], tokens
    pooled by mean(cos * |cos|^(p-1)) for p > 1 (:97-98), arg-max token for p = 0 (:87-94), plain mean for p = 1."""
    img_features = out / out.norm(dim=-1, keepdim=True)
    logits = img_features @ zeroshot_weight
    if attn_unpool:
        logits = logits.reshape(-1, 1)
        if pool_cosine == 0:
            num_features = logits.shape[0]
            logits = logits.reshape(-1, num_features)
            max_locations = logits.argmax(dim=1)
            mask = torch.zeros_like(logits)
            for i in range(logits.shape[0]):
                mask[i, max_locations[i]] = 1.0
            logits = (logits * mask.detach()).reshape(1, num_features)
        if norm_max_cosine:
            logits = logits / logits.abs().detach().max(dim=0, keepdim=True)[0]
        if pool_cosine > 1:
            logits = logits * torch.pow(logits, pool_cosine - 1).abs().detach()
        logits = logits.mean(dim=0)
    if logits.dim() == 1:
        logits = logits.unsqueeze(0)
    return logits.max(1).values

# ----------------------------------------------------------------------------------------------
# fine-tuning step  (bcos/training/trainer.py:666-784 `training_step`, bcos/modules/losses.py:99-139,
# bcos/training/agc.py:28-42; SURVEY.md 8f row 2 / BASELINE config 5)
# ----------------------------------------------------------------------------------------------
def uniform_off_labels_bce(logits: Tensor, labels: Tensor, off_label: Optional[float] = None) -> Tensor:
    """`UniformOffLabelsBCEWithLogitsLoss.forward` losses.py:119-131 (reduction "mean")."""
    num_classes = logits.shape[-1]
    off_value = off_label or (1.0 / num_classes)
    target = F.one_hot(labels.long(), num_classes=num_classes).to(dtype=logits.dtype).clamp(min=off_value)
    return F.binary_cross_entropy_with_logits(logits, target, reduction="mean")


def unitwise_norm(x: Tensor) -> Tensor:
    """agc.py:12-25"""
    if x.squeeze().ndim <= 1:
        return x.norm(2.0)
    if x.ndim in (2, 3):
        return x.norm(2.0, dim=1, keepdim=True)
    return x.norm(2.0, dim=(1, 2, 3), keepdim=True)


def adaptive_clip_grad(p: Tensor, g: Tensor, clip_factor: float = 0.01, eps: float = 1e-3) -> Tensor:
    """agc.py:28-42 for one parameter: returns the clipped gradient."""
    max_norm = unitwise_norm(p).clamp(min=eps) * clip_factor
    grad_norm = unitwise_norm(g)
    clipped = g * (max_norm / grad_norm.clamp(min=1e-6))
    return torch.where(grad_norm < max_norm, g, clipped)


def train_step_reference(model: "OracleResNet", x6: Tensor, labels: Tensor, lr: float = 1e-4, betas=(0.9, 0.999), adam_eps: float = 1e-8,
                         weight_decay: float = 0.0, agc_clip: float = 0.01, agc_eps: float = 1e-3, off_label: Optional[float] = None,
                         momentum: float = 0.1) -> Dict[str, object]:
    """One fine-tuning step in fp32: forward in train mode (batch statistics, scales in the graph), loss, autograd gradients of
    every conv / classifier / norm weight, AGC, the first AdamW step.  Returns loss, logits, gradients, updated weights and the
    updated running variances (all keyed like the reference's state dict)."""
    keys = [k for k, v in model.sd.items() if v.is_floating_point() and (k.endswith(".linear.weight") or (k.endswith(".weight") and v.ndim == 1))]
    saved = {k: model.sd[k] for k in model.sd}
    params = {k: model.sd[k].detach().clone().requires_grad_(True) for k in keys}
    model.sd = {**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in saved.items()}, **params}
    model.training, model.momentum = True, momentum
    try:
        with torch.enable_grad():
            logits = model.forward(x6, detach=False)
            loss = uniform_off_labels_bce(logits, labels, off_label)
            grads = torch.autograd.grad(loss, [params[k] for k in keys])
        running = {k: v.detach().clone() for k, v in model.sd.items() if k.endswith("running_var")}
    finally:
        model.training, model.momentum = False, 0.1
        model.sd = saved
    grads = dict(zip(keys, grads))
    new_w = {}
    for k in keys:
        p = params[k].detach()
        g = adaptive_clip_grad(p, grads[k], agc_clip, agc_eps) if agc_clip > 0 else grads[k]
        m = (1 - betas[0]) * g
        v = (1 - betas[1]) * g * g
        p2 = p * (1 - lr * weight_decay)
        new_w[k] = p2 - lr * (m / (1 - betas[0])) / ((v / (1 - betas[1])).sqrt() + adam_eps)
    return {"loss": loss.detach(), "logits": logits.detach(), "grads": grads, "weights": new_w, "running_var": running}
