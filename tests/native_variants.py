"""Native B-cos-v2 model variants of the reference's model zoo (bcos/models/resnet.py:219-505, bcos/models/densenet.py:196-400),
built through whatever `bcos` package is importable: the real reference (oracle/make_golden.py --only native -> tests/golden/native_*.npz)
or bcos_b200 registered under the reference's import paths (tests/test_native_models_gpu.py: the reference's model FILES on OUR modules).
TEST INFRASTRUCTURE."""
from functools import partial

NUM_CLASSES = 16
IMAGE = 64
BATCH = 2
VARIANTS = ("resnet18_positionnorm", "resnet50_bnu_b1.5_maxout2", "densenet121_gnlayernorm")


def build(name: str):
    import bcos.models.densenet as D
    import bcos.models.resnet as R
    from bcos.modules import BcosConv2d, norms
    if name == "resnet18_positionnorm":          # the zoo's defaults: BcosConv2d(b=2), NoBias(DetachablePositionNorm2d), no activation
        return R.resnet18(num_classes=NUM_CLASSES)
    if name == "resnet50_bnu_b1.5_maxout2":      # uncentred batch norm (eval statistics), b = 1.5, MaxOut over 2 units
        return R.resnet50(num_classes=NUM_CLASSES, norm_layer=norms.NoBias(norms.BatchNormUncentered2d),
                          conv_layer=partial(BcosConv2d, b=1.5, max_out=2))
    if name == "densenet121_gnlayernorm":        # GroupNorm-style LayerNorm over (C, H, W), detachable
        return D.densenet121(num_classes=NUM_CLASSES, norm_layer=norms.NoBias(norms.DetachableGNLayerNorm2d))
    raise KeyError(name)


def explain_batched(m, x6):
    """logits + contribution maps of the top class of every image (explanation_mode, bcos/common.py:312-385, batched)."""
    import torch
    xb = x6.clone().requires_grad_(True)
    with torch.enable_grad(), m.explanation_mode():
        out = m(xb)
        out.max(1).values.sum().backward(inputs=[xb])
    return out.detach(), (xb * xb.grad).sum(1).detach()
