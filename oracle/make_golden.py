"""Generate tests/golden/*.npz from the REFERENCE ITSELF (build container only).

TEST INFRASTRUCTURE.  Imports shrebox/B-cosification read-only through oracle/refload.py, builds the
reference models offline (weights=None; mirrors bcos/experiments/ImageNet/bcosification/model.py:15-57
and experiment_parameters.py:82-129), loads the synthetic state dict from
`bcos_b200.utils.synth`, calibrates BN, runs forward + batched explanation
 and stores inputs / calibrated BN variances / logits / contribution maps.
While doing so it asserts that oracle/bcos_oracle.py reproduces the reference, i.e. it PINS the oracle.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py [--only resnet18,resnet50,modules]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refload  # noqa: E402
import bcos_oracle as O  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_reference_resnet(arch: str):
    refload.load()
    import bcosify
    from bcos.models.standard_models import ResNetBcos
    from torchvision.models.resnet import BasicBlock, Bottleneck

    kind, layers = O.RESNET_ARCH[arch]
    cfg = dict(is_bcos=True, name=arch, last_layer_name="fc", weights=None, bcos_args=dict(b=2, max_out=1),
               bcosify_args=dict(fix_b=True, use_bias=False, norm_layer="BnUncV2", manual_optim=False, gap=True,
                                 act_layer=True))
    tv = ResNetBcos(BasicBlock if kind == "basic" else Bottleneck, layers)
    m = bcosify.BcosifyNetwork(tv, cfg, add_channels=True, logit_layer=True)
    m.model.maxpool = nn.AvgPool2d(3, 2, 1)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
    return m


def reference_calibrate(m, x6):
    m.train()
    for b in m.modules():
        if isinstance(b, nn.BatchNorm2d):
            b.momentum = 1.0
    with torch.no_grad():
        m(x6)
    for b in m.modules():
        if isinstance(b, nn.BatchNorm2d):
            b.momentum = 0.1
    m.eval()


def reference_explain_batched(m, x6):
    xb = x6.clone().requires_grad_(True)
    with torch.enable_grad(), m.explanation_mode():
        out = m(xb)
        out.max(1).values.sum().backward(inputs=[xb])
    return out.detach(), xb.grad.detach(), (xb * xb.grad).sum(1).detach()


def golden_resnet(arch: str, batch: int, seed: int = 0):
    t0 = time.time()
    torch.manual_seed(0)
    m = build_reference_resnet(arch)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_shapes == O.resnet_state_shapes(arch), "oracle key/shape table differs from the reference state dict"
    sd = synth.synth_state_dict(ref_shapes, seed)
    m.load_state_dict(sd, strict=True)
    u8 = synth.synth_images_u8(batch, 224, seed)
    x6 = synth.to_bcos_input(u8)
    reference_calibrate(m, x6)

    # plain eval forward + the official per-sample explain on image 0
    with torch.inference_mode():
        logits_fwd = m(x6)
    logits, grad, cmap = reference_explain_batched(m, x6)
    assert torch.equal(logits, logits_fwd)
    e0 = m.explain(x6[:1].clone().requires_grad_(True))
    assert torch.allclose(e0["contribution_map"], cmap[:1], rtol=1e-4, atol=1e-9), "batched explain != model.explain"

    # ---- pin the oracle against the reference ----
    osd = {k: v.clone() for k, v in sd.items()}
    om = O.OracleResNet(arch, osd)
    om.calibrate_bn(x6)
    cal = {k: v for k, v in m.state_dict().items() if k.endswith("running_var")}
    for k, v in cal.items():
        assert torch.allclose(osd[k], v, rtol=1e-5, atol=0), k
        osd[k] = v.clone()  # continue from the reference's exact calibration
    oe = O.explain_batched(om.forward, x6)
    pm = O.parity_metrics(oe["logits"], oe["contribution_map"], logits, cmap)
    print(f"[{arch}] oracle vs reference: {pm}")
    assert pm["argmax_equal"] and pm["logit_rel_err"] < 1e-5 and pm["map_cos_min"] > 0.99999

    # ---- fp64 evaluation of the (pinned) oracle: the reference's own fp32 rounding-noise floor on this workload.
    # Random-init deep B-cos nets amplify rounding noise by 10^2-10^3, so "distance to the fp32
    # reference" is only meaningful down to the reference's own distance from the exact result.
    osd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in osd.items()}
    e64 = O.explain_batched(O.OracleResNet(arch, osd64).forward, x6.double())
    floor = O.parity_metrics(logits, cmap, e64["logits"], e64["contribution_map"])
    print(f"[{arch}] reference fp32 vs fp64 evaluation (noise floor): {floor}")

    keys = sorted(cal)
    out = dict(
        logits_fp64=e64["logits"].numpy(),
        contribution_map_fp64=e64["contribution_map"].float().numpy(),
        fp32_noise_floor_maxabs_over_range=np.float64(floor["map_maxabs_over_range"]),
        fp32_noise_floor_logit_rel_err=np.float64(floor["logit_rel_err"]),
        images_u8=u8,
        bn_keys=np.array(keys),
        bn_sizes=np.array([cal[k].numel() for k in keys], dtype=np.int64),
        bn_var=torch.cat([cal[k].flatten() for k in keys]).numpy(),
        logits=logits.numpy(),
        contribution_map=cmap.numpy(),
        grad_absmax=grad.abs().amax(dim=(1, 2, 3)).numpy(),
        seed=np.int64(seed),
    )
    path = os.path.join(GOLD, f"{arch}_b{batch}.npz")
    np.savez_compressed(path, **out)
    print(f"[{arch}] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB) in {time.time()-t0:.1f}s; "
          f"logit std {logits.std():.4f}, pred {logits.argmax(1).tolist()}")


def build_reference_densenet(arch: str):
    refload.load()
    import bcosify
    from bcos.models.standard_models import DenseNetBcos
    growth, blocks, init = O.DENSENET_ARCH[arch]
    cfg = dict(is_bcos=True, name=arch, last_layer_name="classifier", weights=None, bcos_args=dict(b=2, max_out=1),
               bcosify_args=dict(fix_b=True, use_bias=False, norm_layer="BnUncV2", manual_optim=False, gap=True,
                                 act_layer=True))
    m = bcosify.BcosifyNetwork(DenseNetBcos(growth, blocks, init), cfg, add_channels=True, logit_layer=True)
    m.model.features[3] = nn.AvgPool2d(kernel_size=3, stride=2, padding=1)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
    return m


def golden_densenet(arch: str, batch: int, seed: int = 0):
    t0 = time.time()
    m = build_reference_densenet(arch)
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_shapes == O.densenet_state_shapes(arch), "oracle key/shape table differs from the reference state dict"
    sd = synth.synth_state_dict(ref_shapes, seed)
    m.load_state_dict(sd, strict=True)
    u8 = synth.synth_images_u8(batch, 224, seed)
    x6 = synth.to_bcos_input(u8)
    reference_calibrate(m, x6)
    logits, grad, cmap = reference_explain_batched(m, x6)
    osd = {k: v.clone() for k, v in sd.items()}
    om = O.OracleDenseNet(arch, osd)
    om.calibrate_bn(x6)
    cal = {k: v for k, v in m.state_dict().items() if k.endswith("running_var")}
    for k, v in cal.items():
        assert torch.allclose(osd[k], v, rtol=1e-4, atol=0), k
        osd[k] = v.clone()
    oe = O.explain_batched(om.forward, x6)
    pm = O.parity_metrics(oe["logits"], oe["contribution_map"], logits, cmap)
    print(f"[{arch}] oracle vs reference: {pm}")
    assert pm["argmax_equal"] and pm["logit_rel_err"] < 1e-5 and pm["map_cos_min"] > 0.99999
    osd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in osd.items()}
    e64 = O.explain_batched(O.OracleDenseNet(arch, osd64).forward, x6.double())
    floor = O.parity_metrics(logits, cmap, e64["logits"], e64["contribution_map"])
    print(f"[{arch}] reference fp32 vs fp64 evaluation (noise floor): {floor}")
    keys = sorted(cal)
    path = os.path.join(GOLD, f"{arch}_b{batch}.npz")
    np.savez_compressed(
        path, images_u8=u8, bn_keys=np.array(keys), bn_sizes=np.array([cal[k].numel() for k in keys], dtype=np.int64),
        bn_var=torch.cat([cal[k].flatten() for k in keys]).numpy(), logits=logits.numpy(), contribution_map=cmap.numpy(),
        logits_fp64=e64["logits"].numpy(), contribution_map_fp64=e64["contribution_map"].float().numpy(),
        fp32_noise_floor_maxabs_over_range=np.float64(floor["map_maxabs_over_range"]),
        fp32_noise_floor_logit_rel_err=np.float64(floor["logit_rel_err"]), seed=np.int64(seed))
    print(f"[{arch}] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB) in {time.time()-t0:.1f}s; logit std {logits.std():.4f}")


def build_reference_vit(arch: str):
    refload.load()
    import bcosify_vit
    import bcos.models.vit as rvit
    from bcos.modules.norms.centered_norms import DetachableGNLayerNorm2d
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        base = getattr(rvit, arch)(linear_layer=nn.Linear, conv2d_layer=nn.Conv2d, norm_layer=nn.LayerNorm,
                                   norm2d_layer=DetachableGNLayerNorm2d, act_layer=nn.GELU, channels=3)
    cfg = dict(is_bcos=True, name=arch, weights=None, logit_layer=True, bcos_args=dict(b=2, max_out=1),
               bcosify_args=dict(use_bias=False), args=dict(gap_reorder=True))
    m = bcosify_vit.BcosifyNetwork(base, cfg, add_channels=True, logit_layer=True)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
    m.model.gap_reorder = True
    return m


def golden_vit(arch: str, batch: int, seed: int = 0):
    t0 = time.time()
    m = build_reference_vit(arch).eval()
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_shapes == O.vit_state_shapes(arch), (set(ref_shapes) ^ set(O.vit_state_shapes(arch)))
    sd = synth.synth_state_dict(ref_shapes, seed)
    m.load_state_dict(sd, strict=True)
    u8 = synth.synth_images_u8(batch, 224, seed)
    x6 = synth.to_bcos_input(u8)
    logits, grad, cmap = reference_explain_batched(m, x6)
    with torch.inference_mode():
        assert torch.allclose(m(x6), logits, rtol=1e-5, atol=1e-6)
    oe = O.explain_batched(O.OracleViT(arch, sd).forward, x6)
    pm = O.parity_metrics(oe["logits"], oe["contribution_map"], logits, cmap)
    print(f"[{arch}] oracle vs reference: {pm}")
    assert pm["argmax_equal"] and pm["logit_rel_err"] < 1e-5 and pm["map_cos_min"] > 0.99999
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    e64 = O.explain_batched(O.OracleViT(arch, sd64).forward, x6.double())
    floor = O.parity_metrics(logits, cmap, e64["logits"], e64["contribution_map"])
    print(f"[{arch}] reference fp32 vs fp64 evaluation (noise floor): {floor}")
    path = os.path.join(GOLD, f"{arch}_b{batch}.npz")
    np.savez_compressed(
        path, images_u8=u8, logits=logits.numpy(), contribution_map=cmap.numpy(), logits_fp64=e64["logits"].numpy(),
        contribution_map_fp64=e64["contribution_map"].float().numpy(),
        fp32_noise_floor_maxabs_over_range=np.float64(floor["map_maxabs_over_range"]),
        fp32_noise_floor_logit_rel_err=np.float64(floor["logit_rel_err"]), seed=np.int64(seed))
    print(f"[{arch}] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB) in {time.time()-t0:.1f}s; logit std {logits.std():.4f}")


def golden_clip_rn50(batch: int, seed: int = 0):
    """config 4: B-cos CLIP RN50 image encoder, explanation target = cos(embedding, fixed unit vector)."""
    import torch.nn.functional as F
    t0 = time.time()
    refload.load()
    import bcosify
    from CLIP.clip.model import ModifiedResNet
    cfg = dict(is_bcos=True, name="resnet50clip", bcos_args=dict(b=2, max_out=1),
               bcosify_args=dict(clip_kd=True, fix_b=True, norm_layer="BnUncV2", use_bias=False))
    m = bcosify.BcosifyNetwork(ModifiedResNet((3, 4, 6, 3), 1024, 32, 224, 64).float(), cfg, add_channels=True, logit_layer=False)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
        if hasattr(mod, "positional_embedding") and mod.positional_embedding is not None:
            mod.positional_embedding = None
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    mine = O.clip_rn_state_shapes()
    assert ref_shapes == mine, (sorted(set(ref_shapes) ^ set(mine))[:10])
    sd = synth.synth_state_dict(ref_shapes, seed)
    m.load_state_dict(sd, strict=True)
    u8 = synth.synth_images_u8(batch, 224, seed)
    x6 = synth.to_bcos_input(u8)
    reference_calibrate(m, x6)
    tvec = O.clip_seed_direction(1024, seed)
    xb = x6.clone().requires_grad_(True)
    with torch.enable_grad(), m.explanation_mode():
        emb = m(xb)
        F.cosine_similarity(emb, tvec[None], dim=1).sum().backward(inputs=[xb])
    cmap = (xb * xb.grad).sum(1).detach()
    emb = emb.detach()
    osd = {k: v.clone() for k, v in sd.items()}
    om = O.OracleCLIPResNet(osd)
    om.calibrate_bn(x6)
    cal = {k: v for k, v in m.state_dict().items() if k.endswith("running_var")}
    for k, v in cal.items():
        assert torch.allclose(osd[k], v, rtol=1e-4, atol=0), k
        osd[k] = v.clone()
    oe = O.explain_cosine(om.forward, x6, tvec)
    erel = ((oe["embedding"] - emb).abs().max() / emb.abs().max()).item()
    mrel = ((oe["contribution_map"] - cmap).abs().max() / cmap.abs().max()).item()
    print(f"[clip_rn50] oracle vs reference: embedding rel err {erel:.2e}, map rel err {mrel:.2e}")
    assert erel < 1e-5 and mrel < 1e-4
    osd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in osd.items()}
    e64 = O.explain_cosine(O.OracleCLIPResNet(osd64).forward, x6.double(), tvec.double())
    rng = (cmap.flatten(1).max(1).values - cmap.flatten(1).min(1).values)
    floor = ((cmap.double() - e64["contribution_map"]).abs().flatten(1).max(1).values / rng).max().item()
    print(f"[clip_rn50] reference fp32 vs fp64 evaluation: map max-abs/range {floor:.2e}; |emb| {emb.norm(dim=1).tolist()}")
    keys = sorted(cal)
    path = os.path.join(GOLD, f"clip_rn50_b{batch}.npz")
    np.savez_compressed(
        path, images_u8=u8, bn_keys=np.array(keys), bn_sizes=np.array([cal[k].numel() for k in keys], dtype=np.int64),
        bn_var=torch.cat([cal[k].flatten() for k in keys]).numpy(), embedding=emb.numpy(), contribution_map=cmap.numpy(),
        contribution_map_fp64=e64["contribution_map"].float().numpy(), embedding_fp64=e64["embedding"].numpy(),
        fp32_noise_floor_maxabs_over_range=np.float64(floor), seed=np.int64(seed))
    print(f"[clip_rn50] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB) in {time.time()-t0:.1f}s")


def golden_clip_vit(batch: int = 2, seed: int = 0, patch: int = 32, width: int = 768, layers: int = 12, heads: int = 12, out_dim: int = 512):
    """north_star's "CLIP ... ViT image encoder": CLIP's VisionTransformer (ViT-B/32 geometry) converted by the reference's
    bcosify.py with clip_kd, biases and positional embedding stripped (clip_bcosification/model.py:17-25); explanation target =
    cos(embedding, fixed unit vector) like the CLIP RN50 golden."""
    import torch.nn.functional as F
    t0 = time.time()
    refload.load()
    import bcosify
    from CLIP.clip.model import VisionTransformer
    cfg = dict(is_bcos=True, name="vitclip", bcos_args=dict(b=2, max_out=1),
               bcosify_args=dict(clip_kd=True, fix_b=True, norm_layer="BnUncV2", use_bias=False))
    m = bcosify.BcosifyNetwork(VisionTransformer(224, patch, width, layers, heads, out_dim).float(), cfg, add_channels=True, logit_layer=False)
    for mod in m.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
        if hasattr(mod, "positional_embedding") and mod.positional_embedding is not None:
            mod.positional_embedding = None
    ref_shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    mine = O.clip_vit_state_shapes(224, patch, width, layers, out_dim)
    assert ref_shapes == mine, (sorted(set(ref_shapes) ^ set(mine))[:10])
    sd = synth.synth_state_dict(ref_shapes, seed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    u8 = synth.synth_images_u8(batch, 224, seed)
    x6 = synth.to_bcos_input(u8)
    tvec = O.clip_seed_direction(out_dim, seed)
    xb = x6.clone().requires_grad_(True)
    with torch.enable_grad(), m.explanation_mode():
        emb = m(xb)
        F.cosine_similarity(emb, tvec[None], dim=1).sum().backward(inputs=[xb])
    cmap = (xb * xb.grad).sum(1).detach()
    emb = emb.detach()
    om = O.OracleCLIPViT({k: v.clone() for k, v in sd.items()}, heads)
    oe = O.explain_cosine(om.forward, x6, tvec)
    erel = ((oe["embedding"] - emb).abs().max() / emb.abs().max()).item()
    mrel = ((oe["contribution_map"] - cmap).abs().max() / cmap.abs().max()).item()
    print(f"[clip_vit] oracle vs reference: embedding rel err {erel:.2e}, map rel err {mrel:.2e}")
    assert erel < 1e-5 and mrel < 1e-4
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    e64 = O.explain_cosine(O.OracleCLIPViT(sd64, heads).forward, x6.double(), tvec.double())
    rng = (cmap.flatten(1).max(1).values - cmap.flatten(1).min(1).values)
    floor = ((cmap.double() - e64["contribution_map"]).abs().flatten(1).max(1).values / rng).max().item()
    print(f"[clip_vit] reference fp32 vs fp64 evaluation: map max-abs/range {floor:.2e}; |emb| {emb.norm(dim=1).tolist()}")
    path = os.path.join(GOLD, f"clip_vit_b32_b{batch}.npz")
    np.savez_compressed(path, images_u8=u8, embedding=emb.numpy(), contribution_map=cmap.numpy(),
                        contribution_map_fp64=e64["contribution_map"].float().numpy(), embedding_fp64=e64["embedding"].numpy(),
                        fp32_noise_floor_maxabs_over_range=np.float64(floor), seed=np.int64(seed),
                        geometry=np.array([224, patch, width, layers, heads, out_dim], dtype=np.int64))
    print(f"[clip_vit] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB) in {time.time()-t0:.1f}s")


def golden_clip_unpool(seed: int = 0):
    """the attn_unpool head (bcosattnpool.py:23-33) and the text-localisation target
    (interpretability/analyses/text_localisation.py:68-105) on the reference CLIP RN50, image 0 of the clip_rn50_b2 golden
    (same weights and BN calibration: the state dict is per-key deterministic).  text_localisation.py imports matplotlib /
    the CLIP tokenizer and cannot be imported here, so lines 77-105 are executed literally below on the reference model."""
    t0 = time.time()
    refload.load()
    import bcosify
    from CLIP.clip.model import ModifiedResNet
    base = np.load(os.path.join(GOLD, "clip_rn50_b2.npz"))
    cfg = dict(is_bcos=True, name="resnet50clip", bcos_args=dict(b=2, max_out=1), attn_unpool=True,
               bcosify_args=dict(clip_kd=True, fix_b=True, norm_layer="BnUncV2", use_bias=False))
    model = bcosify.BcosifyNetwork(ModifiedResNet((3, 4, 6, 3), 1024, 32, 224, 64).float(), cfg, add_channels=True, logit_layer=False)
    for mod in model.modules():
        if hasattr(mod, "bias") and mod.bias is not None:
            mod.bias = None
        if hasattr(mod, "positional_embedding") and mod.positional_embedding is not None:
            mod.positional_embedding = None
    assert model.model.attnpool.attn_unpool
    ref_shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ref_shapes == O.clip_rn_state_shapes(attn_unpool=True), sorted(set(ref_shapes) ^ set(O.clip_rn_state_shapes(attn_unpool=True)))
    sd = synth.synth_state_dict(ref_shapes, seed)
    off = 0
    for k, n in zip(base["bn_keys"].tolist(), base["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(base["bn_var"][off:off + n].copy()); off += n
    model.load_state_dict(sd, strict=True)
    model.eval()
    test_img = synth.to_bcos_input(base["images_u8"][:1])[0]
    zeroshot_weight = O.clip_seed_direction(1024, seed).unsqueeze(1)
    out = {}
    for pool_cosine in (1, 2, 0):
        # ---- text_localisation.py:77-105, verbatim semantics ----
        with torch.enable_grad(), model.explanation_mode():
            imga = test_img[None].requires_grad_()
            outa = model(imga)
            img_features = outa / outa.norm(dim=-1, keepdim=True)
            logits = img_features @ zeroshot_weight
            if model.model.attnpool.attn_unpool:
                logits = logits.reshape(-1, 1)
                if pool_cosine == 0:
                    num_features = logits.shape[0]
                    logits = logits.reshape(-1, num_features)
                    max_locations = logits.argmax(dim=1)
                    mask = torch.zeros_like(logits)
                    for i in range(logits.shape[0]):
                        mask[i, max_locations[i]] = 1.0
                    logits = logits * mask.detach()
                    logits = logits.reshape(1, num_features)
                if pool_cosine > 1:
                    logits = logits * torch.pow(logits, pool_cosine - 1).abs().detach()
                logits = logits.mean(dim=0)
            if logits.dim() == 1:
                logits = logits.unsqueeze(0)
            target = logits.max(1).values
            target.backward(inputs=[imga])
            grada = imga.grad.detach()[0]
        contribs = (test_img * grada).sum(0)
        # ---- oracle ----
        om = O.OracleCLIPResNet(sd)
        xb = test_img[None].clone().requires_grad_(True)
        with torch.enable_grad():
            oo = om.forward(xb, detach=True)
            ot = O.text_localisation_target(oo, zeroshot_weight, True, pool_cosine)
            (og,) = torch.autograd.grad(ot.sum(), [xb])
        erel = ((oo.detach() - outa.detach()).abs().max() / outa.detach().abs().max()).item()
        grel = ((og[0] - grada).abs().max() / grada.abs().max()).item()
        print(f"[clip_unpool p={pool_cosine}] oracle vs reference: tokens rel err {erel:.2e}, target {ot.item():.6f} vs {target.item():.6f}, grad rel err {grel:.2e}")
        assert erel < 1e-5 and grel < 1e-4 and abs(ot.item() - target.item()) < 1e-6
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        x64 = test_img[None].double().clone().requires_grad_(True)
        with torch.enable_grad():
            t64 = O.text_localisation_target(O.OracleCLIPResNet(sd64).forward(x64, detach=True), zeroshot_weight.double(), True, pool_cosine)
            (g64,) = torch.autograd.grad(t64.sum(), [x64])
        c64 = (test_img.double() * g64[0]).sum(0)
        out[f"p{pool_cosine}.target"] = target.detach().numpy()
        out[f"p{pool_cosine}.contribution_map"] = contribs.numpy()
        out[f"p{pool_cosine}.contribution_map_fp64"] = c64.float().numpy()
        if pool_cosine == 1:
            out["tokens"] = outa.detach().numpy()
    path = os.path.join(GOLD, "clip_rn50_unpool_b1.npz")
    np.savez_compressed(path, seed=np.int64(seed), **out)
    print(f"[clip_unpool] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB) in {time.time()-t0:.1f}s")


def golden_modules(seed: int = 0):
    """Known-answer vectors for single modules, produced by the reference classes themselves."""
    refload.load()
    from bcos.modules.bcosconv2d import BcosConv2d
    from bcos.modules.bcosifyconv2d import BcosifyConv2d
    from bcos.modules.bcoslinear import BcosLinear
    from bcos.modules.bcosifylinear import BcosifyLinear
    from bcos.modules.norms.uncentered_norms import BatchNormUncentered2d
    from bcos.modules.norms.centered_norms import DetachableLayerNorm
    from bcos.modules.logitlayer import LogitLayer
    import bcosify_vit

    g = torch.Generator().manual_seed(1234 + seed)
    out = {}
    conv_cases = [
        # name, cls, cin, cout, k, s, p, b, max_out, H
        ("conv_bcosify_3x3", BcosifyConv2d, 16, 32, 3, 1, 1, 2, 1, 12),
        ("conv_bcosify_3x3_s2", BcosifyConv2d, 16, 32, 3, 2, 1, 2, 1, 13),
        ("conv_bcosify_1x1", BcosifyConv2d, 64, 24, 1, 1, 0, 2, 1, 7),
        ("conv_bcosify_1x1_s2", BcosifyConv2d, 32, 16, 1, 2, 0, 2, 1, 8),
        ("conv_bcosify_7x7_s2", BcosifyConv2d, 6, 16, 7, 2, 3, 2, 1, 20),
        ("conv_bcos_3x3_normed", BcosConv2d, 8, 16, 3, 1, 1, 2, 1, 9),
        ("conv_bcos_b1p5", BcosConv2d, 8, 16, 3, 1, 1, 1.5, 1, 9),
        ("conv_bcos_b2p5_mo2", BcosConv2d, 8, 16, 5, 2, 2, 2.5, 2, 11),
        ("conv_bcos_b2_mo3", BcosConv2d, 8, 8, 1, 1, 0, 2, 3, 6),
        ("conv_bcos_b1", BcosConv2d, 8, 8, 3, 1, 1, 1, 1, 6),
    ]
    for name, cls, cin, cout, k, s, p, b, mo, H in conv_cases:
        mod = cls(cin, cout, kernel_size=k, stride=s, padding=p, b=b, max_out=mo)
        w = torch.randn(mod.linear.weight.shape, generator=g) * 0.2
        mod.linear.weight.data = w.clone()
        x = torch.randn(2, cin, H, H, generator=g)
        y = mod(x)
        # explanation mode: detached scale, gradient of a fixed random seed
        seedg = torch.randn(y.shape, generator=g)
        mod.set_explanation_mode(True)
        xg = x.clone().requires_grad_(True)
        ye = mod(xg)
        (gx,) = torch.autograd.grad((ye * seedg).sum(), [xg])
        norm = mod.calc_patch_norms(x) if b != 1 else torch.zeros(1)
        normed = cls is BcosConv2d
        yo = O.bcos_conv2d(x, w, None, s, p, b=b, max_out=mo, normalize_weight=normed)
        assert torch.equal(yo, y.detach()), name
        out.update({f"{name}.w": w, f"{name}.x": x, f"{name}.y": y.detach(), f"{name}.seed": seedg, f"{name}.gx": gx,
                    f"{name}.norm": norm.detach(),
                    f"{name}.meta": torch.tensor([cin, cout, k, s, p, b, mo, float(normed)], dtype=torch.float64)})

    lin_cases = [
        ("lin_bcosify", BcosifyLinear, 48, 40, 2, 1),
        ("lin_bcos_normed", BcosLinear, 48, 40, 2, 1),
        ("lin_bcos_b1p5_mo2", BcosLinear, 32, 24, 1.5, 2),
    ]
    for name, cls, fin, fout, b, mo in lin_cases:
        mod = cls(fin, fout, b=b, max_out=mo)
        w = torch.randn(mod.linear.weight.shape, generator=g) * 0.2
        mod.linear.weight.data = w.clone()
        if getattr(mod.linear, "bias", None) is not None:
            mod.linear.bias = None
        mod.bias = None
        x = torch.randn(3, 5, fin, generator=g)
        y = mod(x)
        seedg = torch.randn(y.shape, generator=g)
        mod.set_explanation_mode(True)
        xg = x.clone().requires_grad_(True)
        (gx,) = torch.autograd.grad((mod(xg) * seedg).sum(), [xg])
        normed = cls is BcosLinear
        yo = O.bcos_linear(x, w, None, b=b, max_out=mo, normalize_weight=normed)
        assert torch.allclose(yo, y.detach(), rtol=1e-6, atol=1e-7), name
        out.update({f"{name}.w": w, f"{name}.x": x, f"{name}.y": y.detach(), f"{name}.seed": seedg, f"{name}.gx": gx,
                    f"{name}.meta": torch.tensor([fin, fout, b, mo, float(normed)], dtype=torch.float64)})

    # BatchNormUncentered2d eval + train + BnUncV2 fold
    bnu = BatchNormUncentered2d(12)
    bnu.weight.data = torch.rand(12, generator=g) + 0.5
    bnu.bias.data = torch.randn(12, generator=g) * 0.1
    bnu.running_var.data = torch.rand(12, generator=g) + 0.2
    x = torch.randn(4, 12, 5, 5, generator=g) + 0.3
    bnu.eval()
    y_eval = bnu(x)
    assert torch.equal(O.batch_norm_uncentered_2d(x, bnu.running_var, bnu.weight, bnu.bias), y_eval.detach())
    rv0 = bnu.running_var.detach().clone()
    bnu.train()
    y_train = bnu(x)
    out.update({"bnu.w": bnu.weight.detach(), "bnu.b": bnu.bias.detach(), "bnu.rv0": rv0, "bnu.x": x,
                "bnu.y_eval": y_eval.detach(), "bnu.y_train": y_train.detach(), "bnu.rv1": bnu.running_var.detach().clone()})
    std_bn = nn.BatchNorm2d(12)
    std_bn.weight.data = torch.rand(12, generator=g) + 0.5
    std_bn.bias.data = torch.randn(12, generator=g)
    std_bn.running_mean.data = torch.randn(12, generator=g)
    std_bn.running_var.data = torch.rand(12, generator=g) + 0.2
    std_bn.eval()
    folded = BatchNormUncentered2d.from_standard_module(std_bn, dict(bcosify_args=dict(norm_layer="BnUncV2"))).eval()
    assert torch.allclose(folded(x), std_bn(x), rtol=1e-5, atol=1e-6)
    fw, fb = O.bn_uncentered_from_standard(std_bn.weight.data, std_bn.bias.data, std_bn.running_mean, std_bn.running_var, std_bn.eps)
    assert torch.allclose(fb, folded.bias.data)
    out.update({"bnfold.w": std_bn.weight.detach(), "bnfold.b": std_bn.bias.detach(), "bnfold.rm": std_bn.running_mean.clone(),
                "bnfold.rv": std_bn.running_var.clone(), "bnfold.y": std_bn(x).detach(), "bnfold.bias_folded": folded.bias.detach()})

    # DetachableLayerNorm / MyGELU / LogitLayer
    ln = DetachableLayerNorm(24)
    ln.weight.data = torch.rand(24, generator=g) + 0.5
    ln.bias = None
    x = torch.randn(2, 7, 24, generator=g)
    y = ln(x)
    seedg = torch.randn(y.shape, generator=g)
    ln.set_explanation_mode(True)
    xg = x.clone().requires_grad_(True)
    ye = ln(xg)
    (gx,) = torch.autograd.grad((ye * seedg).sum(), [xg])
    assert torch.allclose(O.layer_norm_detachable(x, ln.weight, None, ln.eps, True), ye.detach(), rtol=1e-6, atol=1e-6)
    out.update({"ln.w": ln.weight.detach(), "ln.x": x, "ln.y": y.detach(), "ln.y_explain": ye.detach(), "ln.seed": seedg, "ln.gx": gx})

    gelu = bcosify_vit.MyGELU()
    x = torch.randn(3, 11, 16, generator=g) * 2
    y = gelu(x)
    seedg = torch.randn(y.shape, generator=g)
    gelu.set_explanation_mode(True)
    xg = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad((gelu(xg) * seedg).sum(), [xg])
    assert torch.allclose(O.gelu_detachable(x), y, rtol=1e-6, atol=1e-7)
    out.update({"gelu.x": x, "gelu.y": y.detach(), "gelu.seed": seedg, "gelu.gx": gx})

    ll = LogitLayer(logit_temperature=2.0, logit_bias=-1.5)
    x = torch.randn(4, 10, generator=g)
    assert torch.equal(O.logit_layer(x, 2.0, -1.5), ll(x))
    out.update({"logit.x": x, "logit.y": ll(x)})

    path = os.path.join(GOLD, "modules_kat.npz")
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in out.items()})
    print(f"[modules] wrote {path} ({os.path.getsize(path)/1e6:.2f} MB), {len(out)} arrays")


def golden_norms(seed: int = 7):
    """Known-answer vectors for the group / position / all norms, produced by the reference classes themselves:
    forward outside and inside explanation mode, and the explanation gradient of a fixed random seed."""
    refload.load()
    from bcos.modules.norms import centered_norms as CN
    from bcos.modules.norms import uncentered_norms as UN

    g = torch.Generator().manual_seed(4321 + seed)
    out = {}
    cases = [
        # name, factory, oracle(x, w, b, eps, detach), channels, H, W, drop bias
        ("gnu_g4", lambda: UN.GroupNormUncentered2d(4, 16), lambda x, w, b, e, d: O.group_norm_detachable(x, 4, w, b, e, d, False), 16, 6, 6, True),
        ("gnu_g4_odd", lambda: UN.GroupNormUncentered2d(4, 16), lambda x, w, b, e, d: O.group_norm_detachable(x, 4, w, b, e, d, False), 16, 5, 5, False),
        ("gnu_layer", lambda: UN.GNLayerNormUncentered2d(12), lambda x, w, b, e, d: O.group_norm_detachable(x, 1, w, b, e, d, False), 12, 8, 8, True),
        ("gnu_instance", lambda: UN.GNInstanceNormUncentered2d(12), lambda x, w, b, e, d: O.group_norm_detachable(x, 12, w, b, e, d, False), 12, 7, 7, True),
        ("dgn_g2", lambda: CN.DetachableGroupNorm2d(2, 16), lambda x, w, b, e, d: O.group_norm_detachable(x, 2, w, b, e, d, True), 16, 6, 6, True),
        ("dgn_layer", lambda: CN.DetachableGNLayerNorm2d(24), lambda x, w, b, e, d: O.group_norm_detachable(x, 1, w, b, e, d, True), 24, 14, 14, False),
        ("dgn_instance_odd", lambda: CN.DetachableGNInstanceNorm2d(6), lambda x, w, b, e, d: O.group_norm_detachable(x, 6, w, b, e, d, True), 6, 5, 7, True),
        ("pnu", lambda: UN.PositionNormUncentered2d(24), lambda x, w, b, e, d: O.position_norm_detachable(x, w, b, e, d, False), 24, 7, 7, True),
        ("pnu_wide", lambda: UN.PositionNormUncentered2d(80), lambda x, w, b, e, d: O.position_norm_detachable(x, w, b, e, d, False), 80, 9, 8, False),
        ("dpn", lambda: CN.DetachablePositionNorm2d(24), lambda x, w, b, e, d: O.position_norm_detachable(x, w, b, e, d, True), 24, 7, 7, True),
        ("dpn_bias", lambda: CN.DetachablePositionNorm2d(5), lambda x, w, b, e, d: O.position_norm_detachable(x, w, b, e, d, True), 5, 6, 6, False),
    ]
    for name, make, orc, c, h, w_, nobias in cases:
        mod = make()
        mod.weight.data = torch.rand(c, generator=g) + 0.5
        if nobias:
            mod.bias = None
        else:
            mod.bias.data = torch.randn(c, generator=g) * 0.1
        x = torch.randn(3, c, h, w_, generator=g) * (torch.rand(3, c, 1, 1, generator=g) + 0.5) + 0.4
        y = mod(x)
        seedg = torch.randn(y.shape, generator=g)
        mod.set_explanation_mode(True)
        xg = x.clone().requires_grad_(True)
        ye = mod(xg)
        (gx,) = torch.autograd.grad((ye * seedg).sum(), [xg])
        bias = None if mod.bias is None else mod.bias.detach()
        assert torch.allclose(orc(x, mod.weight.detach(), bias, mod.eps, False), y.detach(), rtol=1e-6, atol=1e-6), name
        xo = x.clone().requires_grad_(True)
        yo = orc(xo, mod.weight.detach(), bias, mod.eps, True)
        (go,) = torch.autograd.grad((yo * seedg).sum(), [xo])
        assert torch.equal(yo.detach(), ye.detach()) and torch.equal(go, gx), name      # same ATen calls: bit exact
        out.update({f"{name}.w": mod.weight.detach(), f"{name}.x": x, f"{name}.y": y.detach(), f"{name}.y_explain": ye.detach(),
                    f"{name}.seed": seedg, f"{name}.gx": gx})
        if bias is not None:
            out[f"{name}.b"] = bias

    an = UN.AllNormUncentered2d(10)
    an.weight.data = torch.rand(1, generator=g) + 0.5
    an.bias.data = torch.randn(1, generator=g) * 0.1
    an.running_var.data = torch.rand(1, generator=g) + 0.2
    x = torch.randn(3, 10, 6, 6, generator=g) + 0.3
    an.eval()
    y_eval = an(x)
    assert torch.equal(O.all_norm_uncentered_2d(x, an.running_var, an.weight, an.bias), y_eval.detach())
    rv0 = an.running_var.detach().clone()
    an.train()
    y_train = an(x)
    an.eval()
    an.set_explanation_mode(True)
    seedg = torch.randn(y_eval.shape, generator=g)
    xg = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad((an(xg) * seedg).sum(), [xg])
    out.update({"alln.w": an.weight.detach(), "alln.b": an.bias.detach(), "alln.rv0": rv0, "alln.x": x, "alln.y_eval": y_eval.detach(),
                "alln.y_train": y_train.detach(), "alln.rv1": an.running_var.detach().clone(), "alln.seed": seedg, "alln.gx": gx})

    path = os.path.join(GOLD, "norms_kat.npz")
    np.savez_compressed(path, **{k: v.detach().numpy() for k, v in out.items()})
    print(f"[norms] wrote {path} ({os.path.getsize(path)/1e3:.0f} kB), {len(out)} arrays; oracle == reference bit exact in explanation mode")


def golden_localisation(seed: int = 11):
    """Localisation scores (interpretability/analyses/localisation.py:306-388).  The reference computes them inline in
    `LocalisationAnalyser.analysis`, which needs the experiment / dataset stack and cannot be imported here; the vectors
    below are produced by the very ATen calls of those lines, written out literally (not through the oracle), and the
    oracle is checked against them."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, (nt, c, grid, cell, smooth, neg) in {"g2_s15": (4, 6, 2, 32, 15, False), "g3_s0": (9, 6, 3, 20, 0, False),
                                                   "g2_s3_neg": (4, 1, 2, 24, 3, True), "g2_zero": (4, 6, 2, 16, 5, False)}.items():
        h = grid * cell
        attributions_in = torch.randn(nt, c, h, h, generator=g) * torch.rand(nt, 1, h, h, generator=g)
        if name == "g2_zero":
            attributions_in[1] = -attributions_in[1].abs()          # a target without positive evidence: total = 0
        attributions = attributions_in.sum(1, keepdim=True)
        if smooth:
            attributions = F.avg_pool2d(attributions, smooth, stride=1, padding=(smooth - 1) // 2)
        if neg:
            attributions = -attributions
        attributions = attributions.clamp(min=0)
        single_shape = cell
        contribs = (F.avg_pool2d(attributions, single_shape, stride=single_shape).permute(0, 1, 3, 2)
                    .reshape(attributions.shape[0], -1))
        total = contribs.sum(1, keepdim=True)
        contribs = torch.where(total * contribs > 0, contribs / total, torch.zeros_like(contribs))
        assert torch.equal(O.localisation_scores(attributions_in, cell, smooth, neg), contribs), name
        out[name + ".attr"] = attributions_in.numpy()
        out[name + ".scores"] = contribs.numpy()
        out[name + ".args"] = np.array([cell, smooth, int(neg)], dtype=np.int64)
    path = os.path.join(GOLD, "localisation_kat.npz")
    np.savez_compressed(path, **out)
    print(f"[loc] wrote {path} ({os.path.getsize(path)/1e3:.0f} kB)")


def golden_gradient_to_image(seed: int = 3):
    """RGBA explanation images from the reference's own gradient_to_image (bcos/common.py:387-436), per image."""
    refload.load()
    import bcos.common as BC
    g = torch.Generator().manual_seed(seed)
    nb, S = 3, 48
    x3 = torch.rand(nb, 3, S, S, generator=g)
    x6 = torch.cat([x3, 1 - x3], 1)
    grad6 = torch.randn(nb, 6, S, S, generator=g) * torch.rand(nb, 1, S, S, generator=g)
    out = {"x6": x6.numpy(), "grad6": grad6.numpy()}
    for name, (smooth, pct) in {"default": (15, 99.5), "nosmooth_p90": (0, 90.0), "s3_p100": (3, 100.0)}.items():
        ref = np.stack([BC.gradient_to_image(x6[i], grad6[i], smooth=smooth, alpha_percentile=pct) for i in range(nb)])
        mine = O.gradient_to_image_batched(x6, grad6, smooth, pct).numpy()
        assert np.array_equal(ref, mine), (name, np.abs(ref - mine).max())      # same ATen calls: bit exact
        out[name + ".rgba"] = ref
        out[name + ".args"] = np.array([smooth, pct], dtype=np.float64)
    path = os.path.join(GOLD, "gradient_to_image_kat.npz")
    np.savez_compressed(path, **out)
    print(f"[g2i] wrote {path} ({os.path.getsize(path)/1e3:.0f} kB); oracle == reference bit exact")


def calibration_file(arch: str, batch: int = 16, seed: int = 100):
    """BN running variances for the synthetic benchmark checkpoint (weights seed 0), calibrated on `batch` synthetic
    images by the oracle (== the reference, see golden_resnet).  Shipped with the package so that bench.py needs no
    oracle/reference at run time: random-init B-cos nets collapse to ~1e-13 activations without it."""
    sd = synth.synth_state_dict(O.resnet_state_shapes(arch), 0)
    x6 = synth.to_bcos_input(synth.synth_images_u8(batch, 224, seed))
    om = O.OracleResNet(arch, sd)
    om.calibrate_bn(x6)
    keys = sorted(k for k in sd if k.endswith("running_var"))
    d = os.path.join(ROOT, "b-cosification_b200", "utils", "calib")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"{arch}_bnvar.npz")
    np.savez_compressed(path, bn_keys=np.array(keys), bn_sizes=np.array([sd[k].numel() for k in keys], dtype=np.int64),
                        bn_var=torch.cat([sd[k].flatten() for k in keys]).numpy(), weights_seed=np.int64(0),
                        images_seed=np.int64(seed), batch=np.int64(batch))
    print(f"[calib] wrote {path} ({os.path.getsize(path)/1e3:.0f} kB)")


def calibration_file_clip(batch: int = 8, seed: int = 100):
    """The same for the CLIP RN50 image encoder (bench.py --arch clip_rn50, BASELINE config 4)."""
    sd = synth.synth_state_dict(O.clip_rn_state_shapes(), 0)
    x6 = synth.to_bcos_input(synth.synth_images_u8(batch, 224, seed))
    om = O.OracleCLIPResNet(sd)
    om.calibrate_bn(x6)
    keys = sorted(k for k in sd if k.endswith("running_var"))
    path = os.path.join(ROOT, "b-cosification_b200", "utils", "calib", "clip_rn50_bnvar.npz")
    np.savez_compressed(path, bn_keys=np.array(keys), bn_sizes=np.array([sd[k].numel() for k in keys], dtype=np.int64),
                        bn_var=torch.cat([sd[k].flatten() for k in keys]).numpy(), weights_seed=np.int64(0),
                        images_seed=np.int64(seed), batch=np.int64(batch))
    print(f"[calib] wrote {path} ({os.path.getsize(path)/1e3:.0f} kB)")


def calibration_file_densenet(arch: str = "densenet121", batch: int = 8, seed: int = 100):
    """The same for DenseNet-121 (fused DenseNet plan, scripts/exp_densenet_plan.py)."""
    sd = synth.synth_state_dict(O.densenet_state_shapes(arch), 0)
    x6 = synth.to_bcos_input(synth.synth_images_u8(batch, 224, seed))
    om = O.OracleDenseNet(arch, sd)
    om.calibrate_bn(x6)
    keys = sorted(k for k in sd if k.endswith("running_var"))
    path = os.path.join(ROOT, "b-cosification_b200", "utils", "calib", f"{arch}_bnvar.npz")
    np.savez_compressed(path, bn_keys=np.array(keys), bn_sizes=np.array([sd[k].numel() for k in keys], dtype=np.int64),
                        bn_var=torch.cat([sd[k].flatten() for k in keys]).numpy(), weights_seed=np.int64(0),
                        images_seed=np.int64(seed), batch=np.int64(batch))
    print(f"[calib] wrote {path} ({os.path.getsize(path)/1e3:.0f} kB)")


def golden_groups(seed: int = 5):
    """Known-answer vectors of grouped B-cos convolutions (bcosconv2d.py:201-209, 224-229: per-group patch norms), by the reference classes."""
    refload.load()
    from bcos.modules.bcosconv2d import BcosConv2d
    from bcos.modules.bcosifyconv2d import BcosifyConv2d
    g = torch.Generator().manual_seed(4321 + seed)
    out = {}
    cases = [
        # name, cls, cin, cout, k, s, p, b, max_out, groups, H
        ("conv_bcos_g4", BcosConv2d, 32, 64, 3, 1, 1, 2, 1, 4, 10),
        ("conv_bcosify_g2_s2", BcosifyConv2d, 16, 32, 3, 2, 1, 2, 1, 2, 13),
        ("conv_bcos_g2_b1p5_mo2", BcosConv2d, 16, 32, 3, 1, 1, 1.5, 2, 2, 9),
        ("conv_bcos_depthwise", BcosConv2d, 8, 8, 3, 1, 1, 2, 1, 8, 8),
        ("conv_bcosify_g4_1x1", BcosifyConv2d, 64, 32, 1, 1, 0, 2, 1, 4, 7),
    ]
    for name, cls, cin, cout, k, s, p, b, mo, G, H in cases:
        mod = cls(cin, cout, kernel_size=k, stride=s, padding=p, b=b, max_out=mo, groups=G)
        w = torch.randn(mod.linear.weight.shape, generator=g) * 0.2
        mod.linear.weight.data = w.clone()
        x = torch.randn(2, cin, H, H, generator=g)
        y = mod(x)
        seedg = torch.randn(y.shape, generator=g)
        mod.set_explanation_mode(True)
        xg = x.clone().requires_grad_(True)
        (gx,) = torch.autograd.grad((mod(xg) * seedg).sum(), [xg])
        norm = mod.calc_patch_norms(x)
        assert norm.shape[1] == cout
        out.update({f"{name}.w": w, f"{name}.x": x, f"{name}.y": y.detach(), f"{name}.seed": seedg, f"{name}.gx": gx, f"{name}.norm": norm.detach(),
                    f"{name}.meta": torch.tensor([cin, cout, k, s, p, b, mo, float(cls is BcosConv2d), G], dtype=torch.float64)})
    path = os.path.join(GOLD, "modules_groups_kat.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print(f"[groups] wrote {path} ({os.path.getsize(path)/1e3:.1f} kB), {len(cases)} cases")


def golden_native(seed: int = 0):
    """Native B-cos-v2 variants of the reference's model zoo (tests/native_variants.py), run by the reference itself."""
    refload.load()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import native_variants as V
    u8 = synth.synth_images_u8(V.BATCH, V.IMAGE, seed + 21)
    x6 = synth.to_bcos_input(u8)
    for name in V.VARIANTS:
        t0 = time.time()
        torch.manual_seed(0)
        m = V.build(name)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(synth.synth_state_dict(shapes, seed), strict=True)
        if any(isinstance(b, nn.BatchNorm2d) for b in m.modules()):
            reference_calibrate(m, x6)
        m.eval()
        logits, cmap = V.explain_batched(m, x6)
        m64 = V.build(name)
        m64.load_state_dict(m.state_dict(), strict=True)
        m64 = m64.double().eval()
        l64, c64 = V.explain_batched(m64, x6.double())
        floor = O.parity_metrics(logits, cmap, l64.float(), c64.float())
        cal = {k: v for k, v in m.state_dict().items() if k.endswith("running_var")}
        keys = sorted(cal)
        out = dict(images_u8=u8, logits=logits.numpy(), contribution_map=cmap.numpy(), logits_fp64=l64.numpy(),
                   contribution_map_fp64=c64.float().numpy(), seed=np.int64(seed), num_params=np.int64(sum(v.numel() for v in m.state_dict().values())),
                   fp32_noise_floor_maxabs_over_range=np.float64(floor["map_maxabs_over_range"]),
                   bn_keys=np.array(keys), bn_sizes=np.array([cal[k].numel() for k in keys], dtype=np.int64),
                   bn_var=(torch.cat([cal[k].flatten() for k in keys]).numpy() if keys else np.zeros(0, np.float32)))
        path = os.path.join(GOLD, f"native_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"[native {name}] wrote {path} ({os.path.getsize(path)/1e3:.1f} kB) in {time.time()-t0:.1f}s; logits std {logits.std():.4f} "
              f"pred {logits.argmax(1).tolist()} map range {float(cmap.max()-cmap.min()):.3e}; fp32 vs fp64: {floor}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="modules,resnet18,resnet50")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(True)
    which = args.only.split(",")
    if "modules" in which:
        golden_modules()
    if "resnet18" in which:
        golden_resnet("resnet18", 8)
    if "resnet50" in which:
        golden_resnet("resnet50", 4)
    if "densenet121" in which:
        golden_densenet("densenet121", 2)
    if "vit_ti" in which:
        golden_vit("simple_vit_ti_patch16_224", 2)
    if "vit_b" in which:
        golden_vit("simple_vit_b_patch16_224", 2)
    if "clip_rn50" in which:
        golden_clip_rn50(2)
    if "clip_vit" in which:
        golden_clip_vit(2)
    if "loc" in which:
        golden_localisation()
    if "norms" in which:
        golden_norms()
    if "clip_unpool" in which:
        golden_clip_unpool()
    if "g2i" in which:
        golden_gradient_to_image()
    if "calib" in which:
        calibration_file("resnet18")
        calibration_file("resnet50")
    if "calib_clip" in which:
        calibration_file_clip()
    if "calib_densenet" in which:
        calibration_file_densenet()
    if "native" in which:
        golden_native()
    if "groups" in which:
        golden_groups()
