"""CPU: the oracle (oracle/bcos_oracle.py) against the committed golden vectors that the REFERENCE produced
(oracle/make_golden.py), and - where /root/reference exists - against the live reference modules."""
import os

import numpy as np
import pytest
import torch

import bcos_oracle as OR
import refload
from bcos_b200.utils import synth


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "modules_kat.npz"))


CONV = ["conv_bcosify_3x3", "conv_bcosify_3x3_s2", "conv_bcosify_1x1", "conv_bcosify_1x1_s2", "conv_bcosify_7x7_s2",
        "conv_bcos_3x3_normed", "conv_bcos_b1p5", "conv_bcos_b2p5_mo2", "conv_bcos_b2_mo3", "conv_bcos_b1"]


@pytest.mark.parametrize("name", CONV)
def test_conv_known_answers(kat, name):
    cin, cout, k, s, p, b, mo, normed = kat[name + ".meta"].tolist()
    x, w = _t(kat[name + ".x"]), _t(kat[name + ".w"])
    y = OR.bcos_conv2d(x, w, None, int(s), int(p), b=b, max_out=int(mo), normalize_weight=bool(normed))
    assert torch.equal(y, _t(kat[name + ".y"]))                     # bit exact: same ATen calls in the same order
    xg = x.clone().requires_grad_(True)
    ye = OR.bcos_conv2d(xg, w, None, int(s), int(p), b=b, max_out=int(mo), detach=True, normalize_weight=bool(normed))
    (gx,) = torch.autograd.grad((ye * _t(kat[name + ".seed"])).sum(), [xg])
    assert torch.allclose(gx, _t(kat[name + ".gx"]), rtol=1e-6, atol=1e-7)
    if b != 1:
        n = OR.patch_norms(x, int(k), int(s), int(p))
        assert torch.equal(n, _t(kat[name + ".norm"]))
        # reference's own in-code cross-check (bcosconv2d.py:233-250): slow == fast to 6e-6
        slow = OR.patch_norms_slow(x, w[:1, :, :, :] if False else w, int(s), int(p), 1, 1)[:, :1]
        assert torch.allclose(slow, n, rtol=0, atol=6e-6)


@pytest.mark.parametrize("name", ["lin_bcosify", "lin_bcos_normed", "lin_bcos_b1p5_mo2"])
def test_linear_known_answers(kat, name):
    fin, fout, b, mo, normed = kat[name + ".meta"].tolist()
    x, w = _t(kat[name + ".x"]), _t(kat[name + ".w"])
    y = OR.bcos_linear(x, w, None, b=b, max_out=int(mo), normalize_weight=bool(normed))
    assert torch.allclose(y, _t(kat[name + ".y"]), rtol=1e-6, atol=1e-7)
    xg = x.clone().requires_grad_(True)
    ye = OR.bcos_linear(xg, w, None, b=b, max_out=int(mo), detach=True, normalize_weight=bool(normed))
    (gx,) = torch.autograd.grad((ye * _t(kat[name + ".seed"])).sum(), [xg])
    assert torch.allclose(gx, _t(kat[name + ".gx"]), rtol=1e-5, atol=1e-7)


def test_norms_and_small_layers(kat):
    x = _t(kat["bnu.x"])
    y = OR.batch_norm_uncentered_2d(x, _t(kat["bnu.rv0"]), _t(kat["bnu.w"]), _t(kat["bnu.b"]))
    assert torch.equal(y, _t(kat["bnu.y_eval"]))
    rv = _t(kat["bnu.rv0"]).clone()
    yt = OR.batch_norm_uncentered_2d(x, rv, _t(kat["bnu.w"]), _t(kat["bnu.b"]), training=True, momentum=0.1)
    assert torch.equal(yt, _t(kat["bnu.y_train"])) and torch.equal(rv, _t(kat["bnu.rv1"]))
    w, b2 = OR.bn_uncentered_from_standard(_t(kat["bnfold.w"]), _t(kat["bnfold.b"]), _t(kat["bnfold.rm"]), _t(kat["bnfold.rv"]), 1e-5)
    assert torch.allclose(b2, _t(kat["bnfold.bias_folded"]))
    assert torch.allclose(OR.batch_norm_uncentered_2d(x, _t(kat["bnfold.rv"]), w, b2), _t(kat["bnfold.y"]), rtol=1e-5, atol=1e-6)
    # DetachableLayerNorm: plain and explanation mode
    lx = _t(kat["ln.x"])
    assert torch.allclose(OR.layer_norm_detachable(lx, _t(kat["ln.w"]), None, 1e-5, False), _t(kat["ln.y"]), rtol=1e-6, atol=1e-6)
    xg = lx.clone().requires_grad_(True)
    ye = OR.layer_norm_detachable(xg, _t(kat["ln.w"]), None, 1e-5, True)
    assert torch.allclose(ye, _t(kat["ln.y_explain"]), rtol=1e-6, atol=1e-6)
    (gx,) = torch.autograd.grad((ye * _t(kat["ln.seed"])).sum(), [xg])
    assert torch.allclose(gx, _t(kat["ln.gx"]), rtol=1e-5, atol=1e-6)
    gx_ = _t(kat["gelu.x"]).clone().requires_grad_(True)
    ge = OR.gelu_detachable(gx_, True)
    assert torch.allclose(ge, _t(kat["gelu.y"]), rtol=1e-6, atol=1e-7)
    (gg,) = torch.autograd.grad((ge * _t(kat["gelu.seed"])).sum(), [gx_])
    assert torch.allclose(gg, _t(kat["gelu.gx"]), rtol=1e-6, atol=1e-7)
    assert torch.equal(OR.logit_layer(_t(kat["logit.x"]), 2.0, -1.5), _t(kat["logit.y"]))


# name -> (kind, groups, centred, channels)
NORM_CASES = {"gnu_g4": ("group", 4, False), "gnu_g4_odd": ("group", 4, False), "gnu_layer": ("group", 1, False),
              "gnu_instance": ("group", 12, False), "dgn_g2": ("group", 2, True), "dgn_layer": ("group", 1, True),
              "dgn_instance_odd": ("group", 6, True), "pnu": ("pos", 0, False), "pnu_wide": ("pos", 0, False),
              "dpn": ("pos", 0, True), "dpn_bias": ("pos", 0, True)}


@pytest.mark.parametrize("name", sorted(NORM_CASES))
def test_group_and_position_norm_known_answers(golden_dir, name):
    """oracle group / position norms against the vectors the reference classes produced (tests/golden/norms_kat.npz)."""
    kat = np.load(os.path.join(golden_dir, "norms_kat.npz"))
    kind, groups, centred = NORM_CASES[name]
    w = _t(kat[name + ".w"])
    b = _t(kat[name + ".b"]) if name + ".b" in kat.files else None

    def run(x, detach):
        if kind == "group":
            return OR.group_norm_detachable(x, groups, w, b, 1e-5, detach, centred)
        return OR.position_norm_detachable(x, w, b, 1e-5, detach, centred)
    x = _t(kat[name + ".x"])
    assert torch.allclose(run(x, False), _t(kat[name + ".y"]), rtol=1e-6, atol=1e-6)
    xg = x.clone().requires_grad_(True)
    ye = run(xg, True)
    (gx,) = torch.autograd.grad((ye * _t(kat[name + ".seed"])).sum(), [xg])
    assert torch.equal(ye.detach(), _t(kat[name + ".y_explain"])) and torch.equal(gx, _t(kat[name + ".gx"]))


def test_all_norm_known_answers(golden_dir):
    kat = np.load(os.path.join(golden_dir, "norms_kat.npz"))
    x, w, b = _t(kat["alln.x"]), _t(kat["alln.w"]), _t(kat["alln.b"])
    assert torch.equal(OR.all_norm_uncentered_2d(x, _t(kat["alln.rv0"]), w, b), _t(kat["alln.y_eval"]))
    rv = _t(kat["alln.rv0"]).clone()
    yt = OR.all_norm_uncentered_2d(x, rv, w, b, training=True, momentum=0.1)
    assert torch.equal(yt, _t(kat["alln.y_train"])) and torch.equal(rv, _t(kat["alln.rv1"]))


def _golden_state(arch, gold):
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    return sd


@pytest.mark.parametrize("arch,batch", [("resnet18", 8), ("resnet50", 4)])
def test_resnet_golden(golden_dir, arch, batch):
    gold = np.load(os.path.join(golden_dir, f"{arch}_b{batch}.npz"))
    assert np.array_equal(gold["images_u8"], synth.synth_images_u8(batch, 224, int(gold["seed"])))  # workload is reproducible
    sd = _golden_state(arch, gold)
    nb = 2                                                   # images are independent in eval mode
    x6 = synth.to_bcos_input(gold["images_u8"][:nb])
    e = OR.explain_batched(OR.OracleResNet(arch, sd).forward, x6)
    m = OR.parity_metrics(e["logits"], e["contribution_map"], _t(gold["logits"][:nb]), _t(gold["contribution_map"][:nb]))
    # oneDNN picks batch-size dependent kernels, so sub-batches agree to fp32 noise (amplified by the random deep net)
    assert m["argmax_equal"] and m["logit_rel_err"] < 1e-5 and m["map_cos_min"] > 0.99999, m
    # completeness of the dynamic-linear decomposition (bias-free net): sum(contributions) ~ logit - logit_bias, up to the
    # input-normalisation offset (SURVEY.md section 8c invariant 4)
    tot = e["contribution_map"].flatten(1).sum(1)
    top = e["logits"].max(1).values - OR.LOGIT_BIAS_1000
    assert torch.all((tot - top).abs() < 0.5 * top.abs() + 1e-3)


def test_batched_explain_equals_per_sample(golden_dir):
    gold = np.load(os.path.join(golden_dir, "resnet18_b8.npz"))
    sd = _golden_state("resnet18", gold)
    x6 = synth.to_bcos_input(gold["images_u8"][:3])
    om = OR.OracleResNet("resnet18", sd)
    eb = OR.explain_batched(om.forward, x6)
    for i in range(3):
        ei = OR.explain_batched(om.forward, x6[i:i + 1])
        assert torch.allclose(ei["contribution_map"][0], eb["contribution_map"][i], rtol=1e-3, atol=1e-9)


@pytest.mark.skipif(not refload.live(), reason="reference checkout not present (GPU box)")
def test_oracle_pinned_to_live_reference():
    refload.load()
    from bcos.modules.bcosifyconv2d import BcosifyConv2d
    from bcos.modules.bcoslinear import BcosLinear
    from bcos.modules.norms.uncentered_norms import BatchNormUncentered2d
    g = torch.Generator().manual_seed(99)
    for (cin, cout, k, s, p, b, mo) in [(8, 12, 3, 1, 1, 2, 1), (6, 8, 7, 2, 3, 2, 1), (8, 6, 1, 2, 0, 1.7, 2)]:
        mod = BcosifyConv2d(cin, cout, kernel_size=k, stride=s, padding=p, b=b, max_out=mo)
        x = torch.randn(2, cin, 11, 11, generator=g)
        for detach in (False, True):
            mod.set_explanation_mode(detach)
            xr = x.clone().requires_grad_(True)
            yr = mod(xr)
            xo = x.clone().requires_grad_(True)
            yo = OR.bcos_conv2d(xo, mod.linear.weight.detach(), None, s, p, b=b, max_out=mo, detach=detach)
            assert torch.equal(yr.detach(), yo.detach())
            sg = torch.randn(yr.shape, generator=g)
            assert torch.allclose(torch.autograd.grad((yr * sg).sum(), [xr])[0], torch.autograd.grad((yo * sg).sum(), [xo])[0],
                                  rtol=1e-6, atol=1e-7)
    lin = BcosLinear(20, 10, b=2)
    x = torch.randn(4, 20, generator=g)
    assert torch.allclose(lin(x), OR.bcos_linear(x, lin.linear.weight.detach(), None, normalize_weight=True), rtol=1e-6, atol=1e-7)
    bn = BatchNormUncentered2d(6).eval()
    x = torch.randn(2, 6, 4, 4, generator=g)
    assert torch.equal(bn(x), OR.batch_norm_uncentered_2d(x, bn.running_var, bn.weight, bn.bias))


def test_densenet_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "densenet121_b2.npz"))
    sd = synth.synth_state_dict(OR.densenet_state_shapes("densenet121"), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    x6 = synth.to_bcos_input(gold["images_u8"][:1])
    e = OR.explain_batched(OR.OracleDenseNet("densenet121", sd).forward, x6)
    m = OR.parity_metrics(e["logits"], e["contribution_map"], _t(gold["logits"][:1]), _t(gold["contribution_map"][:1]))
    assert m["argmax_equal"] and m["logit_rel_err"] < 1e-5 and m["map_cos_min"] > 0.99999, m


def test_vit_golden(golden_dir):
    arch = "simple_vit_ti_patch16_224"
    gold = np.load(os.path.join(golden_dir, f"{arch}_b2.npz"))
    sd = synth.synth_state_dict(OR.vit_state_shapes(arch), int(gold["seed"]))
    x6 = synth.to_bcos_input(gold["images_u8"][:1])
    e = OR.explain_batched(OR.OracleViT(arch, sd).forward, x6)
    m = OR.parity_metrics(e["logits"], e["contribution_map"], _t(gold["logits"][:1]), _t(gold["contribution_map"][:1]))
    assert m["argmax_equal"] and m["logit_rel_err"] < 1e-5 and m["map_cos_min"] > 0.99999, m
    # the patch-embedding weight doubling of bcosify_vit.add_channels: [W/2, -W/2] per pixel
    w3 = torch.arange(2 * 12, dtype=torch.float32).view(2, 12)
    w6 = OR.vit_add_channels_linear(w3)
    assert w6.shape == (2, 24) and torch.equal(w6[:, :3], w3[:, :3] / 2) and torch.equal(w6[:, 3:6], -w3[:, :3] / 2)


def test_vit_b_golden(golden_dir):
    arch = "simple_vit_b_patch16_224"
    gold = np.load(os.path.join(golden_dir, f"{arch}_b2.npz"))
    sd = synth.synth_state_dict(OR.vit_state_shapes(arch), int(gold["seed"]))
    x6 = synth.to_bcos_input(gold["images_u8"][:1])
    e = OR.explain_batched(OR.OracleViT(arch, sd).forward, x6)
    m = OR.parity_metrics(e["logits"], e["contribution_map"], _t(gold["logits"][:1]), _t(gold["contribution_map"][:1]))
    assert m["argmax_equal"] and m["logit_rel_err"] < 1e-5 and m["map_cos_min"] > 0.99999, m


def test_clip_rn50_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "clip_rn50_b2.npz"))
    sd = synth.synth_state_dict(OR.clip_rn_state_shapes(), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    x6 = synth.to_bcos_input(gold["images_u8"][:1])
    e = OR.explain_cosine(OR.OracleCLIPResNet(sd).forward, x6, OR.clip_seed_direction(1024, int(gold["seed"])))
    emb, cm = _t(gold["embedding"][:1]), _t(gold["contribution_map"][:1])
    assert ((e["embedding"] - emb).abs().max() / emb.abs().max()).item() < 1e-4
    assert torch.nn.functional.cosine_similarity(e["contribution_map"].flatten(1), cm.flatten(1)).min().item() > 0.9999


def test_clip_vit_golden(golden_dir):
    """B-cos CLIP ViT image encoder (CLIP/clip/model.py:206-241 through bcosify.py with clip_kd): oracle == what the reference
    produced (embedding and contribution map of cos(embedding, fixed unit vector)), image 0 of the golden batch."""
    gold = np.load(os.path.join(golden_dir, "clip_vit_b32_b2.npz"))
    res, patch, width, layers, heads, out_dim = gold["geometry"].tolist()
    sd = synth.synth_state_dict(OR.clip_vit_state_shapes(res, patch, width, layers, out_dim), int(gold["seed"]))
    x6 = synth.to_bcos_input(gold["images_u8"][:1])
    e = OR.explain_cosine(OR.OracleCLIPViT(sd, heads).forward, x6, OR.clip_seed_direction(out_dim, int(gold["seed"])))
    emb, cm = _t(gold["embedding"][:1]), _t(gold["contribution_map"][:1])
    assert ((e["embedding"] - emb).abs().max() / emb.abs().max()).item() < 1e-4
    assert torch.nn.functional.cosine_similarity(e["contribution_map"].flatten(1), cm.flatten(1)).min().item() > 0.9999


def test_gradient_to_image_known_answers(golden_dir):
    """oracle restatement of gradient_to_image (bcos/common.py:387-436) == what the reference produced"""
    kat = np.load(os.path.join(golden_dir, "gradient_to_image_kat.npz"))
    x6, grad6 = _t(kat["x6"]), _t(kat["grad6"])
    for name in ("default", "nosmooth_p90", "s3_p100"):
        smooth, pct = kat[name + ".args"].tolist()
        out = OR.gradient_to_image_batched(x6, grad6, int(smooth), float(pct))
        assert torch.equal(out, _t(kat[name + ".rgba"])), name
        if refload.live():
            refload.load()
            import bcos.common as BC
            ref = np.stack([BC.gradient_to_image(x6[i], grad6[i], smooth=int(smooth), alpha_percentile=float(pct))
                            for i in range(x6.shape[0])])
            assert np.array_equal(out.numpy(), ref), name


LOC_CASES = ["g2_s15", "g3_s0", "g2_s3_neg", "g2_zero"]


@pytest.mark.parametrize("name", LOC_CASES)
def test_localisation_scores_known_answers(golden_dir, name):
    kat = np.load(os.path.join(golden_dir, "localisation_kat.npz"))
    cell, smooth, neg = kat[name + ".args"].tolist()
    got = OR.localisation_scores(_t(kat[name + ".attr"]), cell, smooth, bool(neg))
    assert torch.equal(got, _t(kat[name + ".scores"]))
    assert torch.allclose(got.sum(1)[got.sum(1) > 0], torch.ones(1))          # fractions of the positive evidence


def test_clip_unpool_text_localisation_golden(golden_dir):
    """attn_unpool head + text-localisation target: oracle vs the reference run stored in the golden."""
    base = np.load(os.path.join(golden_dir, "clip_rn50_b2.npz"))
    gold = np.load(os.path.join(golden_dir, "clip_rn50_unpool_b1.npz"))
    sd = synth.synth_state_dict(OR.clip_rn_state_shapes(attn_unpool=True), int(gold["seed"]))
    off = 0
    for k, n in zip(base["bn_keys"].tolist(), base["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(base["bn_var"][off:off + n].copy()); off += n
    x6 = synth.to_bcos_input(base["images_u8"][:1])
    zw = OR.clip_seed_direction(1024, int(gold["seed"])).unsqueeze(1)
    om = OR.OracleCLIPResNet(sd)
    assert om.unpool
    xb = x6.clone().requires_grad_(True)
    tok = om.forward(xb, detach=True)
    assert tok.shape == (49, 1, 1024) and torch.allclose(tok.norm(dim=-1), torch.ones(49, 1), atol=1e-5)
    assert ((tok.detach() - _t(gold["tokens"])).abs().max() / _t(gold["tokens"]).abs().max()).item() < 1e-5
    for p in (1, 2, 0):
        tgt = OR.text_localisation_target(tok, zw, True, p)
        (g,) = torch.autograd.grad(tgt.sum(), [xb], retain_graph=True)
        cmap = (x6 * g).sum(1)[0]
        ref = _t(gold[f"p{p}.contribution_map"])
        assert abs(tgt.item() - float(gold[f"p{p}.target"][0])) < 1e-6
        assert ((cmap - ref).abs().max() / ref.abs().max()).item() < 1e-4, p


@pytest.mark.skipif(not refload.live(), reason="reference checkout not present (GPU box)")
def test_oracle_train_step_pinned_to_live_reference():
    """the oracle's fine-tuning step (train-mode forward, UniformOffLabelsBCE, autograd, AGC, first AdamW step)
    against the reference's own modules, loss and AGC code on the same weights and batch."""
    import make_golden as MG
    import importlib.util
    arch, nb, S = "resnet18", 4, 64
    m = MG.build_reference_resnet(arch)
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 0)
    x6 = synth.to_bcos_input(synth.synth_images_u8(nb, S, 1))
    labels = torch.tensor([3, 500, 999, 0])
    m.load_state_dict(sd, strict=True)
    MG.reference_calibrate(m, x6)
    sd_cal = {k: v.detach().clone() for k, v in m.state_dict().items()}
    om = OR.OracleResNet(arch, {k: v.clone() for k, v in sd_cal.items()})
    ref = OR.train_step_reference(om, x6, labels, lr=1e-4, agc_clip=0.01)

    def load_ref(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(refload.REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    import bcos.modules.losses as losses              # the reference's file (relative import of .common resolves through refload)
    agc = load_ref("_ref_agc", "bcos/training/agc.py")
    m.train()
    params = {k: p for k, p in m.named_parameters()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=0.0)
    out = m(x6)
    loss = losses.UniformOffLabelsBCEWithLogitsLoss()(out, labels)
    loss.backward()
    assert torch.allclose(loss.detach(), ref["loss"], rtol=1e-6, atol=1e-8)
    assert set(ref["grads"]) == set(params)
    for k, p in params.items():
        assert torch.allclose(p.grad, ref["grads"][k], rtol=1e-4, atol=1e-9), k
    agc.adaptive_clip_grad_(list(params.values()), clip_factor=0.01)
    opt.step()
    for k, p in params.items():
        assert torch.allclose(p.detach(), ref["weights"][k], rtol=0, atol=1e-6), k      # 1 % of the 1e-4 step
    for k, v in m.state_dict().items():
        if k.endswith("running_var"):
            assert torch.allclose(v, ref["running_var"][k], rtol=1e-5, atol=1e-8), k


@pytest.mark.skipif(not refload.live(), reason="reference checkout not present (GPU box)")
def test_staged_reference_archive_is_the_reference_and_runs():
    """oracle/stage_ref.py: the archive the GPU box's CPU legs import holds the checkout's files byte for byte, and the model the
    reference builds out of it (no checkout in sight) agrees with the oracle port on a small case."""
    import hashlib
    import json
    import subprocess
    import sys as _sys
    import zipfile
    import stage_ref
    stage_ref.stage(quiet=True)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    man = json.load(open(stage_ref.MANIFEST))
    assert man["count"] >= 20 and "bcos/modules/bcosconv2d.py" in man["files"] and "bcosify.py" in man["files"]
    with zipfile.ZipFile(stage_ref.ARCHIVE) as z:
        assert sorted(z.namelist()) == sorted(man["files"])
        for rel, digest in man["files"].items():
            data = z.read(rel)
            assert hashlib.sha256(data).hexdigest() == digest
            assert data == open(os.path.join(refload.LIVE, rel), "rb").read(), rel
    code = """
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch, refload
assert refload.source() == "archive", refload.source()
import make_golden as G, bcos_oracle as OR
from bcos_b200.utils import synth
m = G.build_reference_resnet("resnet18")
assert ".zip" in sys.modules["bcos.modules.bcosconv2d"].__file__
sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 3)
m.load_state_dict(sd, strict=True); m.eval()
x6 = synth.to_bcos_input(synth.synth_images_u8(2, 64, 5))
out, grad, cmap = G.reference_explain_batched(m, x6)
oe = OR.explain_batched(OR.OracleResNet("resnet18", sd).forward, x6)
pm = OR.parity_metrics(oe["logits"], oe["contribution_map"], out, cmap)
assert pm["argmax_equal"] and pm["logit_rel_err"] < 1e-5 and pm["map_cos_min"] > 0.99999, pm
print("ok")
""" % (root, os.path.join(root, "oracle"))
    env = dict(os.environ, BCOS_REFERENCE_ROOT="/nonexistent")
    r = subprocess.run([_sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
