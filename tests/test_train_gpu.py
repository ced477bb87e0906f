"""-m gpu: the fine-tuning step (SURVEY 8f row 2 / BASELINE config 5) through the C ABI.

Kernel by kernel against the same arithmetic written with torch on the same (16-bit rounded) inputs, then one whole step of a
B-cosified ResNet-18 against the fp32 oracle (`oracle.train_step_reference`: forward in train mode, UniformOffLabelsBCE loss,
autograd gradients, AGC, AdamW).  Whole-step tolerances are those of 16-bit operands on a random-init deep net: the forward
activations already differ from fp32 by ~0.5 % (fp16) / ~4 % (bf16), and every gradient is a sum of mixed-sign terms.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import bcos_oracle as OR
from bcos_b200 import _lib as L
from bcos_b200.engine import ResNetTrainPlan
from bcos_b200.engine import ops as O
from bcos_b200.engine import pack as P
from bcos_b200.engine.train import WgradOp
from bcos_b200.utils import synth

pytestmark = pytest.mark.gpu
DT = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


@pytest.mark.parametrize("case", [
    dict(c=64, o=128, k=3, s=1, p=1, h=14, nb=3),
    dict(c=128, o=64, k=3, s=2, p=1, h=16, nb=2),
    dict(c=256, o=512, k=1, s=2, p=0, h=14, nb=2),
    dict(c=64, o=1000, k=1, s=1, p=0, h=7, nb=5),
    dict(c=32, o=64, k=4, s=1, p=2, h=12, nb=2, pad_hi=1, kch=32),
])
@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
def test_wgrad_matches_torch(bcosk_lib, case, dtype):
    """bcosk_wgrad (tcgen05, MN-major operands, split-K) == autograd's conv weight gradient on the same 16-bit inputs"""
    torch.manual_seed(0)
    c, o, k, s, p, h, nb = (case[x] for x in ("c", "o", "k", "s", "p", "h", "nb"))
    pad_hi, kch = case.get("pad_hi", p), case.get("kch", 64)
    dt = DT[dtype]
    x = torch.randn(nb, h, h, c, device="cuda").to(dt)
    oh = (h + p + pad_hi - k) // s + 1
    g = (torch.randn(nb, oh, oh, o, device="cuda") * 0.1).to(dt)
    taps = P.conv_taps(k, k)
    cpt = (c + kch - 1) // kch
    fwd = O.IgemmOp(name="w", a=x, b=torch.zeros(o, len(taps) * cpt * kch, device="cuda", dtype=dt), n=o, lo=(-p, -p),
                    up=(pad_hi - (k - 1), pad_hi - (k - 1)), stride=(s, s), op=oh, oq=oh, kch=kch, chunks_per_tap=cpt, taps=taps,
                    seg_a_choff=[0], dtype=L.DTYPE_CODE[dtype], y=torch.zeros(1, device="cuda"))
    dw = torch.zeros(o * fwd.ktot, device="cuda")
    WgradOp("w", fwd, g.view(-1, o), dw).run()
    torch.cuda.synchronize()
    xn = x.float().permute(0, 3, 1, 2)
    xp = F.pad(xn, (p, pad_hi, p, pad_hi))
    ref = torch.nn.grad.conv2d_weight(xp, (o, c, k, k), g.float().permute(0, 3, 1, 2), stride=s, padding=0)   # [o, c, k, k]
    ref_packed = torch.zeros(o, len(taps), cpt * kch, device="cuda")
    ref_packed[:, :, :c] = ref.permute(0, 2, 3, 1).reshape(o, k * k, c)
    got = dw.view(o, len(taps), cpt * kch)
    assert _rel(got, ref_packed) < 2e-5, _rel(got, ref_packed)
    assert float(got[:, :, c:].abs().max()) == 0.0 if cpt * kch > c else True


@pytest.mark.parametrize("C,M,relu,use_bn", [(64, 1000, True, True), (512, 77, True, True), (1000, 98, False, False), (256, 640, False, True)])
def test_train_backward_kernels_match_torch(bcosk_lib, C, M, relu, use_bn):
    """train_bwd_reduce / bnu_bwd_finalize / train_bwd_apply == autograd through [B-cos scale, batch-stat norm, ReLU]"""
    torch.manual_seed(1)
    dt, dtc = torch.bfloat16, L.DTYPE_CODE["bf16"]
    dev = "cuda"
    lin = torch.randn(M, C, device=dev)
    n = torch.rand(M, device=dev) + 0.5
    s16 = (lin.abs() / n[:, None]).to(dt)
    out16 = (lin * lin.abs() / n[:, None]).to(dt)
    ga = (torch.randn(M, C, device=dev) * 0.01).to(dt)
    gb = (torch.randn(M, C, device=dev) * 0.01).to(dt)
    tn = torch.randn(M, device=dev) * 0.01
    w = torch.rand(C, device=dev) + 0.5
    # forward statistics with our kernels
    sums = torch.empty(37, 2 * C, device=dev)
    alpha, mean, rstd = (torch.empty(C, device=dev) for _ in range(3))
    rv = torch.ones(C, device=dev)
    L.bnu_stats_nhwc(out16, M, C, dtc, sums)
    L.bnu_finalize(sums, M, C, w, 1e-5, 0.1, rv, alpha, mean, rstd)
    o32 = out16.float()
    var = o32.var(0, unbiased=False)
    assert _rel(mean, o32.mean(0)) < 1e-5 and _rel(rstd, 1 / (var + 1e-5).sqrt()) < 1e-5
    assert _rel(rv, 0.9 * torch.ones(C, device=dev) + 0.1 * var) < 1e-5
    z16 = torch.empty(M, C, device=dev, dtype=dt)
    sq = torch.empty(M, device=dev)
    res16 = (torch.randn(M, C, device=dev) * 0.5).to(dt)
    L.bnu_apply_nhwc(out16, M, C, alpha, res16, relu, z16, sq, dtc)
    zref = o32 * alpha + res16.float()
    zref = zref.clamp(min=0) if relu else zref
    assert _rel(z16.float(), zref) < 5e-3 and _rel(sq, z16.float().pow(2).sum(1)) < 1e-5
    # torch reference of the backward on the same tensors
    g_in = ga.float() + gb.float() + z16.float() * tn[:, None]
    g_y = g_in * (z16.float() > 0) if relu else g_in
    if use_bn:
        S = (g_y * o32).sum(0)
        g_out = g_y * alpha + (o32 - mean) * (-(rstd ** 3) * w * S / M)
    else:
        g_out = g_y
    g_lin_ref = 2 * g_out * s16.float()
    inv_n = 1.0 / n
    gnt_ref = -(g_out * o32).sum(1) * inv_n * inv_n
    s_part = torch.empty(53, C, device=dev)
    kcoef, gw, s_red = torch.empty(C, device=dev), torch.empty(C, device=dev), torch.empty(C, device=dev)
    if use_bn:
        L.train_bwd_reduce(ga, False, gb, z16, tn, relu, out16, False, M, C, s_part, dtc)
        L.bnu_bwd_finalize(s_part, rstd, w, M, C, kcoef, gw, s_red)
        assert _rel(s_red, S) < 1e-4 and _rel(gw, S * rstd) < 1e-4
    g_lin = torch.empty(M, C, device=dev, dtype=dt)
    gnt = torch.empty(M, device=dev)
    g_y16 = torch.empty(M, C, device=dev, dtype=dt)
    L.train_bwd_apply(ga, False, gb, z16, tn, relu, out16, False, s16, alpha if use_bn else None, kcoef if use_bn else None,
                      mean if use_bn else None, inv_n, M, C, g_lin, gnt, g_y16, dtc)
    torch.cuda.synchronize()
    assert _rel(g_lin.float(), g_lin_ref) < 5e-3 and _rel(gnt, gnt_ref) < 1e-4 and _rel(g_y16.float(), g_y) < 5e-3


def test_bcos_backward_formula_is_autograd(bcosk_lib):
    """the arithmetic the kernels implement (g_lin = 2 s g_out, patch-norm path x * T, batch-stat term) IS autograd through
    the reference formulas: checked in fp32 on the CPU with the oracle's own functions"""
    torch.manual_seed(2)
    x = torch.randn(2, 8, 6, 6, requires_grad=True)
    w = torch.randn(16, 8, 3, 3) * 0.2
    out = OR.bcos_conv2d(x, w, None, 2, 1, b=2, detach=False)
    g = torch.randn_like(out)
    (gx_ref,) = torch.autograd.grad((out * g).sum(), [x])
    with torch.no_grad():
        lin = F.conv2d(x, w, None, 2, 1)
        n = OR.patch_norms(x, (3, 3), 2, 1, 1, 16)
        s = lin.abs() / n
        g_lin = 2 * g * s
        gx_direct = torch.nn.grad.conv2d_input(x.shape, w, g_lin, 2, 1)
        gnt = -(g * (lin * s)).sum(1, keepdim=True) / (n * n)
        T = F.conv_transpose2d(gnt, torch.ones(1, 1, 3, 3), stride=2, padding=1, output_padding=(x.shape[2] + 2 - 3) % 2)
        gx = gx_direct + x * T
    assert _rel(gx, gx_ref) < 1e-5
    # the device kernel of the transposed sum-pool
    tn = torch.empty(2 * 6 * 6, device="cuda")
    L.sumpool_transpose(gnt.cuda().contiguous().view(-1), 2, 6, 6, 3, 2, 1, gnt.shape[2], gnt.shape[3], False, tn)
    assert _rel(tn.cpu().view(2, 1, 6, 6), T) < 1e-6


def test_loss_and_optimizer_kernels_match_oracle(bcosk_lib):
    torch.manual_seed(3)
    N, C, npix = 6, 1000, 4
    logits = (torch.randn(N, C) * 2 - 5).cuda()
    labels = torch.randint(0, C, (N,))
    loss = torch.zeros(1, device="cuda")
    g_log = torch.empty(N, C, device="cuda")
    g_fc = torch.empty(N * npix, C, device="cuda", dtype=torch.bfloat16)
    L.bce_uniform_off(logits, labels.to(torch.int32).cuda(), N, C, 1.0 / C, 1.0, npix, 1.0, loss, g_fc, g_log, L.DTYPE_CODE["bf16"])
    lg = logits.cpu().clone().requires_grad_(True)
    ref = OR.uniform_off_labels_bce(lg, labels)
    (gref,) = torch.autograd.grad(ref, [lg])
    assert abs(float(loss) - float(ref.detach())) < 1e-6 * max(1.0, abs(float(ref.detach()))) + 1e-7
    assert _rel(g_log.cpu(), gref) < 1e-5
    assert _rel(g_fc.float().cpu().view(N, npix, C)[:, 2], gref / npix) < 5e-3
    # AGC + AdamW, first and second step, conv weight (unit = output channel) and a 1-D parameter (one unit)
    for shape in ((32, 16, 3, 3), (48,)):
        p0 = torch.randn(shape) * 0.05
        m = torch.zeros(p0.numel(), device="cuda")
        v = torch.zeros(p0.numel(), device="cuda")
        w = p0.clone().reshape(-1).cuda()
        pref, mref, vref = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
        units, cols = (shape[0], p0.numel() // shape[0]) if len(shape) == 4 else (1, shape[0])
        for step in (1, 2):
            g = torch.randn(shape) * (0.5 if step == 1 else 1e-4)          # step 1 clips, step 2 does not
            L.agc_adamw(w, g.reshape(-1).cuda() * 2.0, None, m, v, units, cols, 0.5, 1e-3, 0.9, 0.999, 1e-8, 0.01, 0.01, 1e-3, step)
            gc = OR.adaptive_clip_grad(pref, g, 0.01, 1e-3)
            mref = 0.9 * mref + 0.1 * gc
            vref = 0.999 * vref + 0.001 * gc * gc
            pref = pref * (1 - 1e-3 * 0.01) - 1e-3 * (mref / (1 - 0.9 ** step)) / ((vref / (1 - 0.999 ** step)).sqrt() + 1e-8)
            assert _rel(w.cpu().view(shape), pref) < 1e-5, (shape, step)


@pytest.mark.parametrize("dtype,loss_scale,min_cos", [("fp16", 65536.0, 0.99), ("bf16", 1.0, 0.93)])
def test_resnet18_train_step_matches_oracle(bcosk_lib, dtype, loss_scale, min_cos):
    arch, S, nb = "resnet18", 64, 8
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
    imgs = synth.synth_images_u8(nb, S, 1)
    x6 = synth.to_bcos_input(imgs)
    labels = torch.arange(nb) * 37 % 1000
    om = OR.OracleResNet(arch, sd)
    om.calibrate_bn(x6)
    ref = OR.train_step_reference(om, x6, labels)
    plan = ResNetTrainPlan(arch, sd, nb, dtype=dtype, device="cuda", image_size=S, loss_scale=loss_scale)
    loss = plan.train_step(torch.from_numpy(imgs), labels)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref["loss"])) <= 1e-3 * float(ref["loss"])
    g = plan.gradients()
    cos = {}
    for k, gr in ref["grads"].items():
        a, b = g[k].cpu().double().flatten(), gr.double().flatten()
        cos[k] = float(torch.dot(a, b) / (a.norm() * b.norm()))
    print(f"REPORT train step {dtype}: min cos {min(cos.values()):.5f} median {sorted(cos.values())[len(cos) // 2]:.5f} "
          f"fc {cos['model.fc.linear.weight']:.6f}")
    assert set(g) == set(ref["grads"])
    assert min(cos.values()) >= min_cos and cos["model.fc.linear.weight"] >= (0.9999 if dtype == "fp16" else 0.995)
    # running variances follow the reference's EMA of the biased batch variance
    new = plan.state_dict()
    for k, v in ref["running_var"].items():
        assert _rel(new[k].cpu(), v) < (2e-3 if dtype == "fp16" else 3e-2), k
    # the update moved every weight by about lr (first AdamW step) and the packed operands follow the master weights
    k0 = "model.layer3.0.conv1.linear.weight"
    step = (new[k0].cpu() - sd[k0]).abs()
    assert 0.5e-4 < float(step.median()) < 1.5e-4
    lay = [l for l in plan.layers if l.name == "model.layer3.0.conv1"][0]
    o, c, kh, kw = sd[k0].shape
    packed = lay.fwd.b.float().cpu().view(o, kh * kw, -1)[:, :, :c]
    assert _rel(packed, new[k0].cpu().permute(0, 2, 3, 1).reshape(o, kh * kw, c)) < (1e-3 if dtype == "fp16" else 6e-3)
    # a second step runs on the refreshed operands and lowers the loss on the same batch
    loss2 = float(plan.train_step(torch.from_numpy(imgs), labels))
    assert math.isfinite(loss2)


@pytest.mark.parametrize("arch,nb,min_cos,fc_cos", [("resnet50", 8, 0.85, 0.99), ("resnet34", 4, 0.98, 0.9999)])
def test_other_resnets_train_step_matches_oracle(bcosk_lib, arch, nb, min_cos, fc_cos):
    """bottleneck blocks (1x1 / strided 3x3 / 1x1, strided 1x1 shortcuts - the benchmarked ResNet-50) and the deeper basic-block net:
    loss, every weight gradient and the updated running variances of one fine-tuning step against the oracle's autograd.
    The random-init ResNet-50 with batch-statistics BN is ill conditioned: the reference arithmetic's own gradients differ by
    1.5e-3 .. 4e-3 (relative) between fp32 and fp64 (ResNet-18: 7e-6), i.e. ~500x the sensitivity to rounding - one fp16 plane
    (2^-11) leaves cosines of 0.87 (a BN weight) .. 0.97 there (independent of the loss scale: not a range effect) where ResNet-18 / -34 reach 0.998."""
    S = 64
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
    imgs = synth.synth_images_u8(nb, S, 2)
    x6 = synth.to_bcos_input(imgs)
    labels = torch.arange(nb) * 53 % 1000
    om = OR.OracleResNet(arch, sd)
    om.calibrate_bn(x6)
    ref = OR.train_step_reference(om, x6, labels)
    plan = ResNetTrainPlan(arch, sd, nb, dtype="fp16", device="cuda", image_size=S, loss_scale=65536.0)
    loss = plan.train_step(torch.from_numpy(imgs), labels)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref["loss"])) <= 1e-3 * float(ref["loss"])
    g = plan.gradients()
    assert set(g) == set(ref["grads"])
    cos = {}
    for k, gr in ref["grads"].items():
        a, b = g[k].cpu().double().flatten(), gr.double().flatten()
        cos[k] = float(torch.dot(a, b) / (a.norm() * b.norm()))
    worst = min(cos, key=cos.get)
    print(f"REPORT train step {arch} fp16: min cos {cos[worst]:.5f} ({worst}) median {sorted(cos.values())[len(cos) // 2]:.5f} "
          f"fc {cos['model.fc.linear.weight']:.6f}")
    assert min(cos.values()) >= min_cos and cos["model.fc.linear.weight"] >= fc_cos
    new = plan.state_dict()
    for k, v in ref["running_var"].items():
        assert _rel(new[k].cpu(), v) < (2e-3 if arch != "resnet50" else 1e-2), k     # (forward statistics of the fp16 activations)


def test_captured_train_step_equals_eager(bcosk_lib):
    """ResNetTrainPlan.capture(): the whole step as one CUDA graph (whole-model AGC + AdamW launch, device-side Adam step counter,
    one operand-refresh launch) gives the same weights as the eager step; capturing does not advance the training state."""
    arch, S, nb = "resnet18", 64, 8
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
    imgs = torch.from_numpy(synth.synth_images_u8(nb, S, 1))
    labels = torch.arange(nb) * 37 % 1000
    a = ResNetTrainPlan(arch, sd, nb, dtype="bf16", device="cuda", image_size=S)
    b = ResNetTrainPlan(arch, sd, nb, dtype="bf16", device="cuda", image_size=S)
    b.load_batch(imgs, labels)
    w0 = b.w_flat.clone()
    assert b.capture() and torch.equal(b.w_flat, w0) and float(b.adam_state[0]) == 0.0
    la, lb = [], []
    for _ in range(3):
        la.append(float(a.train_step(imgs, labels)))
        lb.append(float(b.train_step(imgs, labels)))
    torch.cuda.synchronize()
    assert float(b.adam_state[0]) == 3.0 and b.step_count == 3
    assert max(abs(x - y) / abs(x) for x, y in zip(la, lb)) < 1e-4, (la, lb)
    # split-K weight gradients add with fp32 atomics (order-dependent in the last bits), everything else is fixed-order
    assert _rel(b.w_flat, a.w_flat) < 1e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL all-reduce of the gradient buckets)")
def test_two_rank_nccl_allreduce_matches_manual_sum():
    """scripts/exp_train_ddp.py under torchrun: the bucketed side-stream all-reduce equals all_reduce of the ranks' local
    gradients, and both ranks hold identical weights after the step"""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(root, "scripts", "exp_train_ddp.py")], capture_output=True, text=True, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert line, out.stdout + out.stderr
    r = json.loads(line[-1])
    assert r["grad_sum_rel_err_vs_manual_allreduce"] < 1e-5 and r["weights_identical_across_ranks"]
