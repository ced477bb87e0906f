"""-m gpu: the fused SimpleViT plan (engine/vit.py, BASELINE config 3) through the C ABI: every kernel of csrc/bcosk_vit.cu
against its torch restatement (tests/emulator.py) on seeded inputs, and the whole plan against the golden vectors produced by
the reference (tests/golden/simple_vit_{ti,b}_patch16_224_b2.npz).

Tolerances (BASELINE.json north_star): argmax identical, logits <= 2e-3 relative, contribution maps cosine >= 0.999 and
max-abs <= 1e-3 of the map range - asserted for the contract modes ("parity": two-plane fp16 residual stream with one-plane
branch operands, the default; "parity_full": two fp16 planes everywhere; one fp16 plane in the linear explanation pass)."""
import os

import numpy as np
import pytest
import torch

import bcos_oracle as OR
import emulator as E
import opsutil as U
from bcos_b200 import _lib as L
from bcos_b200.engine import ViTPlan
from bcos_b200.engine import ops as O
from bcos_b200.models import synthetic_vit_plan, vit_state_shapes
from bcos_b200.utils import synth

pytestmark = pytest.mark.gpu


def _planes(g, shape, planes, dt, scale=1.0):
    v = torch.randn(*shape, generator=g) * scale
    t = torch.zeros(*shape[:-1], planes * shape[-1], dtype=dt)
    E._split_store(t, v, planes)
    return t


def _check(op, tol16, tol32):
    dev = U.to_device(op, "cuda", {})
    E.run([op])
    dev.run()
    torch.cuda.synchronize()
    errs = {}
    for name in U.OUTPUT_FIELDS[type(op)]:
        t = getattr(op, name)
        if t is not None:
            errs.update(U.compare(op, dev, tol32 if t.dtype == torch.float32 else tol16, [name]))
    return errs


@pytest.mark.parametrize("planes,dt", [(1, torch.bfloat16), (2, torch.float16)])
def test_vit_bandwidth_kernels(bcosk_lib, planes, dt):
    g = torch.Generator().manual_seed(11)
    code = L.DTYPE_CODE["fp16" if dt == torch.float16 else "bf16"]
    tol16 = 1e-2 if planes == 1 else 2e-5
    nb, gh, d = 3, 4, 192
    M = nb * gh * gh
    mean6, istd6 = (0.485, 0.456, 0.406, 0.515, 0.544, 0.594), tuple(1 / s for s in (0.229, 0.224, 0.225, 0.229, 0.224, 0.225))
    for x in (torch.rand(nb, 6, 64, 64, generator=g), torch.randint(0, 256, (nb, 3, 64, 64), generator=g, dtype=torch.uint8)):
        op = O.VitPatchifyOp("patchify", x, 16, mean6, istd6, torch.zeros(nb, 4, 4, planes * 1536, dtype=dt), planes, code, torch.zeros(1, nb * 16))
        print("patchify", _check(op, tol16, 2e-3 if planes == 1 else 2e-5))
        op = O.VitContribMapOp("contrib", torch.randn(nb, 4, 4, 1536, generator=g), x, 16, istd6, 0.25, torch.zeros(nb, 64, 64), torch.zeros(nb, 6, 64, 64))
        print("contrib", _check(op, tol16, 2e-5))
    x = _planes(g, (nb, gh, gh, d), planes, dt)
    w = torch.rand(d, generator=g) + 0.5
    op = O.VitLnFwdOp("ln", x, d, planes, w, 1e-5, torch.zeros_like(x), torch.zeros(M), torch.zeros(1, M), code)
    print("ln_fwd", _check(op, tol16, 2e-3 if planes == 1 else 2e-5))
    for g16 in (False, True):
        gin = torch.randn(nb, gh, gh, d, generator=g)
        op = O.VitLnBwdOp("ln_bwd", gin.to(dt) if g16 else gin, torch.randn(nb, gh, gh, d, generator=g), d, w, torch.rand(M, generator=g) + 0.5,
                          torch.zeros(nb, gh, gh, d), torch.rand(M, d, generator=g).to(dt), torch.zeros(nb, gh, gh, d, dtype=dt), code)
        print("ln_bwd", _check(op, 1e-2 if dt == torch.bfloat16 else 2e-3, 2e-5))
    u = _planes(g, (nb, gh, gh, 768), planes, dt)
    for gdt in (dt, torch.float32):
        op = O.VitGeluFwdOp("gelu", u, 768, planes, torch.zeros_like(u), torch.zeros(1, M), torch.rand(M, 768, generator=g).to(gdt), code)
        print("gelu", _check(op, 1e-2 if (planes == 1 or gdt != torch.float32) and dt == torch.bfloat16 else 2e-3, 2e-3 if planes == 1 else 2e-5))
    op = O.PixelSqsumOp("sq", x, d, planes, code, torch.zeros(1, M))
    print("sqsum", _check(op, tol16, 2e-5))


@pytest.mark.parametrize("tc", [False, True], ids=["cuda_core", "tensor_core"])
@pytest.mark.parametrize("planes,dt,heads,n", [(1, torch.bfloat16, 3, 16), (2, torch.float16, 3, 196), (2, torch.float16, 12, 49),
                                                (1, torch.float16, 2, 196), (3, torch.bfloat16, 2, 144)])
def test_vit_attention_kernel(bcosk_lib, planes, dt, heads, n, tc):
    g = torch.Generator().manual_seed(5)
    code = L.DTYPE_CODE["fp16" if dt == torch.float16 else "bf16"]
    nb, dh = 2, 64
    gh = int(n ** 0.5)
    hd = heads * dh
    qkv = _planes(g, (nb, gh, gh, 3 * hd), planes, dt)
    tol16 = 1e-2 if planes == 1 else 5e-5
    fwd = O.VitAttentionOp("attn", qkv, planes, None, nb, n, heads, dh, dh ** -0.5, False, torch.zeros(nb, gh, gh, planes * hd, dtype=dt), code, tc)
    print("attention fwd", _check(fwd, tol16, 2e-5))
    bwd = O.VitAttentionOp("attn.bwd", qkv, planes, torch.randn(nb, gh, gh, hd, generator=g), nb, n, heads, dh, dh ** -0.5, True,
                           torch.zeros(nb, gh, gh, hd, dtype=dt), code, tc)
    # one 16-bit output plane; the tensor-core kernel also rounds the probabilities and g / row-sum to one plane (explain-pass format)
    print("attention bwd", _check(bwd, (2e-2 if tc else 1e-2) if dt == torch.bfloat16 else (4e-3 if tc else 2e-3), 2e-5))


@pytest.mark.parametrize("arch", ["simple_vit_ti_patch16_224", "simple_vit_b_patch16_224"])
def test_vit_fused_plan_matches_golden(bcosk_lib, golden_dir, arch):
    gold = np.load(os.path.join(golden_dir, f"{arch}_b2.npz"))
    sd = synth.synth_state_dict(vit_state_shapes(arch), int(gold["seed"]))
    x6 = synth.to_bcos_input(gold["images_u8"]).cuda()
    for mode in ("parity", "parity_full", "throughput_fp16", "throughput"):
        plan = ViTPlan(arch, sd, 2, mode=mode, device="cuda")
        assert plan.bp == (1 if mode != "parity_full" else 2) and plan.sp == (2 if mode.startswith("parity") else 1)
        out = plan.explain(x6)
        torch.cuda.synchronize()
        m = OR.parity_metrics(out["logits"].float().cpu(), out["contribution_map"].float().cpu(), torch.from_numpy(gold["logits"]),
                              torch.from_numpy(gold["contribution_map"]))
        print(f"fused {arch} plan ({mode}) vs reference golden:", m)
        if mode.startswith("parity"):     # "parity": two-plane residual stream + one-plane branches (default); "parity_full": all two planes
            assert m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999 and m["map_maxabs_over_range"] <= 1e-3, m
        else:
            assert m["map_cos_min"] >= 0.9, m
        del plan


def test_vit_captured_plan_batch_independent(bcosk_lib, golden_dir):
    """Benchmark configuration (uint8 input, CUDA graphs, autotuned schedules, batch 32): images 0-1 reproduce the batch-2 golden."""
    arch = "simple_vit_ti_patch16_224"
    gold = np.load(os.path.join(golden_dir, f"{arch}_b2.npz"))
    sd = synth.synth_state_dict(vit_state_shapes(arch), int(gold["seed"]))
    batch = torch.from_numpy(synth.synth_images_u8(32, 224, 9))
    batch[:2] = torch.from_numpy(gold["images_u8"])
    plan = ViTPlan(arch, sd, 32, mode="parity", device="cuda", input_u8=True)
    plan.load_input(batch.cuda())
    plan.capture()
    out = plan.explain(batch.cuda())
    torch.cuda.synchronize()
    m = OR.parity_metrics(out["logits"][:2].float().cpu(), out["contribution_map"][:2].float().cpu(), torch.from_numpy(gold["logits"]),
                          torch.from_numpy(gold["contribution_map"]))
    print("captured batch-32 ViT-Ti plan, images 0-1 vs golden:", m)
    assert m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999 and m["map_maxabs_over_range"] <= 1e-3, m


# ---------------------------------------------------------------------------------------------------------------------
# CLIP ViT image encoder (engine/clip_vit.py): true-backward kernels and the fused plan vs the reference golden
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("planes,dt", [(1, torch.float16), (2, torch.float16), (1, torch.bfloat16)])
def test_clip_vit_true_backward_kernels(bcosk_lib, planes, dt):
    g = torch.Generator().manual_seed(29)
    code = L.DTYPE_CODE["fp16" if dt == torch.float16 else "bf16"]
    nb, T, d, heads = 3, 50, 128, 2
    M = nb * T
    tol16 = 1e-2 if dt == torch.bfloat16 else 2e-3
    x = _planes(g, (nb, T, 1, d), planes, dt)
    w = torch.rand(d, generator=g) + 0.5
    for g16 in (False, True):
        gin = torch.randn(nb, T, 1, d, generator=g)
        op = O.VitLnBwdOp("ln_bwd_full", gin.to(dt) if g16 else gin, torch.randn(nb, T, 1, d, generator=g), d, w, torch.rand(M, generator=g) + 0.5,
                          torch.zeros(nb, T, 1, d), torch.rand(M, d, generator=g).to(dt), torch.zeros(nb, T, 1, d, dtype=dt), code, x, planes)
        print("ln_bwd_full", _check(op, tol16, 3e-5))
    u = _planes(g, (nb, T, 1, 256), planes, dt)
    for gdt in (dt, torch.float32):
        op = O.VitGeluFwdOp("quickgelu", u, 256, planes, torch.zeros_like(u), torch.zeros(1, M), torch.rand(M, 256, generator=g).to(gdt), code, True)
        print("quickgelu", _check(op, 1e-2 if dt == torch.bfloat16 else 2e-3, 2e-3 if planes == 1 else 3e-5))
    hd = heads * 64
    qkv = _planes(g, (nb, T, 1, 3 * hd), planes, dt)
    op = O.VitAttentionOp("attn.bwd_full", qkv, planes, torch.randn(nb, T, 1, hd, generator=g), nb, T, heads, 64, 64 ** -0.5, True,
                          torch.zeros(nb, T, 1, 3 * hd, dtype=dt), code, True, True)
    print("attention_bwd_full", _check(op, tol16, 2e-5))
    fwd = O.VitAttentionOp("attn.fwd50", qkv, planes, None, nb, T, heads, 64, 64 ** -0.5, False, torch.zeros(nb, T, 1, planes * hd, dtype=dt), code, True)
    print("attention fwd (tensor cores, 50 tokens)", _check(fwd, 1e-2 if planes == 1 else 5e-5, 2e-5))


def test_clip_vit_fused_plan_matches_golden(bcosk_lib, golden_dir):
    from bcos_b200.engine import CLIPViTPlan
    gold = np.load(os.path.join(golden_dir, "clip_vit_b32_b2.npz"))
    res, patch, width, layers, heads, out_dim = gold["geometry"].tolist()
    sd = synth.synth_state_dict(OR.clip_vit_state_shapes(res, patch, width, layers, out_dim), int(gold["seed"]))
    x6 = synth.to_bcos_input(gold["images_u8"]).cuda()
    t = OR.clip_seed_direction(out_dim, int(gold["seed"]))
    ref_e = torch.from_numpy(gold["embedding"])
    ref, ref64 = torch.from_numpy(gold["contribution_map"]), torch.from_numpy(gold["contribution_map_fp64"])
    rng = ref.flatten(1).max(1).values - ref.flatten(1).min(1).values
    for mode in ("parity", "parity_full", "throughput_fp16"):
        plan = CLIPViTPlan(sd, 2, heads=heads, mode=mode, device="cuda")
        out = plan.explain_direction(x6, t)
        torch.cuda.synchronize()
        emb, cmap = out["embedding"].float().cpu(), out["contribution_map"].float().cpu()
        e_rel = ((emb - ref_e).abs().max() / ref_e.abs().max()).item()
        cos = torch.nn.functional.cosine_similarity(cmap.flatten(1).double(), ref.flatten(1).double()).min().item()
        mar = ((cmap - ref).abs().flatten(1).max(1).values / rng).max().item()
        mar64 = ((cmap - ref64).abs().flatten(1).max(1).values / rng).max().item()
        print(f"fused CLIP ViT-B/32 plan ({mode}): embedding rel err {e_rel:.2e}, map cosine {cos:.8f}, max-abs/range {mar:.2e} (vs fp64 {mar64:.2e}; "
              f"reference's own floor {float(gold['fp32_noise_floor_maxabs_over_range']):.2e})")
        if mode.startswith("parity"):
            assert e_rel <= 2e-3 and cos >= 0.999 and min(mar, mar64) <= 1e-3, (mode, e_rel, cos, mar, mar64)
        else:
            assert cos >= 0.9
        assert torch.allclose(plan.embed(x6), out["embedding"], rtol=1e-5, atol=1e-6)
        del plan


@pytest.mark.gpu
@pytest.mark.parametrize("T", [113, 197, 208])
def test_clip_vit_true_attention_backward_long_sequences(bcosk_lib, T):
    """bcosk_vit_attention_bwd_full beyond 112 tokens (the recomputing kernel: no n x n block in shared memory) vs the emulator"""
    g = torch.Generator().manual_seed(T)
    dt, planes, nb, heads = torch.float16, 1, 2, 2
    code = L.DTYPE_CODE["fp16"]
    hd = heads * 64
    qkv = _planes(g, (nb, T, 1, 3 * hd), planes, dt)
    op = O.VitAttentionOp("attn.bwd_full.long", qkv, planes, torch.randn(nb, T, 1, hd, generator=g), nb, T, heads, 64, 64 ** -0.5, True,
                          torch.zeros(nb, T, 1, 3 * hd, dtype=dt), code, True, True)
    print("attention_bwd_full", T, _check(op, 2e-3, 2e-5))


@pytest.mark.gpu
def test_clip_vit_b16_geometry_197_tokens(bcosk_lib):
    """CLIP ViT-B/16 geometry (196 patches + class token = 197 tokens: both query tiles of the attention kernels, the true attention
    backward with 197 x 197 score blocks), two residual blocks, against the oracle in fp32 and fp64."""
    from bcos_b200.engine import CLIPViTPlan
    res, patch, width, layers, heads, out_dim = 224, 16, 768, 2, 12, 512
    sd = synth.synth_state_dict(OR.clip_vit_state_shapes(res, patch, width, layers, out_dim), 1)
    u8 = synth.synth_images_u8(2, res, 6)
    x6 = synth.to_bcos_input(u8)
    t = OR.clip_seed_direction(out_dim, 1)

    def oracle(dtype):
        o = OR.OracleCLIPViT({k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}, heads=heads)
        xb = x6.to(dtype).clone().requires_grad_(True)
        with torch.enable_grad():
            emb = o.forward(xb, detach=True)
            torch.nn.functional.cosine_similarity(emb, t.to(dtype)[None], dim=1).sum().backward(inputs=[xb])
        return emb.detach().float(), (xb * xb.grad).sum(1).detach().float()

    e32, c32 = oracle(torch.float32)
    e64, c64 = oracle(torch.float64)
    plan = CLIPViTPlan(sd, 2, heads=heads, device="cuda", input_u8=True)
    assert plan.ntok == 197
    out = plan.explain_direction(torch.from_numpy(u8), t)
    torch.cuda.synchronize()
    emb, cmap = out["embedding"].float().cpu(), out["contribution_map"].float().cpu()
    rng = c32.flatten(1).max(1).values - c32.flatten(1).min(1).values
    e_rel = ((emb - e32).abs().max() / e32.abs().max()).item()
    mar = min(((cmap - c).abs().flatten(1).max(1).values / rng).max().item() for c in (c32, c64))
    cos = torch.nn.functional.cosine_similarity(cmap.flatten(1).double(), c32.flatten(1).double()).min().item()
    print(f"CLIP ViT-B/16 geometry, 2 blocks: embedding rel err {e_rel:.2e}, map cosine {cos:.8f}, max-abs/range {mar:.2e}")
    assert e_rel <= 2e-3 and cos >= 0.999 and mar <= 1e-3
