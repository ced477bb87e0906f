"""-m gpu: the fused CLIP ResNet plan (engine/clip_rn.py, BASELINE config 4) through the C ABI against the golden vectors
produced by the reference (tests/golden/clip_rn50_b2.npz: embedding + contribution map of cos(embedding, fixed unit vector)).

Tolerances (BASELINE.json north_star, embedding in the role of the logits): embedding <= 2e-3 relative, contribution maps
cosine >= 0.999 and max-abs <= 1e-3 of the map range against the reference's fp32 run or its fp64 evaluation (on this
random-init net the reference's own fp32 run sits `fp32_noise_floor_maxabs_over_range` = 2.3e-3 away from the exact result)."""
import os

import numpy as np
import pytest
import torch

import bcos_oracle as OR
import emulator as E
import opsutil as U
from bcos_b200.engine import CLIPResNetPlan
from bcos_b200.engine import ops as O
from bcos_b200.models import clip_rn_state_shapes, synthetic_clip_rn50_plan
from bcos_b200.utils import synth

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "clip_rn50_b2.npz"))
    sd = synth.synth_state_dict(clip_rn_state_shapes(), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    return gold, sd


def _metrics(out, gold):
    emb, cmap = out["embedding"].float().cpu(), out["contribution_map"].float().cpu()
    ref_e = torch.from_numpy(gold["embedding"])
    ref, ref64 = torch.from_numpy(gold["contribution_map"]), torch.from_numpy(gold["contribution_map_fp64"])
    rng = ref.flatten(1).max(1).values - ref.flatten(1).min(1).values
    return dict(
        emb_rel=((emb - ref_e).abs().max() / ref_e.abs().max()).item(),
        cos=torch.nn.functional.cosine_similarity(cmap.flatten(1).double(), ref.flatten(1).double()).min().item(),
        mar=((cmap - ref).abs().flatten(1).max(1).values / rng).max().item(),
        mar64=((cmap - ref64).abs().flatten(1).max(1).values / rng).max().item(),
        floor=float(gold["fp32_noise_floor_maxabs_over_range"]))


def test_seed_from_nchw_kernel(bcosk_lib):
    g = torch.Generator().manual_seed(3)
    nb, c, h, w = 3, 72, 5, 7
    for planes, dt, code in ((1, torch.float16, 1), (2, torch.bfloat16, 0)):
        from bcos_b200 import _lib as L
        code = L.DTYPE_CODE["fp16" if dt == torch.float16 else "bf16"]
        gg = torch.randn(nb, c, h, w, generator=g)
        mul1 = torch.rand(nb * h * w, c, generator=g).to(dt if planes == 1 else torch.float32)
        mul2 = torch.rand(nb * h * w, c, generator=g).to(dt)
        mask = torch.randint(-2**31, 2**31 - 1, (nb * h * w, (c + 31) // 32), generator=g, dtype=torch.int64).to(torch.int32)
        op = O.SeedFromNchwOp("seed", gg, 4096.0, mul1, torch.zeros(nb, h, w, planes * c, dtype=dt), mask, mul2,
                              torch.zeros(nb, h, w, planes * c, dtype=dt), planes, code)
        dev = U.to_device(op, "cuda", {})
        E.run([op])
        dev.run()
        torch.cuda.synchronize()
        print(U.compare(op, dev, 1e-3 if planes == 1 else 2e-5, ["out1", "out2"]))


def test_head_kernels(bcosk_lib):
    """csrc/bcosk_head.cu: strided-batched SGEMM (all four transpose forms, ragged sizes, per-head and per-image batching), the token
    gather with the mean token, the row softmax and the token-gradient seed, each against its torch restatement."""
    from bcos_b200 import _lib as L
    g = torch.Generator().manual_seed(23)

    def chk(op, tol16=2e-3, tol32=2e-5):
        dev = U.to_device(op, "cuda", {})
        E.run([op])
        dev.run()
        torch.cuda.synchronize()
        return {n: U.compare(op, dev, tol32 if getattr(op, n).dtype == torch.float32 else tol16, [n])[n]
                for n in U.OUTPUT_FIELDS[type(op)] if getattr(op, n) is not None}

    for (m, n, k, bt) in ((70, 45, 37, 3), (300, 520, 37, 8)):       # few tiles -> 32 x 32 tiles; many -> 64 x 64
      for ta in (False, True):
        for tb in (False, True):
            a = torch.randn(bt, *((k, m) if ta else (m, k)), generator=g)
            b = torch.randn(bt, *((n, k) if tb else (k, n)), generator=g)
            op = O.SgemmOp(f"sgemm_{int(ta)}{int(tb)}", ta, tb, m, n, k, a, 0, a.shape[-1], a[0].numel(), b, 0, b.shape[-1], b[0].numel(),
                           torch.zeros(bt, m, n), 0, n, m * n, bt, 0.5)
            print(op.name, chk(op))
    # per-head batching with strides inside one matrix (q[:, h-block] x W[h-block rows, :] -> out[:, h, :])
    B_, H, dh, c = 5, 4, 16, 64
    q, w = torch.randn(B_, c, generator=g), torch.randn(c, c, generator=g)
    op = O.SgemmOp("per_head", False, False, B_, c, dh, q, 0, c, dh, w, 0, c, dh * c, torch.zeros(B_, H, c), 0, H * c, c, H)
    print(op.name, chk(op))
    x = torch.zeros(3, 4, 4, 2 * 64, dtype=torch.float16)
    E._split_store(x, torch.randn(3, 4, 4, 64, generator=g), 2)
    print("tokens", chk(O.HeadTokensOp("tokens", x, 64, 2, L.DTYPE_CODE["fp16"], torch.zeros(3, 17, 64))))
    print("softmax", chk(O.RowSoftmaxOp("softmax", torch.randn(6, 8, 50, generator=g) * 3)))
    mask = torch.randint(-2**31, 2**31 - 1, (3 * 16, 2), generator=g, dtype=torch.int64).to(torch.int32)
    op = O.SeedFromTokensOp("seed", torch.randn(3, 17, 64, generator=g), 4096.0, torch.rand(48, 64, generator=g).half(),
                            torch.zeros(3, 4, 4, 64, dtype=torch.float16), mask, None, torch.zeros(3, 4, 4, 64, dtype=torch.float16), 1,
                            L.DTYPE_CODE["fp16"])
    print("seed_from_tokens", chk(op))


def test_clip_rn50_module_level_head_path(bcosk_lib, golden_dir):
    """fused_head=False: trunk as fused launches, attention pool on the module-level kernels through autograd (the hand-over path)."""
    gold, sd = _golden(golden_dir)
    x6 = synth.to_bcos_input(gold["images_u8"]).cuda()
    plan = CLIPResNetPlan(sd, 2, mode="parity", device="cuda", fused_head=False)
    out = plan.explain_direction(x6, OR.clip_seed_direction(1024, int(gold["seed"])))
    torch.cuda.synchronize()
    m = _metrics(out, gold)
    print("fused trunk + module-level head vs reference golden:", m)
    assert m["emb_rel"] <= 2e-3 and m["cos"] >= 0.999 and min(m["mar"], m["mar64"]) <= 1e-3, m


@pytest.mark.parametrize("mode", ["parity", "throughput_fp16"])
def test_clip_rn50_fused_plan_matches_golden(bcosk_lib, golden_dir, mode):
    gold, sd = _golden(golden_dir)
    x6 = synth.to_bcos_input(gold["images_u8"]).cuda()
    t = OR.clip_seed_direction(1024, int(gold["seed"]))
    plan = CLIPResNetPlan(sd, 2, mode=mode, device="cuda")
    out = plan.explain_direction(x6, t)
    torch.cuda.synchronize()
    m = _metrics(out, gold)
    print(f"fused CLIP RN50 plan ({mode}) vs reference golden:", m)
    emb2 = plan.embed(x6)
    assert torch.allclose(emb2, out["embedding"], rtol=1e-5, atol=1e-6)         # forward-only API gives the same embedding
    if mode == "parity":
        assert m["emb_rel"] <= 2e-3 and m["cos"] >= 0.999 and min(m["mar"], m["mar64"]) <= 1e-3, m
    else:
        # one fp16 plane on this chaotic random-init net (ReLU decisions flip, SURVEY 7 hard part 1): reported, sanity bounds only
        assert m["emb_rel"] <= 0.5 and m["cos"] >= 0.9, m


def test_clip_rn50_captured_plan_batch_independent(bcosk_lib, golden_dir):
    """The benchmark configuration of config 4 (uint8 input, CUDA graphs, autotuned schedules, batch 64): images 0-1 of the
    batch reproduce the batch-2 golden result (images are independent in eval mode)."""
    gold, sd = _golden(golden_dir)
    u8 = torch.from_numpy(gold["images_u8"])
    batch = torch.from_numpy(synth.synth_images_u8(64, 224, 7))
    batch[:2] = u8
    t = OR.clip_seed_direction(1024, int(gold["seed"]))
    plan = CLIPResNetPlan(sd, 64, mode="parity", device="cuda", input_u8=True)
    plan.load_input(batch.cuda())
    plan.capture()
    out = plan.explain_direction(batch.cuda(), t)
    torch.cuda.synchronize()
    m = _metrics({"embedding": out["embedding"][:2], "contribution_map": out["contribution_map"][:2]}, gold)
    print("captured batch-64 CLIP RN50 plan, images 0-1 vs golden:", m)
    assert m["emb_rel"] <= 2e-3 and m["cos"] >= 0.999 and min(m["mar"], m["mar64"]) <= 1e-3, m


def test_synthetic_clip_plan_builds_and_runs(bcosk_lib):
    plan = synthetic_clip_rn50_plan(4, mode="throughput", device="cuda", input_u8=True)
    x = torch.from_numpy(synth.synth_images_u8(4, 224, 3)).cuda()
    out = plan.explain_direction(x, OR.clip_seed_direction(1024, 0))
    assert torch.isfinite(out["embedding"]).all() and torch.isfinite(out["contribution_map"]).all()
    assert out["embedding"].abs().max() > 1e-6 and out["contribution_map"].abs().max() > 0
