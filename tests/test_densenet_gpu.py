"""-m gpu: the fused DenseNet plan (engine/densenet.py) through the C ABI: the kernels of csrc/bcosk_dense.cu and the slice-writing
igemm launch against their torch restatements (tests/emulator.py), and the whole plan against the golden vectors produced by the
reference (tests/golden/densenet121_b2.npz).  Tolerances (BASELINE.json north_star): argmax identical, logits <= 2e-3 relative,
contribution maps cosine >= 0.999 and max-abs <= 1e-3 of the map range against the reference's fp32 run or its fp64 evaluation."""
import math
import os

import numpy as np
import pytest
import torch

import bcos_oracle as OR
import emulator as E
import opsutil as U
from bcos_b200 import _lib as L
from bcos_b200.engine import DenseNetPlan
from bcos_b200.engine import ops as O
from bcos_b200.engine.base import Act, PlanBase
from bcos_b200.models import densenet_state_shapes, synthetic_densenet_plan
from bcos_b200.utils import synth

pytestmark = pytest.mark.gpu


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
def test_dense_kernels(bcosk_lib, planes, dt):
    g = torch.Generator().manual_seed(13)
    code = L.DTYPE_CODE["fp16" if dt == torch.float16 else "bf16"]
    nb, h, ctot, c = 2, 5, 352, 288
    M = nb * h * h
    F = torch.zeros(nb, h, h, planes * ctot, dtype=dt)
    E._split_store(F, torch.randn(nb, h, h, ctot, generator=g), planes)
    alpha = torch.rand(c, generator=g) + 0.5
    op = O.DenseBnReluFwdOp("bn_relu", F, c, planes, alpha, True, torch.zeros(nb, h, h, planes * c, dtype=dt), torch.zeros(1, M),
                            torch.zeros(M, c // 32, dtype=torch.int32), code)
    print("bn_relu_fwd", _check(op, 1e-2 if planes == 1 else 2e-5, 2e-3 if planes == 1 else 2e-5))
    mask = torch.randint(-2**31, 2**31 - 1, (M, c // 32), generator=g, dtype=torch.int64).to(torch.int32)
    for gdt in (torch.float32, dt):
        for acc in (False, True):
            op = O.DenseBnReluBwdOp("bn_relu_bwd", torch.randn(nb, h, h, c, generator=g).to(gdt), c, alpha, mask,
                                    torch.randn(nb, h, h, ctot, generator=g), acc, code)
            print("bn_relu_bwd", _check(op, 1e-2, 2e-5))
    for gain in (None, torch.rand(M, 32, generator=g).to(dt), torch.rand(M, 32, generator=g)):
        op = O.DenseSliceCastOp("slice", torch.randn(nb, h, h, ctot, generator=g), 288, 32, gain, 0.5, torch.zeros(nb, h, h, 32, dtype=dt), code)
        print("slice_cast", _check(op, 1e-2 if dt == torch.bfloat16 else 2e-3, 2e-5))
    src = torch.zeros(nb, h, h, planes * 64, dtype=dt)
    E._split_store(src, torch.randn(nb, h, h, 64, generator=g), planes)
    op = O.CopyChannelsOp("copy", src, 64, planes, F.clone(), 0)
    print("copy", _check(op, 0.0, 0.0))


@pytest.mark.parametrize("planes", [1, 2])
def test_igemm_writes_into_a_slice(bcosk_lib, planes):
    """3x3 B-cos conv whose 32 output channels land in columns [96, 128) of every plane of a wider tensor (DenseNet growth)"""
    g = torch.Generator().manual_seed(17)
    plan = PlanBase(2, planes=planes, dtype="fp16", device="cpu", explain=True)
    plan.flat_3x3 = False
    dt = torch.float16
    v = torch.randn(2, 6, 6, 128, generator=g)
    t = torch.zeros(2, 6, 6, planes * 128, dtype=dt)
    stored = E._split_store(t, v, planes)
    x = Act(t, 128, (stored ** 2).sum(-1).reshape(1, -1).contiguous(), 1)
    F = torch.zeros(2, 6, 6, planes * 160, dtype=dt)
    E._split_store(F, torch.randn(2, 6, 6, 160, generator=g), planes)
    w = torch.randn(32, 128, 3, 3, generator=g) / math.sqrt(128 * 9)
    plan._conv_fwd("grow", x, w, 1, 1, 1, bn=None, relu=False, want_sq=False, y_buf=F, y_col=96)
    memo = {}
    dev = [U.to_device(o, "cuda", memo) for o in plan.fwd_ops]
    E.run(plan.fwd_ops)
    for o in dev:
        o.run()
    torch.cuda.synchronize()
    ref, got = plan.fwd_ops[-1], dev[-1]
    pst = 160
    for pl in range(planes):      # untouched columns stay bit-identical, the slice matches
        assert torch.equal(got.y.cpu()[..., pl * pst: pl * pst + 96], ref.y[..., pl * pst: pl * pst + 96])
        assert torch.equal(got.y.cpu()[..., pl * pst + 128:(pl + 1) * pst], ref.y[..., pl * pst + 128:(pl + 1) * pst])
    a = sum(got.y.cpu()[..., pl * pst + 96: pl * pst + 128].float() for pl in range(planes))
    b = sum(ref.y[..., pl * pst + 96: pl * pst + 128].float() for pl in range(planes))
    assert ((a - b).abs().max() / b.abs().max()).item() < (2e-3 if planes == 1 else 2e-5)
    assert U.max_rel_err(got.gain, ref.gain) < 2e-3


def _golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "densenet121_b2.npz"))
    sd = synth.synth_state_dict(densenet_state_shapes("densenet121"), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    return gold, sd


def test_densenet121_fused_plan_matches_golden(bcosk_lib, golden_dir):
    gold, sd = _golden(golden_dir)
    x6 = synth.to_bcos_input(gold["images_u8"]).cuda()
    for mode in ("parity", "throughput_fp16"):
        plan = DenseNetPlan("densenet121", sd, 2, mode=mode, device="cuda")
        out = plan.explain(x6)
        torch.cuda.synchronize()
        lg, cm = out["logits"].float().cpu(), out["contribution_map"].float().cpu()
        m = OR.parity_metrics(lg, cm, torch.from_numpy(gold["logits"]), torch.from_numpy(gold["contribution_map"]))
        m64 = OR.parity_metrics(lg, cm, torch.from_numpy(gold["logits_fp64"]), torch.from_numpy(gold["contribution_map_fp64"]))
        print(f"fused densenet121 plan ({mode}) vs reference golden:", m, "vs fp64:", m64["map_maxabs_over_range"])
        if mode == "parity":
            assert m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999, m
            assert min(m["map_maxabs_over_range"], m64["map_maxabs_over_range"]) <= 1e-3, (m, m64)
        else:
            assert m["map_cos_min"] >= 0.9, m
        del plan


def test_densenet_captured_plan_batch_independent(bcosk_lib, golden_dir):
    gold, sd = _golden(golden_dir)
    batch = torch.from_numpy(synth.synth_images_u8(16, 224, 9))
    batch[:2] = torch.from_numpy(gold["images_u8"])
    plan = DenseNetPlan("densenet121", sd, 16, mode="parity", device="cuda", input_u8=True)
    plan.load_input(batch.cuda())
    plan.capture()
    out = plan.explain(batch.cuda())
    torch.cuda.synchronize()
    m = OR.parity_metrics(out["logits"][:2].float().cpu(), out["contribution_map"][:2].float().cpu(), torch.from_numpy(gold["logits"]),
                          torch.from_numpy(gold["contribution_map"]))
    print("captured batch-16 DenseNet-121 plan, images 0-1 vs golden:", m)
    assert m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999, m
