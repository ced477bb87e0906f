"""CPU: host-side plan logic (weight packing, im2col tap tables, space-to-depth stem, zero-insertion, residual and
gradient routing, classifier seed) checked by executing the plan's launch list with the torch emulator
(tests/emulator.py) and comparing with the oracle."""
import pytest
import torch

import bcos_oracle as OR
import emulator as E
from bcos_b200.engine import ResNetPlan, ops as O
from bcos_b200.engine import pack as P
from bcos_b200.models import resnet_state_shapes, synthetic_resnet_plan
from bcos_b200.utils import synth


def _run(arch, planes, size=64, nb=2, **kw):
    sd = synth.synth_state_dict(resnet_state_shapes(arch), 0)
    x6 = synth.to_bcos_input(synth.synth_images_u8(nb, size, 1))
    om = OR.OracleResNet(arch, sd)
    om.calibrate_bn(x6)
    ref = OR.explain_batched(om.forward, x6)
    plan = ResNetPlan(arch, sd, nb, planes=planes, device="cpu", image_size=size, want_grad6=True, **kw)
    plan.x_in.copy_(x6)
    E.run(plan.fwd_ops)
    E.run(plan.bwd_ops)
    return plan, ref


def test_resnet18_plan_matches_oracle_parity_planes():
    plan, ref = _run("resnet18", 3)
    m = OR.parity_metrics(plan.logits, plan.cmap, ref["logits"], ref["contribution_map"])
    assert m["argmax_equal"] and m["logit_rel_err"] < 1e-5 and m["map_cos_min"] > 0.999999 and m["map_maxabs_over_range"] < 1e-4, m
    g = ref["dynamic_linear_weights"]
    assert ((plan.grad6 - g).abs().max() / g.abs().max()).item() < 1e-4


def test_resnet50_plan_matches_oracle_two_planes():
    plan, ref = _run("resnet50", 2, size=96, nb=2)
    m = OR.parity_metrics(plan.logits, plan.cmap, ref["logits"], ref["contribution_map"])
    assert m["argmax_equal"] and m["logit_rel_err"] < 2e-3 and m["map_cos_min"] > 0.999, m


def test_stem_with_64_channel_chunks_is_equivalent():
    plan, ref = _run("resnet18", 3, stem_kch=64)
    m = OR.parity_metrics(plan.logits, plan.cmap, ref["logits"], ref["contribution_map"])
    assert m["map_cos_min"] > 0.999999, m


def test_uint8_input_path():
    arch, size, nb = "resnet18", 32, 2
    sd = synth.synth_state_dict(resnet_state_shapes(arch), 0)
    u8 = torch.from_numpy(synth.synth_images_u8(nb, size, 3))
    a = ResNetPlan(arch, sd, nb, planes=1, device="cpu", image_size=size, input_u8=True)
    b = ResNetPlan(arch, sd, nb, planes=1, device="cpu", image_size=size)
    a.x_in.copy_(u8)
    b.x_in.copy_(synth.to_bcos_input(u8))
    for p in (a, b):
        E.run(p.fwd_ops)
        E.run(p.bwd_ops)
    assert torch.equal(a.logits, b.logits) and torch.equal(a.cmap, b.cmap)


def test_accounting_matches_survey_figures():
    plan = synthetic_resnet_plan("resnet50", 1, device="cpu")
    fwd = sum(o.algo_flops for o in plan.fwd_ops if isinstance(o, O.IgemmOp))
    assert abs(fwd / 1e9 - 8.611) < 0.01          # 8.611 GFLOP / image forward
    assert len([o for o in plan.fwd_ops if isinstance(o, O.IgemmOp)]) == 54


def test_split_planes_reconstruct():
    x = torch.randn(1000) * 3
    for planes, tol in ((1, 2 ** -8), (2, 2 ** -16), (3, 2 ** -23)):
        pl = P.split_planes(x, planes, torch.bfloat16)
        rec = sum(p.float() for p in pl)
        assert ((rec - x).abs() / x.abs()).max() <= tol
