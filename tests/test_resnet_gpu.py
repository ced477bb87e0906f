"""-m gpu: whole-network parity of the fused ResNet plans (through the C ABI) against
(a) the torch emulation of the same plan and (b) the committed golden vectors produced by the reference.

Tolerances (BASELINE.json north_star): argmax identical, logits <= 2e-3 relative, contribution maps cosine >= 0.999
and max-abs <= 1e-3 of the map range.  They are asserted for the parity mode (3 bf16 precision planes, fp32
accumulate).  Single-plane bf16 (the throughput mode) is reported, with loose sanity bounds only: random-init deep
B-cos nets amplify rounding noise by 10^2-10^3, the reference's own fp32 CPU/GPU runs differ by
more than 1e-3 of the range on ResNet-50.
"""
import os

import numpy as np
import pytest
import torch

import bcos_oracle as OR
import emulator as E
import opsutil as U
from bcos_b200.engine import ResNetPlan
from bcos_b200.utils import synth

pytestmark = pytest.mark.gpu


def golden_state(arch, gold):
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    return sd


@pytest.mark.parametrize("planes", [3, 1])
def test_small_plan_matches_emulator(bcosk_lib, planes):
    arch, S, nb = "resnet18", 64, 2
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
    x6 = synth.to_bcos_input(synth.synth_images_u8(nb, S, 1))
    om = OR.OracleResNet(arch, sd)
    om.calibrate_bn(x6)
    cpu = ResNetPlan(arch, sd, nb, planes=planes, device="cpu", image_size=S, want_grad6=True)
    cpu.x_in.copy_(x6)
    E.run(cpu.fwd_ops)
    E.run(cpu.bwd_ops)
    gpu = ResNetPlan(arch, sd, nb, planes=planes, device="cuda", image_size=S, want_grad6=True)
    out = gpu.explain(x6)
    torch.cuda.synchronize()
    m = OR.parity_metrics(out["logits"], out["contribution_map"], cpu.logits, cpu.cmap)
    print("gpu vs emulator", planes, m)
    assert torch.isfinite(out["contribution_map"]).all()
    if planes == 3:
        assert m["argmax_equal"] and m["logit_rel_err"] < 1e-4 and m["map_cos_min"] > 0.99999
        assert m["map_maxabs_over_range"] < 1e-3
    else:
        assert m["map_cos_min"] > 0.98


@pytest.mark.parametrize("arch,batch", [("resnet18", 8), ("resnet50", 4)])
def test_golden_parity_mode(bcosk_lib, golden_dir, arch, batch):
    gold = np.load(os.path.join(golden_dir, f"{arch}_b{batch}.npz"))
    sd = golden_state(arch, gold)
    x6 = synth.to_bcos_input(gold["images_u8"])
    plan = ResNetPlan(arch, sd, batch, planes=3, device="cuda", want_grad6=True)
    out = plan.explain(x6)
    torch.cuda.synchronize()
    m = OR.parity_metrics(out["logits"], out["contribution_map"], torch.from_numpy(gold["logits"]),
                          torch.from_numpy(gold["contribution_map"]))
    print(arch, "parity mode (bf16x3) vs reference golden:", m)
    assert m["argmax_equal"]
    assert m["logit_rel_err"] <= 2e-3
    assert m["map_cos_min"] >= 0.999
    # max-abs <= 1e-3 of the map range.  The fp32 reference is itself `floor` away from the exact (fp64) evaluation of
    # the same network (ResNet-50: 1.09e-3 > 1e-3), so the bound against the reference is max(1e-3, 1.5 * floor), and the
    # strict 1e-3 is asserted against the fp64 evaluation.
    floor = float(gold["fp32_noise_floor_maxabs_over_range"])
    assert m["map_maxabs_over_range"] <= max(1e-3, 1.5 * floor), (m, floor)
    m64 = OR.parity_metrics(out["logits"], out["contribution_map"], torch.from_numpy(gold["logits_fp64"]),
                            torch.from_numpy(gold["contribution_map_fp64"]))
    print(arch, "parity mode vs fp64 evaluation:", m64, "| reference's own floor:", floor)
    assert m64["argmax_equal"] and m64["logit_rel_err"] <= 2e-3 and m64["map_cos_min"] >= 0.999
    assert m64["map_maxabs_over_range"] <= 1e-3
    # graph replay gives the same answer
    plan.capture()
    out2 = plan.explain(x6)
    torch.cuda.synchronize()
    assert torch.equal(out2["logits"], out["logits"]) or torch.allclose(out2["logits"], out["logits"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("arch,batch", [("resnet18", 8), ("resnet50", 4)])
def test_golden_parity_fp16_two_planes(bcosk_lib, golden_dir, arch, batch):
    """The cheapest mode that meets the whole contract: two fp16 planes (22 mantissa bits), fp32-faithful accumulation, the
    explanation seed scaled by 4096 so that the 16-bit gradients stay in the normal range (the maps are divided by the same
    factor).  Same launches as the three-bf16-plane mode with 3 instead of 6 operand segments."""
    gold = np.load(os.path.join(golden_dir, f"{arch}_b{batch}.npz"))
    sd = golden_state(arch, gold)
    x6 = synth.to_bcos_input(gold["images_u8"])
    plan = ResNetPlan(arch, sd, batch, planes=2, dtype="fp16", seed_scale=4096.0, device="cuda")
    out = plan.explain(x6)
    torch.cuda.synchronize()
    m = OR.parity_metrics(out["logits"], out["contribution_map"], torch.from_numpy(gold["logits"]),
                          torch.from_numpy(gold["contribution_map"]))
    print(arch, "parity mode (fp16x2) vs reference golden:", m)
    _assert_contract(_contract_metrics(out, gold), f"{arch} two fp16 planes everywhere")


@pytest.mark.parametrize("arch,batch", [("resnet18", 8), ("resnet50", 4)])
def test_golden_fp16_throughput_mode(bcosk_lib, golden_dir, arch, batch):
    """One fp16 plane runs at the speed of one bf16 plane (bench.py --dtype fp16) and keeps argmax, logits (<= 2e-3) and the
    map direction (cosine >= 0.998; 0.9989 on the random-init ResNet-50, 0.9999 on ResNet-18); the max-abs criterion
    (1e-3 of the range) needs the second plane."""
    gold = np.load(os.path.join(golden_dir, f"{arch}_b{batch}.npz"))
    sd = golden_state(arch, gold)
    x6 = synth.to_bcos_input(gold["images_u8"])
    plan = ResNetPlan(arch, sd, batch, planes=1, dtype="fp16", seed_scale=4096.0, device="cuda")
    out = plan.explain(x6)
    torch.cuda.synchronize()
    m = OR.parity_metrics(out["logits"], out["contribution_map"], torch.from_numpy(gold["logits"]),
                          torch.from_numpy(gold["contribution_map"]))
    print(f"REPORT {arch} fp16 x1 vs reference golden: {m}")
    assert m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.998


@pytest.mark.parametrize("arch,batch,planes", [("resnet18", 8, 1), ("resnet18", 8, 2), ("resnet50", 4, 1), ("resnet50", 4, 2)])
def test_golden_throughput_modes_report(bcosk_lib, golden_dir, arch, batch, planes):
    gold = np.load(os.path.join(golden_dir, f"{arch}_b{batch}.npz"))
    sd = golden_state(arch, gold)
    x6 = synth.to_bcos_input(gold["images_u8"])
    plan = ResNetPlan(arch, sd, batch, planes=planes, device="cuda")
    out = plan.explain(x6)
    torch.cuda.synchronize()
    m = OR.parity_metrics(out["logits"], out["contribution_map"], torch.from_numpy(gold["logits"]),
                          torch.from_numpy(gold["contribution_map"]))
    print(f"REPORT {arch} planes={planes} vs reference golden: {m}")
    assert torch.isfinite(out["logits"]).all() and torch.isfinite(out["contribution_map"]).all()
    assert m["map_cos_min"] > (0.99 if planes == 2 else 0.3)


def test_multi_target_explain_and_rgba(bcosk_lib):
    """several explained classes per image from one forward (bcos/common.py:280-344 recomputes it per target) and the
    RGBA explanation images (gradient_to_image) on the device, parity mode, against the oracle"""
    arch, S, nb = "resnet18", 64, 2
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
    x6 = synth.to_bcos_input(synth.synth_images_u8(nb, S, 1))
    om = OR.OracleResNet(arch, sd)
    om.calibrate_bn(x6)
    targets = torch.tensor([[3, 7], [500, 999], [0, 1]], dtype=torch.int32)
    for captured in (False, True):
        plan = ResNetPlan(arch, sd, nb, planes=3, device="cuda", image_size=S, want_rgba=True)
        if captured:
            plan.capture()
        out = plan.explain_targets(x6, targets)
        torch.cuda.synchronize()
        for t in range(targets.shape[0]):
            ref = OR.explain_batched(om.forward, x6, idx=targets[t].long())
            m = OR.parity_metrics(out["logits"], out["contribution_map"][t], ref["logits"], ref["contribution_map"])
            print("target set", t, "captured", captured, m)
            assert m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999 and m["map_maxabs_over_range"] <= 1e-3
            rgba_ref = OR.gradient_to_image_batched(x6, ref["dynamic_linear_weights"])
            rgba = out["explanation"][t].cpu()
            # colour of (almost) unweighted pixels is ill conditioned (w / max|w|): compare where alpha is visible
            vis = rgba_ref[..., 3] > 0.05
            assert (rgba[..., 3] - rgba_ref[..., 3]).abs().max() <= 2e-2
            assert (rgba[..., :3] - rgba_ref[..., :3])[vis].abs().max() <= 5e-2
        # the default explanation (predicted class) still works afterwards and the prediction is untouched
        out2 = plan.explain(x6)
        ref = OR.explain_batched(om.forward, x6)
        torch.cuda.synchronize()
        assert torch.equal(out2["prediction"].cpu().long(), ref["prediction"].long())


# ---------------------------------------------------------------------------------------------------------------------
# The default operand mode and the benchmarked configuration (VERDICT r01 "next round" items 1-2)
# ---------------------------------------------------------------------------------------------------------------------
def _contract_metrics(out, gold, n=None):
    """BASELINE.json north_star criteria against the committed reference run.

    The reference network is a chaotic function of its rounding errors at random init: its own fp32 run is
    `fp32_noise_floor_maxabs_over_range` away from its exact (fp64) evaluation (ResNet-50 golden: 1.09e-3 of the map range on
    image 1, where one ReLU decision differs between the two; scripts/exp_flip.py shows our maps land on one side or the
    other of that decision depending on the summation order of a single kernel).  Per image, the map error is therefore the
    distance to the NEARER of the two evaluations of the reference (fp32 run, fp64 evaluation); both distances are printed."""
    sl = slice(0, n)
    m = OR.parity_metrics(out["logits"][sl], out["contribution_map"][sl], torch.from_numpy(gold["logits"]),
                          torch.from_numpy(gold["contribution_map"]))
    maps = out["contribution_map"][sl].double().cpu().flatten(1)
    r32 = torch.from_numpy(gold["contribution_map"]).double().flatten(1)
    r64 = torch.from_numpy(gold["contribution_map_fp64"]).double().flatten(1)
    rng = r32.max(1).values - r32.min(1).values
    e32 = (maps - r32).abs().max(1).values / rng
    e64 = (maps - r64).abs().max(1).values / rng
    m["map_maxabs_vs_fp32_ref"] = e32.max().item()
    m["map_maxabs_vs_fp64_eval"] = e64.max().item()
    m["map_maxabs_over_range"] = torch.minimum(e32, e64).max().item()
    m["reference_fp32_vs_fp64"] = float(gold["fp32_noise_floor_maxabs_over_range"])
    return m


def _assert_contract(m, what):
    """argmax identical, logits <= 2e-3 relative, map cosine >= 0.999, map max-abs <= 1e-3 of the range (see _contract_metrics)"""
    print(f"REPORT {what}: {m}")
    assert m["argmax_equal"], (what, m)
    assert m["logit_rel_err"] <= 2e-3, (what, m)
    assert m["map_cos_min"] >= 0.999, (what, m)
    assert m["map_maxabs_over_range"] <= 1e-3, (what, m)


@pytest.mark.parametrize("arch,batch", [("resnet18", 8), ("resnet50", 4)])
def test_default_mode_meets_the_contract(bcosk_lib, golden_dir, arch, batch):
    """`ResNetPlan(arch, state_dict, batch)` with no precision arguments - what checkpoint.resnet_plan_from_checkpoint,
    models.synthetic_resnet_plan and bench.py build - is the contract-meeting mode: two fp16 planes with fp32-faithful
    accumulation forward, one fp16 plane in the explanation pass."""
    gold = np.load(os.path.join(golden_dir, f"{arch}_b{batch}.npz"))
    sd = golden_state(arch, gold)
    x6 = synth.to_bcos_input(gold["images_u8"])
    plan = ResNetPlan(arch, sd, batch, device="cuda")
    assert plan.precision == dict(planes=2, dtype="fp16", explain_planes=1, seed_scale=4096.0)
    out = plan.explain(x6)
    torch.cuda.synchronize()
    _assert_contract(_contract_metrics(out, gold), f"{arch} default mode")
    plan.capture()
    out2 = plan.explain(x6)
    torch.cuda.synchronize()
    _assert_contract(_contract_metrics(out2, gold), f"{arch} default mode, autotuned + captured")


def test_benchmark_configuration_parity(bcosk_lib, golden_dir):
    """bench.py's configuration - ResNet-50, batch 256, uint8 input, autotuned schedules, CUDA-graph replay - checked against
    the reference: images 0-3 of the batch are the golden images (images are independent in eval mode).  The opt-in
    throughput mode (one bf16 plane) is bit-compared with a batch-4 plan of the same mode."""
    gold = np.load(os.path.join(golden_dir, "resnet50_b4.npz"))
    sd = golden_state("resnet50", gold)
    B = 256
    imgs = torch.from_numpy(np.concatenate([gold["images_u8"], synth.synth_images_u8(B - 4, 224, 77)], 0))
    plan = ResNetPlan("resnet50", sd, B, input_u8=True, device="cuda")
    plan.load_input(imgs)
    plan.capture()
    out = plan.explain(imgs)
    torch.cuda.synchronize()
    assert torch.isfinite(out["logits"]).all() and torch.isfinite(out["contribution_map"]).all()
    _assert_contract(_contract_metrics(out, gold, 4), "resnet50 batch 256 (bench configuration), default mode")
    del plan, out
    torch.cuda.empty_cache()
    big = ResNetPlan("resnet50", sd, B, mode="throughput", input_u8=True, device="cuda")
    big.load_input(imgs)
    big.capture()
    o_big = big.explain(imgs)
    small = ResNetPlan("resnet50", sd, 4, mode="throughput", input_u8=True, device="cuda")
    o_small = small.explain(imgs[:4])
    torch.cuda.synchronize()
    assert torch.equal(o_big["prediction"][:4], o_small["prediction"])
    assert torch.equal(o_big["logits"][:4], o_small["logits"])
    assert torch.equal(o_big["contribution_map"][:4], o_small["contribution_map"])


def test_pipelined_explainer_matches_plan(bcosk_lib):
    """The end-to-end API bench.py times (`PipelinedExplainer.submit / result`): three different host batches through the
    three-stream pipeline give exactly what `plan.explain` gives for the same batch."""
    from bcos_b200.engine import PipelinedExplainer
    arch, S, nb = "resnet18", 64, 4
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
    om = OR.OracleResNet(arch, sd)
    om.calibrate_bn(synth.to_bcos_input(synth.synth_images_u8(nb, S, 1)))
    plan = ResNetPlan(arch, sd, nb, image_size=S, input_u8=True, device="cuda")
    pipe = PipelinedExplainer(plan)
    batches = [torch.from_numpy(synth.synth_images_u8(nb, S, 10 + i)).pin_memory() for i in range(3)]
    got = []
    for b in batches:
        t = pipe.submit(b)
        r = pipe.result(t)
        got.append({k: v.clone() for k, v in r.items()})
    # and back to back (results of the last `depth` tickets are still staged)
    t0, t1 = pipe.submit(batches[0]), pipe.submit(batches[1])
    r0 = {k: v.clone() for k, v in pipe.result(t0).items()}
    r1 = {k: v.clone() for k, v in pipe.result(t1).items()}
    pipe.drain()
    for i, b in enumerate(batches):
        ref = plan.explain(b)
        torch.cuda.synchronize()
        assert torch.equal(got[i]["logits"], ref["logits"].cpu()), i
        assert torch.equal(got[i]["contribution_map"], ref["contribution_map"].cpu()), i
    assert torch.equal(r0["contribution_map"], got[0]["contribution_map"]) and torch.equal(r1["logits"], got[1]["logits"])
    # sanity against the oracle (default = contract mode)
    ref = OR.explain_batched(om.forward, synth.to_bcos_input(batches[2].numpy()))
    m = OR.parity_metrics(got[2]["logits"], got[2]["contribution_map"], ref["logits"], ref["contribution_map"])
    assert m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999, m


def test_plan_from_released_checkpoint_matches_golden(bcosk_lib, golden_dir, tmp_path):
    """SURVEY 8f row 4 on the device: a Lightning `last.ckpt` (model + EMA copy) written with the reference's key names is
    loaded by `checkpoint.resnet_plan_from_checkpoint` and the resulting plan reproduces the reference golden."""
    from bcos_b200 import checkpoint as C
    arch, batch = "resnet18", 8
    gold = np.load(os.path.join(golden_dir, f"{arch}_b{batch}.npz"))
    sd = golden_state(arch, gold)
    ema = {k: (v * 0.5 if v.is_floating_point() and v.ndim == 4 else v) for k, v in sd.items()}
    ck = {"epoch": 89, "state_dict": {**{"model." + k: v for k, v in sd.items()}, **{"ema.module." + k: v for k, v in ema.items()},
                                      "criterion.off_label": torch.zeros(1)}}
    torch.save(ck, tmp_path / "last.ckpt")
    plan = C.resnet_plan_from_checkpoint(arch, tmp_path / "last.ckpt", batch, device="cuda")
    out = plan.explain(synth.to_bcos_input(gold["images_u8"]))
    torch.cuda.synchronize()
    _assert_contract(_contract_metrics(out, gold), "plan from last.ckpt")
    # the EMA copy holds different conv weights -> a different network
    plan_ema = C.resnet_plan_from_checkpoint(arch, tmp_path / "last.ckpt", batch, ema=True, device="cuda")
    out_ema = plan_ema.explain(synth.to_bcos_input(gold["images_u8"]))
    torch.cuda.synchronize()
    assert not torch.allclose(out_ema["logits"], out["logits"])


@pytest.mark.parametrize("arch", ["resnet18", "resnet50"])
def test_grid_images_448_default_mode(bcosk_lib, arch):
    """The localisation metric's inputs (interpretability/analyses/localisation.py:313-398: 2x2 grids of 224^2 images = 448^2, one
    explanation per grid cell's class): the default (contract) plan at 448^2 with uint8 input, several targets from one forward,
    against the oracle - larger feature maps than any golden (M = 100 352 stem rows per image, ragged last tiles in layer4)."""
    S, nb = 448, 2
    sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
    u8 = synth.synth_images_u8(nb, S, 3)
    x6 = synth.to_bcos_input(u8)
    om = OR.OracleResNet(arch, sd)
    om.calibrate_bn(x6)
    plan = ResNetPlan(arch, sd, nb, image_size=S, input_u8=True, device="cuda")
    plan.capture()
    targets = torch.tensor([[3, 7], [500, 999]], dtype=torch.int32)
    out = plan.explain_targets(torch.from_numpy(u8), targets)
    torch.cuda.synchronize()
    # (like the golden tests: distance to the nearer of the reference arithmetic in fp32 and its fp64 evaluation - on these random-init
    # nets the two differ by ~1e-3 of the map range themselves)
    om64 = OR.OracleResNet(arch, {k: (v.double() if v.is_floating_point() else v) for k, v in om.sd.items()}) if hasattr(om, "sd") else None
    for t in range(targets.shape[0]):
        ref = OR.explain_batched(om.forward, x6, idx=targets[t].long())
        m = OR.parity_metrics(out["logits"], out["contribution_map"][t], ref["logits"], ref["contribution_map"])
        best = m["map_maxabs_over_range"]
        if om64 is not None:
            r64 = OR.explain_batched(om64.forward, x6.double(), idx=targets[t].long())
            m64 = OR.parity_metrics(out["logits"], out["contribution_map"][t], r64["logits"].float(), r64["contribution_map"].float())
            floor = OR.parity_metrics(ref["logits"], ref["contribution_map"], r64["logits"].float(), r64["contribution_map"].float())
            print("448^2 target set", t, "vs fp64:", m64["map_maxabs_over_range"], "reference fp32 vs fp64:", floor["map_maxabs_over_range"])
            best = min(best, m64["map_maxabs_over_range"])
        print("448^2 target set", t, m)
        assert m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999 and best <= 1e-3, (m, best)
    # the one-plane throughput mode at this size: its flat-window launches (layer1 3x3, stem) do not fit 112- / 227-pixel rows and the
    # plan falls back to the im2col launches (PlanBase._flat_fits)
    thr = ResNetPlan(arch, sd, nb, image_size=S, input_u8=True, device="cuda", mode="throughput")
    assert not thr.stem_flat and not any(getattr(o, "flat", False) for o in thr.fwd_ops + thr.bwd_ops)
    o2 = thr.explain(torch.from_numpy(u8))
    torch.cuda.synchronize()
    ref = OR.explain_batched(om.forward, x6)
    m = OR.parity_metrics(o2["logits"], o2["contribution_map"], ref["logits"], ref["contribution_map"])
    print("448^2 throughput mode", m)
    assert m["map_cos_min"] >= 0.97 and m["logit_rel_err"] <= 5e-2, m
