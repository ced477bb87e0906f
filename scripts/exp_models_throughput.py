"""Module-level (un-fused drop-in) path: forward + batched explanation throughput of the other model families of
BASELINE.json's configs (3: SimpleViT-Ti/16, ViT-B/16; 4: CLIP RN50 image encoder; DenseNet-121 of config 5 in eval mode),
random-init synthetic weights, synthetic images, CUDA-event timing.  Also times the group / position norm kernels at
ResNet-like sizes and reports their HBM GB/s.  Prints one JSON object; run on the GPU box:
    python scripts/exp_models_throughput.py > gpurun_out/models_throughput.json
These are reported numbers of the un-fused path (layout bridges NCHW fp32 <-> NHWC planes around every launch); the fused
whole-network plan exists for the torchvision ResNets only (bench.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcos_b200  # noqa: E402,F401
import bcos_b200.modules as M  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def model_rows():
    from bcos_b200.bcosify import bcosified_densenet
    from bcos_b200.clip_rn import bcosified_clip_rn50
    from bcos_b200.vit import bcosified_simple_vit
    cases = [("simple_vit_ti_patch16_224", lambda: bcosified_simple_vit("simple_vit_ti_patch16_224"), 128, "logit"),
             ("simple_vit_b_patch16_224", lambda: bcosified_simple_vit("simple_vit_b_patch16_224"), 64, "logit"),
             ("clip_rn50", lambda: bcosified_clip_rn50(), 64, "cos"),
             ("densenet121", lambda: bcosified_densenet("densenet121"), 32, "logit")]
    rows = []
    for name, make, nb, target in cases:
        m = make()
        sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 0)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().eval()
        x6 = synth.to_bcos_input(synth.synth_images_u8(nb, 224, 5)).cuda()
        tvec = torch.nn.functional.normalize(torch.randn(1024, device="cuda"), dim=0)
        for mode in ("bf16", "parity"):
            M.set_precision(mode)

            def fwd():
                with torch.inference_mode():
                    return m(x6)

            def fwd_explain():
                xb = x6.clone().requires_grad_(True)
                with torch.enable_grad(), m.explanation_mode():
                    out = m(xb)
                    tgt = out.max(1).values.sum() if target == "logit" else torch.nn.functional.cosine_similarity(out, tvec[None], dim=1).sum()
                    tgt.backward(inputs=[xb])
                return (xb.detach() * xb.grad).sum(1)

            t_f, t_fe = timed(fwd), timed(fwd_explain)
            rows.append({"model": name, "batch": nb, "precision": mode, "fwd_ms": round(t_f, 2), "fwd_img_s": round(nb / t_f * 1e3, 1),
                         "fwd_explain_ms": round(t_fe, 2), "fwd_explain_img_s": round(nb / t_fe * 1e3, 1)})
            print(rows[-1], file=sys.stderr)
        del m
        torch.cuda.empty_cache()
    M.set_precision("parity")
    return rows


def norm_rows():
    rows = []
    for name, make, shape in [("GroupNormUncentered2d(32, 256)", lambda: M.GroupNormUncentered2d(32, 256), (256, 256, 56, 56)),
                              ("DetachableGNLayerNorm2d(256)", lambda: M.DetachableGNLayerNorm2d(256), (256, 256, 56, 56)),
                              ("PositionNormUncentered2d(256)", lambda: M.PositionNormUncentered2d(256), (256, 256, 56, 56)),
                              ("DetachablePositionNorm2d(1024)", lambda: M.DetachablePositionNorm2d(1024), (256, 1024, 14, 14))]:
        mod = make().cuda()
        mod.bias = None
        mod.set_explanation_mode(True)
        x = torch.randn(*shape, device="cuda")
        nbytes = x.numel() * 4
        with torch.no_grad():
            t_f = timed(lambda: mod(x), 3, 10)
        xg = x.clone().requires_grad_(True)
        y = mod(xg)
        g = torch.randn_like(y)
        t_b = timed(lambda: torch.autograd.grad(y, [xg], g, retain_graph=True), 3, 10)
        rows.append({"module": name, "shape": list(shape), "fwd_ms": round(t_f, 3), "fwd_GBs_algorithmic": round(2 * nbytes / t_f / 1e6, 1),
                     "explain_bwd_ms": round(t_b, 3), "explain_bwd_GBs_algorithmic": round(2 * nbytes / t_b / 1e6, 1)})
        print(rows[-1], file=sys.stderr)
    return rows


if __name__ == "__main__":
    out = {"device": torch.cuda.get_device_name(0), "norm_kernels": norm_rows(), "models": model_rows()}
    print(json.dumps(out, indent=1))
