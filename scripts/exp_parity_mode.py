"""Parity-mode (precision planes + fp32-faithful accumulation) accuracy and cost on B200.

  python scripts/exp_parity_mode.py [--chunks 1,2,4] [--boxes 1,0] [--layers out.json]

Accuracy: ResNet-50 batch 4 / ResNet-18 batch 8 against the reference golden vectors.  Cost: ResNet-50 batch 256,
CUDA-graph replays.  One JSON line per configuration.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcos_b200  # noqa: E402,F401
import bcos_oracle as OR  # noqa: E402
from bcos_b200 import _lib  # noqa: E402
from bcos_b200.engine import ResNetPlan, ops as O  # noqa: E402
from bcos_b200.models import resnet_state_shapes  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402


def golden_state(arch, gold):
    sd = synth.synth_state_dict(resnet_state_shapes(arch), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    return sd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", default="1,2,4")
    ap.add_argument("--boxes", default="1,0")
    ap.add_argument("--modes", default="fp16:2")
    ap.add_argument("--layers", default=None)
    ap.add_argument("--batch", type=int, default=256)
    a = ap.parse_args()
    lib = _lib.load()
    chunks = [int(c) for c in a.chunks.split(",")]
    boxes = [int(c) for c in a.boxes.split(",")]
    modes = [(m.split(":")[0], int(m.split(":")[1]), int((m.split(":") + [m.split(":")[1]])[2])) for m in a.modes.split(",")]
    for dtype, planes, eplanes in modes:
        ss = 4096.0 if dtype == "fp16" else 1.0
        for bx in boxes:
            lib.bcosk_set_hp_boxes(bx)
            for ch in chunks:
                lib.bcosk_set_hp_chunk(ch)
                for arch, batch in (("resnet50", 4), ("resnet18", 8)):
                    gold = np.load(os.path.join(ROOT, "tests", "golden", f"{arch}_b{batch}.npz"))
                    sd = golden_state(arch, gold)
                    x6 = synth.to_bcos_input(gold["images_u8"])
                    plan = ResNetPlan(arch, sd, batch, planes=planes, explain_planes=eplanes, dtype=dtype, device="cuda", seed_scale=ss)
                    out = plan.explain(x6)
                    torch.cuda.synchronize()
                    m = OR.parity_metrics(out["logits"].float().cpu(), out["contribution_map"].float().cpu(),
                                          torch.from_numpy(gold["logits"]), torch.from_numpy(gold["contribution_map"]))
                    m = {k: (round(v, 9) if isinstance(v, float) else v) for k, v in m.items()}
                    m64 = OR.parity_metrics(out["logits"].float().cpu(), out["contribution_map"].float().cpu(),
                                            torch.from_numpy(gold["logits_fp64"]), torch.from_numpy(gold["contribution_map_fp64"]))
                    m["vs_fp64_maxabs"] = round(m64["map_maxabs_over_range"], 9)
                    m["ref_vs_fp64_maxabs"] = round(float(gold["fp32_noise_floor_maxabs_over_range"]), 9)
                    print(json.dumps({"arch": arch, "dtype": dtype, "planes": planes, "explain_planes": eplanes, "chunk": ch, "boxes": bx, **m}), flush=True)
                    del plan
                    torch.cuda.empty_cache()
                sd = synth.synthetic_checkpoint("resnet50", resnet_state_shapes("resnet50"))
                x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat(a.batch // 32, 1, 1, 1).cuda()
                plan = ResNetPlan("resnet50", sd, a.batch, planes=planes, explain_planes=eplanes, dtype=dtype, device="cuda", input_u8=True, seed_scale=ss)
                plan.load_input(x)
                plan.capture()
                for _ in range(3):
                    plan.replay_all()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    plan.replay_all()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                print(json.dumps({"timing": f"resnet50 b{a.batch}", "dtype": dtype, "planes": planes, "explain_planes": eplanes, "chunk": ch, "boxes": bx,
                                  "ms_per_step": round(ms, 2), "img_s": round(a.batch / ms * 1e3)}), flush=True)
                if a.layers and bx == boxes[0] and ch == chunks[0]:
                    all_ops = plan.fwd_ops + plan.bwd_ops
                    per = [0.0] * len(all_ops)
                    reps = 3
                    for rep in range(reps + 1):
                        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(all_ops) + 1)]
                        evs[0].record()
                        for i, o in enumerate(all_ops):
                            o.run()
                            evs[i + 1].record()
                        torch.cuda.synchronize()
                        if rep:
                            for i in range(len(all_ops)):
                                per[i] += evs[i].elapsed_time(evs[i + 1]) / reps
                    rows = []
                    for o, t in zip(all_ops, per):
                        row = {"name": o.name, "kind": type(o).__name__, "ms": round(t, 4)}
                        if isinstance(o, O.IgemmOp):
                            row.update(M=o.M, N=o.n, K=o.ktot, gflop=round(o.algo_flops / 1e9, 2), mbytes=round(o.algo_bytes() / 1e6, 1),
                                       tflops=round(o.flops() / (t * 1e-3) / 1e12, 1), gbs=round(o.algo_bytes() / (t * 1e-3) / 1e9))
                        rows.append(row)
                    with open(a.layers.replace(".json", f"_{dtype}x{planes}e{eplanes}.json"), "w") as fh:
                        json.dump({"batch": a.batch, "dtype": dtype, "planes": planes, "step_ms_eager": sum(per), "rows": rows}, fh, indent=1)
                del plan
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
