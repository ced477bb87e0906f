"""Throughput of the fused CLIP RN50 plan (BASELINE config 4: embedding + explanation, batch 512) on one B200.

  python scripts/exp_clip_plan.py [--batches 256,512] [--modes parity,throughput]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcos_b200  # noqa: E402,F401
from bcos_b200.models import synthetic_clip_rn50_plan  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="256,512")
    ap.add_argument("--modes", default="parity,throughput")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--layers", default=None, help="write per-launch timings (eager, CUDA events) of the last configuration here")
    a = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    t = torch.nn.functional.normalize(torch.randn(1024, generator=g), dim=0)
    for mode in a.modes.split(","):
        for B in [int(b) for b in a.batches.split(",")]:
            plan = synthetic_clip_rn50_plan(B, mode=mode, device="cuda", input_u8=True)
            x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat(B // 32, 1, 1, 1).cuda()
            plan.load_input(x)
            plan.capture()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            for _ in range(2):
                plan.explain_direction(None, t)
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(a.reps):
                plan.replay_forward()
            ev[1].record()
            for _ in range(a.reps):
                plan.embed(None)
            ev[2].record()
            for _ in range(a.reps):
                plan.explain_direction(None, t)
            ev[3].record()
            torch.cuda.synchronize()
            trunk, emb, full = (ev[i].elapsed_time(ev[i + 1]) / a.reps for i in range(3))
            print(json.dumps({"mode": mode, "batch": B, "trunk_fwd_ms": round(trunk, 3), "embed_ms": round(emb, 3),
                              "embed_explain_ms": round(full, 3), "embed_img_s": round(B / emb * 1e3, 1),
                              "embed_explain_img_s": round(B / full * 1e3, 1), "launches_fwd": len(plan.fwd_ops),
                              "launches_bwd": len(plan.bwd_ops)}), flush=True)
            if a.layers:
                rows = []
                for op in plan.fwd_ops + plan.bwd_ops:
                    op.run()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        op.run()
                    e1.record()
                    e1.synchronize()
                    row = {"name": op.name, "kind": type(op).__name__, "ms": round(e0.elapsed_time(e1) / 3, 4)}
                    if hasattr(op, "ktot"):
                        row.update(M=op.M, N=op.n, K=op.ktot, hp=bool(getattr(op, "hp", False)))
                    rows.append(row)
                json.dump({"mode": mode, "batch": B, "rows": rows}, open(a.layers, "w"), indent=1)
            del plan
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
