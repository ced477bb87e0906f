"""Throughput of the fused DenseNet-121 plan on one B200: forward + explanation, CUDA-graph replays.

  python scripts/exp_densenet_plan.py [--batches 128,256] [--modes parity,throughput] [--layers out.json]
"""
import argparse
import json
import os
import re
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bcos_b200  # noqa: E402,F401
from bcos_b200.models import synthetic_densenet_plan  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="256")
    ap.add_argument("--modes", default="parity,throughput")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--layers", default=None)
    a = ap.parse_args()
    for mode in a.modes.split(","):
        for B in [int(b) for b in a.batches.split(",")]:
            plan = synthetic_densenet_plan("densenet121", B, mode=mode, device="cuda", input_u8=True)
            x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat((B + 31) // 32, 1, 1, 1)[:B].cuda()
            plan.load_input(x)
            plan.capture()
            for _ in range(2):
                plan.replay_all()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(a.reps):
                plan.replay_forward()
            ev[1].record()
            for _ in range(a.reps):
                plan.replay_all()
            ev[2].record()
            torch.cuda.synchronize()
            fwd, full = (ev[i].elapsed_time(ev[i + 1]) / a.reps for i in range(2))
            print(json.dumps({"arch": "densenet121", "mode": mode, "batch": B, "fwd_ms": round(fwd, 3), "fwd_explain_ms": round(full, 3),
                              "fwd_img_s": round(B / fwd * 1e3, 1), "fwd_explain_img_s": round(B / full * 1e3, 1),
                              "launches": plan.num_launches(), "finite": bool(torch.isfinite(plan.cmap).all()),
                              "cmap_absmax": float(plan.cmap.abs().max())}), flush=True)
            if a.layers:
                agg, cnt = defaultdict(float), defaultdict(int)
                for op in plan.fwd_ops + plan.bwd_ops:
                    op.run()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        op.run()
                    e1.record()
                    e1.synchronize()
                    key = type(op).__name__ + ":" + re.sub(r"\d+", "N", op.name.split("model.features.")[-1])
                    agg[key] += e0.elapsed_time(e1) / 3
                    cnt[key] += 1
                json.dump({"mode": mode, "batch": B, "by_kind": {k: [cnt[k], agg[k]] for k in sorted(agg, key=lambda k: -agg[k])}},
                          open(a.layers, "w"), indent=1)
            del plan
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
