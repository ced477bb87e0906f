"""Throughput of the fused SimpleViT plans (BASELINE config 3) on one B200: forward + explanation, CUDA-graph replays.

  python scripts/exp_vit_plan.py [--archs simple_vit_ti_patch16_224,simple_vit_b_patch16_224] [--batches 256] [--modes parity,throughput]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bcos_b200  # noqa: E402,F401
from bcos_b200.engine import ops as O  # noqa: E402
from bcos_b200.models import synthetic_vit_plan  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--archs", default="simple_vit_ti_patch16_224,simple_vit_b_patch16_224")
    ap.add_argument("--batches", default="256")
    ap.add_argument("--modes", default="parity,throughput")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--layers", default=None, help="write per-launch timings (eager, CUDA events) of the last configuration here")
    a = ap.parse_args()
    for arch in a.archs.split(","):
        for mode in a.modes.split(","):
            for B in [int(b) for b in a.batches.split(",")]:
                plan = synthetic_vit_plan(arch, B, mode=mode, device="cuda", input_u8=True)
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
                flops = plan.gemm_flops()
                print(json.dumps({"arch": arch, "mode": mode, "batch": B, "fwd_ms": round(fwd, 3), "fwd_explain_ms": round(full, 3),
                                  "fwd_img_s": round(B / fwd * 1e3, 1), "fwd_explain_img_s": round(B / full * 1e3, 1),
                                  "launches": plan.num_launches(), "executed_gemm_tflops": round(flops / full / 1e9, 1)}), flush=True)
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
                        rows.append({"name": op.name, "kind": type(op).__name__, "ms": e0.elapsed_time(e1) / 3})
                    json.dump({"arch": arch, "mode": mode, "batch": B, "rows": rows}, open(a.layers, "w"))
                del plan
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
