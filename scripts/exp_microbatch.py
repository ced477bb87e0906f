"""L2-resident micro-batching experiment (SURVEY 7 hard part 3, VERDICT r1 "next" 5).

  python scripts/exp_microbatch.py [--batches 16,32,64,128,256] [--modes fp16:2:1,bf16:1:1]

ResNet-50 forward + explanation of 256 images as 256/B graph replays of a batch-B plan: with B small enough the conv -> conv
tensors of a micro-batch (B x 56 x 56 x 256 x 2 B x planes at the widest point) stay in the 126 MB L2 between the launch that
writes them and the launch that reads them.  Reports ms per 256 images.  One JSON line per (mode, B).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bcos_b200  # noqa: E402,F401
from bcos_b200.engine import ResNetPlan  # noqa: E402
from bcos_b200.models import resnet_state_shapes  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="16,32,64,128,256")
    ap.add_argument("--modes", default="fp16:2:1,bf16:1:1")
    ap.add_argument("--total", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    sd = synth.synthetic_checkpoint("resnet50", resnet_state_shapes("resnet50"))
    for mode in a.modes.split(","):
        dtype, planes, eplanes = mode.split(":")
        planes, eplanes = int(planes), int(eplanes)
        for B in [int(b) for b in a.batches.split(",")]:
            x = torch.from_numpy(synth.synth_images_u8(min(B, 32), 224, 3)).repeat(max(B // 32, 1), 1, 1, 1)[:B].cuda()
            plan = ResNetPlan("resnet50", sd, B, planes=planes, explain_planes=eplanes, dtype=dtype, device="cuda", input_u8=True,
                              seed_scale=4096.0 if dtype == "fp16" else 1.0)
            plan.load_input(x)
            plan.capture()
            n = a.total // B
            for _ in range(3 * n):
                plan.replay_all()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps * n):
                plan.replay_all()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            print(json.dumps({"mode": mode, "micro_batch": B, "replays_per_step": n, "ms_per_%d_images" % a.total: round(ms, 3),
                              "img_per_s": round(a.total / ms * 1e3, 1)}), flush=True)
            del plan
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
