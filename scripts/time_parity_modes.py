"""ResNet-50 forward + explanation at batch 256 in the contract-meeting operand modes (CUDA-graph replay, CUDA events)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcos_b200  # noqa: E402,F401
from bcos_b200.models import synthetic_resnet_plan  # noqa: E402

for dt, pl in (("fp16", 2), ("bf16", 2), ("bf16", 3)):
    plan = synthetic_resnet_plan("resnet50", 256, planes=pl, dtype=dt, device="cuda", input_u8=True,
                                 seed_scale=4096.0 if dt == "fp16" else 1.0)
    plan.capture()
    for _ in range(2):
        plan.replay_all()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.replay_all()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"timing": "resnet50 b256", "dtype": dt, "planes": pl, "hp_accum": True, "ms_per_step": round(ms, 2),
                      "img_s": round(256 / ms * 1e3)}), flush=True)
    del plan
    torch.cuda.empty_cache()
