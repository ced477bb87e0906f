"""Accuracy and cost of the operand formats of the fused ResNet plan against the reference golden vectors:
bf16 / fp16, 1-3 precision planes, with and without the fp32-faithful accumulation, optional seed scaling (fp16 gradients).
Prints one JSON line per configuration; batch-256 timing for the interesting ones."""
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
from bcos_b200.engine import ResNetPlan  # noqa: E402
from bcos_b200.models import resnet_state_shapes  # noqa: E402
from bcos_b200.utils import synth  # noqa: E402

FORCE_HP = None


class Plan(ResNetPlan):
    def __setattr__(self, k, v):
        if k == "hp_accum" and FORCE_HP is not None:
            v = FORCE_HP
        object.__setattr__(self, k, v)


def golden_state(arch, gold):
    sd = synth.synth_state_dict(resnet_state_shapes(arch), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    return sd


def main():
    global FORCE_HP
    cfgs = [("bf16", 1, None, 1.0), ("fp16", 1, None, 1.0), ("fp16", 1, None, 4096.0), ("fp16", 1, True, 4096.0),
            ("bf16", 2, None, 1.0), ("bf16", 2, False, 1.0), ("fp16", 2, None, 4096.0), ("fp16", 2, False, 4096.0),
            ("bf16", 3, None, 1.0)]
    for arch, batch in (("resnet50", 4), ("resnet18", 8)):
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"{arch}_b{batch}.npz"))
        sd = golden_state(arch, gold)
        x6 = synth.to_bcos_input(gold["images_u8"])
        for dtype, planes, hp, ss in cfgs:
            FORCE_HP = hp
            try:
                plan = Plan(arch, sd, batch, planes=planes, dtype=dtype, device="cuda", seed_scale=ss)
                out = plan.explain(x6)
                torch.cuda.synchronize()
                m = OR.parity_metrics(out["logits"].float().cpu(), out["contribution_map"].float().cpu(),
                                      torch.from_numpy(gold["logits"]), torch.from_numpy(gold["contribution_map"]))
                m = {k: (round(v, 9) if isinstance(v, float) else v) for k, v in m.items()}
            except Exception as e:  # noqa: BLE001
                m = {"error": repr(e)[:200]}
            print(json.dumps({"arch": arch, "dtype": dtype, "planes": planes, "hp_accum": hp, "seed_scale": ss, **m}), flush=True)
            del plan
            torch.cuda.empty_cache()
    # cost at the benchmark batch
    sd = synth.synthetic_checkpoint("resnet50", resnet_state_shapes("resnet50"))
    x = torch.from_numpy(synth.synth_images_u8(256, 224, 3)).cuda()
    for dtype, planes, hp in (("bf16", 1, None), ("fp16", 1, None), ("bf16", 2, None), ("bf16", 2, False), ("fp16", 2, False)):
        FORCE_HP = hp
        plan = Plan("resnet50", sd, 256, planes=planes, dtype=dtype, device="cuda", input_u8=True, seed_scale=4096.0 if dtype == "fp16" else 1.0)
        plan.capture()
        plan.load_input(x)
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
        print(json.dumps({"timing": "resnet50 b256", "dtype": dtype, "planes": planes, "hp_accum": hp, "ms_per_step": round(ms, 2),
                          "img_s": round(256 / ms * 1e3)}), flush=True)
        del plan
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
