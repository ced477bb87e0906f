import os, sys, json, numpy as np, torch
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcos_b200
import bcos_oracle as OR
from bcos_b200.engine import ViTPlan
from bcos_b200.models import vit_state_shapes, synthetic_vit_plan
from bcos_b200.utils import synth
for arch in ("simple_vit_ti_patch16_224", "simple_vit_b_patch16_224"):
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"{arch}_b2.npz"))
    sd = synth.synth_state_dict(vit_state_shapes(arch), int(gold["seed"]))
    x6 = synth.to_bcos_input(gold["images_u8"]).cuda()
    for bp in (2, 1):
        plan = ViTPlan(arch, sd, 2, mode="parity", device="cuda", branch_planes=bp)
        out = plan.explain(x6)
        torch.cuda.synchronize()
        m = OR.parity_metrics(out["logits"].float().cpu(), out["contribution_map"].float().cpu(), torch.from_numpy(gold["logits"]), torch.from_numpy(gold["contribution_map"]))
        print(json.dumps({"arch": arch, "branch_planes": bp, **{k: (round(v, 8) if isinstance(v, float) else v) for k, v in m.items()}}), flush=True)
        del plan
    for bp in (2, 1):
        B = 256
        plan = synthetic_vit_plan(arch, B, mode="parity", device="cuda", input_u8=True, branch_planes=bp)
        x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat(8, 1, 1, 1).cuda()
        plan.load_input(x); plan.capture()
        for _ in range(2): plan.replay_all()
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(5): plan.replay_forward()
        e[1].record()
        for _ in range(5): plan.replay_all()
        e[2].record(); torch.cuda.synchronize()
        print(json.dumps({"arch": arch, "branch_planes": bp, "batch": B, "fwd_ms": round(e[0].elapsed_time(e[1]) / 5, 3), "fwd_explain_ms": round(e[1].elapsed_time(e[2]) / 5, 3),
                          "img_s": round(B / (e[1].elapsed_time(e[2]) / 5) * 1e3, 1)}), flush=True)
        del plan; torch.cuda.empty_cache()
