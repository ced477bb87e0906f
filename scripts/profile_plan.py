"""Workload for ncu: one warm eager step of a fused plan, then the launches whose name contains one of --match between
cudaProfilerStart / Stop.

  ncu --set full --clock-control none --profile-from-start off -o /tmp/rep python scripts/profile_plan.py --plan densenet121 --match denselayer12.
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
from bcos_b200 import models as M
from bcos_b200.engine import CLIPViTPlan
from bcos_b200.utils import synth

ap = argparse.ArgumentParser()
ap.add_argument("--plan", default="vit_b", choices=["vit_b", "vit_ti", "densenet121", "clip_rn50", "clip_vit"])
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--mode", default="parity")
ap.add_argument("--match", default="encoder_5.")
a = ap.parse_args()
B = a.batch
if a.plan in ("vit_b", "vit_ti"):
    plan = M.synthetic_vit_plan("simple_vit_b_patch16_224" if a.plan == "vit_b" else "simple_vit_ti_patch16_224", B, mode=a.mode, device="cuda", input_u8=True)
elif a.plan == "densenet121":
    plan = M.synthetic_densenet_plan("densenet121", B, mode=a.mode, device="cuda", input_u8=True)
elif a.plan == "clip_rn50":
    plan = M.synthetic_clip_rn50_plan(B, mode=a.mode, device="cuda", input_u8=True)
else:
    plan = CLIPViTPlan(synth.synth_state_dict(M.clip_vit_state_shapes(), 0), B, mode=a.mode, device="cuda", input_u8=True)
x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat((B + 31) // 32, 1, 1, 1)[:B].cuda()
plan.load_input(x)
plan.autotune()
plan.run_forward(); plan.run_explain()
torch.cuda.synchronize()
keys = [k for k in a.match.split(",") if k]
sel = [o for o in plan.fwd_ops + plan.bwd_ops if any(k in o.name for k in keys)]
print("profiled launches:", [o.name for o in sel], file=sys.stderr)
torch.cuda.cudart().cudaProfilerStart()
for o in sel:
    o.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
