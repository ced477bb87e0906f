"""Workload for ncu: one warm eager step of the fused SimpleViT plan, then the launches of ONE encoder (forward and explanation)
between cudaProfilerStart / Stop.

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_vit python scripts/profile_vit.py
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bcos_b200.models import synthetic_vit_plan
from bcos_b200.utils import synth

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="simple_vit_b_patch16_224")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--mode", default="parity")
ap.add_argument("--encoder", type=int, default=5)
a = ap.parse_args()
plan = synthetic_vit_plan(a.arch, a.batch, mode=a.mode, device="cuda", input_u8=True)
x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat((a.batch + 31) // 32, 1, 1, 1)[:a.batch].cuda()
plan.load_input(x)
plan.autotune()
plan.run_forward(); plan.run_explain()
torch.cuda.synchronize()
sel = [o for o in plan.fwd_ops + plan.bwd_ops if f"encoder_{a.encoder}." in o.name]
print("profiled launches:", [o.name for o in sel], file=sys.stderr)
torch.cuda.cudart().cudaProfilerStart()
for o in sel:
    o.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
