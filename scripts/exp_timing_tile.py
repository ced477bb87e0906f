"""Experiment: where the warps of a per-tile igemm CTA wait (needs a -DBCOSK_TIMING2 build; never shipped)."""
import os, sys, ctypes
os.environ["BCOSK_EXTRA_NVCC_FLAGS"] = "-DBCOSK_TIMING2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bcos_b200 import build as B
B.build(force=True)
from bcos_b200 import _lib as L
from bcos_b200.models import synthetic_resnet_plan
from bcos_b200.utils import synth
lib = L.load()
plan = synthetic_resnet_plan("resnet50", 256, mode="throughput", device="cuda", input_u8=True)
imgs = torch.from_numpy(synth.synth_images_u8(32, 224, 5)).repeat(8, 1, 1, 1).contiguous()
plan.load_input(imgs)
ops = plan.fwd_ops + plan.bwd_ops
for o in ops: o.run()
torch.cuda.synchronize()
cap = 8192
labels = ["producer: wait slot free (total)", "mma: wait stage full (total)", "-", "mma: issue + commit (total)",
          "epi: wait accumulator", "epi: math", "-", "-"]
names = [n for n in sys.argv[1:] if not n.startswith("cl=")] or ["model.layer3.1.conv2", "model.layer3.1.conv1", "model.layer2.1.conv2"]
for mode in (1, 3):
    lib.bcosk_set_cluster(mode)
    for nm in names:
        op = [o for o in ops if o.name == nm][0]
        op.run(); torch.cuda.synchronize()
        buf = torch.zeros(cap * 8, dtype=torch.int64, device="cuda")
        lib.bcosk_debug_set_timing(ctypes.c_void_p(buf.data_ptr()), cap)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); op.run(); ev[1].record(); torch.cuda.synchronize()
        lib.bcosk_debug_set_timing(None, 0)
        t = buf.view(cap, 8).cpu().double()
        stages = op.ktot // 64
        print(f"== cluster mode {mode} {nm}: K stages {stages}, launch {ev[0].elapsed_time(ev[1])*1e3:.0f} us")
        for i in (0, 1, 3, 4, 5):
            col = t[:, i]; col = col[col > 0]
            if col.numel():
                print(f"   {labels[i]:34s} mean {col.mean():9.0f} cyc   per K stage {col.mean()/stages:7.0f}   ({col.numel()} CTAs)")
