"""Experiment: where the warps of the persistent igemm kernel wait (needs a -DBCOSK_TIMING2 build; never shipped)."""
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
cap = 256
labels = ["A/B producer: wait slot free", "mma: wait stage full", "mma: wait accumulator free", "input producer: wait tile free",
          "epi: wait row-side + input tile", "epi: wait accumulator", "epi: math + barriers + store issue", "tiles"]
names = sys.argv[1:] or ["model.layer1.1.conv3", "model.layer1.2.conv1.dgrad", "model.layer2.2.conv1.dgrad"]
for sched in (2, 3):
    for nm in names:
        op = [o for o in ops if o.name == nm][0]
        op.block_n, op.sched = 64, sched
        op.run(); torch.cuda.synchronize()
        buf = torch.zeros(cap * 8, dtype=torch.int64, device="cuda")
        lib.bcosk_debug_set_timing(ctypes.c_void_p(buf.data_ptr()), cap)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(); op.run(); ev[1].record(); torch.cuda.synchronize()
        lib.bcosk_debug_set_timing(None, 0)
        t = buf.view(cap, 8).cpu().double()
        t = t[t[:, 7] > 0]
        tiles = t[:, 7]
        us = ev[0].elapsed_time(ev[1]) * 1e3
        print(f"== sched {sched} {nm}: {t.shape[0]} CTAs x {tiles.mean():.1f} tiles, launch {us:.0f} us = {us*1.965e3/tiles.mean():.0f} cycles per tile")
        for i in range(7):
            print(f"   {labels[i]:36s} {(t[:, i] / tiles).mean():8.0f} cycles per tile")
