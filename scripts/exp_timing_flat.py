"""Experiment: where a flat-window CTA spends its cycles (needs a -DBCOSK_TIMING build; never shipped)."""
import os, sys, ctypes
os.environ["BCOSK_EXTRA_NVCC_FLAGS"] = "-DBCOSK_TIMING"
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
for _ in range(2):
    for o in ops: o.run()
torch.cuda.synchronize()
cap = 256
labels = ["producer: wait window free", "mma: wait window full", "mma: wait accumulator free", "mma: issue + commit",
          "epi: wait accumulator", "epi: tmem+math+stage", "epi: barrier+copy-out", "tiles"]
for nm in sys.argv[1:] or ["stem", "stem.dgrad"]:
    op = [o for o in ops if o.name == nm][0]
    buf = torch.zeros(cap * 8, dtype=torch.int64, device="cuda")
    lib.bcosk_debug_set_timing(ctypes.c_void_p(buf.data_ptr()), cap)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); op.run(); ev[1].record(); torch.cuda.synchronize()
    lib.bcosk_debug_set_timing(None, 0)
    t = buf.view(cap, 8).cpu().double()
    t = t[t[:, 7] > 0]
    tiles = t[:, 7]
    print(f"== {nm}: {t.shape[0]} CTAs, {tiles.mean():.1f} tiles each, launch {ev[0].elapsed_time(ev[1])*1e3:.0f} us")
    for i, lb in enumerate(labels[:7]):
        print(f"   {lb:30s} {(t[:, i] / tiles).mean():8.0f} cycles per tile")
