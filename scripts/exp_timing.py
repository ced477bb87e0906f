"""Experiment: per-CTA phase timing of the per-tile igemm kernel (needs a -DBCOSK_TIMING build)."""
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
names = sys.argv[1:] or ["model.layer1.1.conv3", "model.layer1.1.conv1.dgrad", "model.layer2.1.conv3", "model.layer3.1.conv2", "stem"]
cap = 16384
labels = ["entry->setup", "setup->norm", "norm->in_tile", "in_tile->acc", "acc->math", "math->store_read", "store_read->exit"]
for nm in names:
    op = [o for o in ops if o.name == nm][0]
    buf = torch.zeros(cap * 8, dtype=torch.int64, device="cuda")
    lib.bcosk_debug_set_timing(ctypes.c_void_p(buf.data_ptr()), cap)
    op.run(); torch.cuda.synchronize()
    lib.bcosk_debug_set_timing(None, 0)
    t = buf.view(cap, 8).cpu().double()
    t = t[(t[:, 0] > 0) & (t[:, 7] > 0)]
    n = t.shape[0]
    d = (t[:, 1:] - t[:, :-1])
    life = t[:, 7] - t[:, 0]
    print(f"== {nm}: {n} CTAs sampled, block_n {op.resolved_block_n()}, lifetime mean {life.mean():.0f} cyc (median {life.median():.0f})")
    for i, lb in enumerate(labels):
        col = d[:, i]
        print(f"   {lb:20s} mean {col.mean():8.0f}  median {col.median():8.0f}  p90 {col.quantile(0.9):8.0f} cyc")
