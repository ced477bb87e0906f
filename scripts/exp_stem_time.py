import sys, os, torch
sys.path.insert(0, "/root/repo")
import bcos_b200
from bcos_b200.models import synthetic_resnet_plan
from bcos_b200.utils import synth
for im2col in (True, False):
    plan = synthetic_resnet_plan("resnet50", 256, mode="parity", device="cuda", input_u8=True) if im2col else None
    if not im2col:
        from bcos_b200.engine.base import PlanBase
        PlanBase.stem_im2col = False
        plan = synthetic_resnet_plan("resnet50", 256, mode="parity", device="cuda", input_u8=True)
    x = torch.from_numpy(synth.synth_images_u8(32, 224, 3)).repeat(8, 1, 1, 1).cuda()
    plan.load_input(x)
    plan.run_forward()
    torch.cuda.synchronize()
    for op in plan.fwd_ops[:4]:
        op.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            op.run()
        e1.record(); e1.synchronize()
        print(im2col, op.name, type(op).__name__, round(e0.elapsed_time(e1) / 5, 3), "ms")
    del plan
    torch.cuda.empty_cache()
