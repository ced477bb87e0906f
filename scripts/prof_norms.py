"""ncu target: one forward + one explanation backward of each group / position norm case at ResNet-like sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bcos_b200  # noqa: E402,F401
import bcos_b200.modules as M  # noqa: E402

for make, shape in [(lambda: M.GroupNormUncentered2d(32, 256), (256, 256, 56, 56)),
                    (lambda: M.DetachableGNLayerNorm2d(256), (256, 256, 56, 56)),
                    (lambda: M.PositionNormUncentered2d(256), (256, 256, 56, 56)),
                    (lambda: M.DetachablePositionNorm2d(1024), (256, 1024, 14, 14))]:
    mod = make().cuda()
    mod.bias = None
    mod.set_explanation_mode(True)
    x = torch.randn(*shape, device="cuda").requires_grad_(True)
    for _ in range(2):
        y = mod(x)
        torch.autograd.grad(y, [x], torch.ones_like(y))
    torch.cuda.synchronize()
