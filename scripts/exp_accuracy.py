"""GPU experiment: (1) accuracy of the tcgen05 fp32 accumulation vs fp64; (2) where the parity-mode error grows."""
import sys, os, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import bcos_oracle as OR
import emulator as E
import opsutil as U
from bcos_b200 import _lib as L
from bcos_b200.engine import ops as O, ResNetPlan
from bcos_b200.engine.base import PlanBase, Act
from bcos_b200.engine.pack import join_planes
from bcos_b200.utils import synth

g = torch.Generator().manual_seed(0)
print("== (1) GEMM accumulation accuracy, bf16 operands (exact), fp32 out")
for K in (64, 256, 1152, 4608):
    for positive in (False, True):
        plan = PlanBase(1, planes=1, device="cpu", explain=False, b=1.0)
        x = torch.randn(1, 16, 16, K, generator=g)
        w = torch.randn(64, K, 1, 1, generator=g)
        if positive:
            x, w = x.abs(), w.abs()
        xa = Act(x.to(torch.bfloat16), K)
        plan._conv_fwd("g", xa, w.to(torch.bfloat16).float(), 1, 0, 0, bn=None, relu=False, y_f32=True, want_sq=False)
        op = plan.fwd_ops[-1]
        dop = U.to_device(op, "cuda")
        dop.run(); torch.cuda.synchronize()
        A = xa.t.double().view(256, K); B = w.to(torch.bfloat16).double().view(64, K)
        truth = A @ B.t()
        gpu = dop.y.double().cpu().view(256, 64)
        cpu32 = (A.float() @ B.float().t()).double()
        scale = (A.abs() @ B.abs().t())   # sum |terms|
        e_gpu = ((gpu - truth) / scale); e_cpu = ((cpu32 - truth) / scale)
        print(f"K={K:5d} positive={positive}: gpu err/sum|terms| max {e_gpu.abs().max():.2e} mean(signed) {e_gpu.mean():+.2e} rms {e_gpu.pow(2).mean().sqrt():.2e} | cpu fp32 max {e_cpu.abs().max():.2e} rms {e_cpu.pow(2).mean().sqrt():.2e}")

print("== (2) resnet18 @224 B=2 planes=3: GPU vs CPU emulator, per block")
arch, S, nb = "resnet18", 224, 2
import numpy as np
gold = np.load(os.path.join(ROOT, "tests/golden/resnet18_b8.npz"))
sd = synth.synth_state_dict(OR.resnet_state_shapes(arch), 0)
off = 0
for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
    sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy()); off += n
x6 = synth.to_bcos_input(gold["images_u8"][:nb])
om = OR.OracleResNet(arch, {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()})
om.taps = {}
truth = OR.explain_batched(om.forward, x6.double())
cpu = ResNetPlan(arch, sd, nb, planes=3, device="cpu", image_size=S)
cpu.x_in.copy_(x6); E.run(cpu.fwd_ops); E.run(cpu.bwd_ops)
gpu = ResNetPlan(arch, sd, nb, planes=3, device="cuda", image_size=S)
out = gpu.explain(x6); torch.cuda.synchronize()
def rel(a, b): return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max()).item()
print("stem   gpu-vs-truth %.2e  emu-vs-truth %.2e" % (rel(join_planes(gpu.stem_out.t, 3), om.taps["stem"].permute(0, 2, 3, 1)), rel(join_planes(cpu.stem_out.t, 3), om.taps["stem"].permute(0, 2, 3, 1))))
for bg, bc in zip(gpu.blocks, cpu.blocks):
    t = om.taps[bg.name].permute(0, 2, 3, 1)
    print("%-16s gpu-vs-truth %.2e  emu-vs-truth %.2e" % (bg.name, rel(join_planes(bg.y.t, 3), t), rel(join_planes(bc.y.t, 3), t)))
print("gpu vs truth:", OR.parity_metrics(out["logits"], out["contribution_map"], truth["logits"], truth["contribution_map"]))
print("emu vs truth:", OR.parity_metrics(cpu.logits, cpu.cmap, truth["logits"], truth["contribution_map"]))
print("ref32 vs truth:", OR.parity_metrics(torch.from_numpy(gold["logits"][:nb]), torch.from_numpy(gold["contribution_map"][:nb]), truth["logits"], truth["contribution_map"]))
# per-image error localisation
d = (out["contribution_map"].double().cpu() - truth["contribution_map"]).abs()
for i in range(nb):
    idx = d[i].argmax(); print("img", i, "max err at", divmod(idx.item(), S), "value", d[i].max().item(), "range", (truth["contribution_map"][i].max() - truth["contribution_map"][i].min()).item())
