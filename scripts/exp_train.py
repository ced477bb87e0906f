"""Fine-tuning step on B200 against the fp32 oracle: loss, logits, per-tensor gradient cosine / relative error, updated weights."""
import argparse, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcos_b200  # noqa
import bcos_oracle as OR
from bcos_b200.engine import ResNetTrainPlan
from bcos_b200.utils import synth

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="resnet18")
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--loss-scale", type=float, default=1.0)
a = ap.parse_args()
sd = synth.synth_state_dict(OR.resnet_state_shapes(a.arch), 0)
imgs = synth.synth_images_u8(a.batch, a.size, 1)
x6 = synth.to_bcos_input(imgs)
labels = torch.arange(a.batch) * 37 % 1000
om = OR.OracleResNet(a.arch, sd)
om.calibrate_bn(x6)
ref = OR.train_step_reference(om, x6, labels)
plan = ResNetTrainPlan(a.arch, sd, a.batch, dtype=a.dtype, device="cuda", image_size=a.size, loss_scale=a.loss_scale)
plan.load_batch(torch.from_numpy(imgs), labels)
plan.forward_backward()
torch.cuda.synchronize()
print("loss", float(plan.loss), "ref", float(ref["loss"]))
lg = plan.logits.cpu()
print("logits rel err", float((lg - ref["logits"]).abs().max() / ref["logits"].abs().max()))
g = plan.gradients()
worst = []
for k, gr in ref["grads"].items():
    mine = g[k].cpu().double().flatten(); r = gr.double().flatten()
    cos = float(torch.dot(mine, r) / (mine.norm() * r.norm() + 1e-300))
    rel = float((mine - r).norm() / (r.norm() + 1e-300))
    worst.append((cos, rel, k, float(r.norm())))
for cos, rel, k, n in worst[::-1][:14]:
    print(f"{k:45s} cos {cos:.5f} rel {rel:.3e} |ref| {n:.3e}")
worst.sort()
print("median cos", sorted(w[0] for w in worst)[len(worst) // 2], "min cos", worst[0][0])
plan.optimizer_step()
torch.cuda.synchronize()
new = plan.state_dict()
errs = []
for k, w in ref["weights"].items():
    d = (new[k].cpu() - w).abs().max().item()
    step = (w - sd[k]).abs().max().item()
    errs.append((d / (step + 1e-30), k))
errs.sort(reverse=True)
print("weight update max-err / max-step (worst 5):", [(round(e, 4), k) for e, k in errs[:5]])
rv = max(float((new[k].cpu() - v).abs().max() / v.abs().max()) for k, v in ref["running_var"].items())
print("running_var rel err", rv)

# ---- localise: gradient wrt the last block's output through the classifier alone (autograd on the oracle's pieces)
if os.environ.get("BCOS_TRAIN_DEBUG"):
    import torch.nn.functional as F
    om2 = OR.OracleResNet(a.arch, sd)
    om2.training = True
    om2.taps = {}
    with torch.no_grad():
        om2.forward(x6)
    last_name = [k for k in om2.taps if k.startswith("model.layer4")][-1]
    xl = om2.taps[last_name].clone().requires_grad_(True)
    wfc = sd["model.fc.linear.weight"]
    fc_out = OR.bcos_conv2d(xl, wfc, None, 1, 0, b=2, detach=False)
    lg2 = OR.logit_layer(F.adaptive_avg_pool2d(fc_out, 1).flatten(1), None, OR.LOGIT_BIAS_1000)
    loss2 = OR.uniform_off_labels_bce(lg2, labels)
    (gx_ref,) = torch.autograd.grad(loss2, [xl])
    # parts: direct path only (scale detached)
    xl2 = om2.taps[last_name].clone().requires_grad_(True)
    lin = F.conv2d(xl2, wfc)
    nrm = OR.patch_norms(xl2.detach(), (1, 1), 1, 0, 1, lin.shape[1])
    out_d = lin * lin.abs() / nrm
    lg3 = OR.logit_layer(F.adaptive_avg_pool2d(out_d, 1).flatten(1), None, OR.LOGIT_BIAS_1000)
    (gx_direct,) = torch.autograd.grad(OR.uniform_off_labels_bce(lg3, labels), [xl2])
    fcL = plan.fc
    nb = a.batch
    gx_mine = fcL.gx.float().cpu().permute(0, 3, 1, 2) / a.loss_scale
    zt = plan.blocks[-1]["y"].t.float().cpu().permute(0, 3, 1, 2)
    T = fcL.gnT.float().cpu().view(nb, 1, zt.shape[2], zt.shape[3]) / a.loss_scale
    tot_mine = gx_mine + zt * T
    def rel(a_, b_): return float((a_ - b_).norm() / b_.norm())
    print("x last block: mine vs oracle", rel(zt, om2.taps[last_name]))
    print("fc dgrad direct: mine vs oracle direct", rel(gx_mine, gx_direct))
    print("fc total grad: mine vs oracle", rel(tot_mine, gx_ref), " norm-path share", float((gx_ref - gx_direct).norm() / gx_ref.norm()))
    print("norm path: mine vs oracle", rel(zt * T, gx_ref - gx_direct))
if os.environ.get("BCOS_TRAIN_DEBUG"):
    lastc = plan.blocks[-1]["convs"][-1]
    M, o = lastc.out.shape
    gz = (fcL.gx.view(M, -1).double() + plan.blocks[-1]["y"].t.view(M, -1).double() * fcL.gnT.double()[:, None]) / a.loss_scale
    mask = (lastc.z.t.view(M, -1) > 0).double()
    S_py = (gz * mask * lastc.out.double()).sum(0)
    gw_py = (S_py * lastc.rstd.double()).cpu()
    key = lastc.bn + ".weight"
    print("bn grad: python-from-buffers vs oracle", rel(gw_py, ref["grads"][key].double()), "| kernel vs python", rel(g[key].cpu().double(), gw_py))
    # oracle's own decomposition: grad wrt bn weight with the incoming gradient taken from the oracle
    print("mask density", float(mask.mean()), "rstd rel err vs oracle var:", )
    import torch.nn.functional as F2
    blkname = plan.blocks[-1]["name"]
    prev_name = plan.blocks[-2]["name"]
    xin = om2.taps[prev_name].clone()
    om2.taps = None
    c1 = om2._conv(blkname + ".conv1", xin, 1, 1, False)
    z1 = F2.relu(om2._bn(blkname + ".bn1", c1, False))
    c2 = om2._conv(blkname + ".conv2", z1, 1, 1, False).detach().requires_grad_(True)
    wbn = sd[blkname + ".bn2.weight"].clone().requires_grad_(True)
    var = c2.var(dim=(0, 2, 3), unbiased=False)
    y2 = wbn[None, :, None, None] * c2 / (var + 1e-5).sqrt()[None, :, None, None]
    zz = F2.relu(y2 + xin)
    (gw_o, gc2_o) = torch.autograd.grad((zz * gx_ref).sum(), [wbn, c2])
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).double()
    print("out2 mine vs oracle", rel(lastc.out.double().cpu(), nhwc(c2.detach())))
    print("z mine vs oracle", rel(lastc.z.t.view(M, -1).double().cpu(), nhwc(zz.detach())))
    print("rstd mine vs oracle", rel(lastc.rstd.double().cpu(), 1.0 / (var.detach().double() + 1e-5).sqrt()))
    print("bn w grad: oracle block-local vs oracle full", rel(gw_o.double(), ref["grads"][key].double()), " python vs block-local", rel(gw_py, gw_o.double()))
    mask_o = nhwc((zz > 0).float())
    print("mask agreement", float((mask_o == mask.cpu()).double().mean()))
    S_o = (nhwc(gx_ref) * mask_o * nhwc(c2.detach())).sum(0) / (var.detach().double() + 1e-5).sqrt()
    print("oracle formula S*rstd vs oracle autograd", rel(S_o, gw_o.double()))
