"""Experiment: does the forward pass of the (chaotic, random-init) ResNets tolerate ONE-plane weights (a0 b0 + a1 b0: 2/3 of the tensor
work) while the activations keep two fp16 planes?  ResNet-50 batch 4 / ResNet-18 batch 8 against the reference goldens."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcos_b200  # noqa
import bcos_oracle as OR
from bcos_b200.engine import ResNetPlan
from bcos_b200.engine.base import PlanBase
from bcos_b200.models import resnet_state_shapes
from bcos_b200.utils import synth
for wp in (None, 1):
    PlanBase.fwd_w_planes = wp
    for arch, batch in (("resnet50", 4), ("resnet18", 8)):
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"{arch}_b{batch}.npz"))
        sd = synth.synth_state_dict(resnet_state_shapes(arch), int(gold["seed"]))
        off = 0
        for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
            sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy()); off += n
        plan = ResNetPlan(arch, sd, batch, mode="parity", device="cuda")
        out = plan.explain(synth.to_bcos_input(gold["images_u8"]).cuda())
        torch.cuda.synchronize()
        m = OR.parity_metrics(out["logits"].float().cpu(), out["contribution_map"].float().cpu(), torch.from_numpy(gold["logits"]), torch.from_numpy(gold["contribution_map"]))
        m64 = OR.parity_metrics(out["logits"].float().cpu(), out["contribution_map"].float().cpu(), torch.from_numpy(gold["logits_fp64"]), torch.from_numpy(gold["contribution_map_fp64"]))
        print(json.dumps({"arch": arch, "forward_weight_planes": wp or 2, "argmax_equal": m["argmax_equal"], "logit_rel_err": m["logit_rel_err"], "map_cos_min": m["map_cos_min"],
                          "map_maxabs_vs_fp32_ref": m["map_maxabs_over_range"], "map_maxabs_vs_fp64": m64["map_maxabs_over_range"]}), flush=True)
        del plan
