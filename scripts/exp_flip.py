"""Per-image distance of the default-mode maps from the fp32 reference and from its fp64 evaluation (RN50 golden)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bcos_b200  # noqa
from bcos_b200 import _lib
from bcos_b200.engine import ResNetPlan
from bcos_b200.models import resnet_state_shapes
from bcos_b200.utils import synth
lib = _lib.load()
for arch, batch in (("resnet50", 4), ("resnet18", 8)):
    gold = np.load(os.path.join(ROOT, "tests", "golden", f"{arch}_b{batch}.npz"))
    sd = synth.synth_state_dict(resnet_state_shapes(arch), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy()); off += n
    x6 = synth.to_bcos_input(gold["images_u8"])
    ref, ref64 = torch.from_numpy(gold["contribution_map"]).double(), torch.from_numpy(gold["contribution_map_fp64"]).double()
    rng = (ref.flatten(1).max(1).values - ref.flatten(1).min(1).values)
    print(arch, "ref vs fp64 per image:", ((ref - ref64).abs().flatten(1).max(1).values / rng).tolist())
    for mode, ch in (("parity", 1), ("parity", 2), ("parity", 4), ("parity_full", 2)):
        lib.bcosk_set_hp_chunk(ch)
        plan = ResNetPlan(arch, sd, batch, mode=mode, device="cuda")
        m = plan.explain(x6)["contribution_map"].double().cpu()
        e32 = ((m - ref).abs().flatten(1).max(1).values / rng).tolist()
        e64 = ((m - ref64).abs().flatten(1).max(1).values / rng).tolist()
        print(arch, mode, "chunk", ch, "vs fp32 ref:", ["%.2e" % v for v in e32], "| vs fp64:", ["%.2e" % v for v in e64])
