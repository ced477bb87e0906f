"""Host logic of the fine-tuning plan on CPU: operand index maps, gradient buckets, the all-reduce over gloo (world size 2)."""
import os
import subprocess
import sys

import pytest
import torch

from bcos_b200.engine import ResNetTrainPlan
from bcos_b200.engine import pack as P
from bcos_b200.models import resnet_state_shapes
from bcos_b200.utils import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_operand_index_maps_reproduce_the_packing():
    """every packed operand element points at the master weight the ordinary packing code would have put there"""
    arch = "resnet18"
    sd = synth.synth_state_dict(resnet_state_shapes(arch), 0)
    plan = ResNetTrainPlan(arch, sd, 2, device="cpu", image_size=64)
    flat = plan.w_flat
    lay = [l for l in plan.layers if l.name == "model.layer2.0.conv1"][0]           # strided 3x3
    buf, idx = plan._fwd_pack_of(lay)
    got = torch.where(idx >= 0, flat[idx.clamp(min=0).long()], torch.zeros(()))
    want, _ = P.pack_b(P.fwd_weight_taps(sd["model.layer2.0.conv1.linear.weight"]), 1, 64, torch.float32)
    assert torch.equal(got, want)
    stem = plan.stem
    _, idx0 = plan._fwd_pack_of(stem)
    got0 = torch.where(idx0 >= 0, flat[idx0.clamp(min=0).long()], torch.zeros(()))
    want0, _ = P.pack_b(P.fwd_weight_taps(P.stem_s2d_weight(sd["model.conv1.linear.weight"], 32)), 1, 32, torch.float32)
    assert torch.equal(got0, want0)
    # gradient slots: a bijection between master elements and slots of the flat gradient buffer
    gi = plan.gidx.long()
    assert gi.numel() == flat.numel() and gi.unique().numel() == gi.numel() and int(gi.max()) < plan.g_flat.numel()
    # buckets tile the gradient buffer in backward order
    assert plan.buckets[0][0] == 0 and plan.buckets[-1][1] == plan.g_flat.numel()
    assert all(a[1] == b[0] for a, b in zip(plan.buckets, plan.buckets[1:]))
    assert plan.fc.g_off == 0 and plan.stem.g_off > plan.blocks[0]["convs"][0].g_off


def test_bucketed_allreduce_gloo_world2():
    code = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from bcos_b200.engine import ResNetTrainPlan
from bcos_b200.models import resnet_state_shapes
from bcos_b200.utils import synth
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=2)
sd = synth.synth_state_dict(resnet_state_shapes("resnet18"), 0)
plan = ResNetTrainPlan("resnet18", sd, 2, device="cpu", image_size=64, world_size=2, bucket_mb=8.0)
assert len(plan.buckets) >= 3
plan.g_flat.fill_(float(rank + 1))
for a, b in plan.buckets:
    plan._allreduce(a, b)
assert bool((plan.g_flat == 3.0).all())
dist.destroy_process_group()
print("ALLREDUCE_OK", rank)
""" % ROOT
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    procs = [subprocess.Popen([sys.executable, "-c", code], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all("ALLREDUCE_OK" in o for o in outs), outs
