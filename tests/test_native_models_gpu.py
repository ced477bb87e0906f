"""SURVEY 8(f) row 4, second half: native B-cos-v2 variants of the reference's model zoo (PositionNorm / GN-LayerNorm / uncentred batch
norm, b != 2, MaxOut 2).  The reference's own MODEL FILES (bcos/models/resnet.py, densenet.py -- imported unmodified from the checkout
or from the archive oracle/stage_ref.py packs) are instantiated over OUR modules (bcos_b200.compat.install_as_bcos) and their logits /
contribution maps compared with what the reference computed with its own modules (tests/golden/native_*.npz, oracle/make_golden.py
--only native)."""
import os
import subprocess
import sys

import pytest

import native_variants as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCHIVE = os.path.join(ROOT, "oracle", "_ref", "bcos_reference.zip")
LIVE = "/root/reference"

CODE = r"""
import json, os, sys
import numpy as np
sys.dont_write_bytecode = True
ROOT, REF, NAME = %r, %r, %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from bcos_b200.compat import install_as_bcos
assert install_as_bcos(REF) == os.path.abspath(REF)
import bcos_b200.modules as M
import bcos_oracle as OR
import native_variants as V
from bcos_b200.utils import synth
import bcos.models.resnet as R
assert R.__file__.startswith(os.path.abspath(REF)) and R.BcosConv2d is M.BcosConv2d      # the reference's file over OUR modules
g = np.load(os.path.join(ROOT, "tests", "golden", "native_" + NAME + ".npz"))
m = V.build(NAME)
sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, int(g["seed"]))
off = 0
for k, n in zip(g["bn_keys"].tolist(), g["bn_sizes"].tolist()):
    sd[k] = torch.from_numpy(g["bn_var"][off:off + n].copy()); off += n
m.load_state_dict(sd, strict=True)
assert sum(v.numel() for v in m.state_dict().values()) == int(g["num_params"])
m = m.cuda().eval()
kinds = sorted({type(x).__name__ for x in m.modules() if type(x).__module__.startswith("bcos_b200")})
x6 = synth.to_bcos_input(g["images_u8"]).cuda()
logits, cmap = V.explain_batched(m, x6)
with torch.inference_mode():
    assert torch.equal(m(x6), logits)
logits, cmap = logits.float().cpu(), cmap.float().cpu()
res = {}
for tag, lk, ck in (("fp32", "logits", "contribution_map"), ("fp64", "logits_fp64", "contribution_map_fp64")):
    pm = OR.parity_metrics(logits, cmap, torch.from_numpy(g[lk]).float(), torch.from_numpy(g[ck]).float())
    res[tag] = {k: (float(v) if not isinstance(v, bool) else v) for k, v in pm.items()}
res["kinds"] = kinds
res["floor"] = float(g["fp32_noise_floor_maxabs_over_range"])
print("RESULT " + json.dumps(res))
"""


def _reference_root():
    if os.path.isdir(os.path.join(LIVE, "bcos", "models")):
        return LIVE
    if os.path.isfile(ARCHIVE):
        return ARCHIVE
    return None


@pytest.mark.gpu
@pytest.mark.parametrize("name", V.VARIANTS)
def test_reference_native_model_runs_on_our_modules(name):
    import json
    ref = _reference_root()
    if ref is None:
        pytest.skip("neither the reference checkout nor oracle/_ref/bcos_reference.zip is present")
    r = subprocess.run([sys.executable, "-c", CODE % (ROOT, ref, name)], capture_output=True, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-1500:] + r.stderr[-3000:]
    res = json.loads(lines[-1][7:])
    print(name, res)
    assert "BcosConv2d" in res["kinds"]
    # the contract's tolerances (BASELINE.json north_star), against the nearer of the reference's fp32 run and its fp64 evaluation
    best = min((res["fp32"], res["fp64"]), key=lambda pm: pm["map_maxabs_over_range"])
    assert res["fp32"]["argmax_equal"]
    assert min(res["fp32"]["logit_rel_err"], res["fp64"]["logit_rel_err"]) < 1e-3
    assert best["map_cos_min"] > 0.999
    assert best["map_maxabs_over_range"] < 1e-3
