"""Fused execution plan for B-cosified DenseNets (DenseNet-121: the network of BASELINE config 5 and of the reference's
`densenet_121` bcosification configs): forward + dynamic-linear explanation.

Network (reference): `BcosifyNetwork(DenseNetBcos(...))` bcosify.py:22-113 over torchvision's DenseNet skeleton with the classifier
applied per position before the global average (bcos/models/standard_models.py:56-63), max pool -> AvgPool2d(3, 2, 1), biases
removed (bcos/experiments/ImageNet/bcosification/model.py:47-55).  Layer order inside dense layers and transitions is
norm -> relu -> conv, so every consumer of a dense block's growing feature map applies its own uncentred BN + ReLU.

Layout: ONE feature tensor F [images, h, w, planes * C_total] per dense block.  Each dense layer is three launches:
    bcosk_dense_bn_relu_fwd   t = relu(F[:, :C_l] * alpha_l)  (+ sum t^2 for the 1x1 B-cos norm, ReLU bits for the explanation pass)
    bcosk_igemm (1x1)         u = relu(bn2(bcos(t)))           (BN multiplier + ReLU + gain in the epilogue, like the ResNet plan)
    bcosk_igemm (3x3)         F[:, C_l : C_l + growth] = bcos(u)   (written into its slice: the concatenation never copies)
Explanation pass, in reverse, over an fp32 feature-gradient tensor G per block:
    bcosk_dense_slice_cast    ghat2 = G[:, slice] * gain2  ->  3x3 data gradient (x gain1 in its epilogue)  ->  1x1 data gradient
    bcosk_dense_bn_relu_bwd   G[:, :C_l] += g * alpha_l * mask_l
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L
from . import ops as O
from . import pack as P
from .base import Act, ConvRec, PlanBase
from .resnet import IMAGENET_MEAN_ADDINVERSE, IMAGENET_STD_ADDINVERSE, resolve_precision

DENSENET_ARCH = {"densenet121": (32, (6, 12, 24, 16), 64), "densenet169": (32, (6, 12, 32, 32), 64), "densenet201": (32, (6, 12, 48, 32), 64)}


def dn_names(nblocks: int) -> Dict[str, str]:
    """State-dict prefixes: `BcosSequential.from_standard_module` (bcos/modules/common.py:46-51) rebuilds every nn.Sequential
    positionally, so `features` and the transitions lose their child names; dense blocks are ModuleDicts and keep `denselayerL`."""
    f = "model.features"
    names = {"conv0": f + ".0", "norm0": f + ".1", "norm5": f + f".{3 + 2 * nblocks}"}
    for bi in range(1, nblocks + 1):
        names[f"denseblock{bi}"] = f + f".{2 + 2 * bi}"
        names[f"transition{bi}.norm"] = f + f".{3 + 2 * bi}.0"
        names[f"transition{bi}.conv"] = f + f".{3 + 2 * bi}.2"
    return names


@dataclass
class DenseLayerRec:
    name: str
    c_in: int            # channels of F this layer reads
    col: int             # first channel of F it writes
    alpha1: Tensor
    mask1: Optional[Tensor]
    conv1: ConvRec
    conv2: ConvRec


@dataclass
class DenseBlockRec:
    name: str
    F: Tensor
    c_total: int
    c_in: int
    hw: Tuple[int, int]
    layers: List[DenseLayerRec]
    # transition behind the block (None for the last block)
    t_alpha: Optional[Tensor] = None
    t_mask: Optional[Tensor] = None
    t_conv: Optional[ConvRec] = None
    G: Optional[Tensor] = None


class DenseNetPlan(PlanBase):
    def __init__(self, arch: str, state_dict: Dict[str, Tensor], batch: int, *, mode: Optional[str] = None, planes: Optional[int] = None,
                 dtype: Optional[str] = None, device="cuda", image_size: int = 224, explain: bool = True, want_grad6: bool = False,
                 b: float = 2.0, bn_eps: float = 1e-5, mean=IMAGENET_MEAN_ADDINVERSE, std=IMAGENET_STD_ADDINVERSE,
                 logit_bias: Optional[float] = -math.log(1000 - 1), logit_temperature: Optional[float] = None,
                 seed_scale: Optional[float] = None, stem_kch: int = 32, input_u8: bool = False, explain_planes: Optional[int] = None):
        cfg = resolve_precision(mode, planes, dtype, explain_planes, seed_scale)
        cfg["explain_planes"] = 1          # the (linear) explanation pass of this plan always runs on one 16-bit plane
        self.precision = cfg
        super().__init__(batch, planes=cfg["planes"], dtype=cfg["dtype"], device=device, explain=explain, b=b, bn_eps=bn_eps,
                         state_dict=state_dict, explain_planes=cfg["explain_planes"])
        assert self.bplanes == 1, "the DenseNet plan runs its (linear) explanation pass on one 16-bit plane"
        self.arch = arch
        self.growth, self.block_cfg, self.init_c = DENSENET_ARCH[arch]
        self.mean, self.std = tuple(mean), tuple(std)
        self.inv_std = tuple(1.0 / s for s in std)
        self.logit_bias = 0.0 if logit_bias is None else float(logit_bias)
        self.inv_temp = 1.0 if logit_temperature is None else 1.0 / float(logit_temperature)
        self.seed_scale = float(cfg["seed_scale"])
        self.stem_kch, self.stem_cp = stem_kch, (32 if stem_kch == 32 else 64)
        self.size, self.input_u8 = image_size, input_u8
        self.blocks: List[DenseBlockRec] = []
        self._build_forward()
        if explain:
            self._build_explain(want_grad6)

    # ------------------------------------------------------------------ forward
    def _bn_relu(self, name: str, F: Tensor, c: int, bn: str, hw: Tuple[int, int]):
        """t = relu(bn(F[:, :c])) as its own dense plane tensor (+ sums of squares, + ReLU bits when an explanation follows)."""
        nb, pl = self.nb, self.planes
        alpha, beta = self._bn_alpha(bn)
        assert beta is None, "the factories strip every bias (bcosification/model.py:50-55)"
        t = self._empty(nb, hw[0], hw[1], pl * c)
        sq = self._empty(1, nb * hw[0] * hw[1], dtype=torch.float32)
        mask = self._zeros(nb * hw[0] * hw[1], c // 32, dtype=torch.int32) if self.with_explain else None
        self.fwd_ops.append(O.DenseBnReluFwdOp(name, F, c, pl, alpha, True, t, sq, mask, self.dt_code))
        return Act(t, c, sq, 1), alpha, mask

    def _build_forward(self) -> None:
        nb, S, pl, sd = self.nb, self.size, self.planes, self.sd
        nm = dn_names(len(self.block_cfg))
        self.x_in = self._empty(nb, 3, S, S, dtype=torch.uint8) if self.input_u8 else self._empty(nb, 6, S, S, dtype=torch.float32)
        # ---- stem: normalise + space-to-depth, 7x7/2 conv as a 4x4/1 conv, BN, ReLU (same launches as the ResNet plan)
        h2 = S // 2
        w4 = P.stem_s2d_weight(sd[nm["conv0"] + ".linear.weight"], self.stem_cp)
        if self.hp_accum and self.input_u8 and self.stem_im2col:
            y1, self.stem = self._stem_fwd_im2col("stem", self.x_in, sd[nm["conv0"] + ".linear.weight"], w4, 7, 2, 3, bn=nm["norm0"],
                                                  mean6=self.mean, inv_std6=self.inv_std, s2d_pad=(2, 1), stem_cp=self.stem_cp)
        else:
            a0 = self._empty(nb, h2, h2, pl * self.stem_cp)
            sq0 = self._empty(1, nb * S * S, dtype=torch.float32)
            self.fwd_ops.append(O.InputPrepOp("input_prep", self.x_in, self.mean, self.inv_std, a0, self.stem_cp, pl, self.dt_code, sq0))
            y1, self.stem = self._conv_fwd("stem", Act(a0, self.stem_cp, sq0, 1), w4, 1, 2, 1, bn=nm["norm0"], relu=True, kch=self.stem_kch,
                                           want_sq=False, sq_geom=(S, S, 7, 2, 3))
        hp = (h2 + 2 - 3) // 2 + 1
        pooled = self._empty(nb, hp, hp, pl * self.init_c)
        self.fwd_ops.append(O.AvgPoolFwdOp("pool", y1.t, self.init_c, pl, 3, 2, 1, pooled, self.dt_code, None))
        self.pool_hw = (hp, hp)
        c, hw = self.init_c, (hp, hp)
        for bi, nlayers in enumerate(self.block_cfg, start=1):
            c_total = c + self.growth * nlayers
            F = self._zeros(nb, hw[0], hw[1], pl * c_total)
            self.fwd_ops.append(O.CopyChannelsOp(f"denseblock{bi}.input", pooled, c, pl, F, 0))
            blk = DenseBlockRec(f"denseblock{bi}", F, c_total, c, hw, [])
            cl = c
            for li in range(1, nlayers + 1):
                p = nm[f"denseblock{bi}"] + f".denselayer{li}"
                t, alpha1, mask1 = self._bn_relu(p + ".norm1", F, cl, p + ".norm1", hw)
                u, r1 = self._conv_fwd(p + ".conv1", t, sd[p + ".conv1.linear.weight"], 1, 0, 0, bn=p + ".norm2", relu=True)
                _, r2 = self._conv_fwd(p + ".conv2", u, sd[p + ".conv2.linear.weight"], 1, 1, 1, bn=None, relu=False, want_sq=False,
                                       y_buf=F, y_col=cl)
                blk.layers.append(DenseLayerRec(p, cl, cl, alpha1, mask1, r1, r2))
                cl += self.growth
            self.blocks.append(blk)
            c = c_total
            if bi != len(self.block_cfg):
                t, blk.t_alpha, blk.t_mask = self._bn_relu(nm[f"transition{bi}.norm"], F, c, nm[f"transition{bi}.norm"], hw)
                y, blk.t_conv = self._conv_fwd(nm[f"transition{bi}.conv"], t, sd[nm[f"transition{bi}.conv"] + ".linear.weight"], 1, 0, 0,
                                               bn=None, relu=False, want_sq=False)
                c //= 2
                hw = (hw[0] // 2, hw[1] // 2)
                pooled = self._empty(nb, hw[0], hw[1], pl * c)
                self.fwd_ops.append(O.AvgPoolFwdOp(f"transition{bi}.pool", y.t, c, pl, 2, 2, 0, pooled, self.dt_code, None))
        last = self.blocks[-1]
        t5, self.alpha5, self.mask5 = self._bn_relu(nm["norm5"], last.F, c, nm["norm5"], hw)
        wfc = sd["model.classifier.linear.weight"]
        self.ncls = wfc.shape[0]
        fc, self.fc = self._conv_fwd("classifier", t5, wfc, 1, 0, 0, bn=None, relu=False, y_f32=True, want_sq=False)
        self.npix = hw[0] * hw[1]
        self.c_last = c
        self.fc_out = fc.t.view(nb * self.npix, self.ncls)
        self.logits = self._empty(nb, self.ncls, dtype=torch.float32)
        self.pred = self._zeros(nb, dtype=torch.int32)
        self.fwd_ops.append(O.GapLogitsOp("gap_logits", self.fc_out, nb, self.npix, self.ncls, self.inv_temp, self.logit_bias, self.logits,
                                          self.pred))

    # ------------------------------------------------------------------ explanation pass
    def _build_explain(self, want_grad6: bool) -> None:
        nb = self.nb
        f32 = torch.float32
        for blk in self.blocks:
            blk.G = self._zeros(nb, blk.hw[0], blk.hw[1], blk.c_total, dtype=f32)
            for ly in blk.layers:
                self._alloc_ghat(ly.conv1)
                self._alloc_ghat(ly.conv2)
            if blk.t_conv is not None:
                self._alloc_ghat(blk.t_conv)
        self._alloc_ghat(self.stem)
        last = self.blocks[-1]
        # ---- seed: one-hot logit gradient through GAP and the classifier's detached scale -> gradient wrt relu(norm5(F))
        self.w_fc32 = self._dev(self.sd["model.classifier.linear.weight"].reshape(self.ncls, self.c_last))
        g_t5 = self._zeros(nb, last.hw[0], last.hw[1], self.c_last)
        self.bwd_ops.append(O.FcSeedOp("classifier.seed", self.pred, self.fc.gain, self.w_fc32, nb, self.npix, self.ncls, self.c_last,
                                       self.inv_temp, self.seed_scale, None, g_t5.view(nb * self.npix, self.c_last), None, None, 1, self.dt_code))
        self.bwd_ops.append(O.DenseBnReluBwdOp("norm5.bwd", g_t5, self.c_last, self.alpha5, self.mask5, last.G, False, self.dt_code))
        for bi in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[bi]
            h, w = blk.hw
            scratch = self._zeros(nb * h * w * blk.c_total, dtype=f32)        # 1x1 data gradients of this block (one at a time)
            for ly in reversed(blk.layers):
                self.bwd_ops.append(O.DenseSliceCastOp(ly.name + ".conv2.ghat", blk.G, ly.col, self.growth, ly.conv2.gain, 1.0, ly.conv2.ghat,
                                                       self.dt_code))
                m1, m1s = self._gain_of(ly.conv1)
                self._dgrad(ly.conv2, y=ly.conv1.ghat, mul1=m1, mul1_sqrt_scale=m1s)
                g_t = scratch[:nb * h * w * ly.c_in].view(nb, h, w, ly.c_in)
                self._dgrad(ly.conv1, y=g_t, y_f32=True)
                self.bwd_ops.append(O.DenseBnReluBwdOp(ly.name + ".norm1.bwd", g_t, ly.c_in, ly.alpha1, ly.mask1, blk.G, True, self.dt_code))
            # ---- gradient wrt the block input (its first c_in channels) -> through the pool in front
            gin = self._zeros(nb, h, w, blk.c_in)
            if bi > 0:
                prev = self.blocks[bi - 1]
                self.bwd_ops.append(O.DenseSliceCastOp(blk.name + ".input.grad", blk.G, 0, blk.c_in, None, 1.0, gin, self.dt_code))
                self.bwd_ops.append(O.AvgPoolBwdMulOp(f"transition{bi}.pool.bwd", gin, blk.c_in, 1, 2, 2, 0, prev.t_conv.gain, prev.t_conv.ghat,
                                                      self.dt_code))
                ph, pw = prev.hw
                g_tt = self._zeros(nb, ph, pw, prev.c_total, dtype=f32)
                self._dgrad(prev.t_conv, y=g_tt, y_f32=True)
                self.bwd_ops.append(O.DenseBnReluBwdOp(f"transition{bi}.norm.bwd", g_tt, prev.c_total, prev.t_alpha, prev.t_mask, prev.G, False,
                                                       self.dt_code))
            else:
                self.bwd_ops.append(O.DenseSliceCastOp(blk.name + ".input.grad", blk.G, 0, blk.c_in, None, 1.0, gin, self.dt_code))
                m1, m1s = self._gain_of(self.stem)
                self.bwd_ops.append(O.AvgPoolBwdMulOp("pool.bwd", gin, self.init_c, 1, 3, 2, 1, m1, self.stem.ghat, self.dt_code, m1s))
        h2 = self.size // 2
        self.g0 = self._zeros(nb, h2, h2, self.stem_cp, dtype=f32)
        self._dgrad(self.stem, y=self.g0, y_f32=True, kch=64)
        self.cmap = self._zeros(nb, self.size, self.size, dtype=f32)
        self.grad6 = self._zeros(nb, 6, self.size, self.size, dtype=f32) if want_grad6 else None
        self.bwd_ops.append(O.ContribMapOp("contrib_map", self.g0, self.x_in, self.stem_cp, self.inv_std, 1.0 / self.seed_scale, self.cmap,
                                           self.grad6))

    # ------------------------------------------------------------------ public API (same as ResNetPlan)
    def load_input(self, x6: Tensor) -> None:
        assert tuple(x6.shape) == tuple(self.x_in.shape), (x6.shape, self.x_in.shape)
        self.x_in.copy_(x6, non_blocking=True)

    def forward(self, x6: Optional[Tensor] = None) -> Tensor:
        if x6 is not None:
            self.load_input(x6)
        self.replay_forward()
        return self.logits

    def explain(self, x6: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """Forward + explanation of each image's predicted class (argmax logit)."""
        if not self.with_explain:
            raise RuntimeError("plan was built with explain=False")
        if x6 is not None:
            self.load_input(x6)
        self.replay_all()
        out = {"logits": self.logits, "prediction": self.pred, "contribution_map": self.cmap}
        if self.grad6 is not None:
            out["dynamic_linear_weights"] = self.grad6
        return out
