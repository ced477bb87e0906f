"""Fused execution plan for the B-cos CLIP ResNet image encoder (BASELINE config 4): embedding + explanation.

Network (reference): CLIP/clip/model.py:94-154 `ModifiedResNet` (three-conv stem, anti-aliasing average pools,
CLIP/clip/model.py:10-55 `Bottleneck`: the stride lives in an AvgPool2d AFTER conv2 and in front of the downsample conv)
converted by bcosify.py:74-113 with `clip_kd` (CLIP mean / std, no LogitLayer), biases and positional embedding removed
(bcos/experiments/ImageNet/clip_bcosification/model.py:15-23), `BcosAttentionPool2d` head (bcos/modules/bcosattnpool.py:34-59).

The trunk runs like engine/resnet.py: every conv + BN (+ residual) + ReLU group is ONE `bcosk_igemm` launch, the explanation
pass is the chain of explain-dgrad launches whose epilogues multiply by the producer's saved gain; the average pools are
`bcosk_avgpool_fwd` / `bcosk_avgpool_bwd_mul` (the backward fused with the gain of the conv in front of the pool).  The 3x3/2
stem conv runs as a 2x2/1 conv over the 2x2 space-to-depth input (engine/pack.py stem_s2d_weight), so the input-prep and
contribution-map kernels of the ResNet plan serve here too.

The attention-pool head is part of the plan (`fused_head=True`, csrc/bcosk_head.cu): only the mean token's output is kept
(bcosattnpool.py:52-58), so the projections commute with the pooling - scores s_j = (W_k,h^T q_h) . x_j, output
o_h = W_v,h (sum_j p_j x_j): three token-equivalents of 2048 x 2048 projections per image instead of 150, done in fp32 by a
strided-batched SGEMM; the explanation backward (q, k frozen: p constant) is three more SGEMM calls and `bcosk_seed_from_tokens`
turns the token gradient into the last block's gradient tensors.  The only torch code left is the user's target function on the
[batch, 1024] embedding.  `fused_head=False` keeps the module-level head (modules/bcosattnpool.py, q|k|v of all 50 tokens through
the tcgen05 linear kernel + autograd) for comparison.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L
from . import ops as O
from . import pack as P
from .base import Act, ConvRec, PlanBase
from .resnet import resolve_precision

CLIP_MEAN_ADDINVERSE = (0.48145466, 0.4578275, 0.40821073, 0.51854534, 0.5421725, 0.59178927)  # reference bcosify.py:17
CLIP_STD_ADDINVERSE = (0.26862954, 0.26130258, 0.27577711, 0.26862954, 0.26130258, 0.27577711)  # reference bcosify.py:19


@dataclass
class ClipBlock:
    name: str
    convs: List[ConvRec]
    ds: Optional[ConvRec]
    x: Act
    y: Act
    mask: Tensor
    stride: int
    side: Optional[Tensor] = None   # gradient entering the identity / downsample branch


class CLIPResNetPlan(PlanBase):
    def __init__(self, state_dict: Dict[str, Tensor], batch: int, *, mode: Optional[str] = None, planes: Optional[int] = None,
                 dtype: Optional[str] = None, device="cuda", image_size: int = 224, layers=(3, 4, 6, 3), width: int = 64,
                 heads: int = 32, explain: bool = True, b: float = 2.0, bn_eps: float = 1e-5, mean=CLIP_MEAN_ADDINVERSE,
                 std=CLIP_STD_ADDINVERSE, seed_scale: Optional[float] = None, input_u8: bool = False,
                 explain_planes: Optional[int] = None, want_grad6: bool = False, fused_head: bool = True):
        cfg = resolve_precision(mode, planes, dtype, explain_planes, seed_scale)
        self.precision = cfg
        super().__init__(batch, planes=cfg["planes"], dtype=cfg["dtype"], device=device, explain=explain, b=b, bn_eps=bn_eps,
                         state_dict=state_dict, explain_planes=cfg["explain_planes"])
        self.layers, self.width, self.heads = tuple(layers), width, heads
        self.mean, self.std = tuple(mean), tuple(std)
        self.inv_std = tuple(1.0 / s for s in std)
        self.seed_scale = float(cfg["seed_scale"])
        self.size, self.input_u8 = image_size, input_u8
        self.stem_cp = 32
        self.fused_head = fused_head
        self.blocks: List[ClipBlock] = []
        self._head = None
        self._build_forward()
        if explain:
            self._build_explain(want_grad6)

    # ------------------------------------------------------------------ forward
    def _pool(self, name: str, x: Act, s: int) -> Act:
        nb, pl = self.nb, self.planes
        h, w = x.hw
        y = self._empty(nb, h // s, w // s, pl * x.c)
        sq = self._empty(1, nb * (h // s) * (w // s), dtype=torch.float32)
        self.fwd_ops.append(O.AvgPoolFwdOp(name, x.t, x.c, pl, s, s, 0, y, self.dt_code, sq))
        return Act(y, x.c, sq, 1)

    def _build_forward(self) -> None:
        nb, S, pl, sd = self.nb, self.size, self.planes, self.sd
        assert S % 32 == 0, "CLIP ResNets reduce the resolution 32x"
        self.x_in = self._empty(nb, 3, S, S, dtype=torch.uint8) if self.input_u8 else self._empty(nb, 6, S, S, dtype=torch.float32)
        h2 = S // 2
        a0 = self._empty(nb, h2, h2, pl * self.stem_cp)
        sq0 = self._empty(1, nb * S * S, dtype=torch.float32)
        self.fwd_ops.append(O.InputPrepOp("input_prep", self.x_in, self.mean, self.inv_std, a0, self.stem_cp, pl, self.dt_code, sq0))
        # ---- stem: 3x3/2 (as 2x2/1 over the space-to-depth input; patch norm = the ORIGINAL 3x3/2 pad-1 window), 3x3, 3x3, avg pool 2
        w2 = P.stem_s2d_weight(sd["model.conv1.linear.weight"], self.stem_cp)
        y1, s1 = self._conv_fwd("stem.conv1", Act(a0, self.stem_cp, sq0, 1), w2, 1, 1, 0, bn="model.bn1", relu=True, kch=32,
                                sq_geom=(S, S, 3, 2, 1))
        y2, s2 = self._conv_fwd("stem.conv2", y1, sd["model.conv2.linear.weight"], 1, 1, 1, bn="model.bn2", relu=True, kch=32)
        y3, s3 = self._conv_fwd("stem.conv3", y2, sd["model.conv3.linear.weight"], 1, 1, 1, bn="model.bn3", relu=True, kch=32,
                                want_sq=False)
        self.stem = [s1, s2, s3]
        self.stem_out = y3
        x = self.pool_out = self._pool("stem.pool", y3, 2)
        # ---- residual stages
        inpl = self.width
        for li, (planes_, nblocks) in enumerate(zip([self.width, self.width * 2, self.width * 4, self.width * 8], self.layers), start=1):
            for bi in range(nblocks):
                stride = 2 if (li > 1 and bi == 0) else 1
                pfx = f"model.layer{li}.{bi}"
                has_ds = (pfx + ".downsample.1.linear.weight") in sd
                t1, r1 = self._conv_fwd(pfx + ".conv1", x, sd[pfx + ".conv1.linear.weight"], 1, 0, 0, bn=pfx + ".bn1", relu=True)
                t2, r2 = self._conv_fwd(pfx + ".conv2", t1, sd[pfx + ".conv2.linear.weight"], 1, 1, 1, bn=pfx + ".bn2", relu=True,
                                        want_sq=stride == 1)
                if stride > 1:
                    t2 = self._pool(pfx + ".avgpool", t2, stride)
                idn, ds_rec = x, None
                if has_ds:
                    xp = self._pool(pfx + ".downsample.pool", x, stride) if stride > 1 else x
                    idn, ds_rec = self._conv_fwd(pfx + ".downsample", xp, sd[pfx + ".downsample.1.linear.weight"], 1, 0, 0,
                                                 bn=pfx + ".downsample.2", relu=False, want_sq=False)
                y, r3 = self._conv_fwd(pfx + ".conv3", t2, sd[pfx + ".conv3.linear.weight"], 1, 0, 0, bn=pfx + ".bn3", relu=True,
                                       res=idn, want_mask=True)
                self.blocks.append(ClipBlock(pfx, [r1, r2, r3], ds_rec, x, y, r3.mask, stride))
                x = y
                inpl = planes_ * 4
        self.trunk_out = x
        self.c_out = x.c
        if self.fused_head:
            self._build_head_forward()
        else:
            self.feat = self._empty(nb, x.c, x.hw[0], x.hw[1], dtype=torch.float32)      # NCHW fp32 for the module-level head
            self.fwd_ops.append(O.TrunkOutOp("trunk_out", x.t, nb, x.c, x.hw[0], x.hw[1], pl, self.dt_code, self.feat))

    def _build_head_forward(self) -> None:
        """bcosattnpool.py:34-59 with the projections commuted past the pooling (module docstring); fp32 SGEMM launches."""
        nb, c, H = self.nb, self.c_out, self.heads
        dh = c // H
        T = self.trunk_out.hw[0] * self.trunk_out.hw[1] + 1
        f32 = torch.float32
        sd = self.sd
        self.wq, self.wk, self.wv = (self._dev(sd[f"model.attnpool.{n}.weight"]) for n in ("q_proj", "k_proj", "v_proj"))
        self.wc = self._dev(sd["model.attnpool.c_proj.linear.weight"])
        od = self.wc.shape[0]
        self.tokens = self._empty(nb, T, c, dtype=f32)
        self.h_q = self._empty(nb, c, dtype=f32)
        self.h_qt = self._empty(nb, H, c, dtype=f32)
        self.h_p = self._empty(nb, H, T, dtype=f32)
        self.h_xbar = self._empty(nb, H, c, dtype=f32)
        self.h_o = self._empty(nb, c, dtype=f32)
        self.emb = self._empty(nb, od, dtype=f32)
        ops = self.fwd_ops
        ops.append(O.HeadTokensOp("attnpool.tokens", self.trunk_out.t, c, self.planes, self.dt_code, self.tokens))
        # q = W_q x_0 / sqrt(dh)  (x_0 = mean token)
        ops.append(O.SgemmOp("attnpool.q", False, True, nb, c, c, self.tokens, 0, T * c, 0, self.wq, 0, c, 0, self.h_q, 0, c, 0, 1, dh ** -0.5))
        # qt_h = W_k,h^T q_h
        ops.append(O.SgemmOp("attnpool.qk", False, False, nb, c, dh, self.h_q, 0, c, dh, self.wk, 0, c, dh * c, self.h_qt, 0, H * c, c, H))
        # s[b, h, j] = qt[b, h] . x[b, j];  p = softmax_j
        ops.append(O.SgemmOp("attnpool.scores", False, True, H, T, c, self.h_qt, 0, c, H * c, self.tokens, 0, c, T * c, self.h_p, 0, T, H * T, nb))
        ops.append(O.RowSoftmaxOp("attnpool.softmax", self.h_p))
        # xbar[b, h] = sum_j p[b, h, j] x[b, j];  o_h = W_v,h xbar_h;  emb = W_c o
        ops.append(O.SgemmOp("attnpool.pool", False, False, H, c, T, self.h_p, 0, T, H * T, self.tokens, 0, c, T * c, self.h_xbar, 0, c, H * c, nb))
        ops.append(O.SgemmOp("attnpool.v", False, True, nb, dh, c, self.h_xbar, 0, H * c, c, self.wv, 0, c, dh * c, self.h_o, 0, c, dh, H))
        ops.append(O.SgemmOp("attnpool.c_proj", False, True, nb, od, c, self.h_o, 0, c, 0, self.wc, 0, c, 0, self.emb, 0, od, 0, 1))

    # ------------------------------------------------------------------ explanation pass
    def _build_explain(self, want_grad6: bool) -> None:
        nb, pl = self.nb, self.bplanes
        for blk in self.blocks:
            for r in blk.convs:
                self._alloc_ghat(r)
            if blk.ds is not None:
                self._alloc_ghat(blk.ds)
                blk.side = blk.ds.ghat
            else:
                blk.side = self._zeros(nb, blk.y.hw[0], blk.y.hw[1], pl * blk.y.c)
        for r in self.stem:
            self._alloc_ghat(r)
        last = self.blocks[-1]
        # (the seed scale that keeps fp16 gradients in range is applied to the target itself, in front of the head's backward)
        if self.fused_head:
            # ---- head backward in explanation mode (q, k frozen: p is a constant): g_o = W_c^T g_emb, g_xbar_h = W_v,h^T g_o,h,
            #      g_x[b, j] = sum_h p[b, h, j] g_xbar[b, h]; the mean token's share is spread over the pixels by the seed kernel
            c, H = self.c_out, self.heads
            dh, T, od = c // H, self.tokens.shape[1], self.wc.shape[0]
            f32 = torch.float32
            self.g_emb = self._zeros(nb, od, dtype=f32)
            self.g_o = self._empty(nb, c, dtype=f32)
            self.g_xbar = self._empty(nb, H, c, dtype=f32)
            self.g_tokens = self._empty(nb, T, c, dtype=f32)
            ops = self.bwd_ops
            ops.append(O.SgemmOp("attnpool.c_proj.bwd", False, False, nb, c, od, self.g_emb, 0, od, 0, self.wc, 0, c, 0, self.g_o, 0, c, 0, 1))
            ops.append(O.SgemmOp("attnpool.v.bwd", False, False, nb, c, dh, self.g_o, 0, c, dh, self.wv, 0, c, dh * c, self.g_xbar, 0, H * c, c, H))
            ops.append(O.SgemmOp("attnpool.pool.bwd", True, False, T, c, H, self.h_p, 0, T, H * T, self.g_xbar, 0, c, H * c, self.g_tokens, 0, c,
                                 T * c, nb))
            ops.append(O.SeedFromTokensOp("head.seed", self.g_tokens, 1.0, last.convs[-1].gain, last.convs[-1].ghat, last.mask,
                                          None if last.ds is None else last.ds.gain, last.side, pl, self.dt_code))
        else:
            # ---- seed: d target / d trunk output (NCHW fp32, computed by the module-level head through autograd)
            self.g_feat = self._zeros(nb, self.c_out, last.y.hw[0], last.y.hw[1], dtype=torch.float32)
            self.bwd_ops.append(O.SeedFromNchwOp("head.seed", self.g_feat, 1.0, last.convs[-1].gain, last.convs[-1].ghat,
                                               last.mask, None if last.ds is None else last.ds.gain, last.side, pl, self.dt_code))
        self.g_pool = self._zeros(nb, self.pool_out.hw[0], self.pool_out.hw[1], pl * self.pool_out.c)
        for bi in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[bi]
            c1, c2, c3 = blk.convs
            s = blk.stride
            if s > 1:
                # conv3's data gradient lives at the pooled resolution; the pool's backward spreads it and applies conv2's gain
                gp = self._zeros(nb, c3.in_hw[0], c3.in_hw[1], pl * c3.cin_phys)
                self._dgrad(c3, y=gp)
                self.bwd_ops.append(O.AvgPoolBwdMulOp(blk.name + ".avgpool.bwd", gp, c2.cout, pl, s, s, 0, c2.gain, c2.ghat, self.dt_code))
            else:
                m1, m1s = self._gain_of(c2)
                self._dgrad(c3, y=c2.ghat, mul1=m1, mul1_sqrt_scale=m1s)
            m1, m1s = self._gain_of(c1)
            self._dgrad(c2, y=c1.ghat, mul1=m1, mul1_sqrt_scale=m1s)
            add = blk.side
            if blk.ds is not None:
                dds = self._zeros(nb, blk.x.hw[0], blk.x.hw[1], pl * blk.x.c)
                if s > 1:
                    ddp = self._zeros(nb, blk.ds.in_hw[0], blk.ds.in_hw[1], pl * blk.ds.cin_phys)
                    self._dgrad(blk.ds, y=ddp)
                    self.bwd_ops.append(O.AvgPoolBwdMulOp(blk.name + ".downsample.pool.bwd", ddp, blk.x.c, pl, s, s, 0, None, dds, self.dt_code))
                else:
                    self._dgrad(blk.ds, y=dds)
                add = dds
            if bi > 0:
                prev = self.blocks[bi - 1]
                self._dgrad(c1, y=prev.convs[-1].ghat, mul1=prev.convs[-1].gain, add=add, add_stride=1, out2=prev.side,
                            mul2=None if prev.ds is None else prev.ds.gain, mask2=prev.mask)
            else:
                self._dgrad(c1, y=self.g_pool, add=add, add_stride=1)
        # ---- stem: pool backward x conv3 gain, three data gradients, contribution map
        s1, s2, s3 = self.stem
        self.bwd_ops.append(O.AvgPoolBwdMulOp("stem.pool.bwd", self.g_pool, s3.cout, pl, 2, 2, 0, s3.gain, s3.ghat, self.dt_code))
        self._dgrad(s3, y=s2.ghat, mul1=s2.gain, kch=64)
        self._dgrad(s2, y=s1.ghat, mul1=s1.gain, kch=32)
        h2 = self.size // 2
        self.g0 = self._zeros(nb, h2, h2, self.stem_cp, dtype=torch.float32)
        self._dgrad(s1, y=self.g0, y_f32=True, kch=32)
        self.cmap = self._zeros(nb, self.size, self.size, dtype=torch.float32)
        self.grad6 = self._zeros(nb, 6, self.size, self.size, dtype=torch.float32) if want_grad6 else None
        self.bwd_ops.append(O.ContribMapOp("contrib_map", self.g0, self.x_in, self.stem_cp, self.inv_std, 1.0 / self.seed_scale,
                                           self.cmap, self.grad6))

    # ------------------------------------------------------------------ head (module-level path) and public API
    def head(self):
        """`BcosAttentionPool2d` over the plan's state dict (lazily built; module-level libbcosk kernels)."""
        if self._head is None:
            from ..modules.bcosattnpool import BcosAttentionPool2d
            e = self.c_out
            out_dim = self.sd["model.attnpool.c_proj.linear.weight"].shape[0]
            m = BcosAttentionPool2d(self.trunk_out.hw[0], e, self.heads, out_dim)
            m.positional_embedding = None
            for nme in ("q_proj", "k_proj", "v_proj"):
                lin = getattr(m, nme)
                lin.weight.data = self.sd[f"model.attnpool.{nme}.weight"].clone()
                lin.bias = None
            m.c_proj.weight.data = self.sd["model.attnpool.c_proj.linear.weight"].clone()
            m.c_proj.bias = None
            self._head = m.to(self.device).eval()
        return self._head

    def load_input(self, x6: Tensor) -> None:
        assert tuple(x6.shape) == tuple(self.x_in.shape), (x6.shape, self.x_in.shape)
        self.x_in.copy_(x6, non_blocking=True)

    def embed(self, x6: Optional[Tensor] = None) -> Tensor:
        """Image embeddings [nb, output_dim] (forward only)."""
        if x6 is not None:
            self.load_input(x6)
        self.replay_forward()
        if self.fused_head:
            return self.emb
        from ..modules import _runtime as R
        with torch.no_grad(), R.precision(self.planes, self.precision["dtype"]):
            return self.head()(self.feat)

    def explain_target(self, x6: Optional[Tensor], target_fn) -> Dict[str, Tensor]:
        """Forward + explanation of `target_fn(embedding).sum()` (e.g. the cosine with a text embedding,
        interpretability/analyses/text_localisation.py:77-100).  The head runs in explanation mode (query / keys frozen)."""
        if not self.with_explain:
            raise RuntimeError("plan was built with explain=False")
        if x6 is not None:
            self.load_input(x6)
        self.replay_forward()
        if self.fused_head:
            emb = self.emb.detach().clone().requires_grad_(True)
            with torch.enable_grad():
                (g,) = torch.autograd.grad(target_fn(emb).sum() * self.seed_scale, [emb])     # the user's target on [batch, dim]
            self.g_emb.copy_(g)
            self.replay_explain()
            out = {"embedding": emb.detach(), "contribution_map": self.cmap}
            if self.grad6 is not None:
                out["dynamic_linear_weights"] = self.grad6
            return out
        head = self.head()
        feat = self.feat.detach().requires_grad_(True)
        head.set_explanation_mode(True)
        from ..modules import _runtime as R
        try:
            # the head's launches use the plan's operand format; the target is scaled like the one-hot seed of the ResNet plan
            # (keeps 16-bit gradients in range; the contribution-map kernel divides it out again)
            with torch.enable_grad(), R.precision(self.planes, self.precision["dtype"]):
                emb = head(feat)
                (g,) = torch.autograd.grad(target_fn(emb).sum() * self.seed_scale, [feat])
        finally:
            head.set_explanation_mode(False)
        self.g_feat.copy_(g)
        self.replay_explain()
        out = {"embedding": emb.detach(), "contribution_map": self.cmap}
        if self.grad6 is not None:
            out["dynamic_linear_weights"] = self.grad6
        return out

    def explain_direction(self, x6: Optional[Tensor], direction: Tensor) -> Dict[str, Tensor]:
        """Explanation of cos(embedding, direction) for a fixed unit vector (a text embedding)."""
        d = direction.to(self.device, torch.float32)
        return self.explain_target(x6, lambda emb: torch.nn.functional.cosine_similarity(emb, d[None], dim=1))
