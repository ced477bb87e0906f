"""Fused execution plan for the B-cosified SimpleViT (BASELINE config 3: ViT-Ti/16, ViT-B/16): forward + explanation.

Network (reference): bcos/models/vit.py:64-339 `SimpleViT` as converted by bcosify_vit.py:45-153 - patch-embedding weights
doubled for the 6-channel input, every nn.Linear except `to_qkv` -> BcosifyLinear, GELU -> MyGELU, LayerNorm ->
DetachableLayerNorm, biases removed, `gap_reorder` (classifier per token, then the mean); explanation mode freezes the
attention probabilities (vit.py:148-150), the GELU gate and the LayerNorm variance.

Token tensors live in HBM as [images, 14, 14, planes * d] 16-bit precision planes, so every linear layer is ONE 1x1
`bcosk_igemm` launch (residual add, B-cos scale, saved gain in its epilogue) and the only other launches are the bandwidth
kernels of csrc/bcosk_vit.cu (patchify, LayerNorm, GELU, attention, LayerNorm backward + residual-stream add, contribution map).
No layout bridges, no fp32 NCHW round trips (the module-level path of vit.py pays two per module).

Explanation pass: the residual-stream gradient G is carried in fp32; per encoder, in reverse,
    linear2^T (x gate x gain1 in the epilogue) -> linear1^T -> LN backward (+G, x gain_out) -> to_out^T -> P^T g (q, k frozen)
    -> W_v^T (the q / k thirds of to_qkv get no gradient) -> LN backward (+G, x gain of the linear in front).
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

VIT_ARCH = {  # name: (dim, depth, heads, mlp_dim)   reference bcos/models/vit.py:441-467
    "simple_vit_ti_patch16_224": (192, 12, 3, 768),
    "simple_vit_s_patch16_224": (384, 12, 6, 1536),
    "simple_vit_b_patch16_224": (768, 12, 12, 3072),
}


def posemb_sincos_2d(h: int, w: int, dim: int, temperature: float = 10000.0) -> Tensor:
    """reference bcos/models/vit.py:64-86"""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    omega = torch.arange(dim // 4) / (dim // 4 - 1)
    omega = 1.0 / (temperature ** omega)
    y = y.flatten()[:, None] * omega[None, :]
    x = x.flatten()[:, None] * omega[None, :]
    return torch.cat((x.sin(), x.cos(), y.sin(), y.cos()), dim=1)


@dataclass
class EncoderRec:
    name: str
    qkv: Tensor
    rstd1: Tensor
    rstd2: Tensor
    w_ln1: Tensor
    w_ln2: Tensor
    out: ConvRec
    lin1: ConvRec
    lin2: ConvRec
    w_v: Tensor          # [d, d, 1, 1] fp32: the v third of to_qkv


class ViTPlan(PlanBase):
    fuse_gelu = True     # one-plane branches: MyGELU inside linear1's epilogue (include/bcosk.h `act`); False keeps the separate pass

    def __init__(self, arch: str, state_dict: Dict[str, Tensor], batch: int, *, mode: Optional[str] = None, planes: Optional[int] = None,
                 dtype: Optional[str] = None, device="cuda", image_size: int = 224, patch: int = 16, explain: bool = True,
                 want_grad6: bool = False, b: float = 2.0, ln_eps: float = 1e-5, mean=IMAGENET_MEAN_ADDINVERSE,
                 std=IMAGENET_STD_ADDINVERSE, logit_bias: Optional[float] = -math.log(1000 - 1), logit_temperature: Optional[float] = None,
                 seed_scale: Optional[float] = None, input_u8: bool = False, explain_planes: Optional[int] = None,
                 branch_planes: Optional[int] = None):
        cfg = resolve_precision(mode, planes, dtype, explain_planes, seed_scale)
        cfg["explain_planes"] = 1          # the (linear) explanation pass of this plan always runs on one 16-bit plane
        self.precision = cfg
        super().__init__(batch, planes=cfg["planes"], dtype=cfg["dtype"], device=device, explain=explain, b=b, state_dict=state_dict,
                         explain_planes=cfg["explain_planes"])
        assert self.bplanes == 1, "the ViT plan runs its (linear) explanation pass on one 16-bit plane"
        # Mixed operand format (branch_planes = 1 with planes = 2): the residual stream x, the patch embedding and every residual add
        # keep `planes` precision planes; the branch tensors (LayerNorm outputs, q | k | v, attention output, MLP hidden) and the
        # weights that produce them carry ONE plane, the weights of the launches that add into the stream two (a0 b0 + a0 b1).
        # A ViT has no ReLU decisions to flip: what has to stay exact is the stream that 24 branch outputs are added to.
        # Measured against the reference goldens (scripts/exp_vit_mixed.py): ViT-Ti map max-abs 5.0e-4 of the range (all planes: 3.4e-4),
        # ViT-B 4.5e-4 (2.7e-4), logits 1.5e-4 / 3e-5 - inside the contract with a factor two to spare, at 47 % of the tensor work:
        # the default of the "parity" mode.  `mode="parity_full"` (or branch_planes=planes) keeps every operand at `planes` planes.
        self.sp = self.planes
        if branch_planes is None:
            branch_planes = 1 if (mode in (None, "parity") and planes is None and dtype is None) else self.planes
        self.bp = int(branch_planes)
        assert 1 <= self.bp <= self.sp
        self.arch = arch
        self.dim, self.depth, self.heads, self.mlp = VIT_ARCH[arch]
        self.dh = self.dim // self.heads
        assert self.dh == 64, "bcosk_vit_attention is built for dim_head = 64 (all reference SimpleViT sizes)"
        self.patch, self.size, self.ln_eps = patch, image_size, ln_eps
        self.gh = self.gw = image_size // patch
        self.ntok = self.gh * self.gw
        self.mean, self.std = tuple(mean), tuple(std)
        self.inv_std = tuple(1.0 / s for s in std)
        self.logit_bias = 0.0 if logit_bias is None else float(logit_bias)
        self.inv_temp = 1.0 if logit_temperature is None else 1.0 / float(logit_temperature)
        self.seed_scale = float(cfg["seed_scale"])
        self.input_u8 = input_u8
        self.encoders: List[EncoderRec] = []
        self._build_forward()
        if explain:
            self._build_explain(want_grad6)

    # ------------------------------------------------------------------ helpers
    def _tok(self, c: int, planes: Optional[int] = None, dtype=None) -> Tensor:
        pl = self.planes if planes is None else planes
        return self._empty(self.nb, self.gh, self.gw, pl * c, dtype=dtype)

    def _rows(self) -> Tensor:
        return self._empty(1, self.nb * self.ntok, dtype=torch.float32)

    def _ln(self, name: str, x: Act, wkey: str, want_sq: bool) -> Tuple[Act, Tensor, Tensor]:
        """LayerNorm of the residual stream (`sp` planes) -> a branch operand (`bp` planes)."""
        w = self._dev(self.sd[wkey])
        y = self._tok(x.c, self.bp)
        rstd = self._empty(self.nb * self.ntok, dtype=torch.float32)
        sq = self._rows() if want_sq else None
        self.fwd_ops.append(O.VitLnFwdOp(name, x.t, x.c, self.sp, w, self.ln_eps, y, rstd, sq, self.dt_code, self.bp))
        return Act(y, x.c, sq, 1 if want_sq else 0), rstd, w

    def _lin(self, name: str, x: Act, w2d: Tensor, want_sq: bool = False, **kw):
        """B-cos linear (bcosifylinear.py:42-95: scale from ||x|| + 1e-12) as a 1x1 launch."""
        return self._conv_fwd(name, x, w2d[:, :, None, None], 1, 0, 0, bn=None, relu=False, sq_eps=(0.0, 1e-12), want_sq=want_sq, **kw)

    def _branch(self) -> Dict[str, object]:
        """launch format of a branch-internal linear map: `bp` planes in, out and in the weights"""
        return dict(a_planes=self.bp, w_planes=self.bp, y_planes=self.bp, hp=self.bp > 1)

    def _into_stream(self) -> Dict[str, object]:
        """launch format of a linear map whose output is added to the residual stream: `bp`-plane input, `sp`-plane weights,
        residual and output (the plane-aware kernel)"""
        return dict(a_planes=self.bp, w_planes=self.sp, y_planes=self.sp, res_planes=self.sp, hp=self.sp > 1)

    # ------------------------------------------------------------------ forward
    def _build_forward(self) -> None:
        nb, S, pl, sd, d = self.nb, self.size, self.planes, self.sd, self.dim
        p = self.patch
        self.x_in = self._empty(nb, 3, S, S, dtype=torch.uint8) if self.input_u8 else self._empty(nb, 6, S, S, dtype=torch.float32)
        pd = p * p * 6
        patches = self._tok(pd)
        sqp = self._rows()
        self.fwd_ops.append(O.VitPatchifyOp("patchify", self.x_in, p, self.mean, self.inv_std, patches, pl, self.dt_code, sqp))
        # positional embedding (vit.py:325-326) enters through the residual input of the patch-embedding launch
        pos = posemb_sincos_2d(self.gh, self.gw, d).view(1, self.gh, self.gw, d).expand(nb, -1, -1, -1)
        pos_t = self._dev(torch.cat(P.split_planes(pos, pl, self.dt), dim=-1), self.dt)
        x, self.patch_rec = self._lin("patch_embedding", Act(patches, pd, sqp, 1), sd["model.to_patch_embedding.linear.linear.weight"],
                                      res=Act(pos_t, d))
        for i in range(self.depth):
            pfx = f"model.transformer.encoder_{i}"
            h1, rstd1, w1 = self._ln(pfx + ".attn.norm", x, pfx + ".attn.norm.weight", want_sq=False)
            wqkv = sd[pfx + ".attn.to_qkv.weight"]
            bp = self.bp
            qkv, _ = self._conv_fwd(pfx + ".attn.to_qkv", h1, wqkv[:, :, None, None], 1, 0, 0, bn=None, relu=False, want_sq=False,
                                    scale_mode=L.BCOSK_SCALE_NONE, want_gain=False, **self._branch())     # plain nn.Linear (vit.py:140)
            o = self._tok(d, bp)
            self.fwd_ops.append(O.VitAttentionOp(pfx + ".attn.core", qkv.t, bp, None, nb, self.ntok, self.heads, self.dh, self.dh ** -0.5,
                                                 False, o, self.dt_code))
            sqo = self._rows()
            self.fwd_ops.append(O.PixelSqsumOp(pfx + ".attn.core.sq", o, d, bp, self.dt_code, sqo))
            x1, r_out = self._lin(pfx + ".attn.to_out", Act(o, d, sqo, 1), sd[pfx + ".attn.to_out.linear.weight"], res=x, **self._into_stream())
            h2, rstd2, w2 = self._ln(pfx + ".ff.net.norm", x1, pfx + ".ff.net.norm.weight", want_sq=True)
            if bp == 1 and self.fuse_gelu:
                # MyGELU (vit.py:89-113) inside linear1's epilogue: the launch writes the activation, its per-tile sums of squares and
                # the gain with the (detached) gate folded in - no pass over the [tokens, mlp] tensor in between
                act_in, r1 = self._lin(pfx + ".ff.net.linear1", h2, sd[pfx + ".ff.net.linear1.linear.weight"], want_sq=True, act=1,
                                       **self._branch())
            else:
                u, r1 = self._lin(pfx + ".ff.net.linear1", h2, sd[pfx + ".ff.net.linear1.linear.weight"], **self._branch())
                a = self._tok(self.mlp, bp)
                sqa = self._rows()
                self.fwd_ops.append(O.VitGeluFwdOp(pfx + ".ff.net.act", u.t, self.mlp, bp, a, sqa, r1.gain, self.dt_code))
                act_in = Act(a, self.mlp, sqa, 1)
            x2, r2 = self._lin(pfx + ".ff.net.linear2", act_in, sd[pfx + ".ff.net.linear2.linear.weight"], res=x1,
                               **self._into_stream())
            self.encoders.append(EncoderRec(pfx, qkv.t, rstd1, rstd2, w1, w2, r_out, r1, r2, wqkv[2 * d:3 * d, :, None, None].contiguous()))
            x = x2
        hN, self.rstd_head, self.w_head_ln = self._ln("model.linear_head.norm", x, "model.linear_head.norm.weight", want_sq=True)
        whead = sd["model.linear_head.linear.linear.weight"]
        self.ncls = whead.shape[0]
        fc, self.head_rec = self._lin("model.linear_head.linear", hN, whead, y_f32=True, a_planes=self.bp, w_planes=self.bp, hp=self.bp > 1)
        self.fc_out = fc.t.view(nb * self.ntok, self.ncls)
        self.logits = self._empty(nb, self.ncls, dtype=torch.float32)
        self.pred = self._zeros(nb, dtype=torch.int32)
        self.fwd_ops.append(O.GapLogitsOp("gap_logits", self.fc_out, nb, self.ntok, self.ncls, self.inv_temp, self.logit_bias,
                                          self.logits, self.pred))

    # ------------------------------------------------------------------ explanation pass
    def _build_explain(self, want_grad6: bool) -> None:
        nb, d = self.nb, self.dim
        M = nb * self.ntok
        for e in self.encoders:
            for r in (e.out, e.lin1, e.lin2):
                self._alloc_ghat(r)
        self._alloc_ghat(self.patch_rec)
        f32 = torch.float32
        G = [self._tok(d, 1, f32), self._tok(d, 1, f32)]           # residual-stream gradient, ping-pong
        g_a, g_o, g_h = self._tok(d, 1, f32), self._tok(d, 1, f32), self._tok(d, 1, f32)
        gv = self._tok(d, 1)                                         # P^T g, the A operand of the W_v data gradient
        # ---- seed: one-hot logit gradient through the token mean and the classifier's detached scale (bcos/common.py:166-177)
        self.w_head32 = self._dev(self.sd["model.linear_head.linear.linear.weight"])
        g_hn = self._tok(d, 1)
        self.bwd_ops.append(O.FcSeedOp("head.seed", self.pred, self.head_rec.gain, self.w_head32, nb, self.ntok, self.ncls, d,
                                       self.inv_temp, self.seed_scale, None, g_hn.view(M, d), None, None, 1, self.dt_code))
        cur = 0
        last = self.encoders[-1]
        self.bwd_ops.append(O.VitLnBwdOp("model.linear_head.norm.bwd", g_hn, None, d, self.w_head_ln, self.rstd_head, G[cur],
                                         last.lin2.gain, last.lin2.ghat, self.dt_code))
        for i in range(self.depth - 1, -1, -1):
            e = self.encoders[i]
            self._dgrad(e.lin2, y=e.lin1.ghat, mul1=e.lin1.gain)                 # x (gate x gain1): the GELU kernel folded the gate in
            self._dgrad(e.lin1, y=g_a, y_f32=True)
            self.bwd_ops.append(O.VitLnBwdOp(e.name + ".ff.net.norm.bwd", g_a, G[cur], d, e.w_ln2, e.rstd2, G[1 - cur], e.out.gain,
                                             e.out.ghat, self.dt_code))
            cur = 1 - cur
            self._dgrad(e.out, y=g_o, y_f32=True)
            self.bwd_ops.append(O.VitAttentionOp(e.name + ".attn.core.bwd", e.qkv, self.bp, g_o, nb, self.ntok, self.heads, self.dh,
                                                 self.dh ** -0.5, True, gv, self.dt_code))
            rec_v = ConvRec(e.name + ".attn.to_qkv.v", e.w_v, 1, 0, 0, (self.gh, self.gw), (self.gh, self.gw), d, ghat=gv,
                            algo_flops=2.0 * M * d * d)
            self._dgrad(rec_v, y=g_h, y_f32=True)
            prev = self.encoders[i - 1].lin2 if i > 0 else self.patch_rec
            self.bwd_ops.append(O.VitLnBwdOp(e.name + ".attn.norm.bwd", g_h, G[cur], d, e.w_ln1, e.rstd1, G[1 - cur], prev.gain, prev.ghat,
                                             self.dt_code))
            cur = 1 - cur
        pd = self.patch * self.patch * 6
        self.g_patch = self._tok(pd, 1, f32)
        self._dgrad(self.patch_rec, y=self.g_patch, y_f32=True)
        self.cmap = self._zeros(nb, self.size, self.size, dtype=f32)
        self.grad6 = self._zeros(nb, 6, self.size, self.size, dtype=f32) if want_grad6 else None
        self.bwd_ops.append(O.VitContribMapOp("contrib_map", self.g_patch, self.x_in, self.patch, self.inv_std, 1.0 / self.seed_scale,
                                              self.cmap, self.grad6))

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
