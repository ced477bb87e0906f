"""Fused execution plan for the B-cos CLIP ViT image encoder (north_star: "the CLIP RN50 and ViT image encoders"): embedding +
explanation.

Network (reference): CLIP/clip/model.py:206-241 `VisionTransformer` (conv1 patch embedding, class token, ln_pre, `Transformer` of
`ResidualAttentionBlock`s :171-204, ln_post, proj) converted by bcosify.py:74-113 with `clip_kd`: conv1 -> BcosifyConv2d over the
6-channel input, mlp.c_fc / mlp.c_proj -> BcosifyLinear, attn.out_proj a BcosifyLinear OBJECT whose weight nn.MultiheadAttention uses
as a plain matrix; `.bias` attributes and the positional embedding stripped (clip_bcosification/model.py:17-25; `in_proj_bias`
survives).  LayerNorm, QuickGELU and the attention are stock torch modules in the reference - nothing there is detached, so the
explanation pass differentiates them exactly (csrc/bcosk_vit.cu: `vit_ln_bwd_full`, `vit_quickgelu_fwd`, `vit_attention_bwd_full`);
only the B-cos scales are frozen.

Layout like engine/vit.py: token tensors [images, 50, 1, planes * d]; the patch embedding (a 32x32 / 32 conv = a 1x1 map over the
unfolded patches, K = 6144) writes rows 1..49 of every image through the launch's output map, row 0 holds the class embedding.
The residual stream keeps `planes` planes, branch operands `branch_planes` (default 1), the residual-stream gradient is fp32.
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
from .clip_rn import CLIP_MEAN_ADDINVERSE, CLIP_STD_ADDINVERSE
from .resnet import resolve_precision


@dataclass
class ClipBlockRec:
    name: str
    x_in: Tensor          # stream entering the block (input of ln_1)
    x_mid: Tensor         # stream after the attention branch (input of ln_2)
    qkv: Tensor
    rstd1: Tensor
    rstd2: Tensor
    w_ln1: Tensor
    w_ln2: Tensor
    in_proj: ConvRec
    out_proj: ConvRec
    c_fc: ConvRec
    c_proj: ConvRec


class CLIPViTPlan(PlanBase):
    fuse_gelu = True     # one-plane branches: QuickGELU inside c_fc's epilogue; False keeps the separate pass

    def __init__(self, state_dict: Dict[str, Tensor], batch: int, *, heads: int = 12, mode: Optional[str] = None, planes: Optional[int] = None,
                 dtype: Optional[str] = None, device="cuda", image_size: int = 224, explain: bool = True, want_grad6: bool = False,
                 b: float = 2.0, ln_eps: float = 1e-5, mean=CLIP_MEAN_ADDINVERSE, std=CLIP_STD_ADDINVERSE, seed_scale: Optional[float] = None,
                 input_u8: bool = False, explain_planes: Optional[int] = None, branch_planes: Optional[int] = None):
        cfg = resolve_precision(mode, planes, dtype, explain_planes, seed_scale)
        cfg["explain_planes"] = 1          # the (linear) explanation pass of this plan always runs on one 16-bit plane
        self.precision = cfg
        super().__init__(batch, planes=cfg["planes"], dtype=cfg["dtype"], device=device, explain=explain, b=b, state_dict=state_dict,
                         explain_planes=cfg["explain_planes"])
        assert self.bplanes == 1
        self.sp = self.planes
        if branch_planes is None:
            branch_planes = 1 if (mode in (None, "parity") and planes is None and dtype is None) else self.planes
        self.bp = int(branch_planes)
        sd = self.sd
        w1 = sd["model.conv1.linear.weight"]
        self.dim, self.patch = w1.shape[0], w1.shape[-1]
        self.heads, self.dh = heads, self.dim // heads
        assert self.dh == 64
        self.depth = len([k for k in sd if k.endswith(".attn.in_proj_weight")])
        self.size, self.ln_eps = image_size, ln_eps
        self.gh = image_size // self.patch
        self.ntok = self.gh * self.gh + 1
        self.mean, self.std = tuple(mean), tuple(std)
        self.inv_std = tuple(1.0 / s for s in std)
        self.seed_scale = float(cfg["seed_scale"])
        self.input_u8 = input_u8
        self.blocks: List[ClipBlockRec] = []
        self._build_forward()
        if explain:
            self._build_explain(want_grad6)

    # ------------------------------------------------------------------ helpers
    def _tok(self, c: int, planes: int, dtype=None) -> Tensor:
        return self._empty(self.nb, self.ntok, 1, planes * c, dtype=dtype)

    def _rows(self) -> Tensor:
        return self._empty(1, self.nb * self.ntok, dtype=torch.float32)

    def _ln(self, name: str, x: Tensor, wkey: str, want_sq: bool, out_planes: int):
        w = self._dev(self.sd[wkey])
        y = self._tok(self.dim, out_planes)
        rstd = self._empty(self.nb * self.ntok, dtype=torch.float32)
        sq = self._rows() if want_sq else None
        self.fwd_ops.append(O.VitLnFwdOp(name, x, self.dim, self.sp, w, self.ln_eps, y, rstd, sq, self.dt_code, out_planes))
        return Act(y, self.dim, sq, 1 if want_sq else 0), rstd, w

    def _branch(self) -> Dict[str, object]:
        return dict(a_planes=self.bp, w_planes=self.bp, y_planes=self.bp, hp=self.bp > 1)

    def _into_stream(self) -> Dict[str, object]:
        return dict(a_planes=self.bp, w_planes=self.sp, y_planes=self.sp, res_planes=self.sp, hp=self.sp > 1)

    # ------------------------------------------------------------------ forward
    def _build_forward(self) -> None:
        nb, S, sd, d, p = self.nb, self.size, self.sd, self.dim, self.patch
        NONE = L.BCOSK_SCALE_NONE
        self.x_in = self._empty(nb, 3, S, S, dtype=torch.uint8) if self.input_u8 else self._empty(nb, 6, S, S, dtype=torch.float32)
        pd = p * p * 6
        patches = self._empty(nb, self.gh, self.gh, self.sp * pd)
        sqp = self._empty(1, nb * self.gh * self.gh, dtype=torch.float32)
        self.fwd_ops.append(O.VitPatchifyOp("patchify", self.x_in, p, self.mean, self.inv_std, patches, self.sp, self.dt_code, sqp))
        # tokens: row 0 = class embedding (constant), rows 1.. = B-cos patch embedding (conv eps: the norm is that of a conv patch)
        X = self._tok(d, self.sp)
        cls = torch.cat(P.split_planes(sd["model.class_embedding"].view(1, d), self.sp, self.dt), dim=-1)
        X[:, 0, 0, :] = cls.to(self.device)
        wpe = sd["model.conv1.linear.weight"].permute(0, 2, 3, 1).reshape(d, pd)                   # (p1 p2 c) = the patch rows' column order
        _, self.patch_rec = self._conv_fwd("conv1", Act(patches, pd, sqp, 1), wpe[:, :, None, None], 1, 0, 0, bn=None, relu=False,
                                           want_sq=False, sq_eps=(1e-6, 0.0), y_buf=X, out_map=(1, self.ntok, self.gh, 1))
        xs, self.rstd_pre, self.w_pre = self._ln("ln_pre", X, "model.ln_pre.weight", False, self.sp)
        self.tokens_in = X
        x = xs.t
        for i in range(self.depth):
            pfx = f"model.transformer.resblocks.{i}"
            y1, rstd1, w1 = self._ln(pfx + ".ln_1", x, pfx + ".ln_1.weight", False, self.bp)
            qkv, r_in = self._conv_fwd(pfx + ".attn.in_proj", y1, sd[pfx + ".attn.in_proj_weight"][:, :, None, None], 1, 0, 0, bn=None,
                                       relu=False, want_sq=False, scale_mode=NONE, want_gain=False, lin_bias=sd[pfx + ".attn.in_proj_bias"],
                                       **self._branch())
            o = self._tok(d, self.bp)
            self.fwd_ops.append(O.VitAttentionOp(pfx + ".attn.core", qkv.t, self.bp, None, nb, self.ntok, self.heads, self.dh, self.dh ** -0.5,
                                                 False, o, self.dt_code))
            x1, r_out = self._conv_fwd(pfx + ".attn.out_proj", Act(o, d), sd[pfx + ".attn.out_proj.linear.weight"][:, :, None, None], 1, 0, 0,
                                       bn=None, relu=False, want_sq=False, scale_mode=NONE, want_gain=False, res=Act(x, d), **self._into_stream())
            y2, rstd2, w2 = self._ln(pfx + ".ln_2", x1.t, pfx + ".ln_2.weight", True, self.bp)
            if self.bp == 1 and self.fuse_gelu:
                # QuickGELU inside c_fc's epilogue (include/bcosk.h `act` = 2): activation, its sums of squares and the gain x QuickGELU'
                act_in, r_fc = self._conv_fwd(pfx + ".mlp.c_fc", y2, sd[pfx + ".mlp.0.linear.weight"][:, :, None, None], 1, 0, 0, bn=None,
                                              relu=False, want_sq=True, sq_eps=(0.0, 1e-12), act=2, **self._branch())
            else:
                u, r_fc = self._conv_fwd(pfx + ".mlp.c_fc", y2, sd[pfx + ".mlp.0.linear.weight"][:, :, None, None], 1, 0, 0, bn=None, relu=False,
                                         want_sq=False, sq_eps=(0.0, 1e-12), **self._branch())
                hid = u.c
                a = self._tok(hid, self.bp)
                sqa = self._rows()
                self.fwd_ops.append(O.VitGeluFwdOp(pfx + ".mlp.gelu", u.t, hid, self.bp, a, sqa, r_fc.gain, self.dt_code, True))
                act_in = Act(a, hid, sqa, 1)
            x2, r_pr = self._conv_fwd(pfx + ".mlp.c_proj", act_in, sd[pfx + ".mlp.2.linear.weight"][:, :, None, None], 1, 0, 0,
                                      bn=None, relu=False, want_sq=False, sq_eps=(0.0, 1e-12), res=x1, **self._into_stream())
            self.blocks.append(ClipBlockRec(pfx, x, x1.t, qkv.t, rstd1, rstd2, w1, w2, r_in, r_out, r_fc, r_pr))
            x = x2.t
        self.x_last = x
        yN, self.rstd_post, self.w_post = self._ln("ln_post", x, "model.ln_post.weight", False, self.bp)
        # x[:, 0] @ proj: the projection runs over all tokens (50 x the needed rows of a tiny GEMM), the embedding is row 0 of every image
        proj = sd["model.proj"]                                                                      # [d, out]
        self.out_dim = proj.shape[1]
        e_all, self.proj_rec = self._conv_fwd("proj", yN, proj.t().contiguous()[:, :, None, None], 1, 0, 0, bn=None, relu=False, want_sq=False,
                                              scale_mode=NONE, want_gain=False, y_f32=True, a_planes=self.bp, w_planes=self.bp, hp=self.bp > 1)
        self.emb_all = e_all.t.view(nb, self.ntok, self.out_dim)
        self.proj32 = self._dev(proj)

    # ------------------------------------------------------------------ explanation pass (true gradients except the B-cos scales)
    def _build_explain(self, want_grad6: bool) -> None:
        nb, d, T = self.nb, self.dim, self.ntok
        M = nb * T
        f32 = torch.float32
        for blk in self.blocks:
            for r in (blk.c_fc, blk.c_proj):
                self._alloc_ghat(r)
        self._alloc_ghat(self.patch_rec)
        G = [self._tok(d, 1, f32), self._tok(d, 1, f32)]
        g_y2, g_o, g_y1 = self._tok(d, 1, f32), self._tok(d, 1, f32), self._tok(d, 1, f32)
        g_attn = self._tok(d, 1)                 # A operand of the out_proj data gradient (plain linear: no gain)
        g_qkv = self._tok(3 * d, 1)              # A operand of the in_proj data gradient
        self.g_emb = self._zeros(nb, self.out_dim, dtype=f32)
        # d target / d ln_post output: only token 0 of every image; the SGEMM writes into the strided rows of a zeroed buffer
        g_post = self._zeros(nb, T, 1, d, dtype=f32)
        self.bwd_ops.append(O.SgemmOp("proj.bwd", False, True, nb, d, self.out_dim, self.g_emb, 0, self.out_dim, 0, self.proj32, 0, self.out_dim, 0,
                                      g_post, 0, T * d, 0, 1))
        cur = 0
        last = self.blocks[-1]
        self.bwd_ops.append(O.VitLnBwdOp("ln_post.bwd", g_post, None, d, self.w_post, self.rstd_post, G[cur], last.c_proj.gain, last.c_proj.ghat,
                                         self.dt_code, self.x_last, self.sp))
        for i in range(self.depth - 1, -1, -1):
            blk = self.blocks[i]
            self._dgrad(blk.c_proj, y=blk.c_fc.ghat, mul1=blk.c_fc.gain)            # x (gain_fc x QuickGELU'): folded by the GELU kernel
            self._dgrad(blk.c_fc, y=g_y2, y_f32=True)
            self.bwd_ops.append(O.VitLnBwdOp(blk.name + ".ln_2.bwd", g_y2, G[cur], d, blk.w_ln2, blk.rstd2, G[1 - cur], None, g_attn, self.dt_code,
                                             blk.x_mid, self.sp))
            cur = 1 - cur
            rec_o = ConvRec(blk.name + ".attn.out_proj", blk.out_proj.w, 1, 0, 0, (T, 1), (T, 1), d, ghat=g_attn, algo_flops=2.0 * M * d * d)
            self._dgrad(rec_o, y=g_o, y_f32=True)
            self.bwd_ops.append(O.VitAttentionOp(blk.name + ".attn.core.bwd", blk.qkv, self.bp, g_o, nb, T, self.heads, self.dh, self.dh ** -0.5,
                                                 True, g_qkv, self.dt_code, True, True))
            rec_i = ConvRec(blk.name + ".attn.in_proj", blk.in_proj.w, 1, 0, 0, (T, 1), (T, 1), d, ghat=g_qkv, algo_flops=2.0 * M * d * 3 * d)
            self._dgrad(rec_i, y=g_y1, y_f32=True)
            prev = self.blocks[i - 1].c_proj if i > 0 else None
            self.bwd_ops.append(O.VitLnBwdOp(blk.name + ".ln_1.bwd", g_y1, G[cur], d, blk.w_ln1, blk.rstd1, G[1 - cur],
                                             None if prev is None else prev.gain, None if prev is None else prev.ghat, self.dt_code,
                                             blk.x_in, self.sp))
            cur = 1 - cur
        # ln_pre backward -> gradient wrt the tokens; rows 1.. of every image x conv1's gain -> patch-embedding data gradient
        g_tok = self._tok(d, 1, f32)
        self.bwd_ops.append(O.VitLnBwdOp("ln_pre.bwd", G[cur], None, d, self.w_pre, self.rstd_pre, g_tok, None, None, self.dt_code,
                                         self.tokens_in, self.sp))
        npx = self.gh * self.gh
        self.bwd_ops.append(O.DenseSliceCastOp("conv1.ghat", g_tok.view(nb, 1, 1, T * d), d, npx * d, self.patch_rec.gain.view(nb, npx * d), 1.0,
                                               self.patch_rec.ghat.view(nb, 1, 1, npx * d), self.dt_code))
        pd = self.patch * self.patch * 6
        self.g_patch = self._empty(nb, self.gh, self.gh, pd, dtype=f32)
        self._dgrad(self.patch_rec, y=self.g_patch, y_f32=True)
        self.cmap = self._zeros(nb, self.size, self.size, dtype=f32)
        self.grad6 = self._zeros(nb, 6, self.size, self.size, dtype=f32) if want_grad6 else None
        self.bwd_ops.append(O.VitContribMapOp("contrib_map", self.g_patch, self.x_in, self.patch, self.inv_std, 1.0 / self.seed_scale,
                                              self.cmap, self.grad6))

    # ------------------------------------------------------------------ public API (same as CLIPResNetPlan)
    def load_input(self, x6: Tensor) -> None:
        assert tuple(x6.shape) == tuple(self.x_in.shape), (x6.shape, self.x_in.shape)
        self.x_in.copy_(x6, non_blocking=True)

    def embed(self, x6: Optional[Tensor] = None) -> Tensor:
        if x6 is not None:
            self.load_input(x6)
        self.replay_forward()
        return self.emb_all[:, 0]

    def explain_target(self, x6: Optional[Tensor], target_fn) -> Dict[str, Tensor]:
        if not self.with_explain:
            raise RuntimeError("plan was built with explain=False")
        if x6 is not None:
            self.load_input(x6)
        self.replay_forward()
        emb = self.emb_all[:, 0].detach().clone().requires_grad_(True)
        with torch.enable_grad():
            (g,) = torch.autograd.grad(target_fn(emb).sum() * self.seed_scale, [emb])
        self.g_emb.copy_(g)
        self.replay_explain()
        out = {"embedding": emb.detach(), "contribution_map": self.cmap}
        if self.grad6 is not None:
            out["dynamic_linear_weights"] = self.grad6
        return out

    def explain_direction(self, x6: Optional[Tensor], direction: Tensor) -> Dict[str, Tensor]:
        d = direction.to(self.device, torch.float32)
        return self.explain_target(x6, lambda emb: torch.nn.functional.cosine_similarity(emb, d[None], dim=1))
