"""Execution of single B-cos modules on the C-ABI kernels (the un-fused, drop-in path).

A module call = layout bridge (NCHW fp32 -> NHWC 16-bit planes, + per-pixel sums of squares) -> one `bcosk_igemm`
launch with the B-cos epilogue -> layout bridge back.  In explanation mode (`module.detach`) the autograd node's
backward is the explain-dgrad launch fed with `g_out * gain`.  Per (module, input shape) a small launch plan with its
buffers and packed weights is cached and rebuilt when the weights change.

There is no other execution path: on a CPU tensor, or without libbcosk.so / a B200, the call raises.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .. import _lib as L
from ..engine import ops as O
from ..engine.base import Act, PlanBase


class config:
    """Precision of the module-level path: 3 bf16 planes + fp32-faithful accumulation ("parity", default) or 1 ("bf16")."""
    planes = 3
    dtype = "bf16"


def set_precision(mode: str) -> None:
    config.planes = {"parity": 3, "bf16x2": 2, "bf16": 1}[mode]


def _require_cuda(x: Tensor, who: str) -> None:
    if not x.is_cuda:
        raise L.BcoskError(f"{who}: bcos_b200 modules run only on CUDA tensors (sm_100a kernels); there is no CPU fallback")
    L.require_device()


class LayerPlan(PlanBase):
    """Launches for one conv / linear module at one input shape."""
    hp_chunk = 1      # three bf16 planes: every K stage of the leading segment gets its own accumulation (strictest)

    def __init__(self, weight: Tensor, bias: Optional[Tensor], in_shape: Tuple[int, int, int, int], stride: int, pad: int,
                 b: float, linear_eps: bool, max_out: int = 1):
        nb, cin, h, w = in_shape
        # MaxOut: the launch computes the plain linear map of all O*M units (scale mode NONE); the maximum over each group
        # and the B-cos scale of the kept unit follow in bcosk_maxout_bcos_fwd
        self.max_out, self.b_real = max_out, float(b)
        super().__init__(nb, planes=config.planes, dtype=config.dtype, device=weight.device, explain=True,
                         b=b if max_out == 1 else 1.0)
        self.cin, self.cp = cin, (cin + 7) // 8 * 8
        o, _, kh, kw = weight.shape
        self.x = Act(self._empty(nb, h, w, self.planes * self.cp), self.cp, self._empty(1, nb * h * w, dtype=torch.float32), 1)
        wpad = weight.detach().float()
        if self.cp != cin:
            wpad = torch.cat([wpad, wpad.new_zeros(o, self.cp - cin, kh, kw)], 1)
        self.y, self.rec = self._conv_fwd("module", self.x, wpad, stride, pad, pad, bn=None, relu=False, y_f32=True,
                                          want_sq=False, lin_bias=None if bias is None else bias.detach().float(),
                                          sq_eps=(0.0, 1e-12) if linear_eps else (1e-6, 0.0))
        self.fwd_op = self.fwd_ops[-1]
        self._alloc_ghat(self.rec)
        dense = self.rec.stride > 1 and self.rec.k == 1
        gh, gw = self.rec.out_hw if dense else self.rec.in_hw
        self.gx = self._empty(nb, gh, gw, self.cp, dtype=torch.float32)
        self._dgrad(self.rec, y=self.gx, y_f32=True)
        self.bwd_op = self.bwd_ops[-1]
        self.ghat_dense = self.rec.ghat if self.rec.ghat_map is None else self._empty(nb, *self.rec.out_hw, self.planes * o)
        self.dense_dgrad = dense
        if max_out > 1:
            assert o % max_out == 0
            oh, ow = self.rec.out_hw
            self.mo_rows, self.mo_o = nb * oh * ow, o // max_out
            self.mo_scale = L.BCOSK_SCALE_NONE if self.b_real == 1.0 else (L.BCOSK_SCALE_B2 if self.b_real == 2.0 else L.BCOSK_SCALE_POW)
            self.mo_inv = self._empty(self.mo_rows, dtype=torch.float32)
            eps_in, eps_out = (0.0, 1e-12) if linear_eps else (1e-6, 0.0)
            self.mo_norm = O.PatchNormOp("module.norm", self.x.sq, 1, nb, h, w, kh, stride, pad, eps_in, eps_out, self.mo_inv, oh, ow)
            self.mo_y = self._empty(nb, oh, ow, self.mo_o, dtype=torch.float32)
            self.mo_g16 = self._empty(self.mo_rows, self.planes * self.mo_o)

    # ---- forward: x NCHW fp32 -> y NCHW fp32 (and the gain tensor when an explanation backward may follow)
    def forward(self, x: Tensor, want_gain: bool) -> Tuple[Tensor, Optional[Tensor]]:
        nb, _, h, w = x.shape
        L.nchw_to_nhwc16(x, self.x.t, self.cp, self.planes, self.dt_code, None, self.x.sq)
        o = self.rec.cout
        if self.max_out > 1:
            self.fwd_op.gain = None
            for op in self.fwd_ops:
                op.run()
            if self.mo_scale != L.BCOSK_SCALE_NONE:
                self.mo_norm.run()
            gain = torch.empty(self.mo_rows, self.mo_o, dtype=torch.float32, device=x.device) if want_gain else None
            amax = torch.empty(self.mo_rows, self.mo_o, dtype=torch.uint8, device=x.device) if want_gain else None
            L.maxout_bcos_fwd(self.y.t, self.mo_inv, self.mo_rows, self.mo_o, self.max_out, self.mo_scale, self.b_real, self.mo_y,
                              gain, amax)
            out = torch.empty(nb, self.mo_o, *self.rec.out_hw, dtype=torch.float32, device=x.device)
            L.nhwc_to_nchw_f32(self.mo_y, nb, self.mo_o, self.rec.out_hw[0], self.rec.out_hw[1], 1, self.dt_code, out)
            return out, (None if gain is None else (gain, amax))
        gain = torch.empty(nb * self.rec.out_hw[0] * self.rec.out_hw[1], o, dtype=self.gain_dt, device=x.device) if want_gain else None
        self.fwd_op.gain = gain
        for op in self.fwd_ops:        # [stand-alone patch norm for large windows,] the fused conv launch
            op.run()
        out = torch.empty(nb, o, *self.rec.out_hw, dtype=torch.float32, device=x.device)
        L.nhwc_to_nchw_f32(self.y.t, nb, o, self.rec.out_hw[0], self.rec.out_hw[1], 1, self.dt_code, out)
        return out, gain

    # ---- explanation backward: g_out NCHW fp32, gain [M, o] -> g_in NCHW fp32
    def explain_backward(self, gy: Tensor, gain, amax: Optional[Tensor] = None) -> Tensor:
        nb, o, oh, ow = gy.shape
        if self.max_out > 1:
            # g_out * scale on the O kept units, then routed to the kept unit of every group (zeros elsewhere)
            L.nchw_to_nhwc16(gy.contiguous(), self.mo_g16, o, self.planes, self.dt_code, gain, None)
            L.maxout_scatter(self.mo_g16, amax, self.mo_rows, self.mo_o, self.max_out, self.planes, self.dt_code, self.ghat_dense)
        else:
            L.nchw_to_nhwc16(gy.contiguous(), self.ghat_dense, o, self.planes, self.dt_code, gain.float() if gain.dtype != torch.float32 else gain, None)
        if self.rec.ghat_map is not None:          # strided k>1 conv: zero-inserted gradient at input resolution
            s = self.rec.stride
            self.rec.ghat[:, ::s, ::s][:, :oh, :ow] = self.ghat_dense
        self.bwd_op.run()
        h, w = self.rec.in_hw
        if self.dense_dgrad:                       # 1x1 strided conv: gradient lives on the sampled positions only
            s = self.rec.stride
            small = torch.empty(nb, self.cin, oh, ow, dtype=torch.float32, device=gy.device)
            L.nhwc_to_nchw_f32(self.gx, nb, self.cin, oh, ow, 1, self.dt_code, small)
            gx = torch.zeros(nb, self.cin, h, w, dtype=torch.float32, device=gy.device)
            gx[:, :, ::s, ::s][:, :, :oh, :ow] = small
            return gx
        gx = torch.empty(nb, self.cin, h, w, dtype=torch.float32, device=gy.device)
        L.nhwc_to_nchw_f32(self.gx, nb, self.cin, h, w, 1, self.dt_code, gx)
        return gx


class _PlanCache:
    """Per-module cache of LayerPlans keyed by input shape + precision; dropped when the weights change."""

    def __init__(self):
        self.plans: Dict = {}
        self.stamp = None

    def get(self, weight: Tensor, bias: Optional[Tensor], eff_weight_fn, in_shape, stride, pad, b, linear_eps,
            max_out: int = 1) -> LayerPlan:
        stamp = (weight.data_ptr(), weight._version, None if bias is None else (bias.data_ptr(), bias._version), str(weight.device))
        if stamp != self.stamp:
            self.plans.clear()
            self.stamp = stamp
        key = (tuple(in_shape), config.planes, config.dtype, float(b), int(max_out))
        lp = self.plans.get(key)
        if lp is None:
            # plan buffers outlive this call: they must be ordinary tensors even when the first call happens under
            # torch.inference_mode() (evaluate.py) - inference tensors could not be updated later
            with torch.inference_mode(False), torch.no_grad():
                lp = LayerPlan(eff_weight_fn().detach().clone(), None if bias is None else bias.detach().clone(), in_shape,
                               stride, pad, float(b), linear_eps, int(max_out))
            self.plans[key] = lp
        return lp


class BcosMapFn(torch.autograd.Function):
    """y = B-cos transform of x (conv geometry); backward = dynamic-linear (explanation) gradient."""

    @staticmethod
    def forward(ctx, x: Tensor, lp: LayerPlan, detach: bool, want_grad: bool):
        y, gain = lp.forward(x, want_grad)
        ctx.lp, ctx.detach = lp, detach
        if gain is not None:
            ctx.save_for_backward(*(gain if isinstance(gain, tuple) else (gain,)))
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        if not ctx.detach:
            raise NotImplementedError(
                "bcos_b200: only the explanation-mode backward (detached dynamic scale; reference bcos/common.py:163-177) "
                "is built; the full training backward is outside this round's scope")
        saved = ctx.saved_tensors
        return ctx.lp.explain_backward(gy, *saved), None, None, None


def bcos_map(x: Tensor, cache: _PlanCache, weight: Tensor, bias: Optional[Tensor], eff_weight_fn, stride: int, pad: int,
             b: float, detach: bool, linear_eps: bool = False, max_out: int = 1) -> Tensor:
    _require_cuda(x, "B-cos module")
    x32 = x.float().contiguous()
    lp = cache.get(weight, bias, eff_weight_fn, tuple(x32.shape), stride, pad, b, linear_eps, max_out)
    want_grad = torch.is_grad_enabled() and x.requires_grad
    y = BcosMapFn.apply(x32, lp, detach, want_grad)
    return y if x.dtype == torch.float32 else y.to(x.dtype)


class ChannelAffineFn(torch.autograd.Function):
    """y = x * alpha[c] + beta[c] (NCHW fp32), alpha treated as a constant (detached statistics)."""

    @staticmethod
    def forward(ctx, x: Tensor, alpha: Optional[Tensor], beta: Optional[Tensor], smul: float, sadd: float, const_ok: bool):
        nb, c = x.shape[0], x.shape[1]
        hw = x.numel() // (nb * c)
        out = torch.empty_like(x)
        L.scale_bias_nchw(x, nb, c, hw, alpha, beta, smul, sadd, False, out)
        ctx.alpha, ctx.smul, ctx.const_ok = alpha, smul, const_ok
        return out

    @staticmethod
    def backward(ctx, gy: Tensor):
        if not ctx.const_ok:
            raise NotImplementedError("bcos_b200: gradient through the batch statistics (non-explanation training backward) is not built")
        gy = gy.contiguous()
        nb, c = gy.shape[0], gy.shape[1]
        hw = gy.numel() // (nb * c)
        gx = torch.empty_like(gy)
        L.scale_bias_nchw(gy, nb, c, hw, ctx.alpha, None, ctx.smul, 0.0, False, gx)
        return gx, None, None, None, None, None
