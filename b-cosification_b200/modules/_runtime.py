"""Execution of single B-cos modules on the C-ABI kernels (the un-fused, drop-in path).

A module call = layout bridge (NCHW fp32 -> NHWC 16-bit planes, + per-pixel sums of squares) -> one `bcosk_igemm`
launch with the B-cos epilogue -> layout bridge back.  In explanation mode (`module.detach`) the autograd node's
backward is the explain-dgrad launch fed with `g_out * gain`.  Per (module, input shape) a small launch plan with its
buffers and packed weights is cached and rebuilt when the weights change.

There is no other execution path: on a CPU tensor, or without libbcosk.so / a B200, the call raises.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import contextlib
import weakref

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


@contextlib.contextmanager
def precision(planes: int, dtype: str):
    """Operand format of the module-level launches issued inside the block (a fused plan runs its module-level head in the
    plan's own format).  LayerPlans are cached per format; a backward uses the plan its forward ran on."""
    old = (config.planes, config.dtype)
    config.planes, config.dtype = int(planes), dtype
    try:
        yield
    finally:
        config.planes, config.dtype = old


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
        # MaxOut over 2 / 4 / 8 units runs INSIDE the conv launch's epilogue (include/bcosk.h `max_out`: adjacent-column
        # maximum, scale of the kept unit, kept index).  Other group sizes (3: a group would straddle the 8-column register
        # groups of the epilogue): the launch computes the plain linear map of all O*M units (scale mode NONE) and
        # bcosk_maxout_bcos_fwd reduces and scales in a second pass.
        self.max_out, self.b_real = max_out, float(b)
        self.mo_fused = max_out in (2, 4, 8)
        super().__init__(nb, planes=config.planes, dtype=config.dtype, device=weight.device, explain=True,
                         b=b if (max_out == 1 or self.mo_fused) else 1.0)
        self.cin, self.cp = cin, (cin + 7) // 8 * 8
        o, _, kh, kw = weight.shape
        self.x = Act(self._empty(nb, h, w, self.planes * self.cp), self.cp, self._empty(1, nb * h * w, dtype=torch.float32), 1)
        wpad = weight.detach().float()
        if self.cp != cin:
            wpad = torch.cat([wpad, wpad.new_zeros(o, self.cp - cin, kh, kw)], 1)
        self.y, self.rec = self._conv_fwd("module", self.x, wpad, stride, pad, pad, bn=None, relu=False, y_f32=True,
                                          want_sq=False, lin_bias=None if bias is None else bias.detach().float(),
                                          sq_eps=(0.0, 1e-12) if linear_eps else (1e-6, 0.0),
                                          max_out=max_out if self.mo_fused else 1)
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
            self.mo_y = self.y.t if self.mo_fused else self._empty(nb, oh, ow, self.mo_o, dtype=torch.float32)
            self.mo_g16 = self._empty(self.mo_rows, self.planes * self.mo_o)

    # ---- forward: x NCHW fp32 -> y NCHW fp32 (and the gain tensor when an explanation backward may follow)
    def forward(self, x: Tensor, want_gain: bool) -> Tuple[Tensor, Optional[Tensor]]:
        nb, _, h, w = x.shape
        L.nchw_to_nhwc16(x, self.x.t, self.cp, self.planes, self.dt_code, None, self.x.sq)
        o = self.rec.cout
        if self.max_out > 1 and self.mo_fused:
            gain = torch.empty(self.mo_rows, self.mo_o, dtype=torch.float32, device=x.device) if want_gain else None
            amax = torch.empty(self.mo_rows, self.mo_o, dtype=torch.uint8, device=x.device) if want_gain else None
            self.fwd_op.gain, self.fwd_op.amax = gain, amax
            for op in self.fwd_ops:        # [stand-alone patch norm for large windows,] the fused conv + MaxOut launch
                op.run()
            out = torch.empty(nb, self.mo_o, *self.rec.out_hw, dtype=torch.float32, device=x.device)
            L.nhwc_to_nchw_f32(self.y.t, nb, self.mo_o, self.rec.out_hw[0], self.rec.out_hw[1], 1, self.dt_code, out)
            return out, (None if gain is None else (gain, amax))
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
            L.zero_insert_nhwc(self.ghat_dense, self.rec.ghat, self.rec.stride)
        self.bwd_op.run()
        h, w = self.rec.in_hw
        if self.dense_dgrad:                       # 1x1 strided conv: gradient lives on the sampled positions only
            gx = torch.zeros(nb, self.cin, h, w, dtype=torch.float32, device=gy.device)
            L.nhwc_scatter_nchw_f32(self.gx, nb, self.cin, oh, ow, 1, self.dt_code, gx, self.rec.stride)
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
            max_out: int = 1, extra: tuple = ()) -> LayerPlan:
        # everything the packed (effective) weight is made of: further tensors (weight-norm scale, the q / k / v matrices of a packed
        # projection) and flags come in through `extra`
        stamp = (weight.data_ptr(), weight._version, None if bias is None else (bias.data_ptr(), bias._version), str(weight.device),
                 tuple((e.data_ptr(), e._version) if isinstance(e, Tensor) else e for e in extra))
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


# ---------------------------------------------------------------------------------------------------------------------
# torch custom ops (`torch.library.custom_op` + `register_fake` + `register_autograd`).  A LayerPlan (device
# buffers, packed weights, launch records) is not a tensor: the ops take an integer handle into `_PLANS`; everything the
# fake-tensor tracer needs (output shapes / dtypes) is derived from the plan without touching the device.
# ---------------------------------------------------------------------------------------------------------------------
_PLANS: "weakref.WeakValueDictionary[int, LayerPlan]" = weakref.WeakValueDictionary()
_next_handle = [1]


def _handle_of(lp: LayerPlan) -> int:
    h = getattr(lp, "_handle", None)
    if h is None:
        h = lp._handle = _next_handle[0]
        _next_handle[0] += 1
        _PLANS[h] = lp
    return h


def _out_geometry(lp: LayerPlan):
    nb = lp.nb
    oh, ow = lp.rec.out_hw
    o = lp.mo_o if lp.max_out > 1 else lp.rec.cout
    return nb, o, oh, ow


@torch.library.custom_op("bcos_b200::bcos_map", mutates_args=())
def bcos_map_op(x: Tensor, handle: int, detach: bool, want_gain: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """y = B-cos transform of x (conv geometry of plan `handle`), plus what the explanation backward needs: the gain
    d y / d lin under the detached scale and, for MaxOut, the index of the kept unit (empty tensors when not requested)."""
    lp = _PLANS[handle]
    y, gain = lp.forward(x, want_gain)
    empty = x.new_empty(0)
    if gain is None:
        return y, empty, empty.to(torch.uint8)
    if isinstance(gain, tuple):
        return y, gain[0], gain[1]
    return y, gain, empty.to(torch.uint8)


@bcos_map_op.register_fake
def _bcos_map_fake(x, handle, detach, want_gain):
    lp = _PLANS[handle]
    nb, o, oh, ow = _out_geometry(lp)
    y = x.new_empty((nb, o, oh, ow), dtype=torch.float32)
    if not want_gain:
        return y, x.new_empty(0), x.new_empty(0, dtype=torch.uint8)
    if lp.max_out > 1:
        return y, x.new_empty((nb * oh * ow, o), dtype=torch.float32), x.new_empty((nb * oh * ow, o), dtype=torch.uint8)
    return y, x.new_empty((nb * oh * ow, o), dtype=lp.gain_dt), x.new_empty(0, dtype=torch.uint8)


@torch.library.custom_op("bcos_b200::bcos_map_explain_bwd", mutates_args=())
def bcos_map_explain_bwd_op(gy: Tensor, gain: Tensor, amax: Tensor, handle: int) -> Tensor:
    """Dynamic-linear (explanation) gradient: W^T (g_out * gain), reference bcos/common.py:163-177 with detached scales."""
    lp = _PLANS[handle]
    return lp.explain_backward(gy, gain, amax if amax.numel() else None)


@bcos_map_explain_bwd_op.register_fake
def _bcos_map_explain_bwd_fake(gy, gain, amax, handle):
    lp = _PLANS[handle]
    h, w = lp.rec.in_hw
    return gy.new_empty((lp.nb, lp.cin, h, w), dtype=torch.float32)


def _bcos_map_setup(ctx, inputs, output):
    _, handle, detach, want_gain = inputs
    ctx.handle, ctx.detach, ctx.have_gain = handle, detach, want_gain
    ctx.save_for_backward(output[1], output[2])


def _bcos_map_backward(ctx, gy, _g_gain, _g_amax):
    if not ctx.detach:
        raise NotImplementedError(
            "bcos_b200 modules: only the explanation-mode backward (detached dynamic scale; reference bcos/common.py:163-177) runs "
            "through the module-level ops; the full training backward is engine.ResNetTrainPlan (fused fine-tuning step)")
    if not ctx.have_gain:
        raise RuntimeError("bcos_b200::bcos_map: backward requested but the forward ran without saving the gain")
    gain, amax = ctx.saved_tensors
    return torch.ops.bcos_b200.bcos_map_explain_bwd(gy.contiguous(), gain, amax, ctx.handle), None, None, None


torch.library.register_autograd("bcos_b200::bcos_map", _bcos_map_backward, setup_context=_bcos_map_setup)


def _is_fake(x: Tensor) -> bool:
    try:
        from torch._subclasses.fake_tensor import FakeTensor
        return isinstance(x, FakeTensor)
    except Exception:  # noqa: BLE001
        return False


def bcos_map(x: Tensor, cache: _PlanCache, weight: Tensor, bias: Optional[Tensor], eff_weight_fn, stride: int, pad: int,
             b: float, detach: bool, linear_eps: bool = False, max_out: int = 1, extra: tuple = ()) -> Tensor:
    if not _is_fake(x):
        _require_cuda(x, "B-cos module")
    x32 = x.float().contiguous()
    lp = cache.get(weight, bias, eff_weight_fn, tuple(x32.shape), stride, pad, b, linear_eps, max_out, extra)
    want_grad = torch.is_grad_enabled() and x.requires_grad
    y = torch.ops.bcos_b200.bcos_map(x32, _handle_of(lp), bool(detach), bool(want_grad))[0]
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
