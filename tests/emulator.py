"""CPU restatement of every launch record in bcos_b200.engine.ops -- TEST INFRASTRUCTURE ONLY.

It executes a plan's op list with plain torch on the CPU, following the kernels' documented semantics
(include/bcosk.h) including the TMA im2col traversal, the packed (segment, tap, chunk) K order, precision
planes and 16-bit storage rounding.  Used to (a) check the host-side plan logic (weight packing, tap
tables, zero-insertion, residual/gradient routing) against the oracle without a GPU, and (b) as the
per-kernel expected value in the `-m gpu` tests.  The product never imports this file.
"""
from __future__ import annotations

import torch
from torch import Tensor

from bcos_b200 import _lib as L
from bcos_b200.engine import ops as O


def _split_store(dst: Tensor, val: Tensor, planes: int) -> Tensor:
    """val [..., C] fp32 -> dst [..., planes*C] 16-bit; returns the stored (representable) value."""
    c = val.shape[-1]
    r = val.clone()
    acc = torch.zeros_like(val)
    for p in range(planes):
        h = r.to(dst.dtype)
        dst[..., p * c:(p + 1) * c] = h
        r = r - h.float()
        acc = acc + h.float()
    return acc


def _join(t: Tensor, planes: int) -> Tensor:
    c = t.shape[-1] // planes
    acc = t[..., :c].float()
    for p in range(1, planes):
        acc = acc + t[..., p * c:(p + 1) * c].float()
    return acc


def gather_a(op: O.IgemmOp) -> Tensor:
    """[M, ktot] fp32: what the TMA im2col loads deliver, chunk by chunk."""
    a = op.a.float()
    nb, h, w, ac = a.shape
    # the TMA unit derives the number of base pixels per row/column from the bounding box; the plan must agree
    assert op.oq == (w + op.up[0] - op.lo[0] - 1) // op.stride[0] + 1, (op.name, "oq vs im2col box")
    assert op.op == (h + op.up[1] - op.lo[1] - 1) // op.stride[1] + 1, (op.name, "op vs im2col box")
    assert -128 <= op.lo[0] <= 127 and -128 <= op.up[0] <= 127
    p = torch.arange(op.op)
    q = torch.arange(op.oq)
    cols = []
    for choff in op.seg_a_choff:
        for (off_w, off_h) in op.taps:
            hh = op.lo[1] + p * op.stride[1] + off_h
            ww = op.lo[0] + q * op.stride[0] + off_w
            vh = (hh >= 0) & (hh < h)
            vw = (ww >= 0) & (ww < w)
            t = a[:, hh.clamp(0, h - 1)][:, :, ww.clamp(0, w - 1)]           # [nb, op, oq, ac]
            t = t * (vh[:, None] & vw[None, :])[None, :, :, None]
            for kc in range(op.chunks_per_tap):
                c0 = choff + kc * op.kch
                chunk = torch.zeros(nb, op.op, op.oq, op.kch)
                c1 = min(c0 + op.kch, ac)
                if c1 > c0:
                    chunk[..., :c1 - c0] = t[..., c0:c1]
                cols.append(chunk.reshape(-1, op.kch))
    return torch.cat(cols, dim=1)


def _out_rows(op: O.IgemmOp) -> Tensor:
    nb = op.a.shape[0]
    img = torch.arange(nb).view(-1, 1, 1)
    p = torch.arange(op.op).view(1, -1, 1)
    q = torch.arange(op.oq).view(1, 1, -1)
    if op.out_map is None:
        os0, osn, osp, osq = 0, op.op * op.oq, op.oq, 1
    else:
        os0, osn, osp, osq = op.out_map
    return (os0 + img * osn + p * osp + q * osq).reshape(-1)


def _mask_bits(mask_words: Tensor, n: int) -> Tensor:
    """[M, words] int32 -> [M, n] bool"""
    w = mask_words.to(torch.int64) & 0xFFFFFFFF
    bits = (w.unsqueeze(-1) >> torch.arange(32)) & 1
    return bits.reshape(w.shape[0], -1)[:, :n].bool()


def _pack_mask(pos: Tensor) -> Tensor:
    """[M, n] bool -> [M, ceil(n/32)] int32"""
    M, n = pos.shape
    words = (n + 31) // 32
    padded = torch.zeros(M, words * 32, dtype=torch.int64)
    padded[:, :n] = pos.to(torch.int64)
    v = (padded.view(M, words, 32) << torch.arange(32)).sum(-1)
    v = torch.where(v >= 2**31, v - 2**32, v)
    return v.to(torch.int32)


def run_igemm(op: O.IgemmOp) -> None:
    M, n = op.M, op.n
    assert op.ktot % 64 == 0, "chunk count must fill whole pipeline stages"
    A = gather_a(op)
    D = A @ op.b.float().t()                      # [M, n]
    rows = _out_rows(op)
    if op.mode == L.BCOSK_MODE_FWD:
        if op.lin_bias is not None:
            D = D + op.lin_bias.float()
        inv_norm = op.inv_norm
        if inv_norm is None and op.sq_in is not None:      # in-kernel patch norm from the producer's sums of squares
            sh, sw, k, st, pd = op.sq_geom
            nb = op.a.shape[0]
            sq = op.sq_in.view(op.sq_in.shape[0], nb, 1, sh, sw).sum(0)
            sp = torch.nn.functional.avg_pool2d(sq, k, stride=st, padding=pd, divisor_override=1)
            assert sp.shape[-2:] == (op.op, op.oq), (op.name, sp.shape, op.op, op.oq)
            inv_norm = (1.0 / ((sp + op.sq_eps[0]).sqrt() + op.sq_eps[1])).reshape(-1)
        if op.max_out > 1:     # adjacent-column MaxOut: keep the largest unit of each group (first on ties), scale it
            G = op.max_out
            best, idx = D.view(M, n // G, G).max(dim=2)
            # torch.max may return any index on exact ties; the kernel keeps the first
            idx = (D.view(M, n // G, G) == best[..., None]).float().argmax(dim=2)
            D, n = best, n // G
            if op.amax is not None:
                op.amax.copy_(idx.to(torch.uint8))
        alpha = op.alpha.float() if op.alpha is not None else torch.ones(n)
        beta = op.beta.float() if op.beta is not None else torch.zeros(n)
        if op.scale_mode == L.BCOSK_SCALE_B2:
            t = D.abs() * inv_norm.float()[:, None] * alpha
        elif op.scale_mode == L.BCOSK_SCALE_POW:
            t = (D.abs() * inv_norm.float()[:, None] + 1e-6).pow(op.b_exp - 1.0) * alpha
        else:
            t = alpha.expand(M, n).clone()
        v = D * t + beta
        if op.res is not None:
            v = v + _join(op.res.reshape(M, -1), op.res_planes)
        if op.act == 1:        # MyGELU behind the transform: gate on the fp32 value, folded into y and the gain before rounding
            gate = 0.5 * (1.0 + torch.erf(v * 0.70710678118654752440))
            v, t = v * gate, t * gate
        elif op.act == 2:      # QuickGELU (not detached): the gain takes the derivative
            sg = 1.0 / (1.0 + torch.exp(-1.702 * v))
            v, t = v * sg, t * (sg + 1.702 * v * sg * (1.0 - sg))
        # throughput path (single 16-bit plane): ReLU / mask are decided on the ROUNDED value (packed 16-bit compare) and
        # the sums of squares use the un-rounded fp32 value; the generic path decides on fp32 and squares what it stored
        fast = (not op.y_f32) and op.y_planes == 1 and (op.gain is None or op.gain.dtype != torch.float32) \
            and op.res_planes == 1 and op.scale_mode in (L.BCOSK_SCALE_B2, L.BCOSK_SCALE_NONE) and op.lin_bias is None
        pos = (v.to(op.y.dtype).float() > 0) if fast else (v > 0)
        if op.relu:
            v = torch.where(pos, v, torch.zeros_like(v))
            t = torch.where(pos, t, torch.zeros_like(t))
        if op.maskbits is not None:
            op.maskbits.copy_(_pack_mask(pos if op.relu else torch.ones_like(pos)))
        if op.gain is not None:
            op.gain.copy_(t.to(op.gain.dtype))
        if op.inv_norm_out is not None:
            op.inv_norm_out.copy_(inv_norm.float().reshape(-1))
        y2 = op.y.view(-1, op.y.shape[-1])
        if op.y_f32:
            y2[rows] = v
            stored = v
        elif op.y_col or op.y.shape[-1] != op.y_planes * v.shape[-1]:
            # the launch writes an n-column slice of a wider tensor (DenseNet block features): plane pl at pl * (ld / planes) + y_col
            tmp = torch.zeros(M, op.y_planes * v.shape[-1], dtype=op.y.dtype)
            stored = _split_store(tmp, v, op.y_planes)
            pst, nn = op.y.shape[-1] // op.y_planes, v.shape[-1]
            for pl in range(op.y_planes):
                y2[rows, pl * pst + op.y_col: pl * pst + op.y_col + nn] = tmp[:, pl * nn:(pl + 1) * nn]
        else:
            tmp = torch.zeros(M, op.y.shape[-1], dtype=op.y.dtype)
            stored = _split_store(tmp, v, op.y_planes)
            # (fast path: sums of squares of the stored, i.e. rounded, values - same as the generic path)
            y2[rows] = tmp
        if op.sq_out is not None:
            bn = op.resolved_block_n() // max(op.max_out, 1)     # stored columns per n tile
            for ti in range(op.sq_out.shape[0]):
                op.sq_out[ti] = (stored[:, ti * bn:(ti + 1) * bn] ** 2).sum(1)
    else:
        tot = D
        if op.add is not None:
            nb = op.a.shape[0]
            addv = _join(op.add, op.add_planes)                      # [nb, ap, aq, n]
            s = op.add_stride
            full = torch.zeros(nb, op.op, op.oq, n)
            ap, aq = op.add.shape[1], op.add.shape[2]
            hh = min(op.op, (ap - 1) * s + 1)
            ww = min(op.oq, (aq - 1) * s + 1)
            full[:, 0:hh:s, 0:ww:s] = addv[:, :(hh + s - 1) // s, :(ww + s - 1) // s]
            tot = tot + full.reshape(M, n)
        side = rows if op.side_mapped else torch.arange(M)       # side tensors: mapped output row or dense launch row
        if op.out2 is not None:
            o = tot.clone()
            if op.mul2 is not None:
                o = o * op.mul2.float()[side]
            if op.mask2 is not None:
                o = o * _mask_bits(op.mask2, n)[side]
            o2 = op.out2.view(-1, op.out2.shape[-1])
            tmp2 = torch.zeros(M, op.out2.shape[-1], dtype=op.out2.dtype)
            _split_store(tmp2, o, op.out2_planes)
            o2[side] = tmp2
        if op.mul1 is not None and op.mul1_sqrt_scale is not None:      # gain recomputed from the producer's ReLU output
            gsc = (op.mul1.float().view(-1, op.mul1.shape[-1])[side] * op.mul1_sqrt_scale.float().view(-1)[side][:, None]).sqrt()
            v = tot * gsc
        else:
            v = tot * op.mul1.float().view(-1, op.mul1.shape[-1])[side] if op.mul1 is not None else tot
        y2 = op.y.view(-1, op.y.shape[-1])
        if op.y_f32:
            y2[rows] = v
        else:
            tmp = torch.zeros(M, op.y.shape[-1], dtype=op.y.dtype)
            _split_store(tmp, v, op.y_planes)
            y2[rows] = tmp


def _x6(x: Tensor) -> Tensor:
    if x.dtype == torch.uint8:          # RGB uint8 -> [x, 1-x] (AddInverse)
        v = x.float() / 255.0
        return torch.cat([v, 1.0 - v], 1)
    return x.float()


def run_input_prep(op: O.InputPrepOp) -> None:
    x = _x6(op.x)
    nb, _, h, w = x.shape
    mean = torch.tensor(op.mean6).view(1, 6, 1, 1)
    istd = torch.tensor(op.inv_std6).view(1, 6, 1, 1)
    xn = (x - mean) * istd                                          # [nb,6,h,w]
    # channel (dy*2+dx)*6 + c
    s2d = xn.view(nb, 6, h // 2, 2, w // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(nb, h // 2, w // 2, 24)
    val = torch.zeros(nb, h // 2, w // 2, op.cp)
    val[..., :24] = s2d
    stored = _split_store(op.out, val, op.planes)
    if op.sq is not None:
        st = stored[..., :24].view(nb, h // 2, w // 2, 2, 2, 6).permute(0, 5, 1, 3, 2, 4).reshape(nb, 6, h, w)
        op.sq.view(-1)[:] = (st ** 2).sum(1).reshape(-1)


def run_patch_norm(op: O.PatchNormOp) -> None:
    sq = op.sq.view(op.parts, op.nb, 1, op.h, op.w).sum(0)
    s = torch.nn.functional.avg_pool2d(sq, op.k, stride=op.stride, padding=op.pad, divisor_override=1)
    assert s.shape[-2:] == (op.op, op.oq), (op.name, s.shape, op.op, op.oq)
    op.inv_norm[:] = (1.0 / ((s + op.eps_in).sqrt() + op.eps_out)).reshape(-1)


def run_avgpool_fwd(op: O.AvgPoolFwdOp) -> None:
    x = _join(op.x, op.planes).permute(0, 3, 1, 2)
    y = torch.nn.functional.avg_pool2d(x, op.k, stride=op.stride, padding=op.pad).permute(0, 2, 3, 1)
    stored = _split_store(op.y, y.contiguous(), op.planes)
    if op.sq is not None:
        op.sq.view(-1)[:] = (stored ** 2).sum(-1).reshape(-1)


def run_avgpool_bwd_mul(op: O.AvgPoolBwdMulOp) -> None:
    gy = _join(op.gy, op.planes).permute(0, 3, 1, 2).contiguous()
    nb, h, w, _ = op.gx.shape
    x = torch.zeros(nb, op.c, h, w, requires_grad=True)
    y = torch.nn.functional.avg_pool2d(x, op.k, stride=op.stride, padding=op.pad)
    (gx,) = torch.autograd.grad(y, x, gy)
    gx = gx.permute(0, 2, 3, 1)
    if op.gain is not None and op.gain_sqrt_scale is not None:
        gx = gx * (op.gain.float().view(nb, h, w, op.c) * op.gain_sqrt_scale.float().view(nb, h, w, 1)).sqrt()
    elif op.gain is not None:
        gx = gx * op.gain.float().view(nb, h, w, op.c)
    _split_store(op.gx, gx.contiguous(), op.planes)


def run_gap_logits(op: O.GapLogitsOp) -> None:
    fc = op.fc.view(op.nb, op.npix, op.ncls)
    lg = fc.mean(1) * op.inv_temp + op.bias
    op.logits.copy_(lg)
    op.pred.copy_(lg.argmax(1).to(torch.int32))


def run_fc_seed(op: O.FcSeedOp) -> None:
    rows = op.nb * op.npix
    tgt = op.target.long().repeat_interleave(op.npix)
    gf = op.gain_fc.float()[torch.arange(rows), tgt]
    g = (op.inv_temp * op.seed_scale / op.npix) * gf[:, None] * op.w_fc[tgt]
    if op.out2 is not None:
        o = g * _mask_bits(op.mask2, op.c) if op.mask2 is not None else g
        _split_store(op.out2, o, op.planes)
    v = g * op.mul1.float() if op.mul1 is not None else g
    _split_store(op.out1, v, op.planes)


def run_contrib_map(op: O.ContribMapOp) -> None:
    nb, _, h, w = op.x.shape
    g = op.g[..., :24].view(nb, h // 2, w // 2, 2, 2, 6).permute(0, 5, 1, 3, 2, 4).reshape(nb, 6, h, w)
    g6 = g * torch.tensor(op.inv_std6).view(1, 6, 1, 1) * op.out_scale
    op.cmap.copy_((_x6(op.x) * g6).sum(1))
    if op.grad6 is not None:
        op.grad6.copy_(g6)


def run_explanation_image(op: O.ExplanationImageOp) -> None:
    import bcos_oracle as R   # tests/conftest.py puts oracle/ on the path
    op.out.copy_(R.gradient_to_image_batched(_x6(op.x), op.grad6, op.smooth, op.percentile))


def run_trunk_out(op: O.TrunkOutOp) -> None:
    op.out.copy_(_join(op.y, op.planes)[..., :op.c].permute(0, 3, 1, 2))


def run_seed_from_nchw(op: O.SeedFromNchwOp) -> None:
    nb, c, h, w = op.g.shape
    g = op.g.float().permute(0, 2, 3, 1).reshape(nb * h * w, c) * op.seed_scale
    if op.out1 is not None:
        v = g * op.mul1.float().view(-1, c) if op.mul1 is not None else g
        _split_store(op.out1, v.view(nb, h, w, c), op.planes)
    if op.out2 is not None:
        o = g * op.mul2.float().view(-1, c) if op.mul2 is not None else g
        if op.mask2 is not None:
            o = o * _mask_bits(op.mask2, c)
        _split_store(op.out2, o.view(nb, h, w, c), op.planes)


# ---------------------------------------------------------------------------------------------------------------------
# fused SimpleViT plan (engine/vit.py)
# ---------------------------------------------------------------------------------------------------------------------
def run_vit_patchify(op: O.VitPatchifyOp) -> None:
    x = _x6(op.x)
    nb, _, h, w = x.shape
    p = op.p
    xn = (x - torch.tensor(op.mean6).view(1, 6, 1, 1)) * torch.tensor(op.inv_std6).view(1, 6, 1, 1)
    v = xn.view(nb, 6, h // p, p, w // p, p).permute(0, 2, 4, 3, 5, 1).reshape(nb, h // p, w // p, p * p * 6)
    stored = _split_store(op.out, v.contiguous(), op.planes)
    if op.sq is not None:
        op.sq.view(-1)[:] = (stored ** 2).sum(-1).reshape(-1)


def run_vit_contrib_map(op: O.VitContribMapOp) -> None:
    nb, _, h, w = op.x.shape
    p = op.p
    g = op.g.view(nb, h // p, w // p, p, p, 6).permute(0, 5, 1, 3, 2, 4).reshape(nb, 6, h, w)
    g6 = g * torch.tensor(op.inv_std6).view(1, 6, 1, 1) * op.out_scale
    op.cmap.copy_((_x6(op.x) * g6).sum(1))
    if op.grad6 is not None:
        op.grad6.copy_(g6)


def run_vit_ln_fwd(op: O.VitLnFwdOp) -> None:
    x = _join(op.x, op.planes)
    var, mean = torch.var_mean(x, dim=-1, unbiased=False, keepdim=True)
    rstd = 1.0 / (var + op.eps).sqrt()
    y = (x - mean) * rstd * op.w.float()
    stored = _split_store(op.y, y, op.out_planes or op.planes)
    op.rstd.copy_(rstd.reshape(-1))
    if op.sq is not None:
        op.sq.view(-1)[:] = (stored ** 2).sum(-1).reshape(-1)


def run_vit_ln_bwd(op: O.VitLnBwdOp) -> None:
    gw = op.g.float() * op.w.float()
    r = op.rstd.float().view(*gw.shape[:-1], 1)
    if op.x is not None:      # true backward: the normalised input is in the graph
        x = _join(op.x, op.x_planes).reshape(gw.shape)
        xh = (x - x.mean(-1, keepdim=True)) * r
        gx = r * (gw - gw.mean(-1, keepdim=True) - xh * (gw * xh).mean(-1, keepdim=True))
    else:
        gx = r * (gw - gw.mean(-1, keepdim=True))
    if op.G_in is not None:
        gx = gx + op.G_in
    if op.G_out is not None:
        op.G_out.copy_(gx)
    if op.ghat is not None:
        v = gx * op.gain.float().view(gx.shape) if op.gain is not None else gx
        _split_store(op.ghat, v, 1)


def run_vit_gelu_fwd(op: O.VitGeluFwdOp) -> None:
    u = _join(op.u, op.planes)
    if op.quick:
        sg = torch.sigmoid(1.702 * u)
        stored = _split_store(op.a, u * sg, op.planes)
        gate = sg + 1.702 * u * sg * (1.0 - sg)           # the true derivative goes into the gain
    else:
        gate = 0.5 * (1.0 + torch.erf(u / 2.0 ** 0.5))
        stored = _split_store(op.a, u * gate, op.planes)
    if op.sq is not None:
        op.sq.view(-1)[:] = (stored ** 2).sum(-1).reshape(-1)
    if op.gain is not None:
        op.gain.copy_((op.gain.float() * gate.reshape(op.gain.shape)).to(op.gain.dtype))


def run_vit_attention(op: O.VitAttentionOp) -> None:
    hd = op.heads * op.dh
    qkv = _join(op.qkv, op.planes).reshape(op.nb, op.n, 3, op.heads, op.dh)
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))                 # [nb, heads, n, dh]
    prob = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * op.scale, dim=-1)
    if not op.backward:
        o = torch.matmul(prob, v).transpose(1, 2).reshape(op.nb, op.n, hd)
        _split_store(op.out, o.reshape(op.out.shape[:-1] + (hd,)), op.planes)
    elif op.full_bwd:
        g = op.g.float().reshape(op.nb, op.n, op.heads, op.dh).transpose(1, 2)
        dv = torch.matmul(prob.transpose(-1, -2), g)
        dp = torch.matmul(g, v.transpose(-1, -2))
        dz = prob * (dp - (prob * dp).sum(-1, keepdim=True)) * op.scale
        dq, dk = torch.matmul(dz, k), torch.matmul(dz.transpose(-1, -2), q)
        full = torch.stack([t.transpose(1, 2).reshape(op.nb, op.n, hd) for t in (dq, dk, dv)], 2).reshape(op.nb, op.n, 3 * hd)
        _split_store(op.out, full.reshape(op.out.shape[:-1] + (3 * hd,)), 1)
    else:
        g = op.g.float().reshape(op.nb, op.n, op.heads, op.dh).transpose(1, 2)
        gv = torch.matmul(prob.transpose(-1, -2), g).transpose(1, 2).reshape(op.nb, op.n, hd)
        _split_store(op.out, gv.reshape(op.out.shape[:-1] + (hd,)), 1)


def run_pixel_sqsum(op: O.PixelSqsumOp) -> None:
    op.sq.view(-1)[:] = (_join(op.x, op.planes) ** 2).sum(-1).reshape(-1)


# ---------------------------------------------------------------------------------------------------------------------
# fused DenseNet plan (engine/densenet.py)
# ---------------------------------------------------------------------------------------------------------------------
def run_dense_bn_relu_fwd(op: O.DenseBnReluFwdOp) -> None:
    pst = op.x.shape[-1] // op.planes
    x = sum(op.x[..., pl * pst: pl * pst + op.c].float() for pl in range(op.planes))
    t = x * op.alpha.float()
    pos = (t > 0) if op.relu else torch.ones_like(t, dtype=torch.bool)
    t = torch.where(pos, t, torch.zeros_like(t))
    stored = _split_store(op.y, t, op.planes)
    if op.sq is not None:
        op.sq.view(-1)[:] = (stored ** 2).sum(-1).reshape(-1)
    if op.maskbits is not None:
        op.maskbits.copy_(_pack_mask(pos.reshape(-1, op.c)))


def run_dense_bn_relu_bwd(op: O.DenseBnReluBwdOp) -> None:
    v = op.g.float() * op.alpha.float()
    if op.maskbits is not None:
        v = v * _mask_bits(op.maskbits, op.c).reshape(v.shape)
    if op.accumulate:
        op.G[..., :op.c] += v
    else:
        op.G[..., :op.c] = v


def run_dense_slice_cast(op: O.DenseSliceCastOp) -> None:
    v = op.G[..., op.col0: op.col0 + op.c].float()
    if op.gain is not None:
        v = v * op.gain.float().reshape(v.shape)
    op.out.copy_((v * op.scale).to(op.out.dtype))


def run_copy_channels(op: O.CopyChannelsOp) -> None:
    pst = op.dst.shape[-1] // op.planes
    for pl in range(op.planes):
        op.dst[..., pl * pst + op.dst_col: pl * pst + op.dst_col + op.c] = op.src[..., pl * op.c:(pl + 1) * op.c]


# ---------------------------------------------------------------------------------------------------------------------
# attention-pool head inside the fused CLIP plan
# ---------------------------------------------------------------------------------------------------------------------
def run_sgemm(op: O.SgemmOp) -> None:
    def view(t, off, rows, cols, ld, stride, trans):
        r, c = (cols, rows) if trans else (rows, cols)             # stored shape
        v = torch.as_strided(t.reshape(-1), (op.batch, r, c), (stride, ld, 1), off)
        return v.transpose(1, 2) if trans else v
    a = view(op.a, op.a_off, op.m, op.k, op.lda, op.stride_a, op.trans_a)
    b = view(op.b, op.b_off, op.k, op.n, op.ldb, op.stride_b, op.trans_b)
    c = torch.as_strided(op.c.reshape(-1), (op.batch, op.m, op.n), (op.stride_c, op.ldc, 1), op.c_off)
    c.copy_(op.alpha * torch.matmul(a, b))


def run_head_tokens(op: O.HeadTokensOp) -> None:
    nb, h, w, _ = op.x.shape
    x = _join(op.x, op.planes).reshape(nb, h * w, op.c)
    op.tokens[:, 1:] = x
    op.tokens[:, 0] = x.sum(1) / (h * w)


def run_row_softmax(op: O.RowSoftmaxOp) -> None:
    op.s.copy_(torch.softmax(op.s, dim=-1))


def run_seed_from_tokens(op: O.SeedFromTokensOp) -> None:
    nb, t, c = op.g_tokens.shape
    g = ((op.g_tokens[:, 1:] + op.g_tokens[:, :1] / (t - 1)) * op.scale).reshape(nb * (t - 1), c)
    if op.out1 is not None:
        v = g * op.mul1.float().view(-1, c) if op.mul1 is not None else g
        _split_store(op.out1, v.view(op.out1.shape[0], op.out1.shape[1], op.out1.shape[2], c), op.planes)
    if op.out2 is not None:
        o = g * op.mul2.float().view(-1, c) if op.mul2 is not None else g
        if op.mask2 is not None:
            o = o * _mask_bits(op.mask2, c)
        _split_store(op.out2, o.view(op.out2.shape[0], op.out2.shape[1], op.out2.shape[2], c), op.planes)


def run_stem_im2col(op: O.StemIm2colOp) -> None:
    nb, _, h, w = op.x.shape
    k, st, pd = op.k, op.stride, op.pad
    v = op.x.float()
    v4 = torch.cat([v * op.a_scale, torch.ones(nb, 1, h, w)], 1)                                  # [nb, 4, h, w]
    cols = torch.nn.functional.unfold(v4, k, padding=pd, stride=st)                               # [nb, 4*k*k, L] channel-major (c, tap)
    L_ = cols.shape[-1]
    cols = cols.view(nb, 4, k * k, L_).permute(0, 3, 2, 1).reshape(nb, L_, k * k * 4)             # (tap, c)
    out = torch.zeros(nb * L_, op.out.shape[-1])
    out[:, :k * k * 4] = cols.reshape(nb * L_, -1)
    op.out.copy_(out.view(op.out.shape).to(op.out.dtype))
    x6 = _x6(op.x)
    xn = (x6 - torch.tensor(op.mean6).view(1, 6, 1, 1)) * torch.tensor(op.inv_std6).view(1, 6, 1, 1)
    sq = torch.nn.functional.avg_pool2d((xn ** 2).sum(1, keepdim=True), k, stride=st, padding=pd, divisor_override=1)
    op.inv_norm.copy_((1.0 / (sq + 1e-6).sqrt()).reshape(-1))


_DISPATCH = {
    O.IgemmOp: run_igemm, O.InputPrepOp: run_input_prep, O.PatchNormOp: run_patch_norm, O.AvgPoolFwdOp: run_avgpool_fwd,
    O.AvgPoolBwdMulOp: run_avgpool_bwd_mul, O.GapLogitsOp: run_gap_logits, O.FcSeedOp: run_fc_seed,
    O.ContribMapOp: run_contrib_map, O.ExplanationImageOp: run_explanation_image,
    O.TrunkOutOp: run_trunk_out, O.SeedFromNchwOp: run_seed_from_nchw,
    O.VitPatchifyOp: run_vit_patchify, O.VitContribMapOp: run_vit_contrib_map, O.VitLnFwdOp: run_vit_ln_fwd,
    O.VitLnBwdOp: run_vit_ln_bwd, O.VitGeluFwdOp: run_vit_gelu_fwd, O.VitAttentionOp: run_vit_attention,
    O.PixelSqsumOp: run_pixel_sqsum,
    O.DenseBnReluFwdOp: run_dense_bn_relu_fwd, O.DenseBnReluBwdOp: run_dense_bn_relu_bwd, O.DenseSliceCastOp: run_dense_slice_cast,
    O.CopyChannelsOp: run_copy_channels,
    O.StemIm2colOp: run_stem_im2col, O.SgemmOp: run_sgemm, O.HeadTokensOp: run_head_tokens, O.RowSoftmaxOp: run_row_softmax, O.SeedFromTokensOp: run_seed_from_tokens,
}


def run(ops) -> None:
    with torch.no_grad():
        for o in ops:
            fn = _DISPATCH[type(o)]
            if fn is run_avgpool_bwd_mul:
                with torch.enable_grad():
                    fn(o)
            else:
                fn(o)
