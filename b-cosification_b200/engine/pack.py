"""Weight packing for the implicit GEMM (`bcosk_igemm`): K-major [n][segment][tap][channel chunk].

Precision planes.  A tensor with P planes stores value = sum_p plane_p (each plane 16-bit).  A GEMM on
split operands is evaluated as a sum of plane products, dropping the terms below ~2^-(8P) relative:
    P=1: a0*b0          P=2: a0*b0 + a0*b1 + a1*b0          P=3: + a0*b2 + a2*b0 + a1*b1
Each product is one K "segment": the kernel walks (segment, tap, channel-chunk) and only needs the A
plane's channel offset per segment; the B planes are laid out here in the same order.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor

SEGMENTS = {
    1: [(0, 0)],
    2: [(0, 0), (0, 1), (1, 0)],
    3: [(0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1)],
}


def split_planes(x: Tensor, planes: int, dtype: torch.dtype) -> List[Tensor]:
    """fp32 -> `planes` 16-bit tensors whose (fp32) sum approximates x to ~8*planes mantissa bits (bf16)."""
    r = x.to(torch.float32)
    out = []
    for _ in range(planes):
        h = r.to(dtype)
        out.append(h)
        r = r - h.to(torch.float32)
    return out


def join_planes(t: Tensor, planes: int) -> Tensor:
    """[..., planes*C] 16-bit -> [..., C] fp32 (sum of planes)."""
    c = t.shape[-1] // planes
    acc = t[..., :c].to(torch.float32)
    for p in range(1, planes):
        acc = acc + t[..., p * c:(p + 1) * c].to(torch.float32)
    return acc


def segments(a_planes: int, b_planes: int) -> List[Tuple[int, int]]:
    """Plane products kept for operands with a_planes x b_planes precision planes: the canonical list of max(a, b) planes restricted
    to the planes that exist (1 x 1: a0 b0; 2 x 1: a0 b0 + a1 b0; 1 x 2: a0 b0 + a0 b1; 2 x 2: a0 b0 + a0 b1 + a1 b0; ...)."""
    return [(a, b) for a, b in SEGMENTS[max(a_planes, b_planes)] if a < a_planes and b < b_planes]


def pack_b(wt: Tensor, planes: int, kch: int, dtype: torch.dtype, segs: Sequence[Tuple[int, int]] = None) -> Tuple[Tensor, int]:
    """wt [n, taps, c] fp32 (one [n, c] matrix per tap) -> (B [n, segs*taps*cpt*kch], chunks_per_tap).  `planes` = precision planes
    of the weights; `segs` (default SEGMENTS[planes]) = the (a plane, b plane) products the launch walks."""
    n, taps, c = wt.shape
    cpt = (c + kch - 1) // kch
    pl = split_planes(wt, planes, dtype)
    segs = SEGMENTS[planes] if segs is None else list(segs)
    out = torch.zeros(n, len(segs), taps, cpt * kch, dtype=dtype, device=wt.device)
    for s, (_, bp) in enumerate(segs):
        out[:, s, :, :c] = pl[bp]
    return out.reshape(n, len(segs) * taps * cpt * kch).contiguous(), cpt


def seg_a_offsets(planes: int, a_plane_channels: int) -> List[int]:
    return [ap * a_plane_channels for ap, _ in SEGMENTS[planes]]


def seg_b_planes(planes: int) -> List[int]:
    return [bp for _, bp in SEGMENTS[planes]]


def conv_taps(kh: int, kw: int, dil: int = 1) -> List[Tuple[int, int]]:
    """Tap order (kh-major) as (off_w, off_h)."""
    return [(j * dil, i * dil) for i in range(kh) for j in range(kw)]


def fwd_weight_taps(w: Tensor) -> Tensor:
    """W [o, c, kh, kw] -> [o, kh*kw, c] in conv_taps order (forward / fprop)."""
    o, c, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(o, kh * kw, c)


def dgrad_weight_taps(w: Tensor) -> Tensor:
    """W [o, c, kh, kw] -> [c, kh*kw, o]: tap (off_h, off_w) of the stride-1 data-gradient gather uses
    W[:, :, kh-1-off_h, kw-1-off_w]  (d[h] = sum_off g[h + pad-(k-1) + off] * W[k-1-off])."""
    o, c, kh, kw = w.shape
    wf = torch.flip(w, dims=(2, 3))
    return wf.permute(1, 2, 3, 0).reshape(c, kh * kw, o)


def stem_s2d_weight(w7: Tensor, cp: int) -> Tensor:
    """7x7/2 pad-3 stem on [x,1-x] (6 ch) == 4x4/1 conv (pad 2 low, 1 high) on the 2x2 space-to-depth input:
    W4[o, (dy*2+dx)*6 + c, tr, ts] = W7[o, c, 2*tr+dy-1, 2*ts+dx-1] (zero outside the 7x7 support).
    The same identity turns the 3x3/2 pad-1 stem of the CLIP ResNets (CLIP/clip/model.py:109) into a 2x2/1 conv
    (pad 1 low, 0 high): output pixel p reads rows 2p-1 .. 2p+1 = blocks p-1 (dy=1) and p (dy=0,1)."""
    o, c, kh, kw = w7.shape
    assert c == 6 and kh == kw and kh in (3, 7)
    nt = (kh + 1) // 2
    w4 = torch.zeros(o, cp, nt, nt, dtype=w7.dtype, device=w7.device)
    for dy in range(2):
        for dx in range(2):
            for tr in range(nt):
                i = 2 * tr + dy - 1
                if not 0 <= i < kh:
                    continue
                for ts in range(nt):
                    j = 2 * ts + dx - 1
                    if not 0 <= j < kw:
                        continue
                    base = (dy * 2 + dx) * 6
                    w4[:, base:base + 6, tr, ts] = w7[:, :, i, j]
    return w4


def stem_im2col_weight(w: Tensor, mean6: Sequence[float], inv_std6: Sequence[float], a_scale: float, kp: int) -> Tensor:
    """W [o, 6, k, k] (stem conv over the normalised [x, 1-x] input) -> [o, kp] for the patch matrix of bcosk_stem_im2col_u8:
    column tap*4 + c multiplies the raw byte v_c * a_scale, column tap*4 + 3 the in-image indicator.  With
    xn_c = v/255 * istd_c - mean_c * istd_c and xn_{c+3} = -v/255 * istd_{c+3} + (1 - mean_{c+3}) * istd_{c+3}:
        W'[tap, c] = (W[c] istd_c - W[c+3] istd_{c+3}) / (255 a_scale),   W'[tap, 3] = sum_c (-W[c] mean_c istd_c + W[c+3] (1 - mean_{c+3}) istd_{c+3}).
    Folded in fp64."""
    o, c, kh, kw = w.shape
    assert c == 6 and kp >= 4 * kh * kw
    w64 = w.to(torch.float64).permute(0, 2, 3, 1).reshape(o, kh * kw, 6)            # [o, tap, c]
    m = torch.tensor(list(mean6), dtype=torch.float64)
    s = torch.tensor(list(inv_std6), dtype=torch.float64)
    out = torch.zeros(o, kp, dtype=torch.float64)
    col = (w64[:, :, :3] * s[:3] - w64[:, :, 3:] * s[3:]) / (255.0 * a_scale)     # [o, tap, 3]
    one = (-w64[:, :, :3] * (m[:3] * s[:3]) + w64[:, :, 3:] * ((1.0 - m[3:]) * s[3:])).sum(-1)
    blk = torch.cat([col, one[..., None]], -1).reshape(o, kh * kw * 4)
    out[:, :kh * kw * 4] = blk
    return out.to(torch.float32)
