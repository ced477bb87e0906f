"""Synthetic workload: images and random-init "checkpoints", bit-reproducible on any host.

No dataset or checkpoint can be downloaded, so every benchmark/parity input is synthesised
(SURVEY.md section 8d).  Everything here is plain numpy with element-wise float64 arithmetic and
MT19937 streams, so the GPU box reproduces exactly what the golden generator used.

* images   : smooth random fields (bilinear up-sampling of a coarse U(0,1) grid), quantised to
             uint8 like decoded JPEGs; `to_bcos_input` appends the inverse channels
             (`AddInverse`, reference bcos/data/transforms.py:42-55 -> cat([x, 1-x], 1)).
* weights  : one independent MT19937 stream per state-dict key (seeded by crc32 of the key), so a
             tensor depends only on its name and shape - the same dictionary loads into the
             reference model, the oracle and the B200 engine (`*.linear.weight`, BN buffers ...).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def synth_images_u8(batch: int, size: int = 224, seed: int = 0, grid: int = 14) -> np.ndarray:
    """uint8 [batch, 3, size, size] smooth images."""
    rs = np.random.RandomState(1000003 * (seed + 1) % (2**31 - 1))
    coarse = rs.random_sample((batch, 3, grid, grid))  # float64 U(0,1)
    # bilinear, half-pixel centres, clamp-to-edge; explicit gather + lerp (no BLAS) => exact everywhere
    pos = (np.arange(size, dtype=np.float64) + 0.5) * (grid / size) - 0.5
    pos = np.clip(pos, 0.0, grid - 1.0)
    i0 = np.floor(pos).astype(np.int64)
    i1 = np.minimum(i0 + 1, grid - 1)
    f = pos - i0
    rows = coarse[:, :, i0, :] * (1.0 - f)[None, None, :, None] + coarse[:, :, i1, :] * f[None, None, :, None]
    img = rows[:, :, :, i0] * (1.0 - f)[None, None, None, :] + rows[:, :, :, i1] * f[None, None, None, :]
    # contrast stretch so images use most of [0,1], then quantise
    img = np.clip((img - 0.5) * 1.6 + 0.5, 0.0, 1.0)
    return np.floor(img * 255.0 + 0.5).astype(np.uint8)


def to_bcos_input(images_u8) -> torch.Tensor:
    """uint8 [B,3,H,W] -> float32 [B,6,H,W] = cat([x, 1-x]) with x = u8/255."""
    x = torch.as_tensor(np.asarray(images_u8)).to(torch.float32) / 255.0
    return torch.cat([x, 1.0 - x], dim=1).contiguous()


def _stream(name: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761 & 0xFFFFFFFF)) & 0x7FFFFFFF)


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """Deterministic value for one state-dict entry, chosen by the key's suffix."""
    rs = _stream(name, seed)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.int64)
    if leaf == "running_mean":
        return torch.zeros(shape, dtype=torch.float32)
    if leaf == "running_var":
        return torch.ones(shape, dtype=torch.float32)
    if len(shape) == 1:
        if leaf == "bias":
            return torch.zeros(shape, dtype=torch.float32)
        # norm-layer scale: positive, spread around 1
        return torch.from_numpy((0.5 + rs.random_sample(shape)).astype(np.float32))
    # conv / linear weight: N(0, 2/fan_in)
    fan_in = int(np.prod(shape[1:]))
    w = rs.standard_normal(shape) * np.sqrt(2.0 / fan_in)
    return torch.from_numpy(w.astype(np.float32))


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]] | Iterable, seed: int = 0) -> Dict[str, torch.Tensor]:
    """`shapes`: {key: shape} (e.g. from `{k: tuple(v.shape) for k, v in model.state_dict().items()}`)."""
    if not isinstance(shapes, dict):
        shapes = dict(shapes)
    return {k: synth_tensor(k, tuple(s), seed) for k, s in shapes.items()}


def synthetic_checkpoint(arch: str, shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic weights (seed 0) + the shipped BN calibration (utils/calib/<arch>_bnvar.npz, written by
    oracle/make_golden.py --only calib) -> a state dict with the reference's key names."""
    import os
    sd = synth_state_dict(shapes, seed)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "calib", f"{arch}_bnvar.npz")
    z = np.load(path)
    assert int(z["weights_seed"]) == seed, "calibration file belongs to another weight seed"
    off = 0
    for k, n in zip(z["bn_keys"].tolist(), z["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(z["bn_var"][off:off + n].copy())
        off += n
    return sd
