"""Deterministic synthetic calibration volumes (no dataset, no network).

Shape and value conventions follow what the reference's PTQ path consumes
(SURVEY.md section 8(d)): BraTS -- 4 fp32 modalities with the background exactly
0.0 because ``body_mask = data_batch[:,0] != 0`` (reference src/ptqer.py:338);
LiTS -- 1 fp32 CT channel, body mask all ones (ptqer.py:340).  Labels are uint8
({0,1,2,3} BraTS, {0,1,2} LiTS) and only matter to the evaluation side.

numpy ``Generator(PCG64)`` streams are platform-stable, so ``volume(seed=k)``
is bit-identical in the build container and on the GPU box.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["volume", "batch"]


def _grid(shape):
    d, h, w = shape
    z, y, x = np.meshgrid(np.linspace(-1, 1, d, dtype=np.float32),
                          np.linspace(-1, 1, h, dtype=np.float32),
                          np.linspace(-1, 1, w, dtype=np.float32), indexing="ij")
    return z, y, x


def volume(seed: int, n_mod: int = 4, shape=(128, 128, 128), task: str = "brats"):
    """One (image[n_mod,D,H,W] fp32, label[D,H,W] uint8) pair."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    z, y, x = _grid(shape)
    img = rng.standard_normal((n_mod,) + tuple(shape), dtype=np.float32)
    label = np.zeros(shape, dtype=np.uint8)
    n_lab = 3 if task == "brats" else 2
    # nested blobs: label k lives inside label k-1's sphere
    c = rng.uniform(-0.35, 0.35, size=3).astype(np.float32)
    r = np.float32(rng.uniform(0.35, 0.5))
    for k in range(1, n_lab + 1):
        blob = (z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 < r * r
        label[blob] = k
        img[:, blob] += np.float32(0.8 * k) * rng.uniform(0.5, 1.5, size=(n_mod, 1)).astype(np.float32)
        r = np.float32(r * 0.6)
    if task == "brats":
        ax = rng.uniform(0.75, 0.95, size=3).astype(np.float32)
        body = (z / ax[0]) ** 2 + (y / ax[1]) ** 2 + (x / ax[2]) ** 2 < 1.0
        img *= body[None].astype(np.float32)      # background exactly 0.0
        label[~body] = 0
    return torch.from_numpy(img), torch.from_numpy(label)


def batch(n: int, first_seed: int = 0, n_mod: int = 4, shape=(128, 128, 128), task: str = "brats",
          pin: bool = False):
    """(N, n_mod, D, H, W) fp32 calibration batch, volumes ``first_seed .. first_seed+n-1``."""
    out = torch.empty((n, n_mod) + tuple(shape), dtype=torch.float32, pin_memory=pin)
    for i in range(n):
        out[i] = volume(first_seed + i, n_mod, shape, task)[0]
    return out
