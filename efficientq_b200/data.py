"""Calibration-batch assembly (reference src/ptqer.py:83-111).

Only what the hot path consumes is restated: take ``lwq_batchsz`` volumes of the training
split starting at ``lwq_dataid`` (sequential order, fixed transform = ToTensor + optional
mean/std normalisation, src/dataloader/datahub.py:75-86) and centre-crop / zero-pad each
to ``lwq_patchsz`` (src/dataloader/transforms.py:60-81).  ``data_dir: synthetic`` (or no
data_dir) produces the deterministic synthetic volumes of ``synth.py`` instead.  With
several ranks every rank assembles only its own contiguous shard of the batch.
"""
from __future__ import annotations

import os.path as P
from typing import Optional, Tuple

import numpy as np
import torch

from . import synth
from .dist import DistCtx

MODALITIES = {"brats": ("seg", "flair", "t1", "t1ce", "t2"), "lits": ("seg", "ct")}


def center_crop(t: torch.Tensor, size) -> torch.Tensor:
    """Crop (or symmetrically zero-pad) the last three dims to ``size`` (transforms.py:60-81)."""
    for dim, target in zip((-1, -2, -3), (size[2], size[1], size[0])):
        cur = t.shape[dim]
        if cur < target:
            before = (target - cur) // 2
            pad = [0, 0] * 3
            idx = {-1: 0, -2: 2, -3: 4}[dim]
            pad[idx], pad[idx + 1] = before, target - cur - before
            t = torch.nn.functional.pad(t, pad)
    d, h, w = t.shape[-3:]
    x1, y1, z1 = (d - size[0]) // 2, (h - size[1]) // 2, (w - size[2]) // 2
    return t[..., x1:x1 + size[0], y1:y1 + size[1], z1:z1 + size[2]]


SYNTH_VAL_VOLUMES = 2          # held-out synthetic volumes evaluated when data_dir is synthetic
SYNTH_VAL_FIRST_SEED = 5000


class CalibrationData:
    def __init__(self, args):
        self.args = args
        self.task = args.task.lower()
        self.synthetic = (not getattr(args, "data_dir", None)) or str(args.data_dir).lower() == "synthetic"
        self.valloader = self.testloader = None

    def _crop_shape(self, default):
        a = self.args
        if getattr(a, "lwq_patchsz", None):
            return [int(x) for x in str(a.lwq_patchsz).split(",")]
        return [min(x, 192) // 64 * 64 for x in default]

    def _load(self, sn: str):
        a = self.args
        mods = MODALITIES[self.task]
        ext = getattr(a, "access_type", "npz") or "npz"

        def read(mod, dtype):
            f = P.join(a.data_dir, mod, f"{sn}.{ext}")
            arr = np.load(f, allow_pickle=True)
            arr = arr["arr_0"] if ext == "npz" else arr
            return arr.astype(dtype, copy=False)
        img = torch.from_numpy(np.stack([read(m, "float32") for m in mods[1:]]))
        lab = torch.from_numpy(read(mods[0], "uint8"))
        ms = P.join(a.data_dir, "meanstd.txt")
        if P.exists(ms):
            lines = open(ms).read().splitlines()
            mean = torch.tensor([float(x) for x in lines[0].split()[1:]]).view(-1, 1, 1, 1)
            std = torch.tensor([float(x) for x in lines[1].split()[1:]]).view(-1, 1, 1, 1)
            img = (img - mean) / std
        return img, lab

    def calibration_batch(self, args=None, dist: Optional[DistCtx] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        a = args or self.args
        dist = dist or DistCtx()
        n_total = int(getattr(a, "lwq_batchsz", 1) or 1)
        first = int(getattr(a, "lwq_dataid", 0) or 0)
        if n_total < dist.world:
            raise ValueError(f"lwq_batchsz = {n_total} calibration volume(s) cannot be sharded over {dist.world} ranks: "
                             f"every rank needs at least one volume (raise lwq_batchsz or launch fewer processes)")
        lo, hi = dist.shard(n_total)
        n_mod = int(getattr(a, "nMod", None) or (4 if self.task == "brats" else 1))
        if self.synthetic:
            shape = self._crop_shape((128, 128, 128) if self.task == "brats" else (160, 160, 64))
            return synth.batch(hi - lo, first + lo, n_mod, tuple(shape), self.task, pin=torch.cuda.is_available()), None
        split = P.join(a.split_dir, "round" + str(a.round), "train.txt")
        sns = open(split).read().splitlines()
        if not getattr(a, "data_on_disk", False):
            sns.sort()                                   # Dataset_SEG sorts, the on-disk variant does not
        imgs, labs = [], []
        for sn in sns[first + lo:first + hi]:
            img, lab = self._load(sn)
            shape = self._crop_shape(img.shape[-3:])
            imgs.append(center_crop(img, shape))
            labs.append(center_crop(lab, shape))
        return torch.stack(imgs).float(), torch.stack(labs)

    def slide_window(self):
        """(patch_size, overlap) of the sliding-window evaluation (src/definer.py:41-62, :78-83, :106-107)."""
        a = self.args
        ps = getattr(a, "patch_size", None)
        if ps:
            ps = tuple(int(x) for x in str(ps).split(",")) if "," in str(ps) else (int(ps),) * 3
        else:
            ps = (128, 128, 128) if self.task == "brats" else (128, 128, 64)
        return ps, (16, 16, 16)

    def evaluation_volumes(self, split: str = "val", rank: int = 0, world: int = 1):
        """Iterable of ``(sn, image[C,D,H,W] fp32, label[D,H,W])`` of a split with the fixed transform of the
        reference's val / test loaders (ToTensor + optional normalisation, no crop: src/dataloader/datahub.py:75-110),
        or None when the split does not exist.  Synthetic data: ``SYNTH_VAL_VOLUMES`` held-out volumes as the
        'val' split, one quarter-window longer than the window along D so that the stitching is exercised.
        ``rank`` / ``world``: this rank's share of the split, volumes dealt round-robin (volume i -> rank i % world);
        the other volumes are not read."""
        a = self.args
        n_mod = int(getattr(a, "nMod", None) or (4 if self.task == "brats" else 1))
        if self.synthetic:
            if split != "val":
                return None
            ps, _ = self.slide_window()
            shape = (ps[0] + max(ps[0] // 4, 1), ps[1], ps[2])

            def synth_volumes():
                for i in range(rank, SYNTH_VAL_VOLUMES, world):
                    img, lab = synth.volume(SYNTH_VAL_FIRST_SEED + i, n_mod, shape, self.task)
                    yield f"synthetic_{SYNTH_VAL_FIRST_SEED + i}", img, lab
            return synth_volumes()
        if not getattr(a, "split_dir", None):
            return None
        path = P.join(a.split_dir, "round" + str(a.round), f"{split}.txt")
        if not P.isfile(path):
            return None
        sns = [s for s in open(path).read().splitlines() if s]
        if not sns:
            return None
        if not getattr(a, "data_on_disk", False):
            sns.sort()

        def disk_volumes():
            for sn in sns[rank::world]:
                img, lab = self._load(sn)
                yield sn, img.float(), lab
        return disk_volumes()
