"""Sliding-window evaluation of the FP / quantized net.

The caller on the far side of the hot path: after ``do_ptq`` the reference runs ``tester.test_as_is('ptq')``
(src/ptqer.py:379-380 -> src/utils/trainer.py:272-304 -> ``validate_seg``, src/utils/validate.py:212-264), which
tiles every validation volume into overlapping patches (src/utils/transforms.py:784-809), runs the model on each
patch -- for a calibrated net that is the deployment forward of ``qconv.PTQConv`` on the tcgen05 code path --,
averages the overlapping predictions (transforms.py:811-851) and accumulates per-class Dice / accuracy /
sensitivity / specificity (validate.py:19-205, src/utils/metrics.py:21-48).

Same tiling rule, same summation order per voxel (stitched logits are bit-identical to the reference's on the same
patch predictions), same metric formulas and report format.  What differs is the data movement: the volume, the
running sum and the overlap count stay on the device, each patch prediction is added into its window only (the
reference adds a zero-padded full-volume tensor per patch), and the four metrics of all classes come from ONE
confusion count per volume (one ``bincount``) instead of ~20 full-volume reductions per class.  NIfTI export is
out of scope (no nibabel in this image).
"""
from __future__ import annotations

import os
import os.path as P
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

EPS = 1e-6
METRICS = ("acc", "dsc", "sens", "spec")            # validate.py:31 (ALL_METRIC without the lesion-count extras)


def _triple(v) -> Tuple[int, int, int]:
    return (int(v),) * 3 if isinstance(v, int) else tuple(int(x) for x in v)


def window_starts(extent: int, patch: int, overlap: int) -> List[int]:
    """Start offsets of the windows along one axis (transforms.py:793-797): a regular grid of stride
    ``patch - overlap`` over ``[0, extent - patch)`` plus one window flush with the far edge."""
    if patch > extent:
        raise ValueError(f"window {patch} exceeds the volume extent {extent}")
    if patch - overlap <= 0:
        raise ValueError(f"overlap {overlap} must be smaller than the window {patch}")
    return list(range(0, extent - patch, patch - overlap)) + [extent - patch]


def windows(shape, patch_size, overlap) -> List[Tuple[slice, slice, slice]]:
    """All windows of a D x H x W volume in the reference's order (d slowest, w fastest)."""
    p, o = _triple(patch_size), _triple(overlap)
    sd, sh, sw = (window_starts(e, pp, oo) for e, pp, oo in zip(shape, p, o))
    return [(slice(i, i + p[0]), slice(j, j + p[1]), slice(k, k + p[2])) for i in sd for j in sh for k in sw]


@torch.no_grad()
def sliding_window_forward(model, images: torch.Tensor, patch_size, overlap, device=None) -> torch.Tensor:
    """Overlap-averaged prediction of ``model`` over ``images`` (N x C x D x H x W).  Returns a tensor of shape
    ``pred.shape[:-3] + (D, H, W)`` where ``pred`` is what the model returns for one patch (for the nets of this
    repo: heads x N x classes x d x h x w, src/models/model_blk.py:207).  ``patch_size`` / ``overlap`` None: one
    forward over the whole volume (validate.py passes the volume through in that case)."""
    device = torch.device(device) if device is not None else images.device
    images = images.to(device, non_blocking=True)
    if patch_size is None or overlap is None:
        return model(images)
    total = count = None
    for win in windows(images.shape[-3:], patch_size, overlap):
        pred = model(images[(..., *win)].contiguous())
        if total is None:
            total = torch.zeros(pred.shape[:-3] + images.shape[-3:], dtype=pred.dtype, device=pred.device)
            count = torch.zeros(images.shape[-3:], dtype=torch.int32, device=pred.device)
        total[(..., *win)] += pred
        count[win] += 1
    return total / count


# ---------------------------------------------------------------------------------------------------------------
# labels (src/utils/misc.py:221-285)
# ---------------------------------------------------------------------------------------------------------------
def split_label_brats(label: torch.Tensor) -> torch.Tensor:
    """D x H x W labels {0..3} -> 3 x D x H x W binary maps: whole tumour, tumour core, enhancing tumour."""
    return torch.stack([label > 0, (label == 1) | (label == 3), label == 3]).float()


def split_label_lits(label: torch.Tensor) -> torch.Tensor:
    """D x H x W labels {0,1,2} -> 2 x D x H x W binary maps: liver, tumour."""
    return torch.stack([label > 0, label == 2]).float()


def merge_label_basic(pred: torch.Tensor, fusetype: str) -> torch.Tensor:
    """Make nested binary maps consistent: 'con' clears a map wherever an outer one is clear, 'agg' sets a map
    wherever an inner one is set."""
    kind = fusetype.lower()
    if kind in ("con", "conservative"):
        return torch.cumprod(pred, dim=0).to(pred.dtype)
    if kind in ("agg", "aggressive"):
        return (torch.flip(torch.cumsum(torch.flip(pred, (0,)), dim=0), (0,)) > 0).to(pred.dtype)
    raise RuntimeError("Unknown Multilabel Fusetype: %s" % fusetype)


def label_transform(multi_label: Optional[str]):
    if not multi_label:
        return None
    kind = multi_label.lower()
    if kind == "brats":
        return split_label_brats
    if kind == "lits":
        return split_label_lits
    raise RuntimeError("Unknown multi_label type: %s" % multi_label)


# ---------------------------------------------------------------------------------------------------------------
# metrics (src/utils/metrics.py:21-48, src/utils/validate.py:162-204)
# ---------------------------------------------------------------------------------------------------------------
def predict_mask(seg_out: torch.Tensor, label: torch.Tensor, fusetype: Optional[str] = None) -> torch.Tensor:
    """validate.py:169-175: per-channel sigmoid >= 0.5 when the label has one binary map per class (same rank as
    the logits), arg-max over the classes otherwise."""
    if seg_out.dim() == label.dim():
        assert seg_out.shape == label.shape, "pred shape should match label shape: pred %s vs label %s" % (
            tuple(seg_out.shape), tuple(label.shape))
        pred = (torch.sigmoid(seg_out) >= 0.5).int()
        return merge_label_basic(pred, fusetype) if fusetype else pred
    return torch.max(seg_out, dim=0)[1]


def confusion_counts(pred: torch.Tensor, label: torch.Tensor, n_class: int, per_channel: bool) -> torch.Tensor:
    """[n_class][4] int64 counts (tn, fn, fp, tp) of every class from one pass over the volume."""
    if per_channel:
        idx = (pred.reshape(n_class, -1).long() * 2 + (label.reshape(n_class, -1) != 0).long()) + \
            4 * torch.arange(n_class, device=pred.device).view(-1, 1)
        return torch.bincount(idx.reshape(-1), minlength=4 * n_class).view(n_class, 4)
    cm = torch.bincount(pred.reshape(-1).long() * n_class + label.reshape(-1).long(),
                        minlength=n_class * n_class).view(n_class, n_class)          # [pred][label]
    tp = cm.diagonal()
    fp = cm.sum(1) - tp
    fn = cm.sum(0) - tp
    tn = cm.sum() - tp - fp - fn
    return torch.stack([tn, fn, fp, tp], 1)


def metrics_from_counts(counts: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The reference's fp32 formulas on exact integer counts ([n_class][tn, fn, fp, tp]).  Identical to the
    reference whenever the counts are below 2^24 (its multi-label path sums 0/1 floats) and always on its
    integer path."""
    c = counts.cpu()
    tn, fn, fp, tp = (c[:, i].float() for i in range(4))
    n = c.sum(1).float()
    pred_pos, gt_pos = (c[:, 2] + c[:, 3]).float(), (c[:, 1] + c[:, 3]).float()
    gt_neg = (c[:, 0] + c[:, 2]).float()
    return {"acc": (c[:, 0] + c[:, 3]).float() / n,
            "dsc": (2 * tp + EPS) / (pred_pos + gt_pos + EPS),
            "sens": (tp + EPS) / (gt_pos + EPS),
            "spec": (tn + EPS) / (gt_neg + EPS)}


class SegMetric:
    """Per-subject, per-class metric table with the reference's report format (validate.py:19-160)."""

    def __init__(self, n_class: int, sn_list: Optional[Sequence[str]] = None):
        self.n_class = n_class
        self.sn_list = list(sn_list) if sn_list else []
        self.keys = [k for m in METRICS for k in [m] + [f"{m}/{i}" for i in range(n_class)]]
        self.buffer: Dict[str, List[torch.Tensor]] = {k: [] for k in self.keys}
        self.metric: Dict[str, float] = {k: 0 for k in self.keys}

    def __len__(self):
        return len(self.buffer[METRICS[0] + "/0"])

    def evaluate_append(self, seg_out: torch.Tensor, label: torch.Tensor, sn: Optional[str] = None,
                        multilabel_fusetype: Optional[str] = None) -> torch.Tensor:
        if sn is not None:
            self.sn_list.append(sn)
        per_channel = seg_out.dim() == label.dim()
        pred = predict_mask(seg_out, label, multilabel_fusetype)
        vals = metrics_from_counts(confusion_counts(pred, label, self.n_class, per_channel))
        for m in METRICS:
            v = vals[m]
            for i in range(self.n_class):
                self.buffer[f"{m}/{i}"].append(v[i])
            # mean over the classes; the background class only counts when every class has its own map
            self.buffer[m].append(v.mean() if per_channel else v[1:].mean())
        return pred

    def rows(self) -> List[Tuple[str, List[float]]]:
        """Per-subject rows (name, one fp32 value per key) -- what the ranks of a sharded evaluation exchange."""
        return [(sn, [float(self.buffer[k][i]) for k in self.keys]) for i, sn in enumerate(self.sn_list)]

    @classmethod
    def from_rows(cls, n_class: int, rows) -> "SegMetric":
        sm = cls(n_class)
        for sn, vals in rows:
            sm.sn_list.append(sn)
            for k, v in zip(sm.keys, vals):
                sm.buffer[k].append(torch.tensor(v, dtype=torch.float32))
        return sm

    def get_metric(self) -> Dict[str, float]:
        if len(self):
            for k in self.keys:
                self.metric[k] = float(torch.stack(self.buffer[k]).mean())
        return self.metric

    def lines(self, preline: Optional[str] = None, is_indiv: bool = False) -> List[str]:
        self.get_metric()
        out = [preline] if preline else []
        out.append(", ".join("%s = %.4f" % (k, v) for k, v in self.metric.items()))
        if is_indiv:
            out.append("|%20s|" % "SN" + "".join("%8s|" % k.upper() for k in self.keys))
            for i, sn in enumerate(self.sn_list):
                out.append("|%20s|" % sn + "".join("%8.4f|" % float(self.buffer[k][i]) for k in self.keys))
        return out

    def write_metric(self, fid, preline: Optional[str] = None, is_indiv: bool = False) -> None:
        fid.write("\n".join(self.lines(preline, is_indiv)) + "\n")

    def print_metric(self, preword: Optional[str] = None) -> None:
        print("%s Segmentation Metrics:" % preword if preword else "Segmentation Metrics:")
        self.get_metric()
        print(",\n".join(", ".join("%s = %.4f" % (k, self.metric[k]) for k in [m] + [f"{m}/{i}" for i in range(self.n_class)])
                         for m in METRICS))


@torch.no_grad()
def validate_seg(model, volumes: Iterable, device, num_mo: int = 1, n_class: int = 3, patch_size=64, overlap=16,
                 multilabel_fusetype: Optional[str] = None, label_tfm=None) -> List[SegMetric]:
    """validate.py:212-264 without the NIfTI branch.  ``volumes`` yields ``(sn, image, label)`` with image
    C x D x H x W (or N x C x D x H x W with a list of N names and N x ... labels).  Returns one SegMetric per
    model output (head), the main output last."""
    heads = list(range(-num_mo, 0))
    sm = [SegMetric(n_class) for _ in heads]
    model.to(device)
    model.eval()
    for sn, image, label in volumes:
        if image.dim() == 4:
            image, label, sn = image.unsqueeze(0), label.unsqueeze(0), [sn]
        preds = sliding_window_forward(model, image, patch_size, overlap, device)
        label = label.to(preds.device)
        for h in heads:
            for j in range(preds.shape[1]):
                lab = label_tfm(label[j]) if label_tfm is not None else label[j]
                sm[h].evaluate_append(preds[h, j], lab, sn=sn[j], multilabel_fusetype=multilabel_fusetype)
    for s in sm:
        s.get_metric()
    return sm


class PTQTester:
    """The slice of the reference's ``PTQTester`` (src/utils/tester.py:62-66) that ``do_ptq`` uses:
    ``test_as_is(folder)`` evaluates the model in whatever mode it is in and writes
    ``<root>/<folder>/<split>_seg.txt`` (trainer.py:286-291)."""

    def __init__(self, model, data_cube, root: Optional[str], device, num_mo: int, n_class: int, patch_size,
                 overlap, multi_label: Optional[str] = None, multilabel_fusetype: Optional[str] = None, dist=None):
        self.model, self.data_cube, self.root, self.device = model, data_cube, root, device
        self.dist = dist                                  # DistCtx of a torchrun launch: volumes dealt round-robin
        self.num_mo, self.n_class = num_mo, n_class
        self.patch_size, self.overlap = patch_size, overlap
        self.label_tfm = label_transform(multi_label)
        self.fusetype = multilabel_fusetype if multi_label else None
        self.results: Dict[str, Dict[str, Dict[str, float]]] = {}

    def test_as_is(self, folder: str = "results", is_save_nii: bool = False):
        if is_save_nii:
            raise NotImplementedError("--save_nii needs nibabel, which is outside this package")
        out = {}
        rank, world = (self.dist.rank, self.dist.world) if self.dist is not None else (0, 1)
        for split in ("val", "test"):
            vols = self.data_cube.evaluation_volumes(split, rank, world) if world > 1 else \
                self.data_cube.evaluation_volumes(split)
            if vols is None:
                continue
            sm = validate_seg(self.model, vols, self.device, self.num_mo, self.n_class, self.patch_size,
                              self.overlap, self.fusetype, self.label_tfm)
            if world > 1:
                # every rank scored its own volumes (no data-path collective); the per-subject rows -- a few
                # floats each -- are exchanged and put back into the split's order: volume i sits on rank i % world
                per_rank = self.dist.all_gather_object([s.rows() for s in sm])
                n_total = sum(len(r[0]) for r in per_rank)
                sm = [SegMetric.from_rows(self.n_class, [per_rank[i % world][h][i // world] for i in range(n_total)])
                      for h in range(len(sm))]
                for s in sm:
                    s.get_metric()
                if rank != 0:
                    out[split] = dict(sm[-1].metric)
                    continue
            if self.root:
                os.makedirs(P.join(self.root, folder), exist_ok=True)
                with open(P.join(self.root, folder, "%s_seg.txt" % split), "w") as fid:
                    for i in range(-1, -self.num_mo - 1, -1):
                        sm[i].write_metric(fid, "Output %d:" % i, True)
            sm[-1].print_metric("  " + split)
            out[split] = dict(sm[-1].metric)
        self.results[folder] = out
        return out
