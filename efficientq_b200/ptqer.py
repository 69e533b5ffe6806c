"""PTQ orchestrator: FP pass with hooks -> attention-mask pyramid -> quantizing pass.

Mirrors reference src/ptqer.py (setters :17-80, ``get_att_weight_map`` :210-235,
``get_mask_pyramid`` :141-169, ``do_ptq`` :282-387) with three deliberate differences,
all on the data-movement side: FP targets stay in HBM (the reference's hook moves them
to the host, src/models/hooks.py:5-6), the masks stay on the device, and the calibration
volumes may be sharded over ranks (``dist``), in which case the class counts are
all-reduced.  The one-off mask / pooling arithmetic is stock PyTorch (SURVEY.md K10:
out of scope); the per-layer work happens inside ``EfficientQConv.ptq``.
"""
from __future__ import annotations

import os
import os.path as P
import time
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .dist import DistCtx
from .fold_bn import search_fold_and_remove_bn
from .qconv import PTQConv


# -- setters (ptqer.py:17-80) ---------------------------------------------------------------
def _each(m: nn.Module):
    return [(n, mod) for n, mod in m.named_modules() if isinstance(mod, PTQConv)]


def set_fp(m):
    for _, mod in _each(m):
        mod.set_fp()


def set_quantizing(m):
    for _, mod in _each(m):
        mod.set_quantizing()


def set_init_alpha(m):
    for _, mod in _each(m):
        mod.set_init_act()


def set_quantized(m):
    for _, mod in _each(m):
        mod.set_quantized()


def store_int_weight(m):
    for _, mod in _each(m):
        mod.store_int_weight()


def restore_fp_weight(m):
    for _, mod in _each(m):
        mod.restore_fp_weight()


def set_name(m):
    for n, mod in _each(m):
        mod.name = n


def set_anything(m, attr, value):
    for _, mod in _each(m):
        setattr(mod, attr, value)


def set_snapdir(m, snap_dir):
    set_anything(m, "snap_dir", snap_dir)


def set_mask(m, pyramid):
    set_anything(m, "mask_pyramid", pyramid)


def set_debug(m):
    set_anything(m, "debug", True)


# -- predictions (utils/metrics.py:172-192) ---------------------------------------------------
def get_pred_lits(out):
    return torch.max(out, 1)[1]


def get_pred_brats(out):
    hard = torch.sigmoid(out) >= 0.5
    pred = torch.zeros_like(hard[:, 0]).int()
    for i in range(hard.shape[1]):
        pred[hard[:, i]] = i + 1
    return pred


# -- attention weights (ptqer.py:172-235) -----------------------------------------------------
def get_att_weight_map(output_fp, body_mask, style: str, task: str = "lits", dist: Optional[DistCtx] = None):
    out = output_fp[-1]
    if task == "lits":
        pred = torch.max(out, 1)[1]
        n_class = 3
        nums = [((pred == k) & body_mask).sum() for k in range(n_class)]
    elif task == "brats":
        pred = (torch.sigmoid(out) >= 0.5).int()
        n_class = 4
        nums = [(pred.sum(dim=1) == 0).sum() - (~body_mask).sum()]
        nums += [(pred[:, i] * body_mask).sum() for i in range(n_class - 1)]
    else:
        raise RuntimeError(f"Unknown task {task}")
    cnt = torch.stack([n.to(torch.float64) for n in nums])
    if dist is not None:
        dist.all_reduce_sum(cnt)
    nums = [int(v) for v in cnt.tolist()]
    if "p:" not in style:
        raise RuntimeError(f"Unknown attention weight map style {style}")
    p = float(style[2:])
    wmap = {k: (1.0 if nums[k] == 0 else (1 / nums[k] * max(nums)) ** p) for k in range(n_class)}
    return wmap, nums


def get_mask_pyramid(output_fp, body_mask, weight_map: dict, init_stride, num_lvls: int = 5, task: str = "lits"):
    """Five (N,D,H,W) fp32 masks at strides init*{1,2,4,8,16}.  NB the reference fills an
    INTEGER tensor (``ones_like(pred)``), so the class weights are truncated toward zero
    (ptqer.py:160-163); reproduced."""
    if isinstance(init_stride, str):
        init_stride = tuple(int(x) for x in init_stride.split(",")) if "," in init_stride else (int(init_stride),) * 3
    out = F.avg_pool3d(output_fp[-1], init_stride)
    body = F.max_pool3d(body_mask.float(), init_stride).bool()
    pyramid = []
    for _ in range(num_lvls):
        pred = get_pred_lits(out) if task == "lits" else get_pred_brats(out)
        mask = torch.ones_like(pred)
        for k, v in weight_map.items():
            mask[pred == k] = v
        mask[~body] = 1
        pyramid.append(mask.float())
        out = F.avg_pool3d(out, 2)
        body = F.max_pool3d(body.float(), 2).bool()
    return pyramid


def fp_target_hook(m, i, o):
    """The reference's hook is ``m.output_fp = o.detach().cpu()`` (src/models/hooks.py:5-6): on the device it runs on
    that is a COPY of the layer's pre-activation output.  A bare ``detach()`` would alias the conv output, and with
    the BN folded to identity the next unit's ``ReLU(inplace=True)`` (blk: mid) overwrites it, turning the target
    into the post-ReLU tensor.  The copy stays in HBM (no host round trip)."""
    m.output_fp = o.detach().clone()


def register_fp_hooks(model: nn.Module):
    return [mod.register_forward_hook(fp_target_hook) for _, mod in _each(model)]


# -- the calibration pass (ptqer.py:313-364) -------------------------------------------------
@torch.no_grad()
def calibrate(model: nn.Module, data_batch: torch.Tensor, task: str, init_stride, dist: Optional[DistCtx] = None,
              keep_history: bool = False, n_iter: Optional[int] = None) -> Dict:
    """FP forward + mask pyramid + quantizing forward over ``data_batch`` (this rank's shard).
    The model must already be BN-folded, on the device and in eval mode."""
    dist = dist or DistCtx()
    dev = data_batch.device
    set_name(model)
    set_fp(model)
    handles = register_fp_hooks(model)
    torch.cuda.synchronize(dev)
    t0 = time.time()
    output_fp = model(data_batch).detach()
    if task == "brats":
        body = (data_batch[:, 0] != 0.0).bool()
    else:
        body = torch.ones_like(data_batch[:, 0]).bool()
    # NB the reference passes an all-ones mask here, not body_mask (ptqer.py:342)
    wmap, nums = get_att_weight_map(output_fp, torch.ones_like(data_batch[:, 0]).bool(), "p:0.5", task, dist)
    pyramid = get_mask_pyramid(output_fp, body, wmap, init_stride, 5, task)
    set_mask(model, pyramid)
    for h in handles:
        h.remove()
    layer_loss: List[str] = []
    set_anything(model, "layer_loss", layer_loss)
    set_anything(model, "dist", dist)
    set_anything(model, "keep_history", keep_history)
    if n_iter is not None:
        set_anything(model, "lwq_iter", n_iter)
    torch.cuda.synchronize(dev)
    t1 = time.time()
    set_quantizing(model)
    output_q = model(data_batch)
    torch.cuda.synchronize(dev)
    t2 = time.time()
    set_quantized(model)
    reports = [mod.report for _, mod in _each(model)]
    return dict(output_fp=output_fp, output_q=output_q, layer_loss=layer_loss, reports=reports,
                class_nums=nums, weight_map=wmap, pyramid=pyramid, t_fp=t1 - t0, t_ptq=t2 - t1, t_total=t2 - t0)


def do_ptq(args, model_cube, data_cube, tester, snap_dir, dist: Optional[DistCtx] = None):
    """Reference do_ptq (ptqer.py:282-387): load + fold BN, assemble the calibration batch,
    calibrate, write ``time_cost.txt`` / ``layer_loss.txt`` / ``class_voxel_nums.txt`` and the
    three snapshots.  ``data_cube`` provides ``calibration_batch(args)``; ``tester`` may be None
    (no dataset-level validation in this package)."""
    from . import snapshot
    dist = dist or DistCtx()
    model = model_cube["model"]
    device = torch.device(args.device if not isinstance(args.device, int) else f"cuda:{args.device}")
    if model_cube.get("pretrain"):
        print("pretrain is :", model_cube["pretrain"])
        sd = torch.load(model_cube["pretrain"], map_location="cpu")["state_dict"]
        model.load_state_dict(sd, strict=False)
    model.eval()
    search_fold_and_remove_bn(model)
    model.to(device)
    data_batch, _ = data_cube.calibration_batch(args, dist)
    data_batch = data_batch.to(device, non_blocking=True)
    set_snapdir(model, snap_dir)
    torch.backends.cudnn.allow_tf32 = False            # FP targets in true fp32, like the CPU reference
    torch.backends.cuda.matmul.allow_tf32 = False
    if getattr(args, "test_fp", False) and tester is not None:
        set_fp(model)
        tester.test_as_is(folder="fp", is_save_nii=getattr(args, "save_nii", False))
    res = calibrate(model, data_batch, args.task, args.init_stride, dist)
    print(f"FP forward costs {res['t_fp']:.3f}s, PTQ costs {res['t_ptq']:.3f}s, totally {res['t_total']:.3f}s.")
    n_tune = int(getattr(args, "tune_act_iter", 0) or 0)
    if n_tune > 0:                                       # optional: reference ptqer.py:238-272
        from .tune import tune_activation_range
        res["tune_losses"] = tune_activation_range(model, res["output_fp"], data_batch, max_iter=n_tune, dist=dist)
        print(f"alpha_act refinement: loss {res['tune_losses'][0]:.6g} -> {res['tune_losses'][-1]:.6g} "
              f"in {n_tune} Adam iterations")
    if dist.rank == 0 and snap_dir:
        os.makedirs(snap_dir, exist_ok=True)
        with open(P.join(snap_dir, "class_voxel_nums.txt"), "w") as fid:
            for n in res["class_nums"]:
                fid.write(f"{n}\n")
        with open(P.join(snap_dir, "time_cost.txt"), "w") as fid:
            fid.write(f"{res['t_total'] / 60:.3f} min.")
        with open(P.join(snap_dir, "layer_loss.txt"), "w") as fid:
            fid.write("\n".join(res["layer_loss"]))
    if not getattr(args, "no_test", False) and tester is not None:
        tester.test_as_is("ptq", getattr(args, "save_nii", False))
    dist.close()
    if dist.rank == 0 and snap_dir:
        model.cpu()
        snapshot.save(model, P.join(snap_dir, "state_in_fp.pkl"), compress=False)
        store_int_weight(model)
        snapshot.save(model, P.join(snap_dir, "state_in_int8.pkl"), compress=False)
        snapshot.save(model, P.join(snap_dir, "state_in_int8_compress.npz"), compress=True)
        snapshot.save_packed(model, P.join(snap_dir, "state_in_packed.npz"))     # extension: 2 / 4 / 8-bit packed codes
    return res
