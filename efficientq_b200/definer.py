"""Factories: conv-class registry, model, snapshot dir (reference src/definer.py:130-329).

``get_conv_class`` is the reference's plug-in point: ``--qconv`` -> class.  ``effq`` now maps to
the B200-native ``EfficientQConv`` of this package; nothing else in the model code changes.
"""
from __future__ import annotations

import os
import os.path as P
import sys
import time

import torch.nn as nn

from . import model_blk, qconv


def _ints(s):
    return [int(x) for x in str(s).split(",")]


def get_conv_class(args):
    """definer.py:286-329."""
    name = args.qconv.lower()
    if name == "conv":
        return nn.Conv3d, "FP", {}
    if name not in ("effq", "effq_b200"):
        raise RuntimeError("Unknown QConv name: %s" % args.qconv)
    q_weight, q_act = args.qlvl_w > 0, args.qlvl_a > 0
    qlvl, qlvl_act = args.qlvl_w, (args.qlvl_a if q_act else 256)
    kwQ = {a: getattr(args, a) for a in dir(args) if a[:4] == "lwq_"}
    if getattr(args, "w_per_channel", False):          # extension flag (entrance.py): not one of the reference's lwq_* keys
        kwQ["lwq_channel_wise"] = True
    if q_act and q_weight:
        info = "bothQw{}a{}".format(qlvl, qlvl_act)
    elif q_act:
        info = "actQa{}".format(qlvl_act)
    else:
        info = "weightQw{}".format(qlvl)
    return qconv.EfficientQConv, args.qconv + "_" + info, kwQ


def get_model_cube(args, QConv=nn.Conv3d, kwQ=None):
    """definer.py:130-248."""
    kwQ = kwQ or {}
    task = args.task.lower()
    n_mod = args.nMod if getattr(args, "nMod", None) else (4 if task == "brats" else 1)
    n_class = args.nClass if getattr(args, "nClass", None) else (4 if task == "brats" else 3)
    if getattr(args, "bin_label", None):
        n_class = 2
    if getattr(args, "multi_label", None):
        n_class -= 1
    if args.model not in ("UResQ",):
        raise RuntimeError("Unknown model name: %s" % args.model)
    init_stride = tuple(_ints(args.init_stride)) if "," in str(args.init_stride) else (int(args.init_stride),) * 3
    if args.qconv.lower() == "conv":
        q_weight = q_act = False
        q_first = q_last = qlvl = qlvl_act = None
    else:
        q_weight, q_act = args.qlvl_w > 0, args.qlvl_a > 0
        qlvl, qlvl_act = args.qlvl_w, (args.qlvl_a if q_act else 256)
        q_first = _ints(args.q_first) if getattr(args, "q_first", None) else None
        q_last = _ints(args.q_last) if getattr(args, "q_last", None) else None
    nla_name = (getattr(args, "nla", "relu") or "relu").lower()
    if nla_name not in ("relu", "reluf"):
        raise RuntimeError("Unknown NLA name: %s" % args.nla)
    nla = model_blk.ReLU(nla_name == "relu")
    if (getattr(args, "norm", "bn") or "bn").lower() != "bn":
        raise NotImplementedError("Norm type should be in BN")
    width = _ints(args.width) if getattr(args, "width", None) else [32, 64, 128, 256, 128, 64, 32]
    depth = _ints(args.depth) if getattr(args, "depth", None) else [1] * len(width)
    dilation = _ints(args.dilation) if getattr(args, "dilation", None) else [1] * len(width)
    hetero = {"drop_cut_thres": 128, "ds_depth_limit": 3 if 2 in init_stride else 4}
    if getattr(args, "hetero_dim", False):
        hetero["aniso_pool_depth"] = 9999 if 2 in init_stride else 4
        hetero["aniso_pool_stride"] = (2, 2, 1)
    model = model_blk.UResQ(QConv, n_mod, n_class, depth_config=depth, width_config=width, dilation_config=dilation,
                            init_stride=init_stride, stride=2, drop_rate=args.drop_rate, nla=nla, bn=nn.BatchNorm3d,
                            ds=getattr(args, "ds", None), blk_type=args.blk, q_weight=q_weight, qlvl=qlvl, q_act=q_act,
                            qlvl_act=qlvl_act, q_first=q_first, q_last=q_last, hetero_param=hetero, fuse_bn=True,
                            save_mem=True, init_kernel=getattr(args, "init_kernel", 3), **kwQ)
    num_mo = min(hetero["ds_depth_limit"], len(depth) // 2 + 1) if getattr(args, "ds", None) else 1
    cube = {"model": model, "pretrain": getattr(args, "pretrain", None), "resume": getattr(args, "resume", None),
            "num_mo": num_mo, "nClass": n_class, "nMod": n_mod}
    return cube, args.model + "_BN"


def get_snapshot_config(args, model_info, qinfo, model=None, data_cube=None):
    """definer.py:251-283: <repo>/exp_ptq/<task>/snap/round<r>/<exp_id>."""
    exp_id = getattr(args, "exp_id", None) or f"{model_info}_{time.strftime('%m%d%H%M')}_{qinfo}"
    exp_id += getattr(args, "suffix", "") or ""
    root = P.join(P.dirname(P.abspath(__file__)), "..", "exp_ptq", args.task, "snap", "round" + str(args.round), exp_id)
    print(f"Snapshot to {root}")
    os.makedirs(root, exist_ok=True)
    with open(P.join(root, "cmd.txt"), "w+") as fid:
        fid.write(str(sys.argv) + "\n" + " ".join(sys.argv) + "\n")
        if model is not None:
            fid.write("Number of parameters: %d\n" % sum(p.numel() for p in model.parameters()))
    return {"root": root, "is_train": False, "is_val": False, "is_test": False}
