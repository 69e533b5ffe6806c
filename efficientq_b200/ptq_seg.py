"""PTQ mission driver (reference src/ptq_seg.py:7-30)."""
import torch

from . import definer
from .data import CalibrationData
from .dist import init_from_env
from .ptqer import do_ptq


def ptq(args):
    dist = init_from_env("nccl")
    if isinstance(args.device, int) and dist.world > 1:
        args.device = torch.cuda.current_device()
    data_cube = CalibrationData(args)
    QConv, Qinfo, kwQ = definer.get_conv_class(args)
    model_cube, model_info = definer.get_model_cube(args, QConv, kwQ)
    model = model_cube["model"]
    if model_cube["pretrain"]:
        assert "round" + str(args.round) in model_cube["pretrain"], "round number does not match pretrain model!"
    snap = definer.get_snapshot_config(args, model_info, Qinfo, model, data_cube) if dist.rank == 0 else {"root": None}
    return do_ptq(args, model_cube, data_cube, None, snap["root"], dist=dist)
