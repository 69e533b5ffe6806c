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
    # sliding-window evaluation before / after the calibration (src/ptqer.py:309-310, :379-380).  Every rank holds
    # the same calibrated net, so the validation volumes are dealt round-robin over the ranks; rank 0 writes the report
    tester = None
    wants_test = getattr(args, "test_fp", False) or not getattr(args, "no_test", False)
    if wants_test:
        if getattr(args, "save_nii", False):
            raise NotImplementedError("--save_nii needs nibabel, which is outside this package")
        from .evaluate import PTQTester
        device = torch.device(args.device if not isinstance(args.device, int) else f"cuda:{args.device}")
        patch_size, overlap = data_cube.slide_window()
        tester = PTQTester(model, data_cube, snap["root"], device, model_cube["num_mo"], model_cube["nClass"],
                           patch_size, overlap, getattr(args, "multi_label", None), getattr(args, "merge_type", None),
                           dist=dist if dist.world > 1 else None)
    res = do_ptq(args, model_cube, data_cube, tester, snap["root"], dist=dist)
    if tester is not None:
        res["eval"] = tester.results
    return res
