"""``python -m efficientq_b200.entrance ptq ...`` -- the reference's CLI for the PTQ mission.

Same flags and YAML schema as reference src/entrance.py:33-128 (argparse dests == YAML keys;
``merge_config``: every non-null YAML key overrides the command line).  ``train_fp`` is a
different workload and is not part of this package.  New, optional: ``--data_dir synthetic``
and torchrun launch (one process per GPU shards ``lwq_batchsz`` volumes).
"""
import argparse

import yaml


def merge_config(cfg: str, args: argparse.Namespace):
    """entrance.py:17-28: configuration file first."""
    with open(cfg, "r") as fid:
        config = yaml.load(fid, Loader=yaml.FullLoader)
    for k, v in config.items():
        if v is not None:
            setattr(args, k, v)
    return args


def build_parser():
    p = argparse.ArgumentParser(description="Entrance for Quantization/FP training/Inference")
    p.add_argument("mission", choices=["train_fp", "ptq"])
    p.add_argument("--pretrain")
    p.add_argument("--resume")
    p.add_argument("--device", default=0, type=int, dest="device", help="GPU ID.")
    p.add_argument("--task")
    p.add_argument("--suffix", default="", type=str, dest="suffix", help="folder name suffix.")
    p.add_argument("--test_fp", action="store_true")
    p.add_argument("--config", type=str)
    # data
    p.add_argument("--data_dir")
    p.add_argument("--split_dir")
    p.add_argument("--round", default="1", type=str, dest="round", help="round number.")
    p.add_argument("--patch_size")
    p.add_argument("--batch_size", default=1, type=int)
    p.add_argument("--test_batch_size", default=1, type=int)
    p.add_argument("--crop_type", default="random")
    p.add_argument("--balance_rate", type=float)
    p.add_argument("--data_on_disk", action="store_true")
    p.add_argument("--bin_label", help="convert to binary label")
    p.add_argument("--multi_label", help="multiple labels per pixel")
    p.add_argument("--merge_type", help="how to merge multiple labels")
    p.add_argument("--random_noise_p", type=float)
    p.add_argument("--access_type", default="npy")
    p.add_argument("--num_workers", default=4, type=int)
    p.add_argument("--da_scaling", type=str, default=None)
    p.add_argument("--scal_order", type=int, default=1)
    # model
    p.add_argument("--model", default="UResQ")
    p.add_argument("--nMod", type=int)
    p.add_argument("--nClass", type=int)
    p.add_argument("--init_stride", type=str, default="1")
    p.add_argument("--resblock")
    p.add_argument("--depth")
    p.add_argument("--width")
    p.add_argument("--dilation")
    p.add_argument("--nla", default="relu")
    p.add_argument("--norm", type=str, default="bn")
    p.add_argument("--group_num", type=int, help="GN's group number")
    p.add_argument("--drop_rate", default=0.2, type=float)
    p.add_argument("--no_drop", action="store_true")
    p.add_argument("--ds", type=str, default=None, choices=["simple", "complex", ""])
    p.add_argument("--init_kernel", default=3, type=int)
    p.add_argument("--block_type", default="RBpre")
    p.add_argument("--hetero_dim", action="store_true")
    p.add_argument("--blk", type=str, default="pre")
    # FP training (accepted for CLI compatibility, unused by the ptq mission)
    p.add_argument("--lr", default=0.001, type=float, metavar="LR", dest="lr")
    p.add_argument("--max_epoch", type=int, default=20)
    p.add_argument("--loss", type=str, default="CE")
    p.add_argument("--test_interval", type=int, default=50)
    p.add_argument("--disp_interval", type=int, default=10)
    p.add_argument("--weight_decay", type=str, default="0")
    p.add_argument("--no_test", action="store_true")
    p.add_argument("--exp_id", type=str, default=None)
    # quantization
    p.add_argument("--qconv", default="conv")
    p.add_argument("--qlvl_w", type=int)
    p.add_argument("--qlvl_a", type=int)
    p.add_argument("--q_first", help="whether quantize first layer. e.g., --q_first 256,64 for W8A4")
    p.add_argument("--q_last", help="similar to q_first")
    # PTQ
    p.add_argument("--debug", action="store_true")
    p.add_argument("--lwq_dataid", type=int, default=0)
    p.add_argument("--lwq_batchsz", type=int, default=1)
    p.add_argument("--lwq_patchsz")
    p.add_argument("--lwq_verbose", action="store_true")
    # extension (not a reference flag): > 0 runs the reference's tune_activation_range (src/ptqer.py:238-272,
    # defined there but never called) for that many Adam iterations after the layer-wise calibration
    p.add_argument("--tune_act_iter", type=int, default=0)
    p.add_argument("--save_nii", action="store_true")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    new_args = merge_config(args.config, args) if args.config else args
    if new_args.mission == "ptq":
        from .ptq_seg import ptq
        return ptq(new_args)
    raise NotImplementedError("train_fp is outside this package: it only accelerates the ptq mission")


if __name__ == "__main__":
    main()
