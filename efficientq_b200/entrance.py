"""``python -m efficientq_b200.entrance ptq ...`` -- the command line of the PTQ mission.

The flag set and the YAML schema are the reference's (src/entrance.py:33-128: argparse dests == YAML keys, and
``merge_config`` lets every non-null YAML key override the command line), so existing command lines and config
files keep working.  The flags are declared as a table (name, type, default) instead of one ``add_argument``
call each.  ``train_fp`` is a different workload and is not part of this package.  New, optional:
``--data_dir synthetic``, ``--tune_act_iter N`` and a torchrun launch (one process per GPU shards
``lwq_batchsz`` volumes).
"""
import argparse

import yaml

FLAG = object()          # marks a store_true switch in the tables below

# name -> (type or FLAG, default).  Grouped as the YAML files are; only the ptq mission reads them.
_GENERAL = {"pretrain": (None, None), "resume": (None, None), "device": (int, 0), "task": (None, None),
            "suffix": (str, ""), "test_fp": (FLAG, False), "config": (str, None), "exp_id": (str, None),
            "no_test": (FLAG, False), "save_nii": (FLAG, False), "debug": (FLAG, False)}
_DATA = {"data_dir": (None, None), "split_dir": (None, None), "round": (str, "1"), "patch_size": (None, None),
         "batch_size": (int, 1), "test_batch_size": (int, 1), "crop_type": (None, "random"),
         "balance_rate": (float, None), "data_on_disk": (FLAG, False), "bin_label": (None, None),
         "multi_label": (None, None), "merge_type": (None, None), "random_noise_p": (float, None),
         "access_type": (None, "npy"), "num_workers": (int, 4), "da_scaling": (str, None), "scal_order": (int, 1)}
_MODEL = {"model": (None, "UResQ"), "nMod": (int, None), "nClass": (int, None), "init_stride": (str, "1"),
          "resblock": (None, None), "depth": (None, None), "width": (None, None), "dilation": (None, None),
          "nla": (None, "relu"), "norm": (str, "bn"), "group_num": (int, None), "drop_rate": (float, 0.2),
          "no_drop": (FLAG, False), "init_kernel": (int, 3), "block_type": (None, "RBpre"),
          "hetero_dim": (FLAG, False), "blk": (str, "pre")}
# accepted so that the reference's training command lines parse; unused by the ptq mission
_TRAIN = {"lr": (float, 0.001), "max_epoch": (int, 20), "loss": (str, "CE"), "test_interval": (int, 50),
          "disp_interval": (int, 10), "weight_decay": (str, "0")}
_QUANT = {"qconv": (None, "conv"), "qlvl_w": (int, None), "qlvl_a": (int, None), "q_first": (None, None),
          "q_last": (None, None), "lwq_dataid": (int, 0), "lwq_batchsz": (int, 1), "lwq_patchsz": (None, None),
          "lwq_verbose": (FLAG, False),
          # extension (not a reference flag, and deliberately not named lwq_*: the quantizer classes receive exactly
          # the reference's lwq_* keys): > 0 runs tune_activation_range (src/ptqer.py:238-272) for that many Adam steps
          "tune_act_iter": (int, 0),
          # extension: one weight scale per OUTPUT CHANNEL instead of the reference's per-tensor alpha_w
          # (PTQConv.py:26-27); alpha_w then has shape [C_out] in the snapshots
          "w_per_channel": (FLAG, False)}


def merge_config(cfg: str, args: argparse.Namespace):
    """Overlay the YAML file on the parsed arguments: any key with a non-null value wins (entrance.py:17-28)."""
    with open(cfg, "r") as fid:
        loaded = yaml.load(fid, Loader=yaml.FullLoader) or {}
    for key, value in loaded.items():
        if value is not None:
            setattr(args, key, value)
    return args


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="EfficientQ PTQ calibration on B200 (reference-compatible command line)")
    parser.add_argument("mission", choices=["train_fp", "ptq"])
    for table in (_GENERAL, _DATA, _MODEL, _TRAIN, _QUANT):
        for name, (typ, default) in table.items():
            if typ is FLAG:
                parser.add_argument("--" + name, action="store_true")
            elif typ is None:
                parser.add_argument("--" + name, default=default)
            else:
                parser.add_argument("--" + name, type=typ, default=default)
    parser.add_argument("--ds", type=str, default=None, choices=["simple", "complex", ""])
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    merged = merge_config(args.config, args) if args.config else args
    if merged.mission == "ptq":
        from .ptq_seg import ptq
        return ptq(merged)
    raise NotImplementedError("train_fp is outside this package: it only accelerates the ptq mission")


if __name__ == "__main__":
    main()
