"""Sliding-window evaluation around the calibrated net on the GPU: the reference's
`entrance.py ptq --test_fp` sequence -- FP evaluation, calibration, evaluation of the quantized net on the deployment
forward (tcgen05 code path) -- on synthetic volumes, with the report files of src/utils/trainer.py:286-291.
(Named zz: it runs after the parity tests proper.)"""
import os
import shutil

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_evaluates_fp_and_quantized_net(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from efficientq_b200 import entrance
    cfg = yaml.load(open(os.path.join(ROOT, "config", "brats_ptq.yaml")), Loader=yaml.FullLoader)
    cfg.update(patch_size="64,64,64", lwq_patchsz="64,64,64", lwq_batchsz=2)      # YAML keys win over the command line
    cfg_path = os.path.join(tmp_path, "brats_small.yaml")
    yaml.dump(cfg, open(cfg_path, "w"))
    exp_id = "pytest_cli_eval_w8a8"
    root = os.path.join(ROOT, "exp_ptq", "brats", "snap", "round1", exp_id)
    shutil.rmtree(root, ignore_errors=True)
    torch.manual_seed(0)
    try:
        res = entrance.main(["ptq", "--qlvl_w", "256", "--qlvl_a", "256", "--round", "1", "--device", "0",
                             "--config", cfg_path, "--data_dir", "synthetic", "--exp_id", exp_id, "--test_fp"])
        assert set(res["eval"]) == {"fp", "ptq"}
        for folder in ("fp", "ptq"):
            m = res["eval"][folder]["val"]
            assert all(np.isfinite(v) and 0.0 <= v <= 1.0 for v in m.values()), m
            lines = open(os.path.join(root, folder, "val_seg.txt")).read().splitlines()
            assert lines[0] == "Output -1:" and lines[1].startswith("acc = ")
            assert [ln.split("|")[1].strip() for ln in lines[3:5]] == ["synthetic_5000", "synthetic_5001"]
            assert sum(ln.startswith("Output") for ln in lines) == 3          # ds: simple -> three heads
        print("eval fp", res["eval"]["fp"]["val"], "ptq", res["eval"]["ptq"]["val"])
    finally:
        shutil.rmtree(root, ignore_errors=True)
