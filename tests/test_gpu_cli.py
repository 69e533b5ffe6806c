"""The reference's CLI on the GPU: `entrance.py ptq --config brats_ptq.yaml ...` (src/entrance.py:116-128 ->
ptq_seg.ptq -> ptqer.do_ptq) end to end on synthetic volumes, with the artefacts do_ptq writes
(src/ptqer.py:364-387) and the optional alpha refinement."""
import os
import shutil

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_ptq_end_to_end_writes_reference_artefacts():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from efficientq_b200 import entrance, snapshot
    exp_id = "pytest_cli_w4a4"
    root = os.path.join(ROOT, "exp_ptq", "brats", "snap", "round1", exp_id)
    shutil.rmtree(root, ignore_errors=True)
    torch.manual_seed(0)          # no --pretrain: the net keeps torch's default random initialisation
    try:
        res = entrance.main(["ptq", "--qlvl_w", "16", "--qlvl_a", "16", "--round", "1", "--device", "0",
                             "--config", os.path.join(ROOT, "config", "brats_ptq.yaml"), "--data_dir", "synthetic",
                             "--lwq_patchsz", "64,64,64", "--lwq_batchsz", "2", "--exp_id", exp_id,
                             "--tune_act_iter", "2", "--no_test"])      # evaluation: tests/test_gpu_zz_eval.py
        for f in ("cmd.txt", "time_cost.txt", "layer_loss.txt", "class_voxel_nums.txt", "state_in_fp.pkl",
                  "state_in_int8.pkl", "state_in_int8_compress.npz", "state_in_packed.npz"):
            assert os.path.exists(os.path.join(root, f)), f
        lines = open(os.path.join(root, "layer_loss.txt")).read().strip().splitlines()
        assert len(lines) == 22 and lines[0].startswith("conv0.conv")          # BraTS config
        losses = [float(ln.rsplit(":", 1)[1]) for ln in lines]
        assert all(np.isfinite(losses)) and all(v > 0 for v in losses)
        assert open(os.path.join(root, "time_cost.txt")).read().endswith("min.")
        assert len(res["tune_losses"]) == 2 and all(np.isfinite(res["tune_losses"]))
        sd8 = torch.load(os.path.join(root, "state_in_int8.pkl"))["state_dict"]
        sdf = torch.load(os.path.join(root, "state_in_fp.pkl"))["state_dict"]
        sdp = snapshot.load_packed(os.path.join(root, "state_in_packed.npz"))["state_dict"]
        print("maxcodes", sorted({int(v.max()) for kk, v in sd8.items() if v.dtype == torch.uint8}), "tune", res["tune_losses"])
        k = "u_blocks.UResBlock1.Layer1.block1.conv.weight"
        # (codes come from alpha_w of the LAST iterate while the weights are the BEST iterate's -- reference quirk --
        #  so a code may exceed 15 by one; the packed file widens such a layer instead of failing)
        assert sd8[k].dtype == torch.uint8 and int(sd8[k].max()) <= 17
        assert set(sdp) == set(sd8) and all(torch.equal(sdp[kk], sd8[kk]) for kk in sd8)
        assert torch.unique(sdf[k]).numel() <= 16                               # fake-quant weights: 16 levels
        assert float(sdf[k.replace("weight", "alpha_act")]) > 0
    finally:
        shutil.rmtree(root, ignore_errors=True)
