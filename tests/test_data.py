"""Calibration-batch assembly and evaluation volumes (efficientq_b200/data.py) against the reference's own
``get_calibration_data`` (src/ptqer.py:83-111) and val loader (src/dataloader/datahub.py:93-110) on a five-subject
on-disk dataset: tests/golden/calib_data.npz, written by tests/golden/make_golden.py::gen_calib_data.  The dataset
itself is re-written here from the same seeded numpy stream."""
import numpy as np
import pytest
import torch

from efficientq_b200 import evaluate as E
from efficientq_b200.data import CalibrationData, center_crop
from efficientq_b200.dist import DistCtx
from tests.golden.make_golden import tiny_dataset_args, write_tiny_dataset


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("tiny_brats"))
    write_tiny_dataset(root)
    return root


@pytest.mark.parametrize("disk", [True, False])
def test_calibration_batch_equals_reference(golden, dataset, disk):
    g = golden("calib_data.npz")
    tag = "disk" if disk else "mem"
    a = tiny_dataset_args(dataset, disk)                 # lwq_dataid 1, lwq_batchsz 3, lwq_patchsz 16,16,24
    x, y = CalibrationData(a).calibration_batch(a)
    assert x.dtype == torch.float32 and np.array_equal(x.numpy(), g[f"{tag}_batch"])      # order, crop: bit-exact
    # the reference's labels come out of its multi-label split; ours are split where they are used (evaluation)
    split = torch.stack([E.split_label_brats(lab.long()) for lab in y])
    assert np.array_equal(split.numpy().astype(np.uint8), g[f"{tag}_label_batch"])
    a.lwq_batchsz, a.lwq_dataid = 1, 0
    x1, _ = CalibrationData(a).calibration_batch(a)
    assert np.array_equal(x1.numpy(), g[f"{tag}_single"])


@pytest.mark.parametrize("disk", [True, False])
def test_evaluation_volumes_equal_reference_val_loader(golden, dataset, disk):
    g = golden("calib_data.npz")
    tag = "disk" if disk else "mem"
    cube = CalibrationData(tiny_dataset_args(dataset, disk))
    vols = list(cube.evaluation_volumes("val"))
    assert [v[0] for v in vols] == (["s03", "s01"] if disk else ["s01", "s03"])           # file order vs sorted
    for i, (sn, img, lab) in enumerate(vols):
        assert np.array_equal(img.numpy(), g[f"{tag}_val{i}_img"])
        assert np.array_equal(E.split_label_brats(lab.long()).numpy().astype(np.uint8), g[f"{tag}_val{i}_label"])
    assert cube.evaluation_volumes("test") is None                                        # no test.txt in the split


def test_sharded_batch_is_a_contiguous_slice(dataset):
    a = tiny_dataset_args(dataset, True)
    a.lwq_batchsz, a.lwq_dataid = 4, 0
    full, _ = CalibrationData(a).calibration_batch(a)
    parts = []
    for r in range(2):
        ctx = DistCtx()
        ctx.rank, ctx.world = r, 2                       # a rank's view of the batch; no process group needed for that
        parts.append(CalibrationData(a).calibration_batch(a, ctx)[0])
    assert [p.shape[0] for p in parts] == [2, 2] and torch.equal(torch.cat(parts), full)


def test_center_crop_pads_symmetrically_when_smaller():
    t = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(2, 3, 4, 5)
    out = center_crop(t, (5, 2, 5))
    assert out.shape == (2, 5, 2, 5)
    assert torch.equal(out[:, 1:4], t[:, :, 1:3, :]) and float(out[:, 0].abs().sum()) == 0 and float(out[:, 4].abs().sum()) == 0
