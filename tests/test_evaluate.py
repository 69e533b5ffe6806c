"""Sliding-window evaluation (efficientq_b200/evaluate.py) against fixtures generated with the reference's own
``image_to_patch3d`` / ``patch_to_image3d`` (src/utils/transforms.py:784-851), label split / merge
(src/utils/misc.py:221-285) and ``SegMetricMC`` (src/utils/validate.py:19-205): tests/golden/eval.npz, written by
tests/golden/make_golden.py::gen_eval.  Host logic only -- the model inside the window loop is a stub here; the
GPU test of the same driver around the calibrated net is tests/test_gpu_zz_eval.py."""
import io
import os

import numpy as np
import pytest
import torch

from efficientq_b200 import evaluate as E
from tests.golden.make_golden import EVAL_TILINGS


@pytest.mark.parametrize("ci", range(len(EVAL_TILINGS)))
def test_window_order_and_stitching_bit_exact(golden, ci):
    g = golden("eval.npz")
    dhw, patch, ov = EVAL_TILINGS[ci]
    wins = E.windows(dhw, patch, ov)
    starts = [(w[0].start * dhw[1] + w[1].start) * dhw[2] + w[2].start for w in wins]
    assert starts == g[f"tile{ci}_starts"].tolist()                     # same windows, same order (duplicates included)
    img = torch.from_numpy(g[f"tile{ci}_img"])
    calls = []

    def model(p):                                                       # the k-th patch gets a k-dependent prediction
        k = len(calls)
        calls.append(tuple(p.shape))
        return torch.stack([p * (1.0 + 0.125 * k) + 0.5 * k, p.flip(1) - 0.25 * k])

    out = E.sliding_window_forward(model, img, patch, ov)
    assert len(calls) == len(wins) and all(c[-3:] == tuple(patch) for c in calls)
    assert out.shape == (2,) + tuple(img.shape)
    assert np.array_equal(out.numpy(), g[f"tile{ci}_stitched"])         # same summation order per voxel -> same bits


def test_window_rule_edge_cases():
    assert E.window_starts(128, 128, 16) == [0]
    assert E.window_starts(160, 128, 16) == [0, 32]
    assert E.window_starts(240, 128, 16) == [0, 112]                    # 112 both from the grid and from the edge rule
    assert E.window_starts(241, 128, 16) == [0, 112, 113]
    with pytest.raises(ValueError):
        E.window_starts(100, 128, 16)
    with pytest.raises(ValueError):
        E.window_starts(128, 16, 16)
    x = torch.randn(1, 2, 8, 8, 8)
    assert torch.equal(E.sliding_window_forward(lambda p: p * 2, x, None, None), x * 2)   # no window: one forward


def test_label_split_and_merge(golden):
    g = golden("eval.npz")
    assert np.array_equal(E.split_label_brats(torch.from_numpy(g["label_brats"]).long()).numpy(), g["split_brats"])
    assert np.array_equal(E.split_label_lits(torch.from_numpy(g["label_lits"]).long()).numpy(), g["split_lits"])
    bits = torch.from_numpy(g["merge_in"])
    assert np.array_equal(E.merge_label_basic(bits.clone(), "con").numpy(), g["merge_con"])
    assert np.array_equal(E.merge_label_basic(bits.clone(), "agg").numpy(), g["merge_agg"])
    with pytest.raises(RuntimeError):
        E.merge_label_basic(bits, "other")
    assert E.label_transform(None) is None and E.label_transform("BRATS") is E.split_label_brats


@pytest.mark.parametrize("mode", ["mc", "ml_con", "ml_none"])
def test_metrics_match_reference_segmetric(golden, mode):
    g = golden("eval.npz")
    sm = E.SegMetric(3)
    for j, sn in enumerate(["case_a", "case_b"]):
        logits, label = torch.from_numpy(g[f"{mode}_logits{j}"]), torch.from_numpy(g[f"{mode}_label{j}"])
        pred = sm.evaluate_append(logits, label, sn=sn, multilabel_fusetype="con" if mode == "ml_con" else None)
        assert np.array_equal(pred.numpy(), g[f"{mode}_pred{j}"])
    assert sm.keys == [str(k) for k in g[f"{mode}_keys"]]
    buf = np.array([[float(v) for v in sm.buffer[k]] for k in sm.keys], dtype=np.float32)
    assert np.array_equal(buf, g[f"{mode}_buffer"])                      # fp32 formulas on exact counts: same bits
    metric = sm.get_metric()
    assert np.array_equal(np.array([metric[k] for k in sm.keys]), g[f"{mode}_metric"])
    out = io.StringIO()
    sm.write_metric(out, "Output -1:", True)
    assert out.getvalue() == str(g[f"{mode}_text"])                      # the report file, character for character


class _Cube:
    def __init__(self, vols):
        self.vols = vols

    def evaluation_volumes(self, split):
        return iter(self.vols) if split == "val" else None


class _Heads(torch.nn.Module):
    """Two-head stub net: heads x N x classes x D x H x W, like UResQ with deep supervision."""

    def __init__(self, n_class):
        super().__init__()
        self.a = torch.nn.Conv3d(2, n_class, 3, padding=1)
        self.b = torch.nn.Conv3d(2, n_class, 1)

    def forward(self, x):
        return torch.stack([self.b(x), self.a(x)])


def test_tester_writes_the_reference_report(tmp_path):
    torch.manual_seed(3)
    net = _Heads(3)
    vols = [(f"sn{i}", torch.randn(2, 20, 16, 16), torch.randint(0, 4, (20, 16, 16))) for i in range(3)]
    t = E.PTQTester(net, _Cube(vols), str(tmp_path), "cpu", num_mo=2, n_class=3, patch_size=(16, 16, 16), overlap=8,
                    multi_label="brats", multilabel_fusetype="con")
    res = t.test_as_is("ptq")
    assert set(res) == {"val"} and 0.0 <= res["val"]["dsc"] <= 1.0 and t.results["ptq"] is res
    text = open(os.path.join(tmp_path, "ptq", "val_seg.txt")).read().splitlines()
    assert text[0] == "Output -1:" and text[1].startswith("acc = ") and text[2].startswith("|                  SN|")
    assert [ln.split("|")[1].strip() for ln in text[3:6]] == ["sn0", "sn1", "sn2"]
    assert text[6] == "Output -2:" and len(text) == 12
    # a 1x1x1 head sees no window border: its stitched prediction equals the whole-volume forward, so the
    # windowed Dice of head -2 must equal the direct one
    sm = E.SegMetric(3)
    with torch.no_grad():
        for sn, img, lab in vols:
            sm.evaluate_append(net.b(img.unsqueeze(0))[0], E.split_label_brats(lab), sn=sn, multilabel_fusetype="con")
    direct = ", ".join("%s = %.4f" % kv for kv in sm.get_metric().items())
    assert text[7] == direct
    with pytest.raises(NotImplementedError):
        t.test_as_is("ptq", is_save_nii=True)


def test_synthetic_evaluation_volumes_and_window():
    from efficientq_b200 import entrance
    from efficientq_b200.data import CalibrationData, SYNTH_VAL_VOLUMES
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = entrance.build_parser().parse_args(["ptq", "--config", os.path.join(root, "config", "lits_ptq.yaml"),
                                            "--data_dir", "synthetic"])
    a = entrance.merge_config(a.config, a)
    a.data_dir = "synthetic"
    cube = CalibrationData(a)
    assert cube.slide_window() == ((128, 128, 64), (16, 16, 16))
    vols = list(cube.evaluation_volumes("val"))
    assert len(vols) == SYNTH_VAL_VOLUMES and cube.evaluation_volumes("test") is None
    sn, img, lab = vols[0]
    assert img.shape == (1, 160, 128, 64) and lab.shape == (160, 128, 64) and int(lab.max()) == 2
    assert len(E.windows(img.shape[-3:], *cube.slide_window())) == 2


def test_tester_around_the_brats_unet_in_fp_mode(tmp_path):
    """The mission driver's wiring (ptq_seg.ptq) on the CPU as far as it goes without the CUDA kernels: BraTS-config
    U-Net from the YAML, BN folded, FP mode, synthetic validation volumes, three heads, multi-label BraTS maps."""
    import yaml
    from efficientq_b200 import definer, entrance, ptqer
    from efficientq_b200.data import CalibrationData
    from efficientq_b200.fold_bn import search_fold_and_remove_bn
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = yaml.load(open(os.path.join(root, "config", "brats_ptq.yaml")), Loader=yaml.FullLoader)
    cfg.update(patch_size="64,64,64")
    cfg_path = os.path.join(tmp_path, "c.yaml")
    yaml.dump(cfg, open(cfg_path, "w"))
    a = entrance.build_parser().parse_args(["ptq", "--qlvl_w", "16", "--qlvl_a", "16", "--config", cfg_path])
    a = entrance.merge_config(a.config, a)
    a.data_dir = "synthetic"
    torch.manual_seed(0)
    cube = CalibrationData(a)
    QConv, _, kwQ = definer.get_conv_class(a)
    mc, _ = definer.get_model_cube(a, QConv, kwQ)
    model = mc["model"].eval()
    search_fold_and_remove_bn(model)
    ptqer.set_fp(model)
    patch, overlap = cube.slide_window()
    t = E.PTQTester(model, cube, str(tmp_path), "cpu", mc["num_mo"], mc["nClass"], patch, overlap, a.multi_label, a.merge_type)
    res = t.test_as_is("fp")["val"]
    assert mc["num_mo"] == 3 and mc["nClass"] == 3 and patch == (64, 64, 64)
    assert all(np.isfinite(v) and 0.0 <= v <= 1.0 for v in res.values())
    lines = open(os.path.join(tmp_path, "fp", "val_seg.txt")).read().splitlines()
    assert sum(ln.startswith("Output") for ln in lines) == 3 and lines[3].split("|")[1].strip() == "synthetic_5000"


def test_mission_driver_hands_the_tester_to_do_ptq(monkeypatch):
    """ptq_seg.ptq: evaluation is on by default (reference ptqer.py:379), --no_test alone switches it off, --test_fp
    alone keeps the tester, --save_nii fails before any work."""
    import shutil
    from efficientq_b200 import entrance, ptq_seg
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    seen = []
    monkeypatch.setattr(ptq_seg, "do_ptq", lambda args, mc, dc, tester, snap, dist=None: seen.append(tester) or {})
    base = ["ptq", "--qlvl_w", "16", "--qlvl_a", "16", "--config", os.path.join(root, "config", "lits_ptq.yaml"),
            "--data_dir", "synthetic", "--exp_id", "pytest_wiring", "--device", "0"]
    snap = os.path.join(root, "exp_ptq", "lits", "snap", "round1", "pytest_wiring")
    try:
        res = entrance.main(base)
        assert isinstance(seen[-1], E.PTQTester) and res["eval"] == {}
        t = seen[-1]
        assert (t.patch_size, t.overlap, t.num_mo, t.n_class, t.label_tfm, t.fusetype) == \
            ((128, 128, 64), (16, 16, 16), 3, 3, None, None)                      # LiTS: arg-max over three classes
        assert os.path.samefile(t.root, snap)
        entrance.main(base + ["--no_test"])
        assert seen[-1] is None
        entrance.main(base + ["--no_test", "--test_fp"])
        assert isinstance(seen[-1], E.PTQTester)
        n = len(seen)
        with pytest.raises(NotImplementedError):
            entrance.main(base + ["--save_nii"])
        assert len(seen) == n
    finally:
        shutil.rmtree(snap, ignore_errors=True)
