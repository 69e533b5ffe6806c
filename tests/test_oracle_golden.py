"""Pin the CPU oracle against fixtures produced by the reference itself
(tests/golden/make_golden.py).  Integer codes / level values must be bit-exact;
floating-point linear algebra is compared with tight tolerances because BLAS
and reduction blocking differ between hosts."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import effq_oracle as O

LEVEL_CASES = [(4, 0, 1), (16, 0, 1), (256, 0, 1), (4, -1, 1), (16, -1, 1), (256, -1, 1)]


@pytest.mark.parametrize("L,lo,hi", LEVEL_CASES)
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_discretize_bit_exact(golden, L, lo, hi, dt):
    g = golden("discretize.npz")
    v = torch.from_numpy(g[f"L{L}_lo{lo}_{dt}_in"])
    want = g[f"L{L}_lo{lo}_{dt}_out"]
    got = O.discretize(v, L, lo, hi).numpy()
    assert got.dtype == want.dtype
    assert np.array_equal(got, want)
    codes = O.discretize_codes(v, L, lo, hi).numpy()
    assert codes.min() >= 0 and codes.max() <= L - 1
    delta = (hi - lo) / (L - 1)
    back = (torch.from_numpy(codes).to(v.dtype) * delta + lo).numpy()
    assert np.array_equal(back, want)


@pytest.mark.parametrize("L", [4, 16, 256])
def test_fakequant_module_bit_exact(golden, L):
    g = golden("fakequant_module.npz")
    x = torch.from_numpy(g[f"L{L}_x"])
    w = torch.from_numpy(g[f"L{L}_w"])
    a_act = torch.tensor(g[f"L{L}_alpha_act"])
    a_w = torch.tensor(g[f"L{L}_alpha_w"])
    assert np.array_equal(O.quantize_act(x, a_act, L).numpy(), g[f"L{L}_qact"])
    qw = O.quantize_w(w, a_w, L)
    assert np.array_equal(qw.numpy(), g[f"L{L}_qw"])
    wi = O.weight_to_int(qw, a_w, L)
    assert wi.dtype == torch.uint8
    assert np.array_equal(wi.numpy(), g[f"L{L}_wint"])
    assert np.array_equal(O.int_to_weight(wi, a_w, L).numpy(), g[f"L{L}_wrestored"])


@pytest.mark.parametrize("name,lo,hi", [("act", 0, 1), ("wt", -1, 1)])
@pytest.mark.parametrize("L", [4, 16, 256])
def test_project_by_iter(golden, name, lo, hi, L):
    g = golden("project_by_iter.npz")
    v = torch.from_numpy(g[name])
    a, b, passes = O.project_by_iter(v, L, lo, hi, return_iters=True)
    assert abs(a - float(g[f"{name}_L{L}_a"])) <= 1e-12 * max(1.0, abs(a))
    assert passes == int(g[f"{name}_L{L}_passes"])
    assert np.array_equal(b.numpy(), g[f"{name}_L{L}_b"])


@pytest.mark.parametrize("name", ["k3s1p1", "k3s2p1", "k1s1p0"])
def test_im2col_and_normal_equations(golden, name):
    g = golden("solver.npz")
    k, s, p = [int(t) for t in g[f"{name}_geom"]]
    x = torch.from_numpy(g[f"{name}_x"])
    cols = O.im2col(x, k, k, k, s, p)
    assert np.array_equal(cols.numpy(), g[f"{name}_cols"])        # pure gather: exact
    y = torch.from_numpy(g[f"{name}_y"])
    att = torch.from_numpy(g[f"{name}_att"])
    w0 = torch.from_numpy(g[f"{name}_w0"])
    b0 = torch.from_numpy(g[f"{name}_b0"])
    ne = O.NormalEquations(x, y, (k, k, k), s, p, w0, b0, att)
    np.testing.assert_allclose(ne.a0.numpy(), g[f"{name}_A0"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(ne.b0.numpy(), g[f"{name}_B0"], rtol=1e-5, atol=1e-4)
    ws, bs = ne.solve(3.0, 0.7, torch.from_numpy(g[f"{name}_G"]))
    np.testing.assert_allclose(ws.numpy(), g[f"{name}_wstar"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(bs.numpy(), g[f"{name}_bstar"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", ["w4a4_k3", "w2a2_k3", "w4a4_k1", "first_k3s2", "w4a4_k1_c64"])
def test_admm_layer(golden, name):
    g = golden("layers_wide.npz" if "_c" in name else "layers.npz")
    k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
    x = torch.from_numpy(g[f"{name}_x"])
    w = torch.from_numpy(g[f"{name}_w"])
    b = torch.from_numpy(g[f"{name}_b"])
    y = torch.from_numpy(g[f"{name}_y"])
    att = torch.from_numpy(g[f"{name}_att"])
    pyr = [torch.ones(x.shape[0], 3, 3, 3), att]
    r = O.admm_layer(x, w, b, y, s, p, lw, la, bool(qa), pyr)
    hist = g[f"{name}_out_hist"]
    # same library, same op order: the whole 200-iterate loss trajectory matches
    np.testing.assert_allclose(np.array(r.loss_history), hist, rtol=2e-3)
    assert abs(r.final_loss - float(g[f"{name}_out_final"])) <= 1e-3 * float(g[f"{name}_out_final"])
    assert abs(r.alpha_w - float(g[f"{name}_out_alpha_w"])) <= 1e-3 * abs(r.alpha_w)
    if qa:
        assert abs(r.alpha_act - float(g[f"{name}_out_alpha_act"])) <= 1e-6 * abs(r.alpha_act)


def test_mask_pyramid_matches_reference(golden):
    """att map quirk: the reference builds the mask with ones_like(int pred), so the
    class weights are truncated to integers (ptqer.py:160-163)."""
    g = golden("toy_net.npz")
    torch.manual_seed(0)
    out = torch.randn(1, 2, 3, 16, 16, 16)
    body = torch.rand(2, 16, 16, 16) > 0.3
    wmap = {0: 1.0, 1: 2.7, 2: 3.2, 3: 1.9}
    pyr = O.mask_pyramid(out, body, wmap, (2, 2, 2), num_lvls=3, task="brats")
    assert [tuple(p.shape) for p in pyr] == [(2, 8, 8, 8), (2, 4, 4, 4), (2, 2, 2, 2)]
    vals = torch.unique(torch.cat([p.flatten() for p in pyr]))
    assert set(vals.tolist()) <= {1.0, 2.0, 3.0}
    assert g["pyr0"].max() >= 1


@pytest.mark.parametrize("L", [4, 16, 256])
def test_ste_gradients_match_reference_autograd(golden, L):
    """a15: closed-form STE gradients vs torch autograd through the reference's own discretize
    (clamp boundaries and rounding ties included)."""
    g = golden("toy_tune.npz")
    x, go = torch.from_numpy(g[f"ste{L}::x"]), torch.from_numpy(g[f"ste{L}::g"])
    gx, ga = O.ste_grads(x, 1.37, L, go)
    assert torch.equal(gx, torch.from_numpy(g[f"ste{L}::grad_x"]))
    assert abs(ga - float(g[f"ste{L}::grad_alpha"])) <= 2e-5 * abs(ga) + 1e-5     # reference sums in fp32


def test_adam_update_matches_torch(golden):
    g = golden("toy_tune.npz")
    p = torch.from_numpy(g["adam::p0"]).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for k in range(3):
        p, m, v = O.adam_update(p, torch.from_numpy(g["adam::grads"][k]), m, v, k + 1)
        np.testing.assert_allclose(p.numpy(), g["adam::traj"][k], rtol=2e-7, atol=0)


@pytest.mark.parametrize("task", ["brats", "lits"])
def test_att_map_and_pyramid_match_reference_on_toy_nets(golden, task):
    """a12 on real network outputs: class voxel counts, class weights and the means of the five pyramid
    masks the REFERENCE computed on the toy nets (BraTS: nested sigmoid labels, body = non-zero voxels;
    LiTS: softmax/argmax, all-ones body, init_stride 2,2,1) vs the oracle on the same FP output.  The FP
    forward of the host mirror is stock PyTorch, so it runs on the CPU."""
    import torch.nn as nn
    from efficientq_b200 import fold_bn, model_blk, qconv, synth
    from tests.golden.make_golden import TOY, TOY_LITS
    cfg = TOY if task == "brats" else TOY_LITS
    g = golden("toy_net.npz" if task == "brats" else "toy_net_lits.npz")
    hetero = {"drop_cut_thres": 128, "ds_depth_limit": 3, "aniso_pool_depth": 9999, "aniso_pool_stride": (2, 2, 1)}
    model = model_blk.UResQ(qconv.EfficientQConv, cfg["num_mod"], cfg["num_classes"], depth_config=cfg["depth"],
                            width_config=cfg["width"], dilation_config=cfg["dilation"], init_stride=cfg["init_stride"],
                            stride=2, drop_rate=cfg["drop_rate"], nla=model_blk.ReLU(True), bn=nn.BatchNorm3d,
                            ds=cfg["ds"], blk_type=cfg["blk"], q_weight=True, qlvl=cfg["qlvl"], q_act=True,
                            qlvl_act=cfg["qlvl_act"], q_first=cfg["q_first"], q_last=cfg["q_last"],
                            hetero_param=hetero, fuse_bn=True, save_mem=True, init_kernel=3)
    model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    model.eval()
    fold_bn.search_fold_and_remove_bn(model)
    size = cfg["size"] if isinstance(cfg["size"], tuple) else (cfg["size"],) * 3
    data = synth.batch(cfg["n"], 0, cfg["num_mod"], size, cfg["task"])
    assert abs(data.double().sum().item() - float(g["data_checksum"])) < 1e-6
    with torch.no_grad():
        out_fp = model(data)
    assert abs(out_fp.double().sum().item() - float(g["out_fp_sum"])) <= 1e-5 * abs(float(g["out_fp_sum"])) + 1e-3
    body = (data[:, 0] != 0.0) if task == "brats" else torch.ones_like(data[:, 0]).bool()
    wmap, nums = O.att_weight_map(out_fp, torch.ones_like(data[:, 0]).bool(), 0.5, task)   # reference passes all ones
    assert nums == [int(v) for v in g["class_nums"]]
    np.testing.assert_allclose([wmap[k] for k in sorted(wmap)], g["wmap"], rtol=1e-12)
    pyr = O.mask_pyramid(out_fp, body, wmap, cfg["init_stride"], 5, task)
    np.testing.assert_allclose([p.mean().item() for p in pyr], g["pyr_means"], rtol=1e-6)
    assert np.array_equal(pyr[0].numpy().astype(np.uint8), g["pyr0"])
