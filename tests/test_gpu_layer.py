"""GPU parity at layer and network level against fixtures produced by the reference's own
EfficientQConv.ptq / do_ptq sequence (tests/golden/make_golden.py).

Contract (BASELINE.json north_star): per-layer reconstruction loss within 1e-3 relative.
The ADMM trajectory is chaotic in the last bits (a single flipped code changes the next
solve), so the comparison is on the losses, not on individual codes; scales are compared
with the tolerance of the reference's own 1e-5 stopping rule."""
import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def engine_mod():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from efficientq_b200 import layer_engine, ops
    ops.capi.load()
    return layer_engine


WIDE = ["w4a4_k3_c32", "w4a4_k1_c64", "w2a4_k3_c64"]       # >= 32 input channels: e4m3 operand path


@pytest.mark.parametrize("name", ["w4a4_k3", "w2a2_k3", "w4a4_k1", "first_k3s2"] + WIDE)
@pytest.mark.parametrize("generic", [False, True])
def test_layer_calibration_matches_reference(engine_mod, golden, name, generic):
    if generic and name in WIDE[1:]:
        pytest.skip("generic path covered by the other cases")
    g = golden("layers_wide.npz" if name in WIDE else "layers.npz")
    k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
    x = torch.from_numpy(g[f"{name}_x"]).to(DEV)
    w = torch.from_numpy(g[f"{name}_w"]).to(DEV)
    b = torch.from_numpy(g[f"{name}_b"]).to(DEV)
    y = torch.from_numpy(g[f"{name}_y"]).to(DEV)
    att = torch.from_numpy(g[f"{name}_att"]).to(DEV)
    pyr = [torch.ones(x.shape[0], 3, 3, 3, device=DEV), att]
    eng = engine_mod.LayerCalibrator(torch.device(DEV), keep_history=True, force_generic=generic)
    wq, bq, a_w, a_act, out_q, rep = eng.run(x, w, b, y, s, p, lw, la, bool(qa), pyr, name=name)
    ref_final = float(g[f"{name}_out_final"])
    ref_hist = g[f"{name}_out_hist"]
    hist = np.array(rep.history)
    print(name, "tc" if rep.used_tc else "generic", "final", rep.final_loss, "ref", ref_final,
          "best_it", rep.best_iter, "min hist", hist.min(), "ref min", ref_hist.min())
    assert rep.used_tc == ((not generic) and name != "first_k3s2")
    # first iterate: identical problem, no trajectory divergence yet -> tight
    # (w2a4_k3_c64: 4 weight levels, K' = 1729 -- the reference's own first iterate moves by 1.7e-4
    #  between 1 and 8 CPU threads, profiles/r01_parity.txt)
    assert abs(hist[0] - ref_hist[0]) <= (5e-4 if name == "w2a4_k3_c64" else 1e-4) * ref_hist[0]
    # final / best loss: the north-star bar is 1e-3 relative, but the reference algorithm itself
    # moves by 1.1e-3 (final) / 1.9e-3 (best) on this very fixture when it is run with 1 thread
    # instead of 8 or when its inputs are perturbed by 1e-7 (profiles/r01_parity.txt): the ADMM
    # trajectory is chaotic in the last bits.  3e-3 is the tightest bar the reference meets
    # against itself.
    # The K' >= 865 wide fixtures move by up to 5.2e-3 between 1 and 8 CPU threads -> 1e-2 there.
    tol = 1e-2 if name in ("w4a4_k3_c32", "w2a4_k3_c64") else 3e-3
    assert abs(rep.final_loss - ref_final) <= tol * ref_final
    assert abs(hist.min() - ref_hist.min()) <= tol * ref_hist.min()
    if qa:
        assert abs(rep.alpha_act - float(g[f"{name}_out_alpha_act"])) <= 1e-6 * rep.alpha_act
    # alpha_w is the LAST iterate's scale (reference quirk); with 256 levels the late, large-rho
    # iterates settle in different basins, so only the 4/16-level cases are compared tightly
    # (wide k=3 fixtures: the reference's own alpha_w moves by 3.6e-3 / 2.6e-3 between 1 and 8 threads)
    tol_aw = (1e-2 if name in ("w4a4_k3_c32", "w2a4_k3_c64") else 2e-3) if lw <= 16 else 1e-1
    assert abs(rep.alpha_w - float(g[f"{name}_out_alpha_w"])) <= tol_aw * rep.alpha_w
    # returned tensors are consistent: weight is on the level grid of SOME scale, output = conv(qact, w)+b
    lv = torch.unique(wq)
    assert lv.numel() <= lw
    assert rep.factorizations == 5


@pytest.mark.parametrize("name", WIDE)
def test_layer_calibration_e4m3_equals_bf16(engine_mod, golden, name, monkeypatch):
    """The e4m3 and bf16 operand paths compute the same exact integer sums: the whole 200-iterate
    loss history, the scales and the calibrated weights must be bit-identical."""
    from efficientq_b200 import ops
    g = golden("layers_wide.npz")
    k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
    args = [torch.from_numpy(g[f"{name}_{t}"]).to(DEV) for t in ("x", "w", "b", "y")]
    att = torch.from_numpy(g[f"{name}_att"]).to(DEV)
    pyr = [torch.ones(args[0].shape[0], 3, 3, 3, device=DEV), att]
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("EFFQ_FP8", flag)
        ops.timer.enabled, ops.timer.records = True, {}
        eng = engine_mod.LayerCalibrator(torch.device(DEV), keep_history=True)
        res[flag] = eng.run(args[0], args[1], args[2], args[3], s, p, lw, la, bool(qa), pyr, name=name)
        names = set(ops.timer.records.keys())
        ops.timer.enabled = False
        ops.timer.reset()
        assert any(n.endswith("_e4m3") for n in names) == (flag == "1"), names
    (wq1, bq1, aw1, aa1, o1, r1), (wq0, bq0, aw0, aa0, o0, r0) = res["1"], res["0"]
    assert r1.history == r0.history
    assert torch.equal(wq1, wq0) and torch.equal(bq1, bq0) and torch.equal(o1, o0)
    assert float(aw1) == float(aw0) and float(aa1) == float(aa0)


@pytest.mark.parametrize("name", ["w4a4_k3_c32", "w2a4_k3_c64", "w4a4_k3", "w4a4_k1_c64", "w4a4_k1"])
def test_conv_free_scoring_matches_conv_scoring(engine_mod, golden, name, monkeypatch):
    """Quantised 3x3x3 layers score iterates 1..199 from the residual statistics (csrc/quadform.cu) instead of a
    conv per iterate: the loss history must agree with the conv-scored run to fp32-loss accuracy and the same
    iterate must win (so weights, scales and the final loss are identical)."""
    g = golden("layers_wide.npz" if name in WIDE else "layers.npz")
    k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
    x, w, b, y, att = [torch.from_numpy(g[f"{name}_{t}"]).to(DEV) for t in ("x", "w", "b", "y", "att")]
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("EFFQ_QF", flag)
        eng = engine_mod.LayerCalibrator(torch.device(DEV), keep_history=True)
        res[flag] = eng.run(x, w, b, y, s, p, lw, la, bool(qa), [att], name=name)
    (wq1, bq1, aw1, _, o1, r1), (wq0, bq0, aw0, _, o0, r0) = res["1"], res["0"]
    h1, h0 = np.array(r1.history), np.array(r0.history)
    rel = np.abs(h1 - h0) / h0
    print(name, "max rel diff of the loss history", rel.max(), "best iter", r1.best_iter, r0.best_iter)
    assert h1[0] == h0[0]                              # iterate 0 is conv-scored in both
    # fp32-loss accuracy: the conv-scored loss itself carries ~1e-6 (fp32 epilogue, fp32 mean), the residual
    # statistics T = R X^T ~1e-6 (fp32 accumulation chains of the tcgen05 Gram kernel)
    assert rel.max() <= 1e-5
    assert r1.best_iter == r0.best_iter or abs(h0[r1.best_iter] - h0[r0.best_iter]) <= 1e-5 * h0.min()
    if r1.best_iter == r0.best_iter:
        assert torch.equal(wq1, wq0) and torch.equal(bq1, bq0) and torch.equal(o1, o0)
        assert r1.final_loss == r0.final_loss
    assert float(aw1) == float(aw0)


@pytest.mark.parametrize("name", ["w4a4_k3_c32", "w4a4_k3", "w4a4_k1_c64"])
def test_per_output_channel_weight_scales(engine_mod, golden, name):
    """Extension `w_per_channel` (north star: "weights per output channel"; the reference is per-tensor): every output
    channel's row of w* + dual gets its own project_by_iter.  Checked against the oracle's restatement with the same
    per-row projection: first iterate, the per-channel scales after 20 iterations, the layer output recomputed by
    the oracle's conv from the GPU's weights; and it must not be worse than the per-tensor calibration."""
    import torch.nn.functional as F
    from oracle import effq_oracle as O
    g = golden("layers_wide.npz" if name in WIDE else "layers.npz")
    k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
    x, w, b, y, att = [torch.from_numpy(g[f"{name}_{t}"]) for t in ("x", "w", "b", "y", "att")]
    n_it = 20
    ref = O.admm_layer(x, w, b, y, s, p, lw, la, bool(qa), [att], n_iter=n_it, channel_wise=True, keep_qact=True)
    eng = engine_mod.LayerCalibrator(torch.device(DEV), n_iter=n_it, keep_history=True)
    dv = [t.to(DEV) for t in (x, w, b, y, att)]
    wq, bq, a_w, a_act, out_q, rep = eng.run(dv[0], dv[1], dv[2], dv[3], s, p, lw, la, bool(qa), [dv[4]], name=name,
                                             channel_wise=True)
    assert a_w.shape == (w.shape[0],)
    hist = np.array(rep.history[:n_it])
    print(name, "iterate 0", hist[0], ref.loss_history[0], "final", rep.final_loss, ref.final_loss)
    # one fixed-point search per row: every row has its own chance that the 1e-5 stopping rule of project_by_iter
    # ends a pass earlier or later under fp32 solve noise (tests/test_gpu_parity.py::test_one_step_ahead), so the
    # first iterate agrees to ~1e-3 here, not to the 2e-4 of the per-tensor search
    assert abs(hist[0] - ref.loss_history[0]) <= 3e-3 * ref.loss_history[0]
    assert abs(rep.final_loss - ref.final_loss) <= 2e-2 * ref.final_loss
    # every row of the returned weights lies on its own L-level grid, and the output is conv(qact, W, b) of exactly these
    wq_c = wq.cpu()
    assert all(torch.unique(wq_c[r]).numel() <= lw for r in range(wq_c.shape[0]))
    out_o = F.conv3d(ref.qact, wq_c, bq.cpu(), s, p)
    assert float((out_q.cpu() - out_o).abs().max() / out_o.abs().max()) <= 1e-5
    # per-channel scales can only help the projection: not worse than per-tensor on the same layer (5 % slack: 20 iterations)
    eng2 = engine_mod.LayerCalibrator(torch.device(DEV), n_iter=n_it, keep_history=True)
    _, _, _, _, _, rep_t = eng2.run(dv[0], dv[1], dv[2], dv[3], s, p, lw, la, bool(qa), [dv[4]], name=name)
    print("per-channel", rep.final_loss, "per-tensor", rep_t.final_loss)
    assert rep.final_loss <= 1.05 * rep_t.final_loss


def test_per_channel_module_roundtrip(engine_mod):
    """Module level: calibrate with lwq_channel_wise, alpha_w becomes [C2]; store_int_weight / restore_fp_weight round
    trip; the deployment forward (tcgen05 conv with one scale per output channel) equals conv(quantize_act(x), W)."""
    import torch.nn.functional as F
    from efficientq_b200.qconv import EfficientQConv
    from oracle import effq_oracle as O
    torch.manual_seed(3)
    c1, c2 = 32, 32
    m = EfficientQConv(c1, c2, 3, 1, 1, bias=True, q_weight=True, qlvl=16, q_act=True, qlvl_act=16, lwq_channel_wise=True)
    m.lwq_iter = 10
    x = torch.relu(torch.randn(2, c1, 8, 16, 8)) * 1.1
    y = F.conv3d(x, m.weight.data, m.bias.data, 1, 1)
    m.to(DEV)
    m.name, m.layer_loss, m.output_fp = "pc", [], y.to(DEV)
    m.set_quantizing()
    with torch.no_grad():
        out_cal = m(x.to(DEV))
    assert m.alpha_w.shape == (c2,)
    wq = m.weight.data.clone()
    m.set_quantized()
    with torch.no_grad():
        out_dep = m(x.to(DEV))
    assert m._wcodes_cache[1] is not None and m._wcodes_cache[1][1].numel() == c2        # tcgen05 path, per-channel grid
    want = F.conv3d(O.quantize_act(x, m.alpha_act.data.cpu(), 16).double(), wq.cpu().double(), m.bias.data.cpu().double(), 1, 1).float()
    torch.testing.assert_close(out_dep.cpu(), want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out_cal.cpu(), want, rtol=1e-5, atol=1e-5)
    # integer export with per-channel scales.  alpha_w is the LAST iterate's scale (reference quirk), the weights the
    # best iterate's: re-derive the grid from the weights for an exact round trip
    m.alpha_w.data = m._wcodes_cache[1][1].clone()
    m.cpu()
    m.store_int_weight()
    assert m.weight.data.dtype == torch.uint8 and int(m.weight.data.max()) <= 15
    m.restore_fp_weight()
    torch.testing.assert_close(m.weight.data, wq.cpu(), rtol=1e-6, atol=1e-7)


def build_toy(task="brats"):
    from efficientq_b200 import model_blk, qconv
    from tests.golden.make_golden import TOY, TOY_LITS
    cfg = TOY if task == "brats" else TOY_LITS
    hetero = {"drop_cut_thres": 128, "ds_depth_limit": 3, "aniso_pool_depth": 9999, "aniso_pool_stride": (2, 2, 1)}
    return model_blk.UResQ(qconv.EfficientQConv, cfg["num_mod"], cfg["num_classes"], depth_config=cfg["depth"],
                           width_config=cfg["width"], dilation_config=cfg["dilation"], init_stride=cfg["init_stride"],
                           stride=2, drop_rate=cfg["drop_rate"], nla=model_blk.ReLU(True), bn=nn.BatchNorm3d,
                           ds=cfg["ds"], blk_type=cfg["blk"], q_weight=True, qlvl=cfg["qlvl"], q_act=True,
                           qlvl_act=cfg["qlvl_act"], q_first=cfg["q_first"], q_last=cfg["q_last"],
                           hetero_param=hetero, fuse_bn=True, save_mem=True, init_kernel=3), cfg


@pytest.mark.parametrize("task", ["brats", "lits"])
def test_toy_network_matches_reference(engine_mod, golden, task):
    """do_ptq core on a miniature of the BraTS config (4 modalities, cubic, W4A4) and of the LiTS config
    (1 CT channel, init_stride 2,2,1, anisotropic volume, W2A2, softmax prediction, all-ones body mask)."""
    from efficientq_b200 import fold_bn, ptqer, synth
    g = golden("toy_net.npz" if task == "brats" else "toy_net_lits.npz")
    model, cfg = build_toy(task)
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    model.load_state_dict(sd, strict=False)
    model.eval()
    fold_bn.search_fold_and_remove_bn(model)
    model.to(DEV)
    size = cfg["size"] if isinstance(cfg["size"], tuple) else (cfg["size"],) * 3
    data = synth.batch(cfg["n"], 0, cfg["num_mod"], size, cfg["task"])
    assert abs(data.double().sum().item() - float(g["data_checksum"])) < 1e-6
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = ptqer.calibrate(model, data.to(DEV), task, ",".join(str(v) for v in cfg["init_stride"]))
    assert res["class_nums"] == [int(v) for v in g["class_nums"]]
    np.testing.assert_allclose([p.mean().item() for p in res["pyramid"]], g["pyr_means"], rtol=1e-6)
    names = [ln.rsplit(":", 1)[0].strip() for ln in res["layer_loss"]]
    losses = np.array([float(ln.rsplit(":", 1)[1]) for ln in res["layer_loss"]])
    assert names == [str(s) for s in g["layer_names"]]
    ref = g["layer_losses"]
    lines = [f"{nm:45s} ours {a:.6e} ref {b:.6e} rel {abs(a - b) / b:.2e}" for nm, a, b in zip(names, losses, ref)]
    print("\n".join(lines))
    import os
    if os.path.isdir("gpurun_out"):
        with open("gpurun_out/toy_net_parity.txt" if task == "brats" else "gpurun_out/toy_net_lits_parity.txt", "w") as fid:
            fid.write("\n".join(lines) + f"\nt_fp {res['t_fp']:.3f}s t_ptq {res['t_ptq']:.3f}s\n")
    # Each layer calibrates on the output of the already-quantised prefix, and the ADMM trajectory amplifies
    # last-bit differences, so differences compound through the network (per-layer parity WITHOUT the compounding
    # is tests/test_gpu_parity.py::test_teacher_forced_layer_parity, at 1e-3).  Here the bar is the reference
    # against ITSELF: 8 reference runs from volumes perturbed by 1e-7 and one single-threaded run
    # (``ensemble_losses``); the GPU result must lie within the ensemble's range widened by one range on either
    # side (a ninth sample of the same distribution falls outside the bare range of eight 2 times out of 9),
    # and within 1e-3 wherever the reference does not move at all.
    ens = np.concatenate([g["ensemble_losses"], ref[None]], 0)
    lo, hi = ens.min(0), ens.max(0)
    width = hi - lo
    ok = (losses >= lo - width - 1e-3 * ref) & (losses <= hi + width + 1e-3 * ref)
    for nm, a, l_, h_, o in zip(names, losses, lo, hi, ok):
        print(f"{nm:45s} ours {a:.6e} reference ensemble [{l_:.6e}, {h_:.6e}] {'ok' if o else 'OUTSIDE'}")
    assert abs(losses[0] - ref[0]) <= 1e-3 * ref[0]
    assert ok.all(), [n for n, o in zip(names, ok) if not o]


@pytest.mark.parametrize("c1,c2,k", [(32, 32, 3), (64, 32, 1), (16, 48, 3)])
def test_quantized_forward_matches_reference_semantics(engine_mod, c1, c2, k):
    """Deployment forward (PTQConv.forward, _quantized branch, PTQConv.py:163-167) on the
    tensor-core code path vs the oracle's  conv3d(quantize_act(x), qweight, bias)."""
    import torch.nn.functional as F
    from efficientq_b200.qconv import EfficientQConv
    from oracle import effq_oracle as O
    torch.manual_seed(c1 + k)
    m = EfficientQConv(c1, c2, k, 1, (k - 1) // 2, bias=True, q_weight=True, qlvl=16, q_act=True, qlvl_act=16)
    a_w, a_act = torch.tensor(0.21), torch.tensor(1.7)
    qw = O.quantize_w(torch.randn(c2, c1, k, k, k) * 0.1, a_w, 16)
    m.weight.data, m.alpha_w.data, m.alpha_act.data = qw.clone(), a_w.clone(), a_act.clone()
    m.bias.data = torch.randn(c2) * 0.1
    x = torch.relu(torch.randn(2, c1, 6, 16, 8)) * 1.2
    want = F.conv3d(O.quantize_act(x, a_act, 16).double(), qw.double(), m.bias.data.double(), 1, (k - 1) // 2).float()
    m.to(DEV)
    m.set_quantized()
    with torch.no_grad():
        got = m(x.to(DEV))
    assert m._wcodes_cache is not None and m._wcodes_cache[1] is not None      # the tcgen05 path was taken
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-5)


def test_quantized_forward_uses_stored_weights_not_alpha_w(engine_mod):
    """Reference quirk (EfficientQConv.py:155-158): after calibration alpha_w is the LAST iterate's scale
    while the weights are the BEST iterate's, so alpha_w does not describe the stored weights.  The
    reference's quantized forward convolves with the stored weights as they are (PTQConv.py:163-167); the
    tensor-core path must therefore recover the grid scale from the weights (found by the tune parity test:
    0.4 % loss error when alpha_w was used)."""
    import torch.nn.functional as F
    from efficientq_b200.qconv import EfficientQConv
    from oracle import effq_oracle as O
    torch.manual_seed(5)
    c1, c2, k = 32, 32, 3
    m = EfficientQConv(c1, c2, k, 1, 1, bias=True, q_weight=True, qlvl=16, q_act=True, qlvl_act=16)
    a_best, a_last, a_act = torch.tensor(0.2137), torch.tensor(0.2101), torch.tensor(1.7)
    qw = O.quantize_w(torch.randn(c2, c1, k, k, k) * 0.1, a_best, 16)
    m.weight.data, m.alpha_w.data, m.alpha_act.data = qw.clone(), a_last.clone(), a_act.clone()
    m.bias.data = torch.randn(c2) * 0.1
    x = torch.relu(torch.randn(2, c1, 6, 16, 8)) * 1.2
    want = F.conv3d(O.quantize_act(x, a_act, 16).double(), qw.double(), m.bias.data.double(), 1, 1).float()
    m.to(DEV)
    m.set_quantized()
    with torch.no_grad():
        got = m(x.to(DEV))
    assert m._wcodes_cache[1] is not None                   # tcgen05 path, scale recovered from the weights
    assert abs(float(m._wcodes_cache[1][1]) - float(a_best)) <= 1e-6 * float(a_best)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-5)
    # weights that are on no L-level grid at all -> generic fp32 path, still the stored weights
    m.weight.data = (qw * (1 + 0.01 * torch.randn_like(qw))).to(DEV)
    with torch.no_grad():
        got2 = m(x.to(DEV))
    assert m._wcodes_cache[1] is None
    want2 = F.conv3d(O.quantize_act(x, a_act, 16).double(), m.weight.data.cpu().double(), m.bias.data.cpu().double(), 1, 1).float()
    torch.testing.assert_close(got2.cpu(), want2, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["w4a4_k3_c32", "w4a4_k1_c64", "w4a4_k1", "first_k3s2"])
def test_replayed_iterations_are_bit_identical(engine_mod, golden, name, monkeypatch):
    """The steady-state iterations of a rho block are re-issued from a recorded launch sequence
    (capi.record / ops.replay): same kernels, same arguments -> the whole loss history, the scales and the
    calibrated weights must equal the run that goes through the Python wrappers every iteration."""
    g = golden("layers_wide.npz" if name in WIDE else "layers.npz")
    k, s, p, lw, la, qa = [int(t) for t in g[f"{name}_cfg"]]
    x, w, b, y, att = [torch.from_numpy(g[f"{name}_{t}"]).to(DEV) for t in ("x", "w", "b", "y", "att")]
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("EFFQ_REPLAY", flag)
        eng = engine_mod.LayerCalibrator(torch.device(DEV), keep_history=True)
        res[flag] = eng.run(x, w, b, y, s, p, lw, la, bool(qa), [att], name=name)
    (wq1, bq1, aw1, aa1, o1, r1), (wq0, bq0, aw0, aa0, o0, r0) = res["1"], res["0"]
    assert r1.history == r0.history and r1.best_iter == r0.best_iter
    assert torch.equal(wq1, wq0) and torch.equal(bq1, bq0) and torch.equal(o1, o0)
    assert float(aw1) == float(aw0)


def test_final_dice_on_trained_miniature(engine_mod, golden):
    """North star: "final Dice within 0.1 points".  A BraTS miniature trained with stock PyTorch on the synthetic
    nested-sphere volumes (tests/golden/make_golden.py::gen_toy_dice), calibrated at W4A4 on 2 volumes, Dice of
    the quantised model on 4 held-out volumes (deployment forward on the tcgen05 code path).  The REFERENCE's own
    quantised Dice moves by 0.7 points (foreground mean) / 2.0 points (class 3) between a 1-thread and an 8-thread
    CPU run from the same trained state, so the bar is the reference's own ensemble (below); the FP Dice
    (no calibration involved) must agree to 0.01 points."""
    from efficientq_b200 import fold_bn, ptqer, synth
    from tests.golden.make_golden import dice_table
    g = golden("toy_dice.npz")
    model, cfg = build_toy("brats")
    model.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}, strict=False)
    model.eval()
    fold_bn.search_fold_and_remove_bn(model)
    model.to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ev = [synth.volume(100 + i, cfg["num_mod"], (cfg["size"],) * 3, "brats") for i in range(4)]
    ev_x = torch.stack([v[0] for v in ev]).to(DEV)
    ev_l = torch.stack([v[1] for v in ev]).long()
    ptqer.set_name(model)
    ptqer.set_fp(model)
    with torch.no_grad():
        dice_fp = dice_table(ptqer.get_pred_brats(model(ev_x)[-1]).cpu(), ev_l)
    np.testing.assert_allclose(dice_fp, g["dice_fp"], atol=1e-4)
    # (1) deterministic: the REFERENCE's calibrated state through the GPU deployment forward (tcgen05 conv on codes)
    # must give the reference's quantised Dice itself -- the north star's 0.1 points
    import copy
    from efficientq_b200.qconv import PTQConv
    replay = copy.deepcopy(model)
    for name, m in replay.named_modules():
        if isinstance(m, PTQConv):
            m.weight.data = torch.from_numpy(g[f"cal::{name}.weight"]).to(DEV)
            m.bias.data = torch.from_numpy(g[f"cal::{name}.bias"]).to(DEV)
            m.alpha_w.data = torch.tensor(float(g[f"cal::{name}.alpha_w"]), device=DEV)
            m.alpha_act.data = torch.tensor(float(g[f"cal::{name}.alpha_act"]), device=DEV)
    ptqer.set_quantized(replay)
    with torch.no_grad():
        dice_replay = dice_table(ptqer.get_pred_brats(replay(ev_x)[-1]).cpu(), ev_l)
    print("Dice of the reference's calibrated state on the GPU forward", np.round(dice_replay, 4).tolist(),
          "reference", np.round(g["dice_q"], 4).tolist())
    np.testing.assert_allclose(dice_replay, g["dice_q"], atol=1e-3)
    del replay
    # (2) calibrated here, twice: with the FP targets from the library's fp32 conv (the kind of conv the reference's
    # targets come from) and from the repo's own FP conv (default; closer to an fp64 conv than the library's, test
    # test_conv3d_fp_matches_fp64).  The two sets of targets differ by 5e-7 of their maximum in the two 16-channel
    # layers -- and that moves the quantised Dice by up to 0.8 points: the GPU pipeline's own sensitivity, of the size
    # of the reference's (1 vs 8 threads).
    data = synth.batch(cfg["n"], 0, cfg["num_mod"], (cfg["size"],) * 3, cfg["task"]).to(DEV)
    import os
    ens = np.concatenate([g["dice_q_ensemble"], g["dice_q"][None]], 0)
    ens = np.concatenate([ens, ens.mean(1, keepdims=True)], 1)
    lo, hi = ens.min(0), ens.max(0)
    width = hi - lo
    print("reference ensemble Dice range", np.round(lo, 4).tolist(), np.round(hi, 4).tolist())
    lines = []
    for fp_conv, widen in (("lib", 1.0), ("own", 2.0)):
        m2 = copy.deepcopy(model)
        os.environ["EFFQ_FP_CONV"] = fp_conv
        try:
            res = ptqer.calibrate(m2, data, "brats", "2,2,2")
        finally:
            os.environ.pop("EFFQ_FP_CONV", None)
        with torch.no_grad():
            dice_q = dice_table(ptqer.get_pred_brats(m2(ev_x)[-1]).cpu(), ev_l)
        lines.append(f"Dice FP {np.mean(dice_fp):.4f} | W4A4 ours (FP conv: {fp_conv}) {np.round(dice_q, 4).tolist()} mean "
                     f"{np.mean(dice_q):.4f} | reference (8 threads) {np.round(g['dice_q'], 4).tolist()} mean {np.mean(g['dice_q']):.4f}")
        print(lines[-1])
        # the reference against itself (8 perturbed runs + 1 single-threaded, tests/golden/make_golden.py::gen_toy_dice):
        # ours must lie in the ensemble's range widened by one range on either side (two for the own FP conv, whose
        # targets are not the library's), per class and for the mean
        ours = np.array(list(dice_q) + [np.mean(dice_q)])
        assert ((ours >= lo - widen * width - 1e-3) & (ours <= hi + widen * width + 1e-3)).all(), (fp_conv, ours, lo, hi)
        del m2
    if os.path.isdir("gpurun_out"):
        open("gpurun_out/toy_dice_parity.txt", "w").write("\n".join(lines) + "\n")
    np.testing.assert_allclose([float(ln.rsplit(":", 1)[1]) for ln in res["layer_loss"]][:1], g["layer_losses"][:1], rtol=1e-3)
