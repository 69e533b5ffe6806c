"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs the read-only reference checkout):

    python tests/golden/make_golden.py [--ref /root/reference]

It imports the unmodified reference modules from ``<ref>/src`` (with inert stub
modules for pytz / matplotlib / nibabel / progressbar, none of which touch the
arithmetic), runs the hot-path functions on seeded inputs on the CPU, and writes
small ``.npz`` files.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so these files are what pins ``oracle/effq_oracle.py``; the
GPU parity tests then compare the CUDA path with the oracle and with these
files.  Nothing at test time reads ``/root/reference``.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference(ref):
    sys.path.insert(0, os.path.join(ref, "src"))
    for name in ["matplotlib", "matplotlib.pyplot", "pytz", "nibabel", "progressbar"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import importlib
    import models  # noqa: F401  (models/__init__ re-exports classes under the module names)
    lh, solver, ptqconv, effq, model_blk, fold_bn, hooks = [
        sys.modules.get(f"models.{n}") or importlib.import_module(f"models.{n}")
        for n in ("layer_helper", "solver", "PTQConv", "EfficientQConv", "model_blk", "fold_bn", "hooks")]
    import ptqer
    return dict(lh=lh, solver=solver, ptqconv=ptqconv, effq=effq, model_blk=model_blk,
                fold_bn=fold_bn, hooks=hooks, ptqer=ptqer)


def meta():
    return dict(torch_version=torch.__version__, numpy_version=np.__version__,
                threads=torch.get_num_threads())


def copying_hook(m, i, o):
    """The reference's forward hook is ``m.output_fp = o.detach().cpu()`` (src/models/hooks.py:5-6).  On the device the
    reference is run on (``--device <gpu id>``) ``.cpu()`` is a COPY, so the stored FP target is the layer's
    pre-activation output.  On CPU tensors ``.cpu()`` is a no-op, the target aliases the conv output, and the next
    unit's ``ReLU(inplace=True)`` (blk: mid, BN folded to identity) overwrites it -- an artefact of running the
    reference on the CPU, not its semantics.  The fixtures therefore register this copying form of the same hook."""
    m.output_fp = o.detach().clone()


class PtqTrace:
    """Observes the unmodified ``EfficientQConv.ptq`` from the outside: a wrapper around the class method records the
    layer's input / target / attention map, ``QuadraSolver.solve`` and ``F.mse_loss`` are wrapped to read the ADMM
    state out of the caller's frame (``ptq``'s local variables G, dual, rho, w_star, b_star, a_w, i) -- the state
    right before the proximal step of iteration i and right after its projection."""

    def __init__(self, R, state_iters=(), keep_inputs=True):
        self.R, self.state_iters, self.keep_inputs = R, set(state_iters), keep_inputs
        self.layers = []                     # one dict per ptq() call, in execution order

    def __enter__(self):
        import sys as _sys
        effq, solver = self.R["effq"], self.R["solver"]
        self._orig = (effq.EfficientQConv.ptq, solver.QuadraSolver.solve, effq.F.mse_loss)
        orig_ptq, orig_solve, orig_mse = self._orig
        tr = self

        def ptq(mod, x):
            rec = dict(name=mod.name, hist=[], states={}, n_iter=mod.lwq_iter)
            if tr.keep_inputs:
                rec["x"] = x.detach().clone()
                rec["y"] = mod.output_fp.detach().clone()
                att = None
                if mod.mask_pyramid:
                    for lv, mask in enumerate(mod.mask_pyramid):
                        if mask.shape[1:] == mod.output_fp.shape[2:]:
                            att = lv
                            break
                rec["att_level"] = -1 if att is None else att
            rec["w0"], rec["b0"] = mod.weight.data.clone(), mod.bias.data.clone()
            tr.layers.append(rec)
            out = orig_ptq(mod, x)
            rec["final"] = float(mod.layer_loss[-1].split(":")[-1])
            rec["alpha_w"], rec["alpha_act"] = float(mod.alpha_w.item()), float(mod.alpha_act.item())
            rec["weight"], rec["bias"] = mod.weight.data.clone(), mod.bias.data.clone()
            return out

        def solve(qs, rho, eta, gd):
            loc = _sys._getframe(1).f_locals
            if "admm_iter" in loc and loc["i"] in tr.state_iters:
                tr.layers[-1]["states"][loc["i"]] = dict(G=loc["G"].detach().clone(), dual=loc["dual"].detach().clone(),
                                                        rho=float(rho), eta=float(eta))
            return orig_solve(qs, rho, eta, gd)

        def mse(a, t, *args, **kw):
            v = orig_mse(a, t, *args, **kw)
            loc = _sys._getframe(1).f_locals
            if "admm_iter" in loc and tr.layers:
                rec = tr.layers[-1]
                i = loc["i"]
                if i < loc["admm_iter"] and len(rec["hist"]) == i:
                    rec["hist"].append(float(v.item()))
                    if i in tr.state_iters:
                        rec["states"][i].update(wstar=loc["w_star"].detach().clone(), bstar=loc["b_star"].detach().clone(),
                                                a_w=float(loc["a_w"]), G_next=loc["G"].detach().clone(), loss=float(v.item()))
            return v
        effq.EfficientQConv.ptq, solver.QuadraSolver.solve, effq.F.mse_loss = ptq, solve, mse
        return self

    def __exit__(self, *exc):
        effq, solver = self.R["effq"], self.R["solver"]
        effq.EfficientQConv.ptq, solver.QuadraSolver.solve, effq.F.mse_loss = self._orig
        return False


LEVEL_CASES = [(4, 0, 1), (16, 0, 1), (256, 0, 1), (4, -1, 1), (16, -1, 1), (256, -1, 1)]


def edge_values(n_lvl, lo, hi, dtype):
    """Rounding ties, clamp edges and near-ties for the level grid."""
    delta = (hi - lo) / (n_lvl - 1)
    ks = np.arange(n_lvl - 1, dtype=np.float64)
    ties = lo + (ks + 0.5) * delta
    near = np.concatenate([np.nextafter(ties.astype(dtype), np.array(np.inf, dtype)),
                           np.nextafter(ties.astype(dtype), np.array(-np.inf, dtype))])
    edges = np.array([lo, hi, lo - 1e-3, hi + 1e-3, lo - 7.0, hi + 7.0, 0.0, -0.0], dtype=np.float64)
    return np.concatenate([ties, near.astype(np.float64), edges]).astype(dtype)


def gen_discretize(R, out):
    lh = R["lh"]
    res = {}
    g = torch.Generator().manual_seed(11)
    base = torch.randn(8192, generator=g) * 0.8
    for (L, lo, hi) in LEVEL_CASES:
        for dt, name in [(torch.float32, "f32"), (torch.float64, "f64")]:
            npdt = np.float32 if dt == torch.float32 else np.float64
            v = torch.cat([base.to(dt), torch.from_numpy(edge_values(L, lo, hi, npdt))])
            q = lh.discretize(v, L, lo, hi)
            key = f"L{L}_lo{lo}_{name}"
            res[key + "_in"] = v.numpy()
            res[key + "_out"] = q.numpy()
    np.savez_compressed(os.path.join(out, "discretize.npz"), **res, **{f"meta_{k}": v for k, v in meta().items()})


def gen_fakequant_module(R, out):
    """PTQConv._quantize_act/_quantize_w/store_int_weight/restore_fp_weight."""
    PTQConv = R["ptqconv"].PTQConv
    res = {}
    g = torch.Generator().manual_seed(12)
    for L in (4, 16, 256):
        m = PTQConv(6, 5, 3, 1, 1, bias=True, q_weight=True, qlvl=L, q_act=True, qlvl_act=L)
        x = torch.relu(torch.randn(2, 6, 9, 10, 11, generator=g)) * 1.7
        w = torch.randn(5, 6, 3, 3, 3, generator=g) * 0.2
        a_act, a_w = 1.2345678, 0.31415927
        m.alpha_act.data = torch.tensor(a_act)
        m.alpha_w.data = torch.tensor(a_w)
        m.weight.data = w.clone()
        with torch.no_grad():
            qa = m._quantize_act(x)
            qw = m._quantize_w()
            m.weight.data = qw.clone()
            m.store_int_weight()
            wi = m.weight.data.clone()
            m.restore_fp_weight()
            wr = m.weight.data.clone()
        res.update({f"L{L}_x": x.numpy(), f"L{L}_w": w.numpy(), f"L{L}_alpha_act": np.float32(a_act),
                    f"L{L}_alpha_w": np.float32(a_w), f"L{L}_qact": qa.numpy(), f"L{L}_qw": qw.numpy(),
                    f"L{L}_wint": wi.numpy(), f"L{L}_wrestored": wr.numpy()})
    np.savez_compressed(os.path.join(out, "fakequant_module.npz"), **res)


def gen_project(R, out):
    lh = R["lh"]
    res = {}
    g = torch.Generator().manual_seed(13)
    act = torch.relu(torch.randn(3, 8, 12, 12, 12, generator=g)) * 0.9
    wt = torch.randn(16, 8, 3, 3, 3, generator=g) * 0.07
    res["act"] = act.numpy()
    res["wt"] = wt.numpy()
    for name, v, lo, hi in [("act", act, 0, 1), ("wt", wt, -1, 1)]:
        for L in (4, 16, 256):
            calls = {"n": 0}
            orig = lh.discretize

            def counting(*a, **k):
                calls["n"] += 1
                return orig(*a, **k)
            lh.discretize = counting
            try:
                a, b = lh.project_by_iter(v, L, lo, hi)
            finally:
                lh.discretize = orig
            res[f"{name}_L{L}_a"] = np.float64(a)
            res[f"{name}_L{L}_passes"] = np.int64(calls["n"] - 1)
            res[f"{name}_L{L}_b"] = b.numpy()
    np.savez_compressed(os.path.join(out, "project_by_iter.npz"), **res)


def gen_solver(R, out):
    solver = R["solver"]
    res = {}
    g = torch.Generator().manual_seed(14)
    # k3_wide: K' = 1297 unknowns (V = 6144 voxels): takes the tensor-core inverse and the split-bf16 GEMM on the GPU
    cases = [("k3s1p1", 2, 5, 7, 3, 1, 1, (6, 7, 8)), ("k3s2p1", 2, 4, 6, 3, 2, 1, (8, 10, 12)),
             ("k1s1p0", 3, 6, 5, 1, 1, 0, (4, 5, 6)), ("k3_wide", 2, 48, 24, 3, 1, 1, (12, 16, 16))]
    for name, n, c1, c2, k, s, p, sp in cases:
        x = torch.relu(torch.randn(n, c1, *sp, generator=g))
        w0 = torch.randn(c2, c1, k, k, k, generator=g) * 0.1
        b0 = torch.randn(c2, generator=g) * 0.1
        y = F.conv3d(x, w0, b0, s, p) + 0.01 * torch.randn(1, generator=g)
        att = (torch.rand(n, *y.shape[2:], generator=g) * 3).floor() + 1.0
        cols = solver.im2col_loop(x.numpy().astype("float32"), k, k, k, s, p)
        qs = solver.QuadraSolver(x, y, k, k, k, (s,) * 3, (p,) * 3, device=torch.device("cpu"),
                                 mu=0, eta=0.7, W0=w0, att=att, b0=b0)
        gmat = w0 + 0.01 * torch.randn(w0.shape, generator=g)
        ws, bs = qs.solve(3.0, 0.7, gmat)
        if cols.size > 2 ** 18:
            cols = cols[:, :64]                                    # wide case: the first 64 columns pin the ordering
        res.update({f"{name}_x": x.numpy(), f"{name}_w0": w0.numpy(), f"{name}_b0": b0.numpy(),
                    f"{name}_y": y.numpy(), f"{name}_att": att.numpy(), f"{name}_cols": cols,
                    # wide case: every 5th row / column of A0 (6.7 MB otherwise)
                    f"{name}_A0": qs.A0.numpy() if qs.A0.numel() <= 2 ** 18 else qs.A0[::5, ::5].contiguous().numpy(),
                    f"{name}_B0": qs.B0.numpy(), f"{name}_G": gmat.numpy(),
                    f"{name}_wstar": ws.numpy(), f"{name}_bstar": bs.numpy(),
                    f"{name}_geom": np.array([k, s, p], dtype=np.int64)})
    np.savez_compressed(os.path.join(out, "solver.npz"), **res)


class Fp64Solve:
    """Ensemble member "the reference with an exact linear solver": ``torch.linalg.solve`` (a library function the
    reference calls at solver.py:331, not reference code) computes in fp64 and rounds the result to fp32.  How far the
    reference moves under it is its sensitivity to solver-level rounding -- the kind of difference any other correct
    fp32 implementation (another LAPACK, cuSOLVER, this repo) has against it; input perturbations of 1e-7 mostly
    vanish in the activation quantiser and do not probe that."""

    def __enter__(self):
        self.orig = torch.linalg.solve
        orig = self.orig
        torch.linalg.solve = lambda a, b, **kw: orig(a.double(), b.double(), **kw).to(a.dtype)
        return self

    def __exit__(self, *exc):
        torch.linalg.solve = self.orig
        return False


def run_ref_layer(R, x, w, b, y, stride, pad, qlvl_w, qlvl_a, q_act, pyramid, state_iters=(), n_iter=None):
    """EfficientQConv.ptq on one layer, recording the per-iteration losses (and the ADMM state at ``state_iters``)."""
    effq = R["effq"]
    c2, c1, k = w.shape[0], w.shape[1], w.shape[2]
    m = effq.EfficientQConv(c1, c2, k, stride, pad, bias=True, q_weight=True, qlvl=qlvl_w,
                            q_act=q_act, qlvl_act=qlvl_a)
    m.weight.data = w.clone()
    m.bias.data = b.clone()
    m.name = "layer"
    m.output_fp = y.clone()
    m.mask_pyramid = pyramid
    m.layer_loss = []
    if n_iter is not None:
        m.lwq_iter = n_iter               # an attribute of the reference module (EfficientQConv.py:23), not a code change
    tr = PtqTrace(R, state_iters=state_iters, keep_inputs=False)
    with tr, torch.no_grad():
        m.ptq(x)
    rec = tr.layers[0]
    res = dict(weight=m.weight.data.numpy(), bias=m.bias.data.numpy(),
               alpha_w=np.float32(m.alpha_w.item()), alpha_act=np.float32(m.alpha_act.item()),
               hist=np.array(rec["hist"], dtype=np.float64), final=np.float64(rec["final"]))
    lm1 = qlvl_w - 1
    for i, st in sorted(rec["states"].items()):
        res[f"st{i}_dual"], res[f"st{i}_wstar"], res[f"st{i}_bstar"] = st["dual"].numpy(), st["wstar"].numpy(), st["bstar"].numpy()
        res[f"st{i}_rho"], res[f"st{i}_a_w"], res[f"st{i}_loss"] = np.float64(st["rho"]), np.float64(st["a_w"]), np.float64(st["loss"])
        res[f"st{i}_G"] = st["G"].numpy()                        # <= L distinct values: compresses well
        nxt = torch.round(st["G_next"] / float(st["a_w"]) * lm1)
        res[f"st{i}_codes_next"] = nxt.numpy().astype(np.int16)
    return res


STATE_ITERS = (0, 1, 2, 49, 50, 51, 100, 150, 199)
STATE_ITERS_WIDE = (0, 1, 50, 51, 199)


LAYER_CASES = [
    # name, n, c1, c2, k, s, p, spatial, Lw, La, q_act
    ("w4a4_k3", 2, 16, 16, 3, 1, 1, (8, 16, 8), 16, 16, True),
    ("w2a2_k3", 1, 16, 32, 3, 1, 1, (8, 16, 8), 4, 4, True),
    ("w4a4_k1", 2, 16, 32, 1, 1, 0, (8, 8, 8), 16, 16, True),
    ("first_k3s2", 2, 4, 16, 3, 2, 1, (16, 16, 16), 256, 256, False),
]
# channel counts of the real networks' inner layers (>= 32): these take the e4m3 operand path.
# Voxel counts are kept well above K' = C1 k^3 + 1 (as in the real layers, V >= 38 K'): with
# V ~ K' the normal equations are near-singular and the reference's own trajectory moves by
# 5e-3 between 1 and 8 threads (profiles/r01_parity.txt).
WIDE_LAYER_CASES = [
    ("w4a4_k3_c32", 2, 32, 32, 3, 1, 1, (8, 16, 16), 16, 16, True),
    ("w4a4_k1_c64", 2, 64, 32, 1, 1, 0, (8, 8, 8), 16, 16, True),
    ("w2a4_k3_c64", 2, 64, 16, 3, 1, 1, (8, 16, 16), 4, 16, True),
]


def gen_layers(R, out, cases=LAYER_CASES, seed=15, fname="layers.npz", state_iters=STATE_ITERS):
    res = {}
    g = torch.Generator().manual_seed(seed)
    for name, n, c1, c2, k, s, p, sp, lw, la, qa in cases:
        x = torch.randn(n, c1, *sp, generator=g)
        if qa:
            x = torch.relu(x)
        w = torch.randn(c2, c1, k, k, k, generator=g) * (2.0 / (c1 * k ** 3)) ** 0.5
        b = torch.randn(c2, generator=g) * 0.05
        y = F.conv3d(x, w, b, s, p)
        att = (torch.rand(n, *y.shape[2:], generator=g) * 3).floor() + 1.0
        pyramid = [torch.ones(n, 3, 3, 3), att]
        r = run_ref_layer(R, x, w, b, y, s, p, lw, la, qa, pyramid, state_iters=state_iters)
        res.update({f"{name}_x": x.numpy(), f"{name}_w": w.numpy(), f"{name}_b": b.numpy(),
                    f"{name}_y": y.numpy(), f"{name}_att": att.numpy(),
                    f"{name}_cfg": np.array([k, s, p, lw, la, int(qa)], dtype=np.int64)})
        res.update({f"{name}_out_{k2}": v for k2, v in r.items()})
        print(name, "final", r["final"], "alpha_w", r["alpha_w"], "alpha_act", r["alpha_act"])
    np.savez_compressed(os.path.join(out, fname), **res)


def gen_layers_wide(R, out):
    gen_layers(R, out, WIDE_LAYER_CASES, seed=17, fname="layers_wide.npz", state_iters=STATE_ITERS_WIDE)


# Real widths of the BraTS net's two deepest levels (K' = 3457 and 6913).  Inputs are regenerated from
# these seeds by the test (torch's CPU generator is deterministic for a given torch version, tests/golden/VERSIONS.txt),
# so only checksums and results are stored.  V ~ 7 K' voxels (the real layers have V >= 38 K'; with V = 2 K' the
# reference's own fp32 LU is only good to ~1e-4 and the comparison measures conditioning, not kernels).
REAL_WIDTH_CASES = [
    # name, channels, spatial, reference iterations, reference runs of the sensitivity ensemble
    ("c128", 128, (24, 32, 32), 200, 2),          # V = 24576 = 7.1 K'
    ("c256", 256, (32, 40, 40), 5, 0),            # V = 51200 = 7.4 K'
]


def real_width_inputs(name, c, sp, perturb=None):
    g = torch.Generator().manual_seed(2000 + c)
    x = torch.relu(torch.randn(1, c, *sp, generator=g))
    w = torch.randn(c, c, 3, 3, 3, generator=g) * (2.0 / (c * 27)) ** 0.5
    b = torch.randn(c, generator=g) * 0.05
    att = (torch.rand(1, *sp, generator=g) * 3).floor() + 1.0
    if perturb is not None:
        x = x * (1.0 + 1e-7 * torch.randn(x.shape, generator=torch.Generator().manual_seed(3000 + perturb)))
    y = F.conv3d(x, w, b, 1, 1)
    return x, w, b, y, att


def gen_real_width(R, out):
    res = {}
    for name, c, sp, n_iter, n_ens in REAL_WIDTH_CASES:
        x, w, b, y, att = real_width_inputs(name, c, sp)
        its = tuple(range(min(5, n_iter)))
        r = run_ref_layer(R, x, w, b, y, 1, 1, 16, 16, True, [att], state_iters=its, n_iter=n_iter)
        res[f"{name}_sums"] = np.array([x.double().sum().item(), w.double().sum().item(), y.double().sum().item(),
                                        att.double().sum().item()])
        res[f"{name}_hist"], res[f"{name}_final"] = r["hist"], r["final"]
        res[f"{name}_alpha_w"], res[f"{name}_alpha_act"] = r["alpha_w"], r["alpha_act"]
        res[f"{name}_n_iter"] = np.int64(n_iter)
        for i in its:
            res[f"{name}_st{i}_a_w"], res[f"{name}_st{i}_loss"] = r[f"st{i}_a_w"], r[f"st{i}_loss"]
        res[f"{name}_st0_wstar_sub"] = r["st0_wstar"].reshape(-1)[::101].copy()
        res[f"{name}_st0_wstar_absmax"] = np.float64(np.abs(r["st0_wstar"]).max())
        res[f"{name}_st0_bstar"] = r["st0_bstar"]
        res[f"{name}_st0_codes_next_sub"] = r["st0_codes_next"].reshape(-1)[::101].copy()
        print(name, "final", r["final"], "alpha_w", r["alpha_w"], "alpha_act", r["alpha_act"], "hist[:5]", r["hist"][:5])
        ens = [[float(r["final"]), float(r["hist"].min()), float(r["alpha_w"])]]
        for k in range(n_ens + (1 if n_ens else 0)):          # perturbed inputs, then the fp64-solve member
            xk, wk, bk, yk, attk = real_width_inputs(name, c, sp, perturb=k if k < n_ens else None)
            from contextlib import nullcontext
            with (Fp64Solve() if k == n_ens else nullcontext()):
                rk = run_ref_layer(R, xk, wk, bk, y, 1, 1, 16, 16, True, [attk], n_iter=n_iter)
            ens.append([float(rk["final"]), float(rk["hist"].min()), float(rk["alpha_w"])])
            print(name, "ensemble", k, ens[-1])
        res[f"{name}_ens"] = np.array(ens)
    for k, v in meta().items():
        res["meta_" + k] = np.array(str(v))
    np.savez_compressed(os.path.join(out, "real_width.npz"), **res)


TOY = dict(num_mod=4, num_classes=3, depth=[1, 1, 1], width=[8, 16, 8], dilation=[1, 1, 1],
           init_stride=(2, 2, 2), drop_rate=0.5, ds="simple", blk="mid", qlvl=16, qlvl_act=16,
           q_first=[256, -1], q_last=[256, -1], n=2, size=64, task="brats", seed=16)
# LiTS-config miniature (config/lits_ptq.yaml): 1 CT channel, init_stride 2,2,1, anisotropic volume,
# W2A2 (4/4 levels), softmax/argmax prediction, body mask all ones
TOY_LITS = dict(num_mod=1, num_classes=3, depth=[1, 1, 1], width=[8, 16, 8], dilation=[1, 1, 1],
                init_stride=(2, 2, 1), drop_rate=0.5, ds="simple", blk="mid", qlvl=4, qlvl_act=4,
                q_first=[256, -1], q_last=[256, -1], n=2, size=(96, 64, 32), task="lits", seed=18)


def build_toy(R, QConv, cfg=TOY):
    from models import factoryQ, factory_blk
    hetero = {"drop_cut_thres": 128, "ds_depth_limit": 3, "aniso_pool_depth": 9999,
              "aniso_pool_stride": (2, 2, 1)}
    return R["model_blk"].UResQ(
        QConv, cfg["num_mod"], cfg["num_classes"], depth_config=cfg["depth"], width_config=cfg["width"],
        dilation_config=cfg["dilation"], init_stride=cfg["init_stride"], stride=2,
        drop_rate=cfg["drop_rate"], nla=factoryQ.ReLU(True), bn=nn.BatchNorm3d, ds=cfg["ds"],
        blk_type=cfg["blk"], q_weight=True, qlvl=cfg["qlvl"], q_act=True, qlvl_act=cfg["qlvl_act"],
        q_first=cfg["q_first"], q_last=cfg["q_last"], hetero_param=hetero,
        rb=factory_blk.ResBlockWithType, fuse_bn=True, save_mem=True, init_kernel=3)


def seeded_state(model, seed):
    """kaiming init (reference utils/misc.py:85-102) + perturbed BN statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if k.endswith("alpha_act") or k.endswith("alpha_w"):
            continue
        if v.dim() == 5:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3] * v.shape[4]
            sd[k] = torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(v.shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            sd[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("num_batches_tracked"):
            sd[k] = v.clone()
        elif k.endswith(".weight"):      # BN gamma
            sd[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.05 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = v.clone()
    return sd


def ref_toy_run(R, cfg, sd, perturb_seed=None, trace=None):
    """The do_ptq core (reference src/ptqer.py:289-364) on a toy UResQ, CPU.  ``perturb_seed``: multiply the
    calibration volumes by (1 + 1e-7 * N(0,1)) -- the ensemble member of the reference's own sensitivity study."""
    from efficientq_b200 import synth
    ptqer = R["ptqer"]
    model = build_toy(R, R["effq"].EfficientQConv, cfg)
    model.load_state_dict(sd, strict=False)
    model.eval()
    R["fold_bn"].search_fold_and_remove_bn(model)
    size = cfg["size"] if isinstance(cfg["size"], tuple) else (cfg["size"],) * 3
    data = synth.batch(cfg["n"], 0, cfg["num_mod"], size, cfg["task"])
    if perturb_seed is not None:
        gp = torch.Generator().manual_seed(1000 + perturb_seed)
        data = data * (1.0 + 1e-7 * torch.randn(data.shape, generator=gp))
    ptqer.set_name(model)
    ptqer.set_fp(model)
    handles = []
    for m in model.modules():
        if isinstance(m, R["ptqconv"].PTQConv):
            handles.append(m.register_forward_hook(copying_hook))
    with torch.no_grad():
        output_fp = model(data).detach()
    task = cfg["task"]
    # ptqer.py:337-340: BraTS body = non-zero voxels of modality 0, LiTS body = everything
    body = (data[:, 0] != 0.0).bool() if task == "brats" else torch.ones_like(data[:, 0]).bool()
    wmap, nums = ptqer.get_att_weight_map(output_fp, torch.ones_like(data[:, 0]).bool(), "p:0.5", task=task)
    pyr = ptqer.get_mask_pyramid(output_fp, body, wmap, ",".join(str(v) for v in cfg["init_stride"]),
                                 num_lvls=5, task=task)
    ptqer.set_mask(model, pyr)
    for h in handles:
        h.remove()
    layer_loss = []
    ptqer.set_anything(model, "layer_loss", layer_loss)
    ptqer.set_quantizing(model)
    with torch.no_grad():
        if trace is not None:
            with trace:
                output_q = model(data)
        else:
            output_q = model(data)
    ptqer.set_quantized(model)
    names, losses = [], []
    for line in layer_loss:
        nm, val = line.rsplit(":", 1)
        names.append(nm.strip())
        losses.append(float(val))
    return dict(model=model, data=data, output_fp=output_fp, output_q=output_q, names=names, losses=losses,
                nums=nums, wmap=wmap, pyr=pyr)


ENSEMBLE = 8      # reference runs under a 1e-7 relative perturbation of the calibration volumes (+ one 1-thread run)


def ref_ensemble(R, cfg, sd, extra=None):
    """Per-layer losses of ENSEMBLE perturbed reference runs and of one single-threaded run of the unperturbed
    problem: how far the reference moves against ITSELF (the ADMM trajectory is chaotic in the last bits)."""
    rows = []
    for k in range(ENSEMBLE):
        r = ref_toy_run(R, cfg, sd, perturb_seed=k)
        rows.append(r["losses"] + ([extra(r)] if extra else []))
        print("ensemble", k, " ".join(f"{v:.4e}" for v in rows[-1]))
    nthr = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        r = ref_toy_run(R, cfg, sd)
        rows.append(r["losses"] + ([extra(r)] if extra else []))
    finally:
        torch.set_num_threads(nthr)
    print("1 thread  ", " ".join(f"{v:.4e}" for v in rows[-1]))
    with Fp64Solve():
        r = ref_toy_run(R, cfg, sd)
        rows.append(r["losses"] + ([extra(r)] if extra else []))
    print("fp64 solve", " ".join(f"{v:.4e}" for v in rows[-1]))
    return np.array(rows, dtype=np.float64)


def gen_toy_net(R, out, cfg=TOY, fname="toy_net.npz"):
    model0 = build_toy(R, R["effq"].EfficientQConv, cfg)
    sd = seeded_state(model0, cfg["seed"])
    trace = PtqTrace(R, state_iters=(), keep_inputs=True)
    run = ref_toy_run(R, cfg, sd, trace=trace)
    model, data, output_fp, output_q, pyr = run["model"], run["data"], run["output_fp"], run["output_q"], run["pyr"]
    names, losses, nums, wmap = run["names"], run["losses"], run["nums"], run["wmap"]
    res = {f"sd::{k}": v.numpy() for k, v in sd.items()}
    res["layer_names"] = np.array(names)
    res["layer_losses"] = np.array(losses, dtype=np.float64)
    res["class_nums"] = np.array(nums, dtype=np.int64)
    res["wmap"] = np.array([wmap[k] for k in sorted(wmap)], dtype=np.float64)
    res["pyr_means"] = np.array([p.mean().item() for p in pyr], dtype=np.float64)
    res["pyr0"] = pyr[0].numpy().astype(np.uint8)
    res["data_checksum"] = np.float64(data.double().sum().item())
    res["out_fp_mse_vs_q"] = np.float64(F.mse_loss(output_q, output_fp).item())
    res["out_fp_sum"] = np.float64(output_fp.double().sum().item())
    for name, m in model.named_modules():
        if isinstance(m, R["ptqconv"].PTQConv):
            res[f"q::{name}.alpha_w"] = np.float32(m.alpha_w.item())
            res[f"q::{name}.alpha_act"] = np.float32(m.alpha_act.item())
    # first conv's FP target: pre-activation (copying hook), so it must have negative entries
    res["conv0_target_min"] = np.float64(min(float(l["y"].min()) for l in trace.layers[:1]))
    R["ptqer"].store_int_weight(model)
    for name, m in model.named_modules():
        if isinstance(m, R["ptqconv"].PTQConv):
            res[f"q::{name}.wint"] = m.weight.data.numpy()
    for k, v in zip(names, losses):
        print(f"{k:45s} {v:.6e}")
    res["ensemble_losses"] = ref_ensemble(R, cfg, sd)
    np.savez_compressed(os.path.join(out, fname), **res)
    # teacher-forced fixture: what every layer of the REFERENCE run produced (calibrated fake-quant weights, bias,
    # scales, 200-iterate loss history, final loss).  A test rebuilds each layer's input by running the reference's
    # calibrated prefix (these weights) in deployment mode, so the 16 MB of per-layer inputs need not be stored; the
    # exact inputs of the small (<= 2^17 elements) layers ARE stored to check that replay.  ``tf_ens``: the reference
    # re-calibrating the SAME layer from its own input perturbed by 1e-7 (and single-threaded) -- its own per-layer
    # sensitivity, which bounds how tight a per-layer bar can be.
    tf = {"layer_names": np.array(names)}
    for lv, p in enumerate(pyr):
        assert float(p.max()) <= 65535 and bool((p == p.round()).all())
        tf[f"pyr{lv}"] = p.numpy().astype(np.uint16)
    gp = torch.Generator().manual_seed(4242)
    for rec in trace.layers:
        nm = rec["name"]
        mod = dict(model.named_modules())[nm]
        if rec["x"].numel() <= 2 ** 17:
            tf[f"x::{nm}"] = rec["x"].numpy()
        tf[f"x_sum::{nm}"] = np.float64(rec["x"].double().sum().item())
        tf[f"y_sum::{nm}"] = np.float64(rec["y"].double().sum().item())
        tf[f"weight::{nm}"], tf[f"bias::{nm}"] = rec["weight"].numpy(), rec["bias"].numpy()
        tf[f"att_level::{nm}"] = np.int64(rec["att_level"])
        tf[f"hist::{nm}"] = np.array(rec["hist"], dtype=np.float64)
        tf[f"final::{nm}"] = np.float64(rec["final"])
        tf[f"alpha_w::{nm}"], tf[f"alpha_act::{nm}"] = np.float32(rec["alpha_w"]), np.float32(rec["alpha_act"])
        ens = []
        for k in range(TF_ENSEMBLE + 2):             # perturbed inputs | one thread | fp64 linear solves
            xk = rec["x"] * (1.0 + 1e-7 * torch.randn(rec["x"].shape, generator=gp)) if k < TF_ENSEMBLE else rec["x"]
            nthr = torch.get_num_threads()
            if k == TF_ENSEMBLE:
                torch.set_num_threads(1)
            try:
                from contextlib import nullcontext
                with (Fp64Solve() if k == TF_ENSEMBLE + 1 else nullcontext()):
                    r = run_ref_layer(R, xk, rec["w0"], rec["b0"], rec["y"], mod.stride, mod.padding, mod.qlvl_w,
                                      mod.qlvl_act, mod.q_act, pyr)
            finally:
                torch.set_num_threads(nthr)
            ens.append([float(r["final"]), float(r["hist"].min()), float(r["alpha_w"])])
        tf[f"tf_ens::{nm}"] = np.array(ens, dtype=np.float64)
        e = tf[f"tf_ens::{nm}"]
        print(f"TF {nm:45s} ref final {rec['final']:.6e}  ensemble spread final {np.ptp(e[:, 0]) / rec['final']:.2e} "
              f"best {np.ptp(e[:, 1]) / min(rec['hist']):.2e} alpha_w {np.ptp(e[:, 2]) / rec['alpha_w']:.2e}")
    np.savez_compressed(os.path.join(out, fname.replace(".npz", "_tf.npz")), **tf)


TF_ENSEMBLE = 4


def gen_toy_tune(R, out, n_tune=3):
    """Reference tune_activation_range (src/ptqer.py:238-272) on the BraTS miniature, right after its
    do_ptq core: the calibrated state it starts from, d loss / d alpha_act of the first iteration (one
    manual forward/backward identical to the reference's loop body), then the reference function itself
    for n_tune Adam iterations -> losses and the refined alpha_act."""
    from efficientq_b200 import synth
    ptqer = R["ptqer"]
    cfg = TOY
    model = build_toy(R, R["effq"].EfficientQConv, cfg)
    model.load_state_dict(seeded_state(model, cfg["seed"]), strict=False)
    model.eval()
    R["fold_bn"].search_fold_and_remove_bn(model)
    data = synth.batch(cfg["n"], 0, cfg["num_mod"], (cfg["size"],) * 3, cfg["task"])
    ptqer.set_name(model)
    ptqer.set_fp(model)
    handles = [m.register_forward_hook(copying_hook) for m in model.modules()
               if isinstance(m, R["ptqconv"].PTQConv)]
    with torch.no_grad():
        output_fp = model(data).detach()
    body = (data[:, 0] != 0.0).bool()
    wmap, _ = ptqer.get_att_weight_map(output_fp, torch.ones_like(data[:, 0]).bool(), "p:0.5", task="brats")
    ptqer.set_mask(model, ptqer.get_mask_pyramid(output_fp, body, wmap, "2,2,2", num_lvls=5, task="brats"))
    for h in handles:
        h.remove()
    ptqer.set_anything(model, "layer_loss", [])
    ptqer.set_quantizing(model)
    with torch.no_grad():
        model(data)
    ptqer.set_quantized(model)
    res = {"out_fp_sum": np.float64(output_fp.double().sum().item())}
    mods = [(n, m) for n, m in model.named_modules() if isinstance(m, R["ptqconv"].PTQConv)]
    for name, m in mods:                                   # the calibrated state the tuning starts from
        res[f"pre::{name}.weight"] = m.weight.data.numpy().copy()
        res[f"pre::{name}.bias"] = m.bias.data.numpy().copy()
        res[f"pre::{name}.alpha_w"] = np.float32(m.alpha_w.item())
        res[f"pre::{name}.alpha_act"] = np.float32(m.alpha_act.item())
    out_q = model(data)                                    # loop body of ptqer.py:262-268, once, by hand
    loss = F.mse_loss(out_q, output_fp)
    model.zero_grad()
    loss.backward()
    res["loss0"] = np.float64(loss.item())
    for name, m in mods:
        if m.alpha_act.grad is not None:
            res[f"grad0::{name}"] = np.float64(m.alpha_act.grad.item())
    model.zero_grad(set_to_none=True)
    losses = ptqer.tune_activation_range(model, output_fp, data, max_iter=n_tune)
    res["tune_losses"] = np.array(losses, dtype=np.float64)
    for name, m in mods:
        res[f"post::{name}.alpha_act"] = np.float32(m.alpha_act.item())
    # kernel-level known answers: STE gradients of the reference's discretize (incl. values on the clamp
    # boundaries and rounding ties) and three torch.optim.Adam(lr=5e-4) steps on a small vector
    g = torch.Generator().manual_seed(21)
    for lv in (4, 16, 256):
        alpha = torch.tensor(1.37, requires_grad=True)
        x = torch.randn(4099, generator=g) * 1.2
        x[:8] = torch.tensor([0.0, 1.37, 1.37 * 0.5, -0.0, 2.0, -1.0, 1.37 / (lv - 1) * 0.5, 1.37 / (lv - 1) * 1.5])
        x.requires_grad_(True)
        go = torch.randn(4099, generator=g)
        q = R["lh"].discretize(x / alpha, lv, 0, 1) * alpha
        q.backward(go)
        res[f"ste{lv}::x"], res[f"ste{lv}::g"] = x.detach().numpy(), go.numpy()
        res[f"ste{lv}::grad_x"] = x.grad.numpy().copy()
        res[f"ste{lv}::grad_alpha"] = np.float64(alpha.grad.item())
    p = torch.nn.Parameter(torch.tensor([3.68, 6.06, 19.58, 0.5]))
    opt = torch.optim.Adam([p], lr=5e-4)
    gs = torch.tensor([[0.19, -0.028, -1.2e-3, 3.0e-9], [0.17, 0.031, -1.0e-3, -2.0e-9], [-0.05, 0.030, 4.0e-4, 1.0e-9]])
    traj = []
    for k in range(3):
        p.grad = gs[k].clone()
        opt.step()
        traj.append(p.data.clone().numpy())
    res["adam::grads"], res["adam::p0"] = gs.numpy(), np.array([3.68, 6.06, 19.58, 0.5], dtype=np.float32)
    res["adam::traj"] = np.stack(traj)
    print("loss0", res["loss0"], "tune losses", losses)
    for name, m in mods:
        print(f"{name:45s} grad0 {res.get(f'grad0::{name}', float('nan')):+.6e} alpha {res[f'pre::{name}.alpha_act']:.6f} -> "
              f"{res[f'post::{name}.alpha_act']:.6f}")
    np.savez_compressed(os.path.join(out, "toy_tune.npz"), **res)


def dice_table(pred, label, n_cls=3):
    """Per-class Dice with the reference's formula (src/utils/metrics.py:21-25)."""
    out = []
    for k in range(1, n_cls + 1):
        pb, tb = (pred == k), (label == k)
        out.append(float((2 * (pb & tb).sum().float() + 1e-6) / (pb.sum().float() + tb.sum().float() + 1e-6)))
    return out


def gen_toy_dice(R, out, steps=160):
    """Final-Dice parity case (BASELINE north star: "final Dice within 0.1 points"): a BraTS miniature TRAINED
    with stock PyTorch on the synthetic nested-sphere volumes (so that Dice is non-degenerate), then the
    reference's do_ptq core at W4A4 on 2 calibration volumes, and Dice of the FP and of the quantised model
    on 4 held-out volumes (full-volume forward, nested-sigmoid prediction of metrics.py:182-192)."""
    from efficientq_b200 import synth
    ptqer = R["ptqer"]
    cfg = TOY
    torch.manual_seed(cfg["seed"] + 100)
    model = build_toy(R, R["effq"].EfficientQConv, cfg)
    model.load_state_dict(seeded_state(model, cfg["seed"]), strict=False)
    ptqer.set_fp(model)
    vols = [synth.volume(i, cfg["num_mod"], (cfg["size"],) * 3, "brats") for i in range(8)]
    opt = torch.optim.Adam([p for n, p in model.named_parameters() if "alpha" not in n], lr=2e-3)
    model.train()
    for step in range(steps):
        idx = [(2 * step) % 8, (2 * step + 1) % 8]
        x = torch.stack([vols[i][0] for i in idx])
        lab = torch.stack([vols[i][1] for i in idx]).long()
        tgt = torch.stack([(lab >= k).float() for k in (1, 2, 3)], 1)          # nested regions <-> get_pred_brats
        o = model(x)                                                            # (M, N, 3, D, H, W)
        loss = sum(F.binary_cross_entropy_with_logits(o[m], tgt) for m in range(o.shape[0]))
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 20 == 0:
            print("train step", step, float(loss))
    model.eval()
    trained = {k: v.detach().clone() for k, v in model.state_dict().items() if not k.endswith(("alpha_act", "alpha_w"))}
    ev = [synth.volume(100 + i, cfg["num_mod"], (cfg["size"],) * 3, "brats") for i in range(4)]
    ev_x = torch.stack([v[0] for v in ev])
    ev_l = torch.stack([v[1] for v in ev]).long()
    from utils import metrics as M_                                            # reference's get_pred_brats

    def dice_of(m):
        with torch.no_grad():
            return dice_table(M_.get_pred_brats(m(ev_x)[-1]), ev_l)
    fp_model = build_toy(R, R["effq"].EfficientQConv, cfg)
    fp_model.load_state_dict(trained, strict=False)
    fp_model.eval()
    R["fold_bn"].search_fold_and_remove_bn(fp_model)
    ptqer.set_fp(fp_model)
    dice_fp = dice_of(fp_model)
    run = ref_toy_run(R, cfg, trained)
    dice_q = dice_of(run["model"])
    res = {f"sd::{k}": v.numpy() for k, v in trained.items()}
    res["dice_fp"], res["dice_q"] = np.array(dice_fp), np.array(dice_q)
    res["layer_losses"] = np.array(run["losses"])
    # the reference's calibrated state: replayed through the GPU deployment forward it must give dice_q itself
    for name, m in run["model"].named_modules():
        if isinstance(m, R["ptqconv"].PTQConv):
            res[f"cal::{name}.weight"], res[f"cal::{name}.bias"] = m.weight.data.numpy().copy(), m.bias.data.numpy().copy()
            res[f"cal::{name}.alpha_w"] = np.float32(m.alpha_w.item())
            res[f"cal::{name}.alpha_act"] = np.float32(m.alpha_act.item())
    print("Dice FP", dice_fp, "mean", np.mean(dice_fp), "| Dice W4A4 (reference)", dice_q, "mean", np.mean(dice_q))
    # the reference against itself: ENSEMBLE calibrations from volumes perturbed by 1e-7 + one single-threaded run
    rows = []
    for k in range(ENSEMBLE + 2):                    # perturbed volumes | one thread | fp64 linear solves
        nthr = torch.get_num_threads()
        if k == ENSEMBLE:
            torch.set_num_threads(1)
        try:
            from contextlib import nullcontext
            with (Fp64Solve() if k == ENSEMBLE + 1 else nullcontext()):
                r = ref_toy_run(R, cfg, trained, perturb_seed=k if k < ENSEMBLE else None)
            rows.append(dice_of(r["model"]))
        finally:
            torch.set_num_threads(nthr)
        print("dice ensemble", k, rows[-1], "mean", np.mean(rows[-1]))
    res["dice_q_ensemble"] = np.array(rows, dtype=np.float64)
    np.savez_compressed(os.path.join(out, "toy_dice.npz"), **res)


def gen_toy_net_lits(R, out):
    gen_toy_net(R, out, TOY_LITS, "toy_net_lits.npz")


EVAL_TILINGS = [((20, 24, 12), (8, 16, 12), (4, 8, 8)), ((18, 18, 18), (18, 18, 18), (8, 8, 8)),
                ((25, 21, 17), (12, 12, 12), (8, 4, 0)), ((32, 16, 16), (16, 16, 16), (8, 8, 8))]


def gen_eval(R, out):
    """Sliding-window tiling / stitching (src/utils/transforms.py:784-851), label split / merge
    (src/utils/misc.py:221-285) and SegMetricMC (src/utils/validate.py:19-205) of the reference's evaluation."""
    import io
    import utils.transforms as tfm
    import utils.misc as misc
    from utils.validate import SegMetricMC
    res = {}
    g = torch.Generator().manual_seed(41)
    for ci, (dhw, patch, ov) in enumerate(EVAL_TILINGS):
        n_vox = dhw[0] * dhw[1] * dhw[2]
        index_img = torch.arange(n_vox, dtype=torch.float32).reshape(1, 1, *dhw)          # value = linear voxel index
        starts = [int(p[0, 0, 0, 0, 0]) for p in tfm.image_to_patch3d(index_img, patch, ov)]
        img = torch.randn(2, 2, *dhw, generator=g)
        patches = tfm.image_to_patch3d(img, patch, ov)
        preds = [torch.stack([p * (1.0 + 0.125 * k) + 0.5 * k, p.flip(1) - 0.25 * k]) for k, p in enumerate(patches)]
        stitched = tfm.patch_to_image3d(img, preds, patch, ov)
        res[f"tile{ci}_starts"] = np.array(starts, dtype=np.int64)
        res[f"tile{ci}_img"] = img.numpy()
        res[f"tile{ci}_stitched"] = stitched.numpy()
    # labels
    lab = torch.randint(0, 4, (10, 12, 14), generator=g)
    res["label_brats"] = lab.numpy().astype(np.uint8)
    res["split_brats"] = misc.split_label_brats(lab).numpy()
    lab2 = torch.randint(0, 3, (10, 12, 14), generator=g)
    res["label_lits"] = lab2.numpy().astype(np.uint8)
    res["split_lits"] = misc.split_label_lits(lab2).numpy()
    bits = (torch.rand(3, 10, 12, 14, generator=g) > 0.5).int()
    res["merge_in"] = bits.numpy()
    res["merge_con"] = misc.merge_label_basic(bits.clone(), "con").numpy()
    res["merge_agg"] = misc.merge_label_basic(bits.clone(), "agg").numpy()
    res["merge_brats_con"] = misc.merge_label_brats(bits.clone(), "con").numpy()
    # metrics: two subjects each; multi-class (argmax over 3 logits, integer label) and multi-label
    # (sigmoid >= 0.5 per channel, optional fusion)
    for mode in ("mc", "ml_con", "ml_none"):
        sm = SegMetricMC(3, ["case_a", "case_b"])
        for j in range(2):
            logits = torch.randn(3, 12, 10, 8, generator=g)
            if mode == "mc":
                label = torch.randint(0, 3, (12, 10, 8), generator=g)
                pred = sm.evaluate_append(logits, label)
            else:
                label = (torch.rand(3, 12, 10, 8, generator=g) > 0.6).float()
                pred = sm.evaluate_append(logits, label, multilabel_fusetype="con" if mode == "ml_con" else None)
            res[f"{mode}_logits{j}"] = logits.numpy()
            res[f"{mode}_label{j}"] = label.numpy()
            res[f"{mode}_pred{j}"] = pred.numpy()
        sm.get_metric()
        keys = list(sm.metric.keys())
        res[f"{mode}_keys"] = np.array(keys)
        res[f"{mode}_metric"] = np.array([sm.metric[k] for k in keys], dtype=np.float64)
        res[f"{mode}_buffer"] = np.array([[float(v) for v in sm.buffer[k]] for k in keys], dtype=np.float32)
        buf = io.StringIO()
        sm.write_metric(buf, "Output -1:", True)
        res[f"{mode}_text"] = np.array(buf.getvalue())
    for k, v in meta().items():
        res["meta_" + k] = np.array(str(v))
    np.savez_compressed(os.path.join(out, "eval.npz"), **res)


TINY_SNS = ["s03", "s01", "s10", "s02", "s07"]                 # deliberately unsorted: Dataset_SEG sorts, the on-disk variant does not
TINY_SHAPES = [(20, 18, 26), (18, 20, 24), (22, 16, 25), (17, 19, 28), (16, 16, 24)]


def write_tiny_dataset(root, seed=77):
    """A five-subject BraTS-layout dataset (src/definer.py:42, src/dataloader/datasets.py:24-29) written from a
    platform-stable numpy stream; used by gen_calib_data here and, with the same seed, by tests/test_data.py."""
    import pickle
    rng = np.random.Generator(np.random.PCG64(seed))
    for mod in ("seg", "flair", "t1", "t1ce", "t2"):
        os.makedirs(os.path.join(root, mod), exist_ok=True)
    for sn, shp in zip(TINY_SNS, TINY_SHAPES):
        for mod in ("flair", "t1", "t1ce", "t2"):
            np.savez(os.path.join(root, mod, sn + ".npz"), rng.standard_normal(shp).astype(np.float32))
        np.savez(os.path.join(root, "seg", sn + ".npz"), rng.integers(0, 4, shp).astype(np.uint8))
    os.makedirs(os.path.join(root, "split", "round1"), exist_ok=True)
    with open(os.path.join(root, "split", "round1", "train.txt"), "w") as fid:
        fid.write("\n".join(TINY_SNS) + "\n")
    with open(os.path.join(root, "split", "round1", "val.txt"), "w") as fid:
        fid.write("\n".join(TINY_SNS[:2]) + "\n")
    with open(os.path.join(root, "sn_fn.txt"), "w") as fid:
        fid.write("\n".join(f"{s},{s}.nii.gz" for s in TINY_SNS) + "\n")
    with open(os.path.join(root, "restore_shape_infokw.pickle"), "wb") as fid:
        pickle.dump({s: {} for s in TINY_SNS}, fid)


def tiny_dataset_args(root, data_on_disk):
    from efficientq_b200 import entrance
    a = entrance.build_parser().parse_args(["ptq", "--config", os.path.join(ROOT, "config", "brats_ptq.yaml")])
    a = entrance.merge_config(a.config, a)
    a.data_dir, a.split_dir, a.num_workers = root, os.path.join(root, "split"), 0
    a.lwq_patchsz, a.lwq_batchsz, a.lwq_dataid, a.data_on_disk = "16,16,24", 3, 1, data_on_disk
    return a


def gen_calib_data(R, out):
    """Calibration-batch assembly (src/ptqer.py:83-111 over src/definer.py:14-127 / src/dataloader) and the val
    loader's volumes, for the in-memory (sorted) and the on-disk (file order) dataset classes."""
    import tempfile
    import definer as rdef
    root = tempfile.mkdtemp()
    write_tiny_dataset(root)
    res = {}
    for disk in (True, False):
        a = tiny_dataset_args(root, disk)
        cube = rdef.get_data_cube(a)[0]
        x, y = R["ptqer"].get_calibration_data(a, cube)
        tag = "disk" if disk else "mem"
        res[f"{tag}_batch"] = x.numpy()
        res[f"{tag}_label_batch"] = y.numpy().astype(np.uint8)
        for i, (im, lb) in enumerate(cube.valloader):
            res[f"{tag}_val{i}_img"] = im[0].numpy()
            res[f"{tag}_val{i}_label"] = lb[0].numpy().astype(np.uint8)
        a.lwq_batchsz, a.lwq_dataid = 1, 0                       # the single-volume branch (:93-100)
        cube = rdef.get_data_cube(a)[0]
        res[f"{tag}_single"] = R["ptqer"].get_calibration_data(a, cube)[0].numpy()
    for k, v in meta().items():
        res["meta_" + k] = np.array(str(v))
    np.savez_compressed(os.path.join(out, "calib_data.npz"), **res)


def gen_interface(R, out):
    """The reference's command line (src/entrance.py:31-113: every flag with its type, default and action) and the
    state-dict layout (keys + shapes, module order of the quantizer layers) of its BraTS- and LiTS-config nets
    (src/definer.py:130-248, :286-329 over config/*_ptq.yaml) -> interface.json."""
    import argparse as ap
    import json
    import definer as rdef
    src = open(os.path.join(sys.path[0] if os.path.exists(os.path.join(sys.path[0], "entrance.py"))
                            else os.path.join("/root/reference", "src"), "entrance.py")).read()
    part = src[src.index("parser = argparse.ArgumentParser"):src.index("args = parser.parse_args()")]
    ns = {"argparse": ap}
    exec(part, ns)                                              # builds the reference's parser, runs nothing else
    flags = {}
    for act in ns["parser"]._actions:
        if not act.option_strings:
            flags[act.dest] = {"positional": True, "choices": list(act.choices) if act.choices else None}
            continue
        if act.dest == "help":
            continue
        flags[act.dest] = {"option": act.option_strings[0], "type": act.type.__name__ if act.type else None,
                           "default": act.default, "flag": isinstance(act, ap._StoreTrueAction),
                           "choices": list(act.choices) if act.choices else None}
    res = {"flags": flags, "nets": {}}
    from efficientq_b200 import entrance
    for task, lw, la in (("brats", 16, 16), ("lits", 4, 4)):
        a = entrance.build_parser().parse_args(["ptq", "--qlvl_w", str(lw), "--qlvl_a", str(la), "--config",
                                                os.path.join(ROOT, "config", f"{task}_ptq.yaml")])
        a = entrance.merge_config(a.config, a)
        QConv, qinfo, kwQ = rdef.get_conv_class(a)
        cube, info = rdef.get_model_cube(a, QConv, kwQ)
        model = cube["model"]
        qmods = [(n, m.in_channels, m.out_channels, list(m.kernel_size), list(m.stride), list(m.padding),
                  int(m.qlvl_w), int(m.qlvl_act), bool(m.q_act))
                 for n, m in model.named_modules() if isinstance(m, R["ptqconv"].PTQConv)]
        res["nets"][task] = {"qinfo": qinfo, "model_info": info, "num_mo": cube["num_mo"], "nClass": cube["nClass"],
                             "nMod": cube["nMod"], "kwQ": sorted(kwQ),
                             "state": [[k, list(v.shape)] for k, v in model.state_dict().items()],
                             "quantizers": qmods}
    res["meta"] = {k: str(v) for k, v in meta().items()}
    with open(os.path.join(out, "interface.json"), "w") as fid:
        json.dump(res, fid, indent=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.manual_seed(0)
    R = import_reference(args.ref)
    gens = dict(discretize=gen_discretize, fakequant_module=gen_fakequant_module, project=gen_project,
                solver=gen_solver, layers=gen_layers, layers_wide=gen_layers_wide, real_width=gen_real_width, toy_net=gen_toy_net,
                toy_net_lits=gen_toy_net_lits, toy_tune=gen_toy_tune, toy_dice=gen_toy_dice, eval=gen_eval, calib_data=gen_calib_data, interface=gen_interface)
    for name, fn in gens.items():
        if args.only and name not in args.only.split(","):
            continue
        print("==", name)
        fn(R, HERE)
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as fid:
        for k, v in meta().items():
            fid.write(f"{k}: {v}\n")


if __name__ == "__main__":
    main()
