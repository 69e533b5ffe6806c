"""CPU oracle for the EfficientQ PTQ calibration hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product path (``efficientq_b200``)
never imports this module and has no CPU fallback.

It is a from-scratch restatement, in torch-CPU / numpy arithmetic, of the
reference algorithm (rongzhao-zhang/EfficientQ).  Every function cites the
reference file:line it follows.  The arithmetic library is the reference's own
(PyTorch CPU ops: true IEEE division, round-half-even ``torch.round``,
``F.conv3d``, ``torch.linalg.solve``), so the oracle reproduces the reference's
*CPU* semantics: the CUDA build of PyTorch turns
``tensor / python_scalar`` into a multiply by the reciprocal, the CPU build does
not, and the CPU build is the parity anchor.

Pinning: the reference ships no tests and no golden vectors (SURVEY.md section
4), so the oracle is pinned against outputs of the *reference itself*, generated
in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/src``) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "discretize", "discretize_codes", "quantize_act", "quantize_w",
    "project_by_iter", "im2col", "gram_stats", "NormalEquations",
    "admm_layer", "LayerResult", "weight_to_int", "int_to_weight",
    "fold_bn_pair", "class_counts_brats", "class_counts_lits",
    "att_weight_map", "mask_pyramid", "pred_brats", "pred_lits",
    "select_att", "rho_scale_of",
    "glue_relu", "glue_add", "glue_maxpool3d", "glue_upsample_trilinear",
]


def _triple(v):
    return (v, v, v) if isinstance(v, int) else tuple(int(t) for t in v)


# ----------------------------------------------------------------------------
# a1: fake-quant (reference: src/models/layer_helper.py:25-37, :13-22)
# ----------------------------------------------------------------------------
def discretize(v: torch.Tensor, num_lvl: int, lo: float, hi: float) -> torch.Tensor:
    """Clamp to [lo, hi], snap to ``num_lvl`` uniformly spaced levels.

    Same op order and dtypes as layer_helper.py:30-36: ``delta`` is a Python
    double that PyTorch casts to the tensor dtype, the division is a true
    division, ``torch.round`` is round-half-to-even.
    """
    steps = num_lvl - 1
    clipped = torch.clamp(v, lo, hi)
    delta = (hi - lo) / steps
    idx = torch.round((clipped - lo) / delta)
    return idx * delta + lo


def discretize_codes(v: torch.Tensor, num_lvl: int, lo: float, hi: float) -> torch.Tensor:
    """The integer level index in [0, num_lvl-1] that ``discretize`` rounds to
    (the argument of ``round`` at layer_helper.py:34)."""
    steps = num_lvl - 1
    clipped = torch.clamp(v, lo, hi)
    delta = (hi - lo) / steps
    return torch.round((clipped - lo) / delta).to(torch.int32)


def quantize_act(x: torch.Tensor, alpha_act: torch.Tensor, num_lvl: int) -> torch.Tensor:
    """PTQConv._quantize_act (src/models/PTQConv.py:114-116): per-tensor,
    unsigned range [0, alpha]."""
    return discretize(x / alpha_act, num_lvl, 0, 1) * alpha_act


def quantize_w(w: torch.Tensor, alpha_w: torch.Tensor, num_lvl: int) -> torch.Tensor:
    """PTQConv._quantize_w (src/models/PTQConv.py:110-112): per-tensor,
    symmetric range [-alpha, alpha]."""
    return discretize(w / alpha_w, num_lvl, -1, 1) * alpha_w


# ----------------------------------------------------------------------------
# a3: scale search (reference: src/models/layer_helper.py:40-70)
# ----------------------------------------------------------------------------
def project_by_iter(var: torch.Tensor, num_lvl: int, lo: float = -1.0, hi: float = 1.0,
                    return_iters: bool = False):
    """Alternating fixed point a <- <b,v>/<b,b>, b <- Q(v/a) in fp64.

    Start a0 = mean|v|; stop when |a - a_prev| <= 1e-5 or after num_lvl*100
    passes (then the reference raises RuntimeWarning, layer_helper.py:62-64).
    Returns (a: float, b: fp32 tensor of level values[, passes]).
    """
    v64 = var.detach().double()
    limit = num_lvl * 100
    a = v64.abs().mean().item()
    a_prev = -999.0
    passes = 0
    while abs(a - a_prev) > 1e-5 and passes < limit:
        b = discretize(v64 / a, num_lvl, lo, hi)
        a_prev = a
        a = ((b * v64).sum() / (b * b).sum()).item()
        passes += 1
    if passes == limit:
        raise RuntimeWarning(
            f"Exceed maximum iteration ({limit}) for alpha optimization in var_init_iter")
    b = discretize(v64 / a, num_lvl, lo, hi).float()
    if return_iters:
        return a, b, passes
    return a, b


# ----------------------------------------------------------------------------
# a7: im2col (reference: src/models/solver.py:86-111)
# ----------------------------------------------------------------------------
def im2col(x: torch.Tensor, kd: int, kh: int, kw: int, stride, pad) -> torch.Tensor:
    """Patch matrix (C*kd*kh*kw, N*Do*Ho*Wo), fp32.

    Row order (c, kd, kh, kw) row-major -- matches ``weight.reshape(C2, -1)`` --
    and sample-major columns, exactly as solver.py:101-111, but gathered with
    strided views instead of a Python loop over every output voxel.
    """
    s = _triple(stride)
    p = _triple(pad)
    n, c, d, h, w = x.shape
    do = (d + 2 * p[0] - kd) // s[0] + 1
    ho = (h + 2 * p[1] - kh) // s[1] + 1
    wo = (w + 2 * p[2] - kw) // s[2] + 1
    # NB solver.py:94 calls F.pad/np.pad with (p0,p0,p1,p1,p2,p2); for numpy
    # (the live path, solver.py:97) that pads D by p0, H by p1, W by p2.
    xp = F.pad(x.float(), (p[2], p[2], p[1], p[1], p[0], p[0]))
    win = xp.unfold(2, kd, s[0]).unfold(3, kh, s[1]).unfold(4, kw, s[2])
    # win: (n, c, do, ho, wo, kd, kh, kw) -> rows (c,kd,kh,kw), cols (n,do,ho,wo)
    cols = win.permute(1, 5, 6, 7, 0, 2, 3, 4).reshape(c * kd * kh * kw, n * do * ho * wo)
    return cols.contiguous()


# ----------------------------------------------------------------------------
# a8: normal-equation statistics (reference: src/models/solver.py:282-314)
# ----------------------------------------------------------------------------
def gram_stats(x_col: torch.Tensor, y: torch.Tensor, att: Optional[torch.Tensor], n: int):
    """A0 = 2 sum_n X_n (att_n.X_n)^T, B0 = 2 sum_n Y_n (att_n.X_n)^T.

    ``x_col`` (K', V) already carries the ones row for the bias
    (solver.py:255-256); ``y`` is (C2, V); ``att`` is (N, Do, Ho, Wo) or None.
    Per-sample accumulation order as solver.py:305-310.
    """
    xh = x_col * att.reshape(1, -1) if att is not None else x_col
    per = x_col.shape[1] // n
    a_acc = 0
    b_acc = 0
    for i in range(n):
        sl = slice(i * per, (i + 1) * per)
        a_acc = a_acc + x_col[:, sl] @ xh[:, sl].T
        b_acc = b_acc + y[:, sl] @ xh[:, sl].T
    return 2 * a_acc, 2 * b_acc


class NormalEquations:
    """Restatement of ``QuadraSolver`` (src/models/solver.py:201-345).

    Minimises  ||att^(1/2) (W^ X^ - Y)||^2 + rho/2 ||W - Gd||^2 + eta/2 ||W^ - W0^||^2
    where ^ means "with the bias column / ones row appended".
    """

    def __init__(self, x_q: torch.Tensor, y: torch.Tensor, ksize, stride, padding,
                 w0: torch.Tensor, b0: Optional[torch.Tensor], att: Optional[torch.Tensor],
                 mu: float = 0.0):
        kd, kh, kw = ksize
        self.c2 = y.shape[1]
        self.c1 = x_q.shape[1]
        self.ksize = (kd, kh, kw)
        self.k = self.c1 * kd * kh * kw
        self.has_bias = b0 is not None
        self.mu = mu
        w0m = w0.reshape(self.c2, -1).float()
        x_col = im2col(x_q, kd, kh, kw, stride, padding)
        if self.has_bias:
            w0m = torch.cat([w0m, b0.reshape(-1, 1).float()], dim=1)   # solver.py:247
            x_col = torch.cat([x_col, torch.ones(1, x_col.shape[1])], dim=0)  # solver.py:255-256
        self.w0 = w0m
        self.kp = x_col.shape[0]
        self.eye = torch.eye(self.kp)
        self.quasi_eye = torch.eye(self.kp)
        if self.has_bias:
            self.quasi_eye[-1, -1] = 0.0                                # solver.py:249-250
        # solver.py:266-267: (N,C2,D,H,W) -> (C2, N*D*H*W), sample-major columns
        ymat = y.float().permute(1, 0, 2, 3, 4).reshape(self.c2, -1)
        self.a0, self.b0 = gram_stats(x_col, ymat, att, x_q.shape[0])

    def assemble(self, rho: float, eta: float, g: torch.Tensor):
        """solver.py:316-325."""
        if self.has_bias:
            a = self.a0 + (rho + self.mu) * self.quasi_eye + eta * self.eye
            b = self.b0 + eta * self.w0
            b[:, : self.kp - 1] += rho * g.reshape(self.c2, -1)
        else:
            a = self.a0 + (rho + self.mu + eta) * self.eye
            b = self.b0 + rho * g.reshape(self.c2, -1) + eta * self.w0
        return a, b

    def solve(self, rho: float, eta: float, g: torch.Tensor):
        """solver.py:327-345: w* = solve(A, B^T)^T, bias split off the last column."""
        a, b = self.assemble(rho, eta, g)
        sol = torch.linalg.solve(a, b.T).T
        kd, kh, kw = self.ksize
        if self.has_bias:
            return sol[:, :-1].reshape(self.c2, self.c1, kd, kh, kw), sol[:, -1]
        return sol.reshape(self.c2, self.c1, kd, kh, kw), None


# ----------------------------------------------------------------------------
# a6: the per-layer ADMM driver (reference: src/models/EfficientQConv.py:33-166)
# ----------------------------------------------------------------------------
@dataclass
class LayerResult:
    weight: torch.Tensor
    bias: Optional[torch.Tensor]
    alpha_w: float
    alpha_act: Optional[float]
    loss_history: List[float] = field(default_factory=list)
    final_loss: float = 0.0
    best_iter: int = 0
    rho_scale: float = 1.0
    qact: Optional[torch.Tensor] = None


def select_att(mask_pyramid: Optional[Sequence[torch.Tensor]], out_spatial) -> Optional[torch.Tensor]:
    """EfficientQConv.py:53-59: first pyramid level whose (D,H,W) equals the output's."""
    if not mask_pyramid:
        return None
    for m in mask_pyramid:
        if tuple(m.shape[1:]) == tuple(out_spatial):
            return m
    return None


def rho_scale_of(out_fp: torch.Tensor, weight: torch.Tensor, att: Optional[torch.Tensor]) -> float:
    """EfficientQConv.py:44-49,60-61 (torch.std is the unbiased estimator)."""
    rs = max(out_fp.numel() * out_fp.std().item() / (weight.numel() * weight.std().item()), 1.0)
    if att is not None:
        rs *= att.mean().item()
    return rs


def admm_layer(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor],
               out_fp: torch.Tensor, stride, padding, qlvl_w: int, qlvl_act: int,
               q_act: bool = True, mask_pyramid=None, n_iter: int = 200,
               rho0: float = 10.0, rho_max: float = 1000.0, eta0: float = 1.0,
               rho_period: int = 50, keep_qact: bool = False, timers: Optional[dict] = None,
               channel_wise: bool = False) -> LayerResult:
    """One layer of EfficientQ calibration, EfficientQConv.py:33-166.

    Quirks reproduced on purpose: best iterate picked by strict ``<`` on the
    unweighted fp32 MSE with iterate 0 always seeding (:139-142); ``alpha_w``
    comes from the LAST iterate while the weights come from the best one
    (:155-158); rho doubles after iterations 0, 50, 100, ... with the dual
    halved (:129-137); the logged loss is attention-weighted (:161-166).

    ``channel_wise`` (extension, not in the reference's live path): the projection runs the reference's
    project_by_iter on every output channel's row of w* + dual separately; alpha_w is then a [C2] tensor.
    """
    import time as _time

    def _tick(key, t_start):
        if timers is not None:
            timers[key] = timers.get(key, 0.0) + (_time.perf_counter() - t_start)

    stride = _triple(stride)
    padding = _triple(padding)
    g = weight.detach().clone().float()
    dual = torch.zeros_like(g)
    out_fp = out_fp.detach().float()
    x = x.detach().float()
    att = select_att(mask_pyramid, out_fp.shape[2:])
    rs = rho_scale_of(out_fp, g, att)

    alpha_act = None
    t_ = _time.perf_counter()
    if q_act:
        a_act, b_act = project_by_iter(x, qlvl_act, 0, 1)
        alpha_act = a_act
        q_x = a_act * b_act
    else:
        q_x = x
    _tick("act_search", t_)
    rho = rho0 * rs
    rho_m = rho_max * rs
    eta = eta0 * rs

    w0 = g.clone()
    b0 = bias.detach().clone().float() if bias is not None else None
    t_ = _time.perf_counter()
    ne = NormalEquations(q_x, out_fp, tuple(weight.shape[2:]), stride, padding, w0, b0, att)
    _tick("im2col_gram", t_)

    b_star = b0
    best = (None, None, 1e10, 0)
    a_w = None
    hist: List[float] = []
    for it in range(n_iter):
        t_ = _time.perf_counter()
        w_star, b_star_new = ne.solve(rho, eta, g - dual)
        _tick("solve", t_)
        if bias is not None:
            b_star = b_star_new
        t_ = _time.perf_counter()
        if channel_wise:
            v = w_star + dual
            rows = [project_by_iter(v[r], qlvl_w, -1, 1) for r in range(v.shape[0])]
            a_w = torch.tensor([a for a, _ in rows], dtype=torch.float64)
            g = torch.stack([a * b for a, b in rows])
        else:
            a_w, b_w = project_by_iter(w_star + dual, qlvl_w, -1, 1)
            g = a_w * b_w
        dual = w_star - g + dual
        _tick("w_project", t_)
        t_ = _time.perf_counter()
        out_q = F.conv3d(q_x, g.float(), b_star, stride, padding)
        loss = F.mse_loss(out_q, out_fp).item()
        _tick("conv_mse", t_)
        hist.append(loss)
        if it % rho_period == 0:
            if rho * 2 <= rho_m:
                rho *= 2
                dual = dual / 2
            else:
                dual = dual / (rho_m / rho)
                rho = rho_m
        if it == 0 or loss < best[2]:
            best = (g, b_star if bias is not None else None, loss, it)

    g_best, b_best, _, best_it = best
    out_q = F.conv3d(q_x, g_best, b_best, stride, padding)
    final = F.mse_loss(out_q, out_fp).item()
    if att is not None:
        final = (att.unsqueeze(1) * (out_q - out_fp) ** 2).mean().item()
    return LayerResult(weight=g_best, bias=b_best, alpha_w=a_w, alpha_act=alpha_act,
                       loss_history=hist, final_loss=final, best_iter=best_it, rho_scale=rs,
                       qact=q_x if keep_qact else None)


# ----------------------------------------------------------------------------
# a13: integer export (reference: src/models/PTQConv.py:125-152)
# ----------------------------------------------------------------------------
def weight_to_int(q: torch.Tensor, alpha_w: torch.Tensor, num_lvl: int) -> torch.Tensor:
    """PTQConv.store_int_weight: round((q/alpha + 1)/delta) as uint8 (int32 if L>256)."""
    delta = 2 / (num_lvl - 1)
    codes = torch.round((q / alpha_w + 1) / delta)
    return codes.to(torch.uint8 if num_lvl <= 256 else torch.int32)


def int_to_weight(codes: torch.Tensor, alpha_w: torch.Tensor, num_lvl: int) -> torch.Tensor:
    """PTQConv.restore_fp_weight: alpha * (codes*delta - 1)."""
    delta = 2 / (num_lvl - 1)
    return alpha_w * (codes.float() * delta - 1)


# ----------------------------------------------------------------------------
# a14: BN folding (reference: src/models/fold_bn.py:14-34)
# ----------------------------------------------------------------------------
def fold_bn_pair(w: torch.Tensor, b: Optional[torch.Tensor], gamma, beta, mean, var, eps: float):
    """W <- W*gamma/sqrt(var+eps); b <- beta - gamma*mean/sqrt(var+eps) (+ scaled old bias)."""
    std = torch.sqrt(var + eps)
    w_f = w * (gamma / std).view(-1, 1, 1, 1, 1)
    shift = beta - gamma * mean / std
    b_f = gamma * b / std + shift if b is not None else shift
    return w_f, b_f


# ----------------------------------------------------------------------------
# a12: attention-mask pyramid (reference: src/ptqer.py:141-235, utils/metrics.py:172-192)
# ----------------------------------------------------------------------------
def pred_lits(out: torch.Tensor) -> torch.Tensor:
    """metrics.py:172-179: argmax over channels."""
    return torch.max(out, 1)[1]


def pred_brats(out: torch.Tensor) -> torch.Tensor:
    """metrics.py:182-192: nested sigmoid labels, later channels overwrite."""
    hard = torch.sigmoid(out) >= 0.5
    pred = torch.zeros_like(hard[:, 0]).int()
    for i in range(hard.shape[1]):
        pred[hard[:, i]] = i + 1
    return pred


def class_counts_lits(pred: torch.Tensor, body: torch.Tensor) -> List[int]:
    """ptqer.py:172-178."""
    return [int(((pred == k) & body).sum().item()) for k in range(3)]


def class_counts_brats(pred: torch.Tensor, body: torch.Tensor) -> List[int]:
    """ptqer.py:181-188 (pred is N x 3 x D x H x W of {0,1})."""
    nums = [int((pred.sum(dim=1) == 0).sum().item() - (~body).sum().item())]
    for i in range(3):
        nums.append(int((pred[:, i] * body).sum().item()))
    return nums


def att_weight_map(output_fp: torch.Tensor, body: torch.Tensor, p: float, task: str):
    """ptqer.py:210-235: weight_k = (max(nums)/nums_k)^p, 1.0 for an empty class."""
    out = output_fp[-1]
    if task == "lits":
        nums = class_counts_lits(torch.max(out, 1)[1], body)
        n_class = 3
    elif task == "brats":
        nums = class_counts_brats((torch.sigmoid(out) >= 0.5).int(), body)
        n_class = 4
    else:
        raise RuntimeError(f"Unknown task {task}")
    wmap = {}
    for k in range(n_class):
        wmap[k] = 1.0 if nums[k] == 0 else (1 / nums[k] * max(nums)) ** p
    return wmap, nums


def mask_pyramid(output_fp: torch.Tensor, body: torch.Tensor, wmap: dict, init_stride,
                 num_lvls: int = 5, task: str = "lits") -> List[torch.Tensor]:
    """ptqer.py:141-169: five (N,D,H,W) masks at strides init*{1,2,4,8,16}."""
    init_stride = _triple(init_stride)
    out = F.avg_pool3d(output_fp[-1], init_stride)
    body = F.max_pool3d(body.float(), init_stride).bool()
    levels = []
    for _ in range(num_lvls):
        pred = pred_lits(out) if task == "lits" else pred_brats(out)
        m = torch.ones_like(pred)
        for k, v in wmap.items():
            m[pred == k] = v
        m[~body] = 1
        levels.append(m.float())
        out = F.avg_pool3d(out, 2)
        body = F.max_pool3d(body.float(), 2).bool()
    return levels


# ----------------------------------------------------------------------------
# a15: end-to-end activation-range refinement (reference: src/ptqer.py:238-272)
# ----------------------------------------------------------------------------
def ste_grads(x: torch.Tensor, alpha: float, num_lvl: int, grad_out: torch.Tensor, lo: float = 0.0, hi: float = 1.0):
    """Gradients of  qact = discretize(x/alpha, L, lo, hi) * alpha  (PTQConv.py:114-116) under the
    reference's straight-through estimator (layer_helper.py:13-37): round() has identity gradient,
    torch.clamp passes the gradient where lo <= u <= hi.  With u = x/alpha, m = 1[lo <= u <= hi]:
    grad_x = m*(((g*alpha)*delta)/delta)/alpha,  grad_alpha = sum g*(D(u) - m*u).  Returns (grad_x, grad_alpha as fp64)."""
    a = torch.tensor(float(alpha), dtype=torch.float32)
    u = x.float() / a
    m = (u >= lo) & (u <= hi)
    d = discretize(u, num_lvl, lo, hi)
    # autograd's op order through Qvar*alpha, t*delta+lo, round (identity), (var-lo)/delta, clamp, x/alpha:
    # (((g*alpha)*delta)/delta)/alpha in fp32
    delta = torch.tensor((hi - lo) / (num_lvl - 1), dtype=torch.float32)
    gx = torch.where(m, (((grad_out.float() * a) * delta) / delta) / a, torch.zeros_like(grad_out, dtype=torch.float32))
    ga = (grad_out.double() * (d.double() - torch.where(m, u, torch.zeros_like(u)).double())).sum().item()
    return gx, ga


def adam_update(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int, lr: float = 5e-4,
                beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """One torch.optim.Adam step (the optimiser of ptqer.py:259; no weight decay, no amsgrad) on fp32
    vectors, restated from torch/optim/adam.py::_single_tensor_adam.  Returns (p, m, v)."""
    m = m + (1 - beta1) * (g - m)                         # exp_avg.lerp_(grad, 1 - beta1)
    v = v * beta2 + (1 - beta2) * g * g                   # exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


# ----------------------------------------------------------------------------
# f.4: glue ops between the quantizer layers.  The reference runs them as stock modules -- nn.ReLU
# (src/models/factoryQ.py:66-81), nn.MaxPool3d(k, k) (src/models/factory_blk.py:18-42), nn.Upsample(scale_factor,
# mode="trilinear") followed by "+ skip" (:45-93) and the residual add (:147-166); restated here in explicit numpy
# fp32 arithmetic (index formulas and op order of the library kernels) and pinned against the library in
# tests/test_glue_cpu.py.
# ----------------------------------------------------------------------------
def glue_relu(x: torch.Tensor) -> torch.Tensor:
    a = x.detach().numpy()
    return torch.from_numpy(np.where(a < 0, np.float32(0), a).astype(np.float32))      # NaN passes through


def glue_add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return torch.from_numpy(a.detach().numpy().astype(np.float32) + b.detach().numpy().astype(np.float32))


def glue_maxpool3d(x: torch.Tensor, kernel) -> torch.Tensor:
    """MaxPool3d(kernel, stride=kernel), no padding, floor mode; a NaN in the window wins."""
    kd, kh, kw = _triple(kernel)
    a = x.detach().numpy()
    n, c, d, h, w = a.shape
    od, oh, ow = d // kd, h // kh, w // kw
    win = a[:, :, :od * kd, :oh * kh, :ow * kw].reshape(n, c, od, kd, oh, kh, ow, kw)
    win = win.transpose(0, 1, 2, 4, 6, 3, 5, 7).reshape(n, c, od, oh, ow, kd * kh * kw)
    out = np.full((n, c, od, oh, ow), -np.inf, dtype=np.float32)
    for t in range(kd * kh * kw):
        v = win[..., t]
        take = (v > out) | np.isnan(v)
        out = np.where(take, v, out)
    return torch.from_numpy(out.astype(np.float32))


def _linear_taps(n_in: int, factor: int):
    """align_corners=False source positions of the n_in * factor outputs: src = (1/factor) (o + 0.5) - 0.5 clamped at 0,
    i0 = floor(src), i1 = i0 + (i0 < n_in - 1), lambda1 = src - i0, lambda0 = 1 - lambda1 (all fp32)."""
    o = np.arange(n_in * factor, dtype=np.float32)
    src = np.float32(1.0 / factor) * (o + np.float32(0.5)) - np.float32(0.5)
    src = np.maximum(src, np.float32(0))
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = (src - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def glue_upsample_trilinear(x: torch.Tensor, factor, skip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Upsample(scale_factor=factor, mode="trilinear") [+ skip]: the lambdas nested w -> h -> d as in the library."""
    fd, fh, fw = _triple(factor)
    a = x.detach().numpy().astype(np.float32)
    d0, d1, ld0, ld1 = _linear_taps(a.shape[2], fd)
    h0, h1, lh0, lh1 = _linear_taps(a.shape[3], fh)
    w0, w1, lw0, lw1 = _linear_taps(a.shape[4], fw)

    def along_w(plane):                 # plane: [..., W] -> [..., W * fw]
        return lw0 * plane[..., w0] + lw1 * plane[..., w1]

    def along_h(vol, di):               # vol: [N, C, D, H, W] at depth taps di -> [N, C, OD, OH, OW]
        rows0 = along_w(vol[:, :, di][:, :, :, h0])
        rows1 = along_w(vol[:, :, di][:, :, :, h1])
        return lh0[:, None] * rows0 + lh1[:, None] * rows1

    out = ld0[:, None, None] * along_h(a, d0) + ld1[:, None, None] * along_h(a, d1)
    out = out.astype(np.float32)
    if skip is not None:
        out = out + skip.detach().numpy().astype(np.float32)
    return torch.from_numpy(out)
