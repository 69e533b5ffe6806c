"""Per-layer EfficientQ calibration on the GPU (the hot loop).

Mirrors reference ``EfficientQConv.ptq`` (src/models/EfficientQConv.py:33-166) and
``QuadraSolver`` (src/models/solver.py:201-345) step for step, with every tensor
operation replaced by a launch of the C-ABI kernels (``ops``).  What stays in
PyTorch is plumbing only: memory, the dense SPD factorisation of the K'xK' normal
matrix (once per distinct rho, 5 per layer -- the reference re-factorises all 200
times, solver.py:331) and the small GEMM  w* = B A^-1  per iteration (library
linear algebra, DESIGN.md section "solve").

The 200-iteration loop enqueues without a single host synchronisation: scales,
losses and the best-iterate decision live in device structs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import os

import torch

from . import ops
from .dist import DistCtx


@dataclass
class LayerReport:
    name: str = ""
    final_loss: float = 0.0          # attention-weighted, as logged in layer_loss.txt
    best_loss: float = 0.0           # best per-iteration (unweighted) MSE
    best_iter: int = 0
    alpha_w: float = 0.0
    alpha_act: Optional[float] = None
    act_passes: int = 0
    rho_scale: float = 1.0
    used_tc: bool = False
    history: Optional[List[float]] = None
    factorizations: int = 0
    fp64_factor: bool = False        # fp32 Cholesky hit a non-positive pivot -> this layer's systems went to fp64
    lu_factor: bool = False          # ... and fp64 Cholesky failed too -> fp64 LU (the reference's own method)


def select_att(mask_pyramid: Optional[Sequence[torch.Tensor]], out_spatial) -> Optional[torch.Tensor]:
    """EfficientQConv.py:53-59."""
    if not mask_pyramid:
        return None
    for m in mask_pyramid:
        if tuple(m.shape[1:]) == tuple(out_spatial):
            return m
    return None


def _moments(t: torch.Tensor, dist: DistCtx):
    """(numel, unbiased std, sum of squares) over all ranks' shards (torch.std is unbiased,
    EfficientQConv.py:45-48)."""
    td = t.double()
    s = torch.stack([td.sum(), (td * td).sum(), torch.tensor(float(t.numel()), dtype=torch.float64, device=t.device)])
    s = dist.all_reduce_sum(s)
    tot, sq, n = [float(v) for v in s.tolist()]
    var = max(sq - tot * tot / n, 0.0) / max(n - 1.0, 1.0)
    return n, var ** 0.5, sq


class LayerCalibrator:
    """Holds the reusable device scratch of one process and calibrates layers one at a time."""

    def __init__(self, device, dist: Optional[DistCtx] = None, n_iter: int = 200, rho0: float = 10.0,
                 rho_max: float = 1000.0, eta0: float = 1.0, rho_period: int = 50, keep_history: bool = False,
                 force_generic: bool = False):
        self.device = device
        self.dist = dist or DistCtx()
        self.n_iter, self.rho0, self.rho_max, self.eta0, self.rho_period = n_iter, rho0, rho_max, eta0, rho_period
        self.keep_history = keep_history
        self.force_generic = force_generic
        self.xstate = ops.ScaleState(device)
        self.wstate = ops.ScaleState(device)
        self.st = ops.AdmmState(device)
        self.sse = torch.zeros(1, dtype=torch.float64, device=device)
        self.sums = torch.zeros(2, dtype=torch.float64, device=device)
        self.sp_ws = ops.workspace(ops.capi.load().effq_scale_search_workspace(0), device)
        self.gram_ws = None
        self.tc_ws = ops.workspace(16 + 8 * 1024, device)

    # -- activation scale search ------------------------------------------------------------
    def _act_scale(self, x: torch.Tensor, nlvl: int) -> None:
        if self.dist.world == 1:
            ops.scale_search(x, nlvl, 0.0, 1.0, self.xstate)
            return
        peer = self.dist.peer_link(self.device)
        if peer is not None:
            # sharded volumes, one node: the same single launch, the sums of every pass all-reduced
            # in-kernel over NVLink peer memory
            ops.scale_search(x, nlvl, 0.0, 1.0, self.xstate, comm=peer.comm_ptr)
            return
        # NCCL form: one pass of local sums, all-reduce 2 doubles, device-side update
        ops.scale_partial(x, nlvl, 0.0, 1.0, self.xstate, 0, self.sums, self.sp_ws)
        self.dist.all_reduce_sum(self.sums)
        ops.scale_step(self.xstate, self.sums, 0, nlvl)
        done = False
        while not done:
            for _ in range(16):
                ops.scale_partial(x, nlvl, 0.0, 1.0, self.xstate, 1, self.sums, self.sp_ws)
                self.dist.all_reduce_sum(self.sums)
                ops.scale_step(self.xstate, self.sums, 1, nlvl)
            s = self.xstate.read()
            done = bool(s["converged"] or s["failed"])

    # -- a9: (A0 + rho*quasi_eye + eta*I)^-1 ------------------------------------------------------
    def inverse_of(self, a0: torch.Tensor, rho: float, eta: float, has_bias: bool, solve_tc: bool, fstate: dict,
                   rep: Optional[LayerReport] = None, check_first: bool = True, slot: int = 0):
        """One normal matrix of the layer (solver.py:316-331), assembled, factorised and inverted on the current
        stream.  Returns (A^-1 -- as three bf16 planes when ``solve_tc`` --, info tensor of the factorisation).
        ``fstate['use64']`` carries the layer's decision to factorise in fp64 (taken on its first, worst-conditioned
        system when ``check_first``).  ``slot`` selects one of the independent inverters (own buffers and recorded
        launch sequences), so that the systems of a layer can be factorised concurrently on different streams."""
        kp = a0.shape[0]
        dev = a0.device
        rep = rep or LayerReport()
        a_r = torch.empty_like(a0)
        ops.admm_lhs(a0, rho, eta, has_bias, a_r)
        own = not self.force_generic and os.environ.get("EFFQ_SPD", "1") != "0"
        inv_r = None
        if fstate.get("force64") and not fstate["use64"]:
            fstate["use64"] = True
            rep.fp64_factor = True
        if not fstate["use64"]:
            if own:
                # blocked Cholesky + block triangular inverse + W^T W on the tensor cores (spd_inverse.py): no library
                if self._spds is None:
                    self._spds = {}
                if slot not in self._spds:
                    from .spd_inverse import SpdInverter
                    self._spds[slot] = SpdInverter(dev)
                spd = self._spds[slot]
                if solve_tc and kp >= self.FACTOR_FORM_MIN and os.environ.get("EFFQ_FACTOR_FORM", "1") != "0":
                    # large systems: keep W = L^-1 and apply A^-1 = W^T W as TWO triangular products per iteration.
                    # An explicit fp32 A^-1 carries an error ~cond(A) eps -- on the K' = 6913 level (cond ~1e6) the
                    # calibrated layers ended 10 % (32 x 128^3) to 47 % (8 x 64^3) above the loss reached with an fp64
                    # inverse or with the reference's (backward-stable) LU solve; W is good to ~sqrt(cond) eps, and
                    # the two half-empty products cost what the one full product did (profiles/r02_factor_form.txt)
                    wpl, info = spd.invert(a_r, want_inverse=False)
                    if not (check_first and int(info.item()) != 0):
                        return ("w", wpl[0], wpl[1]), info
                    del wpl                            # non-positive pivot in fp32: the fp64 fallback below
                inv_r, info = spd.invert(a_r, copy=not solve_tc)
            else:
                chol, info = ops.timer.run("lib_cholesky", {"flops": kp ** 3 / 3.0},
                                           lambda: torch.linalg.cholesky_ex(a_r))
            # The smallest rho is the worst-conditioned system.  With fewer voxels than unknowns
            # (LiTS deepest level: V = 400, K' = 13825) cond(A) = lambda_max(A0)/(rho+eta) reaches
            # 1e8+ and an fp32 pivot can come out negative: check it once (one sync per layer, on
            # the side stream) and, if so, factorise and invert every A of this layer in fp64 (library: the
            # robustness fallback, like the reference's own CPU retry at solver.py:329-337).
            if check_first and int(info.item()) != 0:
                fstate["use64"] = True
                rep.fp64_factor = True
        if fstate["use64"]:
            a64 = a_r.double()
            chol64, info = ops.timer.run("lib_cholesky_f64", {"flops": kp ** 3 / 3.0},
                                         lambda: torch.linalg.cholesky_ex(a64))
            if int(info.item()) != 0 or fstate.get("force_lu"):
                # not positive definite even in fp64 (A0 itself is off by more than rho + eta):
                # LU, which is what the reference's torch.linalg.solve does (solver.py:331)
                inv64, info = ops.timer.run("lib_lu_inverse_f64", {"flops": 2.0 * kp ** 3},
                                            lambda: torch.linalg.inv_ex(a64))
                rep.lu_factor = True
            else:
                inv64 = ops.timer.run("lib_cholesky_inverse_f64", {"flops": 2.0 * kp ** 3 / 3.0},
                                      lambda: torch.cholesky_inverse(chol64))
            inv_r = inv64.float()
            del a64, chol64, inv64
        elif inv_r is not None:
            pass                                       # own factorisation + inverse, done above
        elif solve_tc and kp >= 1024 and os.environ.get("EFFQ_INV_TC", "1") != "0":
            # A^-1 = L^-T L^-1: one library TRSM for W = L^-1, then W^T W on the tensor cores
            # (the library's potri runs at ~5 TFLOP/s and was the largest item of the step)
            eye = self._eye(kp, dev)
            w_inv = ops.timer.run("lib_trsm", {"flops": float(kp) ** 3},
                                  lambda: torch.linalg.solve_triangular(chol, eye, upper=False))
            wt = ops.split3_bf16(w_inv.T)              # rows of W^T: K-major operand of (W^T W)[i][j]
            inv_r, self._sg_ws2 = ops.solve_gemm_tc(wt, wt, kp, ws=self._sg_ws2)
            del w_inv, wt
        else:
            inv_r = ops.timer.run("lib_cholesky_inverse", {"flops": 2.0 * kp ** 3 / 3.0},
                                  lambda: torch.cholesky_inverse(chol))
        if solve_tc:
            # A^-1 is symmetric: a column-major result is read as its (row-major) transpose, no copy
            inv_rm = inv_r if inv_r.stride(1) == 1 else inv_r.T
            inv_r = ops.timer.run("split3_bf16", {"bytes": 10 * kp * kp}, lambda: ops.split3_bf16(inv_rm))
        return inv_r, info

    @staticmethod
    def use_solve_tc(kp: int, force_generic: bool = False) -> bool:
        """The per-iteration product B A^-1 runs on the tensor cores from bf16 split planes (fp32-class accuracy,
        csrc/solve_gemm_tc.cu; TMA zero-fills partial tiles, so every K' qualifies).  ``force_generic`` /
        EFFQ_SOLVE_TC=0 select the library SGEMM (bring-up comparison only)."""
        return kp >= 8 and not force_generic and os.environ.get("EFFQ_SOLVE_TC", "1") != "0"

    def proximal_step(self, a0, b0, w0p, g, dual, rho: float, eta: float, has_bias: bool, fstate=None):
        """One stand-alone proximal step  w* = (B0 + eta W0' + rho (G - dual)) A^-1  (solver.py:316-345) through
        the same kernels as the loop in ``run``; returns the C2 x K' solution.  Used by the solve-chain tests."""
        c2, kp = b0.shape
        solve_tc = self.use_solve_tc(kp, self.force_generic)
        inv_r, info = self.inverse_of(a0, rho, eta, has_bias, solve_tc, fstate if fstate is not None else {"use64": False})
        if int(info.item()) != 0:
            raise ops.EffqError("normal matrix is numerically singular")
        if solve_tc:
            planes = torch.empty((3, c2, ops.split3_ld(kp)), dtype=torch.bfloat16, device=a0.device)
            ops.admm_rhs(b0, w0p, g, dual, rho, eta, None, planes=planes)
            if isinstance(inv_r, tuple):
                sol = torch.empty((c2, (kp + 3) // 4 * 4), dtype=torch.float32, device=a0.device)[:, :kp]
                z = torch.empty((c2, (kp + 3) // 4 * 4), dtype=torch.float32, device=a0.device)[:, :kp]
                return self._apply_factor_form(planes, inv_r, kp, z, torch.empty_like(planes), sol)
            sol, self._sg_ws = ops.solve_gemm_tc(planes, inv_r, kp, ws=self._sg_ws)
            return sol
        bmat = torch.empty((c2, kp), dtype=torch.float32, device=a0.device)
        ops.admm_rhs(b0, w0p, g, dual, rho, eta, bmat)
        return torch.matmul(bmat, inv_r)

    def _apply_factor_form(self, bplanes, fac, kp, z_buf, zplanes, sol_buf):
        """w* = (B W^T) W with W = L^-1 (A = L L^T): two tensor-core products that skip the structurally zero half of
        the triangular operand, and a split of the intermediate in between."""
        _, wp, wtp = fac
        _, self._sg_ws = ops.gemm_tc_planes(bplanes, wp, kp, out=z_buf, ws=self._sg_ws, tri=1)
        ops.split3_bf16(z_buf, out=zplanes)
        _, self._sg_ws = ops.gemm_tc_planes(zplanes, wtp, kp, out=sol_buf, ws=self._sg_ws, tri=2)
        return sol_buf

    # -- the layer -----------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], out_fp: torch.Tensor,
            stride, padding, qlvl_w: int, qlvl_act: int, q_act: bool, mask_pyramid=None, name: str = "",
            resume=None, stop_iter: Optional[int] = None, channel_wise: bool = False):
        """Returns (weight*, bias*, alpha_w (0-dim device fp32), alpha_act or None, layer output, report).

        ``resume = (G, dual, start_iter)`` enters the ADMM loop at iteration ``start_iter`` with that state (rho follows
        from the schedule) and ``stop_iter`` leaves it early: the one-step-ahead parity tests run single iterations
        from states recorded in the reference (tests/test_gpu_parity.py).  Device-side bookkeeping (history index,
        best-iterate seed) then counts from the first executed iteration.

        ``channel_wise``: one weight scale per OUTPUT CHANNEL -- the projection runs project_by_iter on every row of
        w* + dual separately (optional extension, north star "per output channel"; the reference's live path is
        per-tensor, PTQConv.py:26-27, which stays the default and the parity path).  alpha_w is then a [C2] vector."""
        dev = self.device
        dist = self.dist
        x = x.detach().contiguous().float()
        out_fp = out_fp.detach().contiguous().float()
        w0 = weight.detach().contiguous().float()
        c2, c1 = w0.shape[:2]
        ksize = tuple(w0.shape[2:])
        taps = ksize[0] * ksize[1] * ksize[2]
        k = c1 * taps
        has_bias = bias is not None
        kp = k + (1 if has_bias else 0)
        rep = LayerReport(name=name)

        att = select_att(mask_pyramid, out_fp.shape[2:])
        if att is not None:
            att = att.to(dev).contiguous().float()

        # rho_scale (EfficientQConv.py:44-49, :60-61) -- global statistics over all shards
        y_n, y_std, y_sq = _moments(out_fp, dist)
        w_std = float(w0.std().item())
        rs = max(y_n * y_std / (w0.numel() * w_std), 1.0)
        if att is not None:
            a_s = dist.all_reduce_sum(torch.stack([att.double().sum(),
                                                   torch.tensor(float(att.numel()), dtype=torch.float64, device=dev)]))
            rs *= float(a_s[0] / a_s[1])
        rep.rho_scale = rs
        rho, rho_m, eta = self.rho0 * rs, self.rho_max * rs, self.eta0 * rs

        # activations (EfficientQConv.py:64-72)
        use_tc = (not self.force_generic) and q_act and qlvl_act <= 256 and qlvl_w <= 256 and \
            ops.conv3d_tc_supported(x.shape, c2, ksize, stride, padding)
        rep.used_tc = use_tc
        # <= 16 levels on both sides: the conv runs on e4m3 codes (exact, K = 32 per tcgen05.mma)
        use_fp8 = use_tc and ops.fp8_codes_enabled() and qlvl_act <= 16 and qlvl_w <= 16 and \
            ops.conv3d_tc_supported(x.shape, c2, ksize, stride, padding, ops.CODE_E4M3)
        # normal-equation statistics on the tensor cores from the integer codes whenever the geometry
        # allows, also when the scoring conv cannot take the tcgen05 path (C2 > 256: LiTS 512-channel
        # level).  Besides speed this keeps A0 = 2 s^2 (exact integer Gram): an fp32 Gram of the scaled
        # values rounds every product of the same code pair the same way, a coherent perturbation of
        # ~1e-7 lambda_max that made A indefinite at that level (V = 400 voxels, K' = 13825 unknowns).
        need_gram_tc = (not self.force_generic) and q_act and qlvl_act <= 256 and \
            ops.gram_tc_supported(x.shape, c2, ksize, stride, padding)
        alpha_act = None
        xcodes = xcodes_conv = None
        if q_act:
            self._act_scale(x, qlvl_act)
            alpha_act = self.xstate.a_f32()
            qx = ops.fakequant_state(x, self.xstate, qlvl_act, 0.0, 1.0)
            if use_fp8:
                xcodes, xcodes_conv = ops.quantize_act_ndhwc(x, qlvl_act, state=self.xstate, bf16=need_gram_tc,
                                                             e4m3=True)
            elif use_tc:
                xcodes = xcodes_conv = ops.quantize_act_ndhwc(x, qlvl_act, state=self.xstate)
            elif need_gram_tc:
                xcodes = ops.quantize_act_ndhwc(x, qlvl_act, state=self.xstate)
        else:
            qx = x

        # normal-equation statistics (solver.py:253-272), summed over shards
        gram_flag = None
        # quantised 3x3x3 layers up to 64 channels (K' <= 1729): the quantised input is the same tensor in all 200
        # iterations, so after the first iterate (scored by the conv, which also leaves its output) the other 199 are
        # scored from fp64 statistics of the RESIDUAL R = Y - conv(first iterate) -- csrc/quadform.cu: the unweighted
        # S = X^ X^T rides along with the normal equations in a second TMEM accumulator of the same Gram pass, and
        # T = R X^T costs a rows-only pass later.  Beyond 64 channels the C2 K'^2 fp64 form costs more than the
        # tensor-core conv it would replace.
        qf_delta = use_tc and need_gram_tc and has_bias and tuple(ksize) == (3, 3, 3) and kp <= 2048 and \
            not self.force_generic and os.environ.get("EFFQ_QF", "1") != "0"
        stats_qf = None
        if need_gram_tc and qf_delta:
            code_scale = (self.xstate.a_f32() / float(qlvl_act - 1)).reshape(1)
            a0, b0, stats_qf, self.gram_ws, gram_flag = ops.gram_tc_dual(
                xcodes, code_scale, out_fp, att, ws=self.gram_ws, att_exact=ops.att_is_exact(att, qlvl_act - 1))
        elif need_gram_tc:
            code_scale = (self.xstate.a_f32() / float(qlvl_act - 1)).reshape(1)
            a0, b0, self.gram_ws, gram_flag = ops.gram_tc(xcodes, code_scale, out_fp, att, has_bias=has_bias,
                                                          ws=self.gram_ws,
                                                          att_exact=ops.att_is_exact(att, qlvl_act - 1))
        else:
            a0, b0 = ops.gram(qx, out_fp, att, ksize, stride, padding, has_bias=has_bias, ws=self.gram_ws)
        if dist.world > 1:
            # ONE all-reduce for A0 || B0 (latency of the second call: ~30 us x 22 layers, and NVLS prefers big buffers)
            flat = torch.cat([a0.reshape(-1), b0.reshape(-1)])
            dist.all_reduce_sum(flat)
            a0.copy_(flat[:a0.numel()].view_as(a0))
            b0.copy_(flat[a0.numel():].view_as(b0))
            del flat
        if self.probe is not None:
            for tag, t in (("x", x), ("out_fp", out_fp), ("att", att), ("alpha_act", self.xstate.buf[:8].view(torch.float64).clone() if q_act else None),
                           ("qx", qx if q_act else None), ("xcodes", xcodes), ("a0", a0), ("b0", b0)):
                self._probe(name, tag, t)
        # un-quantised input (conv0 / final_cls): the conv input never changes, so the per-iterate
        # loss comes from fp64 sufficient statistics instead of 200 fp32 convs
        stats64 = None
        if not use_tc and not q_act and kp <= 512:
            stats64 = ops.gram_f64(qx, out_fp, ksize, stride, padding, has_bias=has_bias)
            if dist.world > 1:
                dist.all_reduce_sum(stats64)
        yy_dev = torch.tensor([y_sq], dtype=torch.float64, device=dev) if stats64 is not None else None
        g_ref = b_ref = None
        gram_flag2 = None
        # quantised 1x1x1 layers (K' = 33 .. 257): the same residual-form scoring, with the statistics of the residual
        # from the generic fp64 Gram kernel (a pass over V x K' values: cheap) -- their tensor-core conv is a
        # memory-bound re-read of the target 200 times, and in a sharded run it needs an exchange per iterate
        qf_generic = use_tc and not qf_delta and has_bias and tuple(ksize) == (1, 1, 1) and kp <= 512 and \
            not self.force_generic and os.environ.get("EFFQ_QF", "1") != "0"
        w0p = torch.cat([w0.reshape(c2, k), bias.detach().float().reshape(c2, 1)], 1).contiguous() if has_bias \
            else w0.reshape(c2, k).contiguous()

        it_first, it_end = 0, self.n_iter if stop_iter is None else min(int(stop_iter), self.n_iter)
        if resume is not None:
            g = resume[0].detach().to(dev).float().reshape(c2, k).clone()
            dual = resume[1].detach().to(dev).float().reshape(c2, k).clone()
            it_first = int(resume[2])
        else:
            g = w0.reshape(c2, k).clone()
            dual = torch.zeros_like(g)
        bstar = bias.detach().float().clone() if has_bias else None
        best_g = torch.empty_like(g)
        best_b = torch.empty_like(bstar) if has_bias else None
        bmat = torch.empty((c2, kp), dtype=torch.float32, device=dev)
        hist = torch.zeros(self.n_iter, dtype=torch.float32, device=dev)
        wcodes = best_wcodes = None
        if use_tc:
            wcodes = torch.empty(taps * c1 * c2, dtype=xcodes_conv.dtype, device=dev)
            best_wcodes = torch.empty_like(wcodes)
        self.st.reset()
        wrows = pc = best_pc = None
        if channel_wise:
            wrows = ops.ScaleStateRows(dev, c2)
            pc = torch.zeros(2 * c2, dtype=torch.float32, device=dev)          # [a_w per channel | conv scale per channel]
            best_pc = torch.zeros_like(pc)
        numel_total = y_n                      # mse over every rank's outputs
        g4 = g.view(c2, c1, *ksize)

        # A = A0 + rho*quasi_eye + eta*I takes 5 distinct values per layer (the rho schedule is
        # known up front) and is SPD (A0 is PSD, eta > 0).  The reference re-factorises it in every
        # one of the 200 iterations (solver.py:331); here each value is factorised and inverted
        # once, on a side stream, so the later ones overlap the ADMM iterations of the earlier ones.
        # rho of every iteration (EfficientQConv.py:129-137: doubled after iterations 0, 50, 100, ... up to rho_max)
        rho_seq, r_ = [], rho
        for it in range(self.n_iter):
            rho_seq.append(r_)
            if it % self.rho_period == 0:
                r_ = r_ * 2 if r_ * 2 <= rho_m else rho_m
        rho_seq.append(r_)
        rhos = []
        for r_ in rho_seq[it_first:it_end]:              # only the systems the executed iterations need
            if r_ not in rhos:
                rhos.append(r_)
        rho = rho_seq[it_first]
        main = torch.cuda.current_stream(dev)
        solve_tc = self.use_solve_tc(kp, self.force_generic)
        # The systems of a layer are independent, and one blocked factorisation is a chain of ~12 dependent small
        # launches per 128-wide panel (latency-bound: 31 ms at K' = 6913, 4 % of the SMs busy): each rho gets its own
        # stream and inverter, so the five chains run side by side instead of one after the other -- the K' >= 3457
        # layers were waiting for them at every rho change (profiles/r02_factor_streams.txt).
        n_slots = self.factor_streams if kp <= 8192 else min(self.factor_streams, 2)    # a K' = 13825 plan holds 13 GB
        n_slots = max(1, min(n_slots, len(rhos)))
        if self.force_generic or os.environ.get("EFFQ_SPD", "1") == "0":
            n_slots = 1                                # the library bring-up path shares one GEMM workspace
        if self._sides is None:
            self._sides = []
        while len(self._sides) < n_slots:
            # (a higher stream priority for the two systems needed first was measured: level-4 layers 159 -> 156 ms, but
            # the un-instrumented step went from 1585 to 1667 ms -- the main stream's loop was starved; not used)
            self._sides.append(torch.cuda.Stream(device=dev))
        self._side = self._sides[0]
        for sd in self._sides[:n_slots]:
            sd.wait_stream(main)
        inverses, infos = {}, []
        fstate = {"use64": False, "force64": self.force_fp64_factor or self.force_lu_factor,
                  "force_lu": self.force_lu_factor}

        def factor_all(first_checked: bool):
            inverses.clear()
            infos.clear()
            for idx, r_ in enumerate(rhos):
                slot = idx % n_slots
                with torch.cuda.stream(self._sides[slot]):
                    inv_r, info = self.inverse_of(a0, r_, eta, has_bias, solve_tc, fstate, rep,
                                                  check_first=first_checked and idx == 0, slot=slot)
                    for t_ in (inv_r[1:] if isinstance(inv_r, tuple) else (inv_r,)):
                        t_.record_stream(main)
                    ev = torch.cuda.Event()
                    ev.record(self._sides[slot])
                inverses[r_] = (inv_r, ev)
                infos.append(info)

        if n_slots > 1 and not fstate["force64"]:
            # enqueue every system first, THEN look at the first (worst-conditioned) one's pivots: the check is a host
            # synchronisation, and taken inside the loop it would serialise the chains behind it
            factor_all(first_checked=False)
            if int(infos[0].item()) != 0:          # non-positive fp32 pivot: all systems of this layer again, in fp64
                fstate["use64"] = True
                rep.fp64_factor = True
                for sd in self._sides[1:n_slots]:
                    self._sides[0].wait_stream(sd)
                n_slots = 1
                factor_all(first_checked=False)
        else:
            n_slots = 1 if fstate["force64"] else n_slots
            factor_all(first_checked=True)
        rep.factorizations += len(rhos)
        if self.probe is not None:
            for sd in self._sides:
                sd.synchronize()
            for i_, r_ in enumerate(rhos):
                iv = inverses[r_][0]
                self._probe(name, f"inv{i_}", iv[1] if isinstance(iv, tuple) else iv)
        ainv = None
        rho_built = None
        peer = dist.peer_link(dev) if dist.world > 1 else None
        if solve_tc:
            bplanes = torch.empty((3, c2, ops.split3_ld(kp)), dtype=torch.bfloat16, device=dev)
            sol_buf = torch.empty((c2, (kp + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :kp]
            z_buf = torch.empty((c2, (kp + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :kp]     # factor form: B W^T ...
            zplanes = torch.empty_like(bplanes)                                                      # ... and its split

        loop_prof = os.environ.get("EFFQ_LOOP_PROF") == "1"     # bring-up: CPU enqueue time vs GPU time of the loop
        if loop_prof:
            import time as _t
            ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            ev_a.record()
            t_cpu0 = _t.perf_counter()
        # Iterations inside one rho block launch the same kernels on the same buffers (the iteration counter,
        # scales, losses and the best-iterate decision live in device structs), so the launch sequence of the
        # block's second iteration is recorded once (capi.record) and re-issued for the rest of the block
        # without the Python wrappers: ~17 us -> ~2 us of host time per launch.  The 1x1x1 layers were
        # host-bound (19 ms of enqueue for 19 ms of GPU time per layer, profiles/r01_loop_prof.txt).
        from contextlib import nullcontext
        from . import capi as _capi
        replayable = not (dist.world > 1 and stats64 is None and not (qf_delta or qf_generic) and peer is None) and \
            os.environ.get("EFFQ_REPLAY", "1") != "0"            # an NCCL all-reduce inside the loop cannot be re-issued
        keep_bufs = (best_g, best_b, best_wcodes, best_pc)
        steady = None
        sol_small = None if solve_tc else torch.empty((c2, kp), dtype=torch.float32, device=dev)
        score_overlap = os.environ.get("EFFQ_SCORE_STREAM", "1") != "0"
        score_pending = False
        if score_overlap:
            if self._score_stream is None:
                self._score_stream = torch.cuda.Stream(device=dev)
                self._score_events = (torch.cuda.Event(), torch.cuda.Event())
            score_stream = self._score_stream
            ev_proj, ev_score = self._score_events
        for it in range(it_first, it_end):
            if rho_built != rho:
                ainv, ev = inverses[rho]
                main.wait_event(ev)
                rho_built = rho
                steady = None
            special = it == it_first or it % self.rho_period == 0 or it + 1 == it_end or self.probe is not None
            if steady is not None and not special:
                pre, mm, post = steady
                ops.replay(pre)
                if mm is not None:
                    mm()
                # Conv-free scoring off the critical path: the score of iterate i (quadform + best-iterate decision) is
                # needed only by the NEXT projection (which keeps iterate i if it was the best, then overwrites G / b*),
                # not by the next proximal step or scale search.  It runs on a second stream between two events:
                #   main : solve_i -> search_i -> [wait score_{i-1}] -> project_i -> (event P_i) -> solve_{i+1} -> ...
                #   score:                                        [wait P_i] -> quadform_i -> (event S_i)
                # Same kernels, same arguments, same order of every dependent pair: results are bit-identical.
                split = self._split_score(post) if score_overlap else None
                if split is None:
                    ops.replay(post)
                    continue
                head, proj, score = split
                ops.replay(head)
                if score_pending:
                    main.wait_event(ev_score)
                ops.replay(proj)
                ev_proj.record(main)
                score_stream.wait_event(ev_proj)
                ops.replay(score)
                ev_score.record(score_stream)
                score_pending = True
                continue
            if score_pending:                      # a non-replayed iteration runs entirely on the main stream
                main.wait_event(ev_score)
                score_pending = False
            recording = replayable and not special
            with (_capi.record() if recording else nullcontext()) as rec:
                n_pre = 0
                mm = None
                # proximal step (solver.py:316-345): w* = solve(A, B^T)^T = B A^-1
                if solve_tc:
                    if it == it_first:     # later right-hand sides come out of admm_project of the previous iteration
                        ops.timer.run("admm_rhs", {"bytes": 22 * c2 * kp},
                                      lambda: ops.admm_rhs(b0, w0p, g, dual, rho, eta, None, planes=bplanes))
                    if isinstance(ainv, tuple):
                        sol = self._apply_factor_form(bplanes, ainv, kp, z_buf, zplanes, sol_buf)
                    else:
                        sol, self._sg_ws = ops.solve_gemm_tc(bplanes, ainv, kp, out=sol_buf, ws=self._sg_ws)
                else:
                    ops.timer.run("admm_rhs", {"bytes": 20 * c2 * kp}, lambda: ops.admm_rhs(b0, w0p, g, dual, rho, eta, bmat))
                    n_pre = len(rec.calls) if recording else 0
                    ainv_now = ainv

                    def mm(ainv_now=ainv_now):
                        ops.timer.run("lib_sgemm_B_Ainv", {"flops": 2.0 * c2 * kp * kp},
                                      lambda: torch.matmul(bmat, ainv_now, out=sol_small))
                    mm()
                    sol = sol_small
                # projection + dual update (EfficientQConv.py:107-111)
                wview = sol[:, :k] if has_bias else sol
                if channel_wise:
                    ops.scale_search_rows(wview, qlvl_w, -1.0, 1.0, wrows, v2=dual)
                else:
                    ops.scale_search(wview, qlvl_w, -1.0, 1.0, self.wstate, v2=dual)
                div = 1.0
                new_rho = rho
                if it % self.rho_period == 0:          # EfficientQConv.py:129-137
                    if rho * 2 <= rho_m:
                        new_rho, div = rho * 2, 2.0
                    else:
                        new_rho, div = rho_m, rho_m / rho
                nxt = (b0, w0p, new_rho, eta, bplanes) if (solve_tc and it + 1 < it_end) else None
                if self.probe is not None:
                    self._probe(name, f"it{it}_wstar", sol)
                # the previous iterate, if it was the best so far, is saved by this launch (keep) before G, b* and the
                # weight codes are overwritten
                ops.timer.run("admm_project", {"bytes": 16 * c2 * k}, lambda: ops.admm_project(
                    sol, dual, wrows if channel_wise else self.wstate, self.xstate if q_act else None, qlvl_w, qlvl_act, c2,
                    c1, taps, has_bias, div, g, bstar, wcodes, self.st, next_rhs=nxt, keep=keep_bufs, per_channel_out=pc))
                # score the iterate (EfficientQConv.py:118-122) and do the best-iterate bookkeeping (:139-142)
                if stats64 is not None:
                    ops.quadform_delta(stats64, yy_dev, g, bstar, self.sse, g_ref, b_ref, st=self.st,
                                       numel=numel_total, history=hist)                      # statistics are global
                else:
                    out0 = None
                    if use_tc:
                        out0, _ = ops.conv3d_tc(xcodes_conv, wcodes, bstar, self.st.conv_scale_ptr(), c2, ksize,
                                                want_out=qf_delta or qf_generic, target=out_fp, ws=self.tc_ws, sse=self.sse,
                                                scale_vec=pc[c2:] if channel_wise else None)
                    else:
                        ops.conv3d_f32(qx, g4, bstar, stride, padding, want_out=False, target=out_fp,
                                       ws=self._conv_ws(qx, c2, ksize, stride, padding), sse=self.sse)
                    track_comm = None
                    if dist.world > 1:                                    # this rank's share of the squared error
                        if peer is not None:
                            track_comm = peer.comm_ptr                    # summed inside the decide kernel over NVLink
                        else:
                            dist.all_reduce_sum(self.sse)
                    ops.timer.run("admm_decide", {"bytes": 64}, lambda: ops.admm_decide(
                        self.st, self.sse, numel_total, hist, comm=track_comm))
                    if qf_generic:
                        torch.sub(out_fp, out0, out=out0)
                        stats64 = ops.gram_f64(qx, out0, ksize, stride, padding, has_bias=has_bias)
                        yy_dev = self.sse.clone()
                        if dist.world > 1:
                            dist.all_reduce_sum(stats64)
                            if peer is not None:
                                dist.all_reduce_sum(yy_dev)
                        g_ref, b_ref = g.clone(), bstar.clone()
                        del out0
                    elif qf_delta:
                        # residual statistics of this (first executed) iterate: R = Y - out0, T = R X^T, sum R^2
                        torch.sub(out_fp, out0, out=out0)
                        code_scale = (self.xstate.a_f32() / float(qlvl_act - 1)).reshape(1)
                        self.gram_ws, gram_flag2 = ops.gram_tc_rows_f64(xcodes, code_scale, out0, stats_qf, ws=self.gram_ws)
                        stats64 = stats_qf
                        yy_dev = self.sse.clone()
                        if dist.world > 1:
                            dist.all_reduce_sum(stats64)
                            if peer is not None:                          # with NCCL scoring self.sse is already global
                                dist.all_reduce_sum(yy_dev)
                        g_ref, b_ref = g.clone(), (bstar.clone() if has_bias else None)
                        del out0
                if recording:
                    steady = (rec.calls[:n_pre], mm if not solve_tc else None, rec.calls[n_pre:])
                if self.probe is not None:
                    for tag, t in (("g", g), ("dual", dual), ("bstar", bstar), ("a_w", self.st.a_w_tensor()), ("wcodes", wcodes)):
                        self._probe(name, f"it{it}_{tag}", t)
            rho = new_rho

        if score_pending:
            main.wait_event(ev_score)
        ops.admm_keep(self.st, g, bstar, best_g, best_b, wcodes, best_wcodes, pc, best_pc)   # the last iterate, if it was the best
        if loop_prof:
            t_cpu = _t.perf_counter() - t_cpu0
            ev_b.record()
            torch.cuda.synchronize(dev)
            print(f"[loop] {name:45s} K'={kp:5d} C2={c2:3d} tc={int(use_tc)} cpu enqueue {1e3 * t_cpu:8.2f} ms   gpu {ev_a.elapsed_time(ev_b):8.2f} ms")
        # final forward with the best iterate: layer output + attention-weighted loss (:161-166)
        if use_tc:
            out_q, _ = ops.conv3d_tc(xcodes_conv, best_wcodes, best_b, self.st.best_conv_scale_ptr(), c2, ksize,
                                     want_out=True, target=out_fp, att=att, ws=self.tc_ws, sse=self.sse,
                                     scale_vec=best_pc[c2:] if channel_wise else None)
        else:
            out_q, _ = ops.conv3d_f32(qx, best_g.view(c2, c1, *ksize), best_b, stride, padding, want_out=True,
                                      target=out_fp, att=att, ws=self._conv_ws(qx, c2, ksize, stride, padding),
                                      sse=self.sse)
        if dist.world > 1:
            dist.all_reduce_sum(self.sse)
        for sd in self._sides:                     # every factorisation (also an unused last one) is ordered before the read-back
            main.wait_stream(sd)
        # LAST iterate's scale (reference quirk, :158); per channel: the [C2] vector of the last iterate
        alpha_w = pc[:c2].clone() if channel_wise else self.st.a_w_tensor().clone()
        s = self.st.read()                          # the layer's one result read-back
        final_sse = float(self.sse.item())
        if any(int(i.item()) != 0 for i in infos):
            raise ops.EffqError(f"{name}: normal matrix is numerically singular (factorisation failed"
                                f"{', fp64 LU included' if rep.lu_factor else ''})")
        if solve_tc and int(self._sg_ws[:4].view(torch.int32)[0].item()) != 0:
            raise ops.EffqError(f"{name}: tcgen05 solve GEMM aborted (barrier timeout)")
        if gram_flag2 is not None and int(gram_flag2.item()) != 0:
            raise ops.EffqError(f"{name}: tcgen05 Gram kernel (residual statistics) aborted (barrier timeout)")
        if gram_flag is not None and int(gram_flag.item()) != 0:
            raise ops.EffqError(f"{name}: tcgen05 Gram kernel aborted (barrier timeout)")
        if final_sse != final_sse:
            raise ops.EffqError(f"{name}: non-finite reconstruction error"
                                + (" (tcgen05 conv aborted on a barrier timeout, or non-finite inputs)" if use_tc
                                   else " (non-finite inputs or proximal step)"))
        rep.final_loss = final_sse / numel_total
        if s["last_loss"] != s["last_loss"]:
            raise ops.EffqError(f"{name}: NVLink peer exchange timed out (a rank is missing or out of step)")
        rep.best_loss, rep.best_iter, rep.alpha_w = s["best_loss"], s["best_iter"], s["a_w"]
        if q_act:
            xs = self.xstate.read()
            rep.alpha_act, rep.act_passes = float(xs["a"]), xs["passes"]
            if xs["failed"]:
                raise RuntimeWarning(f"Exceed maximum iteration ({qlvl_act * 100}) for alpha optimization in var_init_iter")
        if channel_wise:
            rep.alpha_w = float(alpha_w.mean().item())
            if any(r_["failed"] for r_ in wrows.read()):
                raise RuntimeWarning(f"Exceed maximum iteration ({qlvl_w * 100}) for alpha optimization in var_init_iter")
        elif self.wstate.read()["failed"]:
            raise RuntimeWarning(f"Exceed maximum iteration ({qlvl_w * 100}) for alpha optimization in var_init_iter")
        if self.keep_history:
            rep.history = hist.cpu().tolist()
        if self.probe is not None:
            for tag, t in (("hist", hist), ("best_g", best_g), ("best_b", best_b), ("alpha_w", alpha_w), ("out_q", out_q)):
                self._probe(name, tag, t)
        return best_g.view(c2, c1, *ksize), best_b, alpha_w, alpha_act, out_q, rep

    # bring-up hook (tools/repro_check.py): callable(layer name, tag, tensor) handed the intermediate results of a
    # layer; None in production (no call, no synchronisation)
    probe = None

    def _probe(self, name, tag, t):
        if self.probe is not None and t is not None:
            self.probe(name, tag, t)

    _spds = None            # slot -> SpdInverter
    _sides = None           # factorisation streams, one per slot
    FACTOR_FORM_MIN = 1024         # K' from which the proximal step multiplies with W^T and W instead of an explicit A^-1
    force_fp64_factor = False      # tests: take the fp64-Cholesky / fp64-LU fallbacks of inverse_of on any layer
    force_lu_factor = False
    _cws = None
    _sg_ws = None
    _sg_ws2 = None          # split-K partials of the K' x K' inverse product (side stream)
    _eyes = None

    def _eye(self, n, dev):
        if self._eyes is None:
            self._eyes = {}
        if n not in self._eyes:
            self._eyes[n] = torch.eye(n, dtype=torch.float32, device=dev)
        return self._eyes[n]

    _side = None
    factor_streams = int(os.environ.get("EFFQ_FACTOR_STREAMS", "5"))   # concurrent factorisation chains per layer
    _score_stream = None    # conv-free scoring of iterate i beside the proximal step of iterate i + 1
    _score_events = None
    _split_cache = None

    def _split_score(self, post):
        """Recorded tail of an iteration -> (calls up to the projection, the projection, the scoring launch re-targeted
        at the score stream), or None when the iteration is not conv-free scored (or the timer brackets the scoring
        kernel on the main stream).  Cached per recorded sequence."""
        import ctypes as C
        if self._split_cache is not None and self._split_cache[0] is post:
            split = self._split_cache[1]
        else:
            split = None
            names = [c[3] for c in post]
            side = C.c_void_p(self._score_stream.cuda_stream)
            n_score = 0
            if names and names[-1] == "effq_quadform_delta":
                n_score = 1
            elif len(names) >= 2 and names[-2:] == ["effq_conv3d_tc", "effq_admm_decide"] and \
                    os.environ.get("EFFQ_SCORE_STREAM_CONV", "1") != "0":
                n_score = 2                  # conv-scored iterate: conv + decide (the decide kernel's NVLink exchange
                                             # is the only peer traffic inside a loop, so it may run on either stream)
            if n_score and names.count("effq_admm_project") == 1 and \
                    names.index("effq_admm_project") == len(names) - n_score - 1:
                score = [(fn, tuple(args[:-1]) + (side,), tag, name) for fn, args, tag, name in post[-n_score:]]
                split = (post[:-n_score - 1], post[-n_score - 1:-n_score], score)
            self._split_cache = (post, split)
        if split is not None and any(c[2] is not None and ops.timer.wants(c[2][0]) for c in split[2]):
            return None
        return split

    def _conv_ws(self, x, c2, ksize, stride, padding):
        import ctypes as C
        g = ops.Geom.make(x.shape, c2, ksize, stride, padding)
        need = ops.capi.load().effq_conv3d_f32_workspace(C.byref(g))
        if self._cws is None or self._cws.numel() < need:
            self._cws = ops.workspace(need, self.device)
        return self._cws
