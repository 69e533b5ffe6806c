/*
 * effq_b200.h -- C-ABI of the B200-native EfficientQ PTQ calibration kernels.
 *
 * This is the drop-in boundary below the reference's per-layer quantizer module
 * (reference: src/models/PTQConv.py:11-175, src/models/EfficientQConv.py:33-166).
 * The reference has no FFI of its own (it is pure Python calling ATen library
 * kernels); each entry point below replaces the ATen call sequence at the cited
 * reference file:line, and INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add to bind it.
 *
 * Conventions
 *   - every pointer is a caller-owned DEVICE pointer unless the name ends in _host;
 *   - nothing is allocated inside the library: scratch comes in through explicit
 *     workspace pointers whose size is returned by the matching *_workspace() query;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     no call synchronises the device unless documented;
 *   - return value: 0 = ok, non-zero = CUDA or argument error, text via
 *     effq_last_error() (thread-local);
 *   - scalars that are produced on the device (scales, losses) stay on the device
 *     so that the 200-iteration ADMM loop needs no host round trip.
 *
 * Layouts
 *   activations fp32 : NCDHW (the reference's / PyTorch's native layout)
 *   activation codes : N,D,H,W,C  bf16 (channels-last-3d), integer values 0..L-1
 *   weights fp32     : [C2][C1][kd][kh][kw]  (== reference weight.reshape(C2,-1))
 *   weight codes     : bf16 values 2c-(L-1) (odd integers), [tap][C1/CG][C2][CG] with
 *                      CG = min(C1,64) and the 16-byte chunks of every row pre-swizzled
 *                      for the UMMA 128/64/32-byte swizzle (csrc/tc_layout.cuh)
 *   normal equations : K' = C1*kd*kh*kw (+1 bias slot last), row order (c,kd,kh,kw)
 *                      exactly as reference src/models/solver.py:104-108.
 */
#ifndef EFFQ_B200_H
#define EFFQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EFFQ_ABI_VERSION 1

/* 3D convolution geometry (dilation 1, groups 1 -- all the reference's quantizer
 * layers use, src/models/model_blk.py:98-181). */
typedef struct effq_geom {
  int32_t n, c1, d, h, w;   /* input  N x C1 x D x H x W        */
  int32_t c2;               /* output channels                  */
  int32_t kd, kh, kw;       /* kernel                           */
  int32_t sd, sh, sw;       /* stride                           */
  int32_t pd, ph, pw;       /* zero padding                     */
} effq_geom;

/* Device-resident state of one scale search (reference project_by_iter,
 * src/models/layer_helper.py:40-70). */
typedef struct effq_scale_state {
  double  a;          /* current / final scale                               */
  double  a_prev;     /* previous iterate                                    */
  double  s_bv;       /* last  sum(b*v)                                      */
  double  s_bb;       /* last  sum(b*b)                                      */
  int32_t passes;     /* fixed-point passes executed (reference's counter c) */
  int32_t converged;  /* 1 when |a-a_prev| <= 1e-5                           */
  int32_t failed;     /* 1 when passes hit num_lvl*100 (reference raises)    */
  int32_t pad_;
} effq_scale_state;

/* Device-resident bookkeeping of one layer's ADMM run
 * (reference EfficientQConv.py:92-158). */
typedef struct effq_admm_state {
  double  sse;          /* last sum of squared errors written by the conv kernel */
  float   best_loss;    /* smallest per-iteration MSE so far (fp32, as reference) */
  float   last_loss;
  int32_t best_iter;
  int32_t iter;         /* iterations tracked so far                           */
  float   conv_scale;   /* a_x*a_w/((La-1)(Lw-1)) for the code-domain conv      */
  float   a_w;          /* fp32(a_w) of the latest projection                   */
  int32_t take_;        /* scratch: 1 when the last tracked iterate became the best */
  float   best_conv_scale; /* conv_scale of the best iterate                     */
} effq_admm_state;

/* Peer-memory communicator of a sharded run (one process per GPU of ONE node): slots[r] is rank
 * r's slot buffer (effq_peer_alloc on rank r, effq_peer_open elsewhere; slots[rank] is the
 * local one).  Kernels that take a `comm` all-reduce their few doubles over it in-kernel
 * (NVLink loads/stores, csrc/peer.cuh) instead of returning to the host for an NCCL call.
 * comm == NULL means a single-GPU run. */
#define EFFQ_PEER_MAX 8
#define EFFQ_PEER_CHANNELS 2        /* 0: scale search, 1: ADMM iterate score */
typedef struct effq_peer_comm {
  void*   slots[EFFQ_PEER_MAX];
  int32_t rank;
  int32_t world;
} effq_peer_comm;

/* ---- library ------------------------------------------------------------- */
int         effq_abi_version(void);
const char* effq_last_error(void);
/* kernels launched by this library since the last reset (bench gpu_launches). */
uint64_t    effq_launch_count(void);
/* Slot buffer of this rank (zero-initialised, effq_peer_bytes() bytes) + its 64-byte CUDA IPC
 * handle; peers map it with effq_peer_open.  Host-side, synchronous. */
int64_t     effq_peer_bytes(void);
int         effq_peer_alloc(void** dev_ptr, uint8_t* handle64);
int         effq_peer_open(const uint8_t* handle64, void** dev_ptr);
int         effq_peer_close(void* dev_ptr);
int         effq_peer_free(void* dev_ptr);
void        effq_reset_launch_count(void);

/* ---- (a) fake-quant: reference layer_helper.py:25-37, PTQConv.py:110-116 ---- */
/* y = discretize(x/alpha, nlvl, lo, hi)*alpha in fp32, op-for-op as the reference's
 * CPU path (IEEE division, round-half-even).  y_out and/or code_out may be NULL.
 * code_out[i] in [0, nlvl-1] (uint8, nlvl <= 256).  alpha is a device fp32 scalar. */
int effq_fakequant_f32(const float* x, int64_t numel, const float* alpha, float lo, float hi,
                       int32_t nlvl, float* y_out, uint8_t* code_out, void* stream);

/* Qact = fp32(a) * fp32(level) with the fp64 discretize and scale of a finished scale
 * search -- the `a_act * b_act` of EfficientQConv.py:68-70 (layer_helper.py:67). */
int effq_fakequant_state(const float* x, int64_t numel, const effq_scale_state* state, float lo,
                         float hi, int32_t nlvl, float* y_out, void* stream);

/* Element type of the integer codes the tensor-core kernels consume.  Both are exact:
 * bf16 holds every code of <= 256 levels, e4m3 (1 byte, twice the K per tcgen05.mma)
 * every code of <= 16 levels; the fp32 TMEM accumulation of the integer products is exact
 * either way, so the two give bit-identical results. */
#define EFFQ_CODE_BF16 0
#define EFFQ_CODE_E4M3 1

/* Activation codes for the tensor-core kernels: NCDHW fp32 -> NDHWC integer codes, as bf16
 * (codes_bf16_out) and / or e4m3 bytes (codes_e4m3_out; nlvl <= 16, c % 16 == 0); either
 * may be NULL.  use_f64 = 1 reproduces project_by_iter's final fp64 discretize with scale
 * state->a (EfficientQConv.py:68-70); use_f64 = 0 reproduces _quantize_act with the
 * fp32 scale alpha_f32 (PTQConv.py:114-116). */
int effq_quantize_act_ndhwc(const float* x, int32_t n, int32_t c, int64_t dhw, int32_t nlvl,
                            const effq_scale_state* state, const float* alpha_f32, int32_t use_f64,
                            void* codes_bf16_out, void* codes_e4m3_out, void* stream);

/* ---- (a3) scale search: reference layer_helper.py:40-70 --------------------- */
/* v = v1 (+ v2 if non-NULL, added in fp32 first as the reference's `w_star + dual`),
 * viewed as rows x cols with leading dimensions ld1 / ld2.  Runs the whole fp64 fixed
 * point on the device in ONE cooperative launch; result in *state. */
/* Workspace: effq_scale_search_workspace(numel) bytes (numel = rows*cols; 0 gives the fixed part that
 * effq_scale_partial needs).  Tensors too large to keep on chip get room for the list of
 * "ambiguous" elements of the interval-stable passes (csrc/scale_search.cu); with a smaller
 * workspace the search falls back to plain passes (same result up to fp64 summation order). */
int64_t effq_scale_search_workspace(int64_t numel);
/* comm != NULL: v is this rank's shard; the sums of every pass are all-reduced in-kernel, so all
 * ranks run the same passes and end with the same scale (the sharded form of the search). */
int effq_scale_search(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows,
                      int64_t cols, int32_t nlvl, float lo, float hi, effq_scale_state* state,
                      void* workspace, int64_t workspace_bytes, const effq_peer_comm* comm,
                      void* stream);
/* Multi-GPU building blocks: one pass of local sums, then (after the caller has
 * all-reduced sums[0..1]) the scale update.  mode 0: sums = {sum|v|, numel};
 * mode 1: sums = {sum(b*v), sum(b*b)} for the scale in *state. */
int effq_scale_partial(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows,
                       int64_t cols, int32_t nlvl, float lo, float hi, const effq_scale_state* state,
                       int32_t mode, double* sums, void* workspace, void* stream);
int effq_scale_step(effq_scale_state* state, const double* sums, int32_t mode, int32_t nlvl,
                    void* stream);

/* ---- (a10) conv forward + reconstruction error: EfficientQConv.py:118-122 ---- */
/* Generic fp32 direct conv (any geometry): out = conv3d(x, w) + bias.
 * out may be NULL; if target != NULL the kernel also reduces
 * sum(att * (out - target)^2) (att NULL -> 1) into *sse (double, deterministic). */
int64_t effq_conv3d_f32_workspace(const effq_geom* g);
int effq_conv3d_f32(const float* x, const float* w, const float* bias, const effq_geom* g,
                    float* out, const float* target, const float* att, double* sse,
                    void* workspace, void* stream);

/* tcgen05 implicit-GEMM conv on integer codes (k = 3/stride 1/pad 1, or k = 1):
 * out = conv_scale * sum(xcode*wcode) + bias, fp32 accumulation in TMEM,
 * fused squared-error reduction against `target` in the epilogue.
 * Requires c2 % 16 == 0, c2 <= 256 and c1 in {16, 32} or a multiple of 64 (bf16 codes),
 * c1 in {32, 64} or a multiple of 128 (e4m3 codes).  Both operands use `code_dtype`. */
int effq_conv3d_tc_supported(const effq_geom* g, int32_t code_dtype);
int64_t effq_conv3d_tc_workspace(const effq_geom* g);
int effq_conv3d_tc(const void* xcodes_ndhwc, const void* wcodes, int32_t code_dtype, const float* bias,
                   const float* conv_scale, const effq_geom* g, float* out, const float* target,
                   const float* att, double* sse, void* workspace, void* stream);
/* One scale per OUTPUT CHANNEL (scale_vec[C2], device): the per-output-channel weight-scale extension. */
int effq_conv3d_tc_pc(const void* xcodes_ndhwc, const void* wcodes, int32_t code_dtype, const float* bias,
                      const float* scale_vec, const effq_geom* g, float* out, const float* target,
                      const float* att, double* sse, void* workspace, void* stream);
/* Accumulating form: out_inout += conv_scale * conv(xcodes, wcodes) + bias.  Used by the host side to run the FP
 * (un-quantised) convolution of the calibration's first pass (reference src/ptqer.py:333-335, F.conv3d in fp32) on
 * the same tcgen05 kernel: x and w become three planes of fixed-point digits each (effq_fixdigits_ndhwc) and the six
 * significant plane products are summed, first launch through effq_conv3d_tc_pc, the rest through this one.
 * scale_vec (optional, [c2]): one scale per output channel instead of *conv_scale. */
int effq_conv3d_tc_acc(const void* xcodes_ndhwc, const void* wcodes, int32_t code_dtype, const float* bias,
                       const float* conv_scale, const float* scale_vec, const effq_geom* g, float* out_inout,
                       void* workspace, void* stream);

/* [C2][C1][taps] fp32 integer weight codes (values 2c-(L-1)) -> codes of `code_dtype` in the
 * layout effq_conv3d_tc consumes (the same layout effq_admm_project emits). */
int effq_pack_wcodes(const float* codes, int32_t c2, int32_t c1, int32_t taps, int32_t code_dtype,
                     void* out, void* stream);

/* ---- (a7+a8) normal-equation statistics: solver.py:86-111, :282-314 ---------- */
/* A0 = 2 X^ diag(att) X^T (K' x K'), B0 = 2 Y diag(att) X^T (C2 x K'), X^ the
 * im2col matrix of `x` with a ones row appended when has_bias; never materialised.
 * x_scale (device fp32 scalar or NULL = 1): x holds integer codes and the true
 * activation is x_scale * x (scaling applied once to the finished sums). */
int64_t effq_gram_workspace(const effq_geom* g, int32_t has_bias);
int effq_gram_f32(const float* x, const float* x_scale, const float* y, const float* att,
                  const effq_geom* g, int32_t has_bias, float* a0_out, float* b0_out,
                  void* workspace, void* stream);

/* Scoring without re-running the conv, for layers whose input is NOT quantised (the input is
 * the same tensor in all 200 iterations): unweighted S = X^ X^T, T = Y X^T accumulated once in
 * fp64 into acc64_out[(K'+C2) x K'] (effq_gram_f64), then per iterate
 *   sse = sum_r [ w_r S w_r^T - 2 w_r.T_r ] + sum_y2      (effq_quadform_sse)
 * which equals sum((conv3d(x, G) + b* - y)^2) of EfficientQConv.py:118-122 up to fp64 rounding.
 * quadform workspace: 16 + 8*c2 bytes, zero-initialised once. */
int effq_gram_f64(const float* x, const float* y, const effq_geom* g, int32_t has_bias, double* acc64_out,
                  void* stream);
int effq_quadform_sse(const double* acc64, double sum_y2, const float* gw, const float* bstar, int32_t c2,
                      int32_t k, int32_t has_bias, double* sse, void* workspace, void* stream);
/* The same scoring for layers WITH quantised activations (the quantised input is just as constant), in
 * residual form so that nothing cancels against sum y^2: for a reference iterate (g_ref, b_ref) whose
 * conv output has been subtracted from the target, R = Y - conv3d(X^, g_ref) - b_ref, with
 * acc64 = [S ; T], T = R X^T (effq_gram_tc_f64 with y = R, att = NULL) and *sum_sq = sum R^2:
 *   sse = sum_r [ u_r S u_r^T - 2 u_r.T_r ] + *sum_sq ,   u = [gw - g_ref | bstar - b_ref]
 * g_ref == NULL gives the plain form above (u = [gw | bstar], *sum_sq = sum y^2).  Tiled fp64 kernel for
 * K' up to a few thousand (replaces EfficientQConv.py:118-122's conv + MSE per iterate).  When st != NULL
 * the launch also does the best-iterate bookkeeping of effq_admm_decide for the iterate it scored
 * (loss = fp32(sse / numel), history[iter]).  workspace: effq_quadform_delta_workspace() bytes, zeroed once. */
int64_t effq_quadform_delta_workspace(int32_t c2, int32_t kp);
int effq_quadform_delta(const double* acc64, const double* sum_sq, const float* gw, const float* bstar,
                        const float* g_ref, const float* b_ref, int32_t c2, int32_t k, int32_t has_bias,
                        double* sse, void* workspace, effq_admm_state* st, double numel, float* history,
                        void* stream);

/* Same statistics on the tensor cores for 3x3x3 / stride 1 / pad 1 layers with quantised
 * activations, from the NDHWC integer codes: A0 (with its bias row / column) and B0 in one
 * tcgen05 kernel (att and att*y enter as bf16 hi+lo splits of the left operand, the codes are
 * exact); code_scale (device fp32 scalar, activation = code_scale * code) is applied in the
 * finalize pass.  If the kernel aborts (barrier timeout) the word after the accumulator in the
 * workspace is non-zero.
 * att_exact != 0: the caller guarantees that every att * code is exactly representable in bf16
 * (e.g. the reference's integer-valued attention masks with max(att) * max(code) <= 256, or
 * att == NULL); the kernel then uses ONE bf16 term for the weighted codes instead of hi + lo
 * (half the tensor work, same exact result). */
int effq_gram_tc_supported(const effq_geom* g);
int effq_gram_tc(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y, const float* att,
                 const effq_geom* g, int32_t has_bias, int32_t att_exact, float* a0_out, float* b0_out,
                 void* workspace, void* stream);
int effq_gram_tc_accumulate(const void* xcodes_ndhwc_bf16, const float* att, const float* y,
                            const effq_geom* g, int32_t has_bias, int32_t att_exact, double* acc64,
                            int32_t ld, void* flags, void* stream);
int effq_gram_tc_accumulate2(const void* xcodes_ndhwc_bf16, const float* att, const float* y,
                             const effq_geom* g, int32_t has_bias, int32_t att_exact, double* acc64,
                             int32_t ld, void* flags, double* acc64_unweighted, int32_t rows_only, void* stream);
/* A0, B0 as effq_gram_tc AND the unweighted S = X^ X^T (K' x K' fp64, real units, bias row / column included) from
 * ONE pass (second TMEM accumulator fed with the raw codes); needs a bias.  workspace: 2 x effq_gram_workspace. */
int effq_gram_tc_dual(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y, const float* att,
                      const effq_geom* g, int32_t att_exact, float* a0_out, float* b0_out,
                      double* s_unweighted_out, void* workspace, void* stream);
/* Unweighted T = Y X^T (C2 x K') for a new target (rows-only pass of the same kernel): writes rows [K', K'+C2) of
 * the [(K'+C2) x K'] statistics matrix at stats_base.  workspace: effq_gram_workspace. */
int effq_gram_tc_rows_f64(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y,
                          const effq_geom* g, double* stats_base, void* workspace, void* stream);
/* The tcgen05 statistics as a full fp64 matrix in real units, acc64_out[(K'+C2) x K'] = [X^ diag(att) X^T ;
 * Y diag(att) X^T] (no factor 2, mirrored, code_scale applied): the operand of effq_quadform_delta.
 * workspace as effq_gram_workspace. */
int effq_gram_tc_f64(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y, const float* att,
                     const effq_geom* g, int32_t has_bias, int32_t att_exact, double* acc64_out,
                     void* workspace, void* stream);

/* ---- (a9,a11) ADMM parameter update: solver.py:316-325, EfficientQConv.py:99-144 */
/* B = B0 + eta*W0' ;  B[:, :K] += rho*(G - dual)          (solver.py:317-320) */
int effq_admm_rhs(const float* b0, const float* w0p, const float* g, const float* dual,
                  float rho, float eta, int32_t c2, int32_t k, int32_t has_bias, float* b_out,
                  void* planes_out, void* stream);
/* b_out (fp32 [C2][K']) and planes_out (three bf16 terms [3][C2][effq_split3_ld(K')] for
 * effq_solve_gemm_tc) may each be NULL, not both. */

/* ---- (a9) proximal step w* = B A^-1 on the tensor cores: solver.py:331-342 ----------- */
/* fp32 -> three bf16 terms x = x0 + x1 + x2 (24 significand bits), planes [3][rows][ldk] with
 * ldk = effq_split3_ld(cols) (multiple of 64, tail zero-filled). */
int64_t effq_split3_ld(int64_t cols);
int effq_split3_bf16(const float* src, int32_t rows, int32_t cols, int64_t ld, void* planes_out,
                     void* stream);
/* out[m][n] (leading dimension ldo) = A B^T with A = a_planes (m x k), B = b_planes (n x k),
 * evaluated as the six bf16 products of combined order <= 2 with fp32 accumulation: fp32-class
 * accuracy at the tensor-core rate.  For the solve, A = the right-hand side B of solver.py:317-320
 * and B = A^-1 (symmetric, so its rows are the K-major operand).  workspace: zero-initialised once,
 * effq_solve_gemm_tc_workspace(m, n, k, ldo) bytes (split-K partials). */
int64_t effq_solve_gemm_tc_workspace(int32_t m, int32_t n, int32_t k, int64_t ldo);
int effq_solve_gemm_tc(const void* a_planes, const void* b_planes, int32_t m, int32_t n, int32_t k,
                       float* out, int64_t ldo, void* workspace, void* stream);
/* A = A0 + rho*quasi_eye + eta*eye                         (solver.py:317,323) */
int effq_admm_lhs(const float* a0, float rho, float eta, int32_t kp, int32_t has_bias,
                  float* a_out, void* stream);
/* After the scale search on (w* + dual): G = a_w*b_w ; dual = (w* - G + dual)/dual_div;
 * b* = last column of w*; emits fp32 G (reference layout) and, if wcodes_out != NULL,
 * weight codes of `code_dtype` in the tensor-core layout; updates st->conv_scale / st->a_w. */
/* next != NULL: the same pass also assembles the NEXT iteration's right-hand side
 * B = B0 + eta*W0' (+ rho*(G - dual) on the weight columns) from the G and dual it just produced and
 * writes it as the three bf16 planes effq_solve_gemm_tc consumes (what effq_admm_rhs(..., planes_out)
 * would write; saves that launch and a re-read of G and dual). */
typedef struct effq_next_rhs {
  const float* b0;      /* [C2][K'] */
  const float* w0p;     /* [C2][K'] */
  float        rho;     /* rho of the NEXT iteration */
  float        eta;
  void*        planes;  /* [3][C2][effq_split3_ld(K')] bf16 */
} effq_next_rhs;
/* Optional `keep` of effq_admm_project: when the PREVIOUS iterate was the best so far (st->take_, written by
 * the step that scored it: effq_admm_decide / effq_admm_track / effq_quadform_delta) its G, b* and weight
 * codes -- still in g_out / bstar_out / wcodes_out on entry -- are saved before being overwritten
 * (reference EfficientQConv.py:139-142, without a launch of its own). */
typedef struct effq_admm_keep_bufs {
  float* best_g;        /* [C2][K]  */
  float* best_b;        /* [C2] or NULL */
  void*  best_wcodes;   /* same size / type as wcodes_out, or NULL */
  float* best_pc;       /* [2*C2] per-channel scales of the best iterate (per_channel_out), or NULL */
} effq_admm_keep_bufs;
int effq_admm_project(const float* wstar, int64_t ldw, float* dual, const effq_scale_state* wscale,
                      const effq_scale_state* xscale, int32_t nlvl_w, int32_t nlvl_a, int32_t c2,
                      int32_t c1, int32_t taps, int32_t has_bias, float dual_div, float* g_out,
                      float* bstar_out, void* wcodes_out, int32_t code_dtype, effq_admm_state* st,
                      const effq_next_rhs* next, const effq_admm_keep_bufs* keep, float* per_channel_out,
                      void* stream);
/* per_channel_out != NULL selects per-OUTPUT-CHANNEL weight scales (optional extension `lwq_channel_wise`; the
 * reference's live path is per-tensor, PTQConv.py:26-27): wscale is then an array of C2 states filled by
 * effq_scale_search_rows, and per_channel_out[0..C2) receives fp32(a_w) of every channel, [C2..2*C2) the conv
 * scale a_x * a_w[c] / ((La-1)(Lw-1)) that effq_conv3d_tc_pc applies per output channel. */
int effq_scale_search_rows(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows, int64_t cols,
                           int32_t nlvl, float lo, float hi, effq_scale_state* states, void* stream);
/* The two halves of effq_admm_track as launches of their own.  effq_admm_decide: loss = fp32(sse/numel)
 * (sse all-reduced in-kernel when comm != NULL), history[iter] = loss, best-iterate bookkeeping, st->take_.
 * effq_admm_keep: copy G, b* (and aux) to the best-iterate buffers if st->take_ (after the loop's last iterate). */
int effq_admm_decide(effq_admm_state* st, const double* sse, double numel, float* history,
                     const effq_peer_comm* comm, void* stream);
int effq_admm_keep(effq_admm_state* st, const float* g, const float* bstar, int64_t g_numel, int32_t c2,
                   float* best_g, float* best_b, const void* aux_src, void* aux_dst, int64_t aux_bytes,
                   const float* pc_src, float* pc_dst, void* stream);
/* loss = fp32(sse/numel); history[iter] = loss; if (iter==0 || loss < best) keep G, b*
 * (and, when aux_bytes > 0, the 16B-aligned side buffer aux_src -> aux_dst, e.g. the
 * tensor-core weight codes of the same iterate). */
/* comm != NULL: *sse is this rank's share; it is all-reduced in-kernel (numel is the global count). */
int effq_admm_track(effq_admm_state* st, const double* sse, double numel, const float* g,
                    const float* bstar, int64_t g_numel, int32_t c2, float* best_g, float* best_b,
                    float* history, const void* aux_src, void* aux_dst, int64_t aux_bytes,
                    const effq_peer_comm* comm, void* stream);

/* ---- (a9) dense SPD factorisation and inverse on the repo's own kernels -------------------------------------
 * The reference solves A x = B^T with an fp32 LU of the K' x K' normal matrix in every iteration
 * (solver.py:331).  A takes 5 values per layer; each is factorised (blocked right-looking Cholesky) and inverted
 * (block triangular inverse, then W^T W) once, the O(n^3) work on the tensor cores through effq_gemm_tc_ex
 * (efficientq_b200/spd_inverse.py drives the block loop). */
/* out = alpha * A B^T + beta * c_in on SUB-BLOCKS of split-plane matrices (three bf16 terms per fp32 value,
 * effq_split3_bf16 / effq_split3_block): A = m x k block at a_planes with row pitch a_ld and plane stride
 * a_plane_stride (elements, multiples of 8), B = n x k likewise.  c_in may be NULL (beta ignored) and may alias
 * out.  `lower_only` is a bit field: bit 0 skips the 128 x 128 tiles strictly above the diagonal; bits 1-2 declare a
 * triangular B operand whose structurally zero K range is skipped (2: B = rows of a lower triangular matrix,
 * 4: B = rows of its transpose).  fp32-class accuracy. */
int64_t effq_gemm_tc_ex_workspace(int32_t m, int32_t n, int32_t k, int64_t ldo);
int effq_gemm_tc_ex(const void* a_planes, int64_t a_ld, int64_t a_plane_stride, const void* b_planes,
                    int64_t b_ld, int64_t b_plane_stride, int32_t m, int32_t n, int32_t k, float alpha,
                    float beta, const float* c_in, int64_t ldc, float* out, int64_t ldo, int32_t lower_only,
                    void* workspace, void* stream);
/* One diagonal block (nb <= 128): a[nb][lda] <- its Cholesky factor (lower triangle); w_out[nb][ldw] <- the
 * factor's inverse (lower triangular, zero above), wt_out[128][128] <- that inverse transposed (dense, zero
 * padded).  *info stays 0 or receives 128 * block_index + (1-based failing pivot), first failure wins. */
int effq_potrf_tile(float* a, int64_t lda, int32_t nb, float* w_out, int64_t ldw, float* wt_out, int32_t* info,
                    int32_t block_index, void* stream);
/* fp32 block src[rows][cols] (pitch ld) -> three bf16 terms at dst inside a plane matrix (row pitch dst_ld,
 * plane stride dst_plane, elements); transpose != 0 writes src^T; the K direction is zero-filled up to pad_k. */
int effq_split3_block(const float* src, int32_t rows, int32_t cols, int64_t ld, void* dst, int64_t dst_ld,
                      int64_t dst_plane, int32_t transpose, int32_t pad_k, void* stream);

/* ---- end-to-end activation-range refinement: reference ptqer.py:238-272 ------------------
 * (tune_activation_range: Adam on every alpha_act through the straight-through estimator of
 * layer_helper.py:13-37; defined in the reference, not called by its do_ptq.) */
/* Backward of qact = discretize(x/alpha, nlvl, lo, hi)*alpha given grad_out = dL/dqact:
 *   grad_x_out[i]   = grad_out[i] * 1[lo <= x[i]/alpha <= hi]           (may be NULL)
 *   *grad_alpha_acc += sum_i grad_out[i] * (D(x[i]/alpha) - 1[..] * x[i]/alpha)   (fp64, deterministic)
 * workspace: effq_ste_bwd_workspace() bytes, zeroed once by the caller (self-resetting). */
int64_t effq_ste_bwd_workspace(void);
int effq_fakequant_ste_bwd(const float* x, const float* grad_out, int64_t numel, const float* alpha,
                           float lo, float hi, int32_t nlvl, float* grad_x_out, double* grad_alpha_acc,
                           void* workspace, void* stream);
/* Operand of the tensor-core dgrad: NCDHW fp32 -> three NDHWC bf16 planes, hi + mid + lo == x exactly.
 * effq_split3_ndhwc_supported returns the tile size (> 0) when the shape is handled. */
int effq_split3_ndhwc_supported(int32_t c, int64_t dhw);
int effq_split3_ndhwc(const float* x, int32_t n, int32_t c, int64_t dhw, void* hi_out, void* mid_out,
                      void* lo_out, void* stream);
/* Operands of the FP (un-quantised) convolution of the calibration's first pass on the tensor cores (reference
 * src/ptqer.py:333-335: F.conv3d in fp32): effq_channel_absmax gives max|x| per channel (out_zeroed: c floats, zeroed
 * by the caller); with the power-of-two channel scales 2^ch_exp[c] >= max|x| effq_fixdigits_ndhwc writes
 * X = rint(x * 2^(23 - ch_exp[c])) as three balanced base-256 digit planes (NDHWC bf16, X = d0 2^16 + d1 2^8 + d2).
 * Digit products accumulate exactly in the fp32 tensor-core accumulator (same shapes as effq_split3_ndhwc). */
int effq_channel_absmax(const float* x, int32_t n, int32_t c, int64_t dhw, float* out_zeroed, void* stream);
int effq_fixdigits_ndhwc(const float* x, int32_t n, int32_t c, int64_t dhw, const int32_t* ch_exp, void* d0_out,
                         void* d1_out, void* d2_out, void* stream);
/* torch.optim.Adam update (no weight decay / amsgrad) of n fp32 scalars from fp64 gradients scaled by
 * grad_scale (1/world after an all-reduce); step counts from 1. */
int effq_adam_step(float* params, const double* grads, float grad_scale, float* exp_avg, float* exp_avg_sq,
                   int32_t n, float lr, float beta1, float beta2, float eps, int32_t step, void* stream);

/* ---- glue ops between the quantizer layers, NCDHW fp32, one HBM pass each -------------------------
 * Reference: the stock modules of src/models/factoryQ.py:66-81 (ReLU of a unit), factory_blk.py:18-42 (MaxPool3d ->
 * unit), :45-93 (trilinear Upsample, "+ skip"), :147-166 (residual add).
 * effq_glue_elementwise:  y = relu(a) (b == NULL, relu != 0) | a + b | relu(a + b); y may alias a.
 * effq_glue_maxpool3d:    MaxPool3d(kernel = stride = (kd, kh, kw)), no padding, floor mode, over nc = N*C volumes of
 *                         d x h x w; relu != 0 applies the ReLU of the unit that follows in the same pass.
 * effq_glue_upsample_trilinear: nn.Upsample(scale_factor = (fd, fh, fw), mode = "trilinear", align_corners = False)
 *                         in the library's fp32 op order; skip != NULL adds the skip connection (output shape) in the
 *                         same pass. */
int effq_glue_elementwise(const float* a, const float* b, int64_t numel, int32_t relu, float* y_out, void* stream);
int effq_glue_maxpool3d(const float* x, int64_t nc, int32_t d, int32_t h, int32_t w, int32_t kd, int32_t kh,
                        int32_t kw, int32_t relu, float* y_out, void* stream);
int effq_glue_upsample_trilinear(const float* x, const float* skip, int64_t nc, int32_t d, int32_t h, int32_t w,
                                 int32_t fd, int32_t fh, int32_t fw, float* y_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EFFQ_B200_H */
