// Fused elementwise ADMM parameter-update kernels (HBM/L2-bound, C2 x K' elements).
// Replace the ATen elementwise chains at reference
//   src/models/solver.py:316-325          (getAB: A and B assembly)
//   src/models/EfficientQConv.py:107-111  (projection G = a_w*b_w, dual update)
//   src/models/EfficientQConv.py:129-142  (rho rescale of the dual, best-iterate pick)
// All scalars that the reference pulls to the host with .item() stay in device
// structs, so the 200-iteration loop of one layer enqueues without a host sync.
#include "common.cuh"
#include "tc_layout.cuh"
#include "peer.cuh"
#include "admm_decide.cuh"
#include <cuda_fp8.h>

namespace effq {

constexpr int AD_THREADS = 256;

static inline int grid_for(long long n) {
  long long b = (n + AD_THREADS - 1) / AD_THREADS;
  const long long cap = (long long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// B = (B0 + eta*W0') ; B[:, :K] += rho*(G - dual)            solver.py:317-320
__global__ void __launch_bounds__(AD_THREADS)
admm_rhs_kernel(const float* __restrict__ b0, const float* __restrict__ w0p, const float* __restrict__ g,
                const float* __restrict__ dual, float rho, float eta, int c2, int k, int kp,
                float* __restrict__ b_out, __nv_bfloat16* __restrict__ planes, int ldk) {
  // planes != NULL: also (or only) emit B as three bf16 terms [3][c2][ldk] for effq_solve_gemm_tc
  const int ldj = planes ? ldk : kp;
  const long long total = (long long)c2 * ldj;
  for (long long e = (long long)blockIdx.x * AD_THREADS + threadIdx.x; e < total;
       e += (long long)gridDim.x * AD_THREADS) {
    const int r = (int)(e / ldj), j = (int)(e % ldj);
    float v = 0.f;
    if (j < kp) {
      const long long be = (long long)r * kp + j;
      v = __fadd_rn(b0[be], __fmul_rn(eta, w0p[be]));
      if (j < k) {
        const long long ge = (long long)r * k + j;
        v = __fadd_rn(v, __fmul_rn(rho, __fsub_rn(g[ge], dual[ge])));
      }
      if (b_out) b_out[be] = v;
    }
    if (planes) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v);
      const float r1 = __fsub_rn(v, __bfloat162float(h0));
      const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
      planes[e] = h0;
      planes[total + e] = h1;
      planes[2 * total + e] = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(h1)));
    }
  }
}

// A = A0 + rho*quasi_eye + eta*eye                            solver.py:317 / :323
__global__ void __launch_bounds__(AD_THREADS)
admm_lhs_kernel(const float* __restrict__ a0, float rho, float eta, int kp, int has_bias,
                float* __restrict__ a_out) {
  const long long total = (long long)kp * kp;
  for (long long e = (long long)blockIdx.x * AD_THREADS + threadIdx.x; e < total;
       e += (long long)gridDim.x * AD_THREADS) {
    const int i = (int)(e / kp), j = (int)(e % kp);
    float v = a0[e];
    if (i == j) {
      if (has_bias) {
        const float qe = (i == kp - 1) ? 0.f : 1.f;
        v = __fadd_rn(__fadd_rn(v, __fmul_rn(rho, qe)), eta);
      } else {
        v = __fadd_rn(v, __fadd_rn(rho, eta));      // (rho + mu + eta) * eye, mu = 0
      }
    }
    a_out[e] = v;
  }
}

// integer code -> operand element (bf16, or e4m3 when |code| <= 16: both exact)
__device__ __forceinline__ void store_code(void* base, long long idx, float code, int eb) {
  if (eb == 2) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(code);
  else reinterpret_cast<uint8_t*>(base)[idx] = (uint8_t)__nv_cvt_float_to_fp8(code, __NV_SATFINITE, __NV_E4M3);
}

// x = x0 + x1 + x2 in bf16 (24 bits), planes `plane` elements apart (as split3_kernel / admm_rhs_kernel)
__device__ __forceinline__ void store_split3(__nv_bfloat16* planes, long long at, long long plane, float v) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v);
  const float r1 = __fsub_rn(v, __bfloat162float(h0));
  const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
  planes[at] = h0;
  planes[plane + at] = h1;
  planes[2 * plane + at] = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(h1)));
}

// Projection + dual update + operand export.           EfficientQConv.py:107-111,131-137
__global__ void __launch_bounds__(AD_THREADS)
admm_project_kernel(const float* __restrict__ wstar, long long ldw, float* __restrict__ dual,
                    const effq_scale_state* __restrict__ wscale, const effq_scale_state* __restrict__ xscale,
                    int nlvl_w, int nlvl_a, int c2, int c1, int taps, int has_bias, float dual_div,
                    float* g_out, float* bstar_out,
                    void* wcodes, TcLayout lay, effq_admm_state* st, effq_next_rhs nx, int ldk,
                    effq_admm_keep_bufs kp_, float* __restrict__ pc_out) {
  // pc_out != NULL: per-output-channel scales -- wscale is an array of c2 states (effq_scale_search_rows) and
  // pc_out[0..c2) receives fp32(a_w) of every row, pc_out[c2..2c2) the conv scale a_x a_w_r / ((La-1)(Lw-1))
  const int k = c1 * taps;
  const long long total = (long long)c2 * k;
  // Fused "keep": when the PREVIOUS iterate was the best so far (st->take_, set by the decide step
  // that scored it) its G / b* / weight codes are still in g_out / bstar_out / wcodes -- every thread
  // saves the element it is about to overwrite (EfficientQConv.py:139-142 without a launch of its own).
  const bool keep_prev = kp_.best_g != nullptr && st != nullptr && *((volatile int*)&st->take_) != 0;
  const bool per_row = pc_out != nullptr;
  const double a64_all = wscale->a;
  const QParamD q = make_qparam_d(-1.f, 1.f, nlvl_w);
  for (long long e = (long long)blockIdx.x * AD_THREADS + threadIdx.x; e < total;
       e += (long long)gridDim.x * AD_THREADS) {
    const int r = (int)(e / k), j = (int)(e % k);
    const double a64 = per_row ? wscale[r].a : a64_all;
    const float a32 = (float)a64;
    const float ws = wstar[(long long)r * ldw + j];
    const float du = dual[e];
    const float v = __fadd_rn(ws, du);                                   // w_star + dual (fp32)
    const double idx = level_index_d(__ddiv_rn((double)v, a64), q);      // fp64 discretize
    const float b32 = (float)level_value_d(idx, q);                      // .float()
    const float gq = __fmul_rn(a32, b32);                                // G = a_w * b_w
    if (keep_prev) kp_.best_g[e] = g_out[e];
    g_out[e] = gq;
    const float dn = __fdiv_rn(__fadd_rn(__fsub_rn(ws, gq), du), dual_div);   // (w*-G+dual) [/2 on rho steps]
    dual[e] = dn;
    if (nx.planes) {
      // next iteration's B on the weight columns, op for op as admm_rhs_kernel, split into three bf16 terms
      const long long be = (long long)r * (k + has_bias) + j;
      float v = __fadd_rn(nx.b0[be], __fmul_rn(nx.eta, nx.w0p[be]));
      v = __fadd_rn(v, __fmul_rn(nx.rho, __fsub_rn(gq, dn)));
      store_split3((__nv_bfloat16*)nx.planes, (long long)r * ldk + j, (long long)c2 * ldk, v);
    }
    if (wcodes) {
      const int c = j / taps, t = j % taps;
      const float code = (float)(2.0 * idx - (double)(nlvl_w - 1));      // odd integer in [-(L-1), L-1]
      const long long wi = tc_wcode_index(r, c, t, c1, c2, lay);
      if (keep_prev && kp_.best_wcodes) {
        if (lay.eb == 2) reinterpret_cast<uint16_t*>(kp_.best_wcodes)[wi] = reinterpret_cast<const uint16_t*>(wcodes)[wi];
        else reinterpret_cast<uint8_t*>(kp_.best_wcodes)[wi] = reinterpret_cast<const uint8_t*>(wcodes)[wi];
      }
      store_code(wcodes, wi, code, lay.eb);
    }
  }
  if (has_bias && bstar_out) {
    for (int r = blockIdx.x * AD_THREADS + threadIdx.x; r < c2; r += gridDim.x * AD_THREADS) {
      if (keep_prev && kp_.best_b) kp_.best_b[r] = bstar_out[r];
      bstar_out[r] = wstar[(long long)r * ldw + k];
    }
  }
  if (nx.planes) {
    // bias column (constant B0 + eta*W0') and the zero tail up to the plane pitch
    const int kp = k + has_bias, tail = ldk - k;
    for (long long t = (long long)blockIdx.x * AD_THREADS + threadIdx.x; t < (long long)c2 * tail;
         t += (long long)gridDim.x * AD_THREADS) {
      const int r = (int)(t / tail), j = k + (int)(t % tail);
      float v = 0.f;
      if (j < kp) { const long long be = (long long)r * kp + j; v = __fadd_rn(nx.b0[be], __fmul_rn(nx.eta, nx.w0p[be])); }
      store_split3((__nv_bfloat16*)nx.planes, (long long)r * ldk + j, (long long)c2 * ldk, v);
    }
  }
  const double ax = xscale ? (double)(float)xscale->a / (double)(nlvl_a - 1) : 1.0;
  if (per_row) {
    for (int r = blockIdx.x * AD_THREADS + threadIdx.x; r < c2; r += gridDim.x * AD_THREADS) {
      if (keep_prev && kp_.best_pc) { kp_.best_pc[r] = pc_out[r]; kp_.best_pc[c2 + r] = pc_out[c2 + r]; }
      const float a32 = (float)wscale[r].a;
      pc_out[r] = a32;
      pc_out[c2 + r] = (float)(ax * (double)a32 / (double)(nlvl_w - 1));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && st) {
    const float a32 = (float)a64_all;                    // per-row mode: row 0's scale (bookkeeping only)
    st->a_w = a32;
    st->conv_scale = (float)(ax * (double)a32 / (double)(nlvl_w - 1));
  }
}

__global__ void admm_decide_kernel(effq_admm_state* st, const double* __restrict__ sse, double numel,
                                   float* __restrict__ history, int* __restrict__ take, effq_peer_comm comm) {
  double total = *sse;
  if (comm.world > 1) {                                   // sharded volumes: sum the ranks' shares over NVLink
    double v[3] = {total, 0.0, 0.0};
    if (!peer_allreduce3(comm, 1, v)) v[0] = __longlong_as_double(0x7ff8000000000000ll);   // NaN: exchange timed out
    total = v[0];
  }
  admm_decide_dev(st, total, numel, history);
  (void)take;
}

__global__ void __launch_bounds__(AD_THREADS)
admm_keep_kernel(const int* __restrict__ take, const float* __restrict__ g, const float* __restrict__ bstar,
                 long long g_numel, int c2, float* __restrict__ best_g, float* __restrict__ best_b,
                 const uint4* __restrict__ aux_src, uint4* __restrict__ aux_dst, long long aux_vec,
                 const float* __restrict__ pc_src = nullptr, float* __restrict__ pc_dst = nullptr) {
  if (*take == 0) return;
  if (pc_src && pc_dst)
    for (int r = blockIdx.x * AD_THREADS + threadIdx.x; r < 2 * c2; r += gridDim.x * AD_THREADS) pc_dst[r] = pc_src[r];
  for (long long e = (long long)blockIdx.x * AD_THREADS + threadIdx.x; e < aux_vec;
       e += (long long)gridDim.x * AD_THREADS) aux_dst[e] = aux_src[e];
  for (long long e = (long long)blockIdx.x * AD_THREADS + threadIdx.x; e < g_numel;
       e += (long long)gridDim.x * AD_THREADS) best_g[e] = g[e];
  if (bstar && best_b)
    for (int r = blockIdx.x * AD_THREADS + threadIdx.x; r < c2; r += gridDim.x * AD_THREADS)
      best_b[r] = bstar[r];
}

// [C2][C1][taps] fp32 integer codes -> bf16 / e4m3 codes in the tensor-core weight layout.
__global__ void __launch_bounds__(AD_THREADS)
pack_wcodes_kernel(const float* __restrict__ codes, int c2, int c1, int taps, TcLayout lay,
                   void* __restrict__ out) {
  const long long total = (long long)c2 * c1 * taps;
  for (long long e = (long long)blockIdx.x * AD_THREADS + threadIdx.x; e < total;
       e += (long long)gridDim.x * AD_THREADS) {
    const int t = (int)(e % taps);
    const int c = (int)((e / taps) % c1);
    const int r = (int)(e / ((long long)taps * c1));
    store_code(out, tc_wcode_index(r, c, t, c1, c2, lay), codes[e], lay.eb);
  }
}

}  // namespace effq

extern "C" int effq_pack_wcodes(const float* codes, int32_t c2, int32_t c1, int32_t taps, int32_t code_dtype,
                                void* out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(codes && out && c2 > 0 && c1 > 0 && taps > 0, "bad argument");
  EFFQ_CHECK_ARG(code_dtype == CODE_BF16 ? c1 % 8 == 0 : (code_dtype == CODE_E4M3 && c1 % 16 == 0),
                 "c1 must be a multiple of 8 (bf16) / 16 (e4m3)");
  pack_wcodes_kernel<<<grid_for((long long)c2 * c1 * taps), AD_THREADS, 0, (cudaStream_t)stream>>>(
      codes, c2, c1, taps, tc_layout(c1, code_dtype), out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_admm_rhs(const float* b0, const float* w0p, const float* g, const float* dual,
                             float rho, float eta, int32_t c2, int32_t k, int32_t has_bias, float* b_out,
                             void* planes_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(b0 && w0p && g && dual && (b_out || planes_out), "null pointer");
  EFFQ_CHECK_ARG(c2 > 0 && k > 0, "bad shape");
  EFFQ_CHECK_ARG(((uintptr_t)planes_out & 15) == 0, "planes must be 16B aligned");
  const int kp = k + (has_bias ? 1 : 0);
  const int ldk = (int)effq_split3_ld(kp);
  admm_rhs_kernel<<<grid_for((long long)c2 * (planes_out ? ldk : kp)), AD_THREADS, 0, (cudaStream_t)stream>>>(
      b0, w0p, g, dual, rho, eta, c2, k, kp, b_out, (__nv_bfloat16*)planes_out, ldk);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_admm_lhs(const float* a0, float rho, float eta, int32_t kp, int32_t has_bias,
                             float* a_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(a0 && a_out && kp > 0, "bad argument");
  admm_lhs_kernel<<<grid_for((long long)kp * kp), AD_THREADS, 0, (cudaStream_t)stream>>>(
      a0, rho, eta, kp, has_bias, a_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_admm_project(const float* wstar, int64_t ldw, float* dual, const effq_scale_state* wscale,
                                 const effq_scale_state* xscale, int32_t nlvl_w, int32_t nlvl_a, int32_t c2,
                                 int32_t c1, int32_t taps, int32_t has_bias, float dual_div, float* g_out,
                                 float* bstar_out, void* wcodes_out, int32_t code_dtype, effq_admm_state* st,
                                 const effq_next_rhs* next, const effq_admm_keep_bufs* keep, float* per_channel_out,
                                 void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(!keep || (keep->best_g && st), "keep: best_g and the ADMM state are required");
  effq_admm_keep_bufs kp_;
  if (keep) kp_ = *keep; else { kp_.best_g = nullptr; kp_.best_b = nullptr; kp_.best_wcodes = nullptr; kp_.best_pc = nullptr; }
  EFFQ_CHECK_ARG(wstar && dual && wscale && g_out, "null pointer");
  EFFQ_CHECK_ARG(c2 > 0 && c1 > 0 && taps > 0 && ldw >= (int64_t)c1 * taps + (has_bias ? 1 : 0), "bad shape");
  EFFQ_CHECK_ARG(!wcodes_out || (code_dtype == CODE_BF16 && c1 % 8 == 0 && nlvl_w <= 256) ||
                     (code_dtype == CODE_E4M3 && c1 % 16 == 0 && nlvl_w <= 16),
                 "weight codes need c1 % 8 == 0, nlvl_w <= 256 (bf16) or c1 % 16 == 0, nlvl_w <= 16 (e4m3)");
  EFFQ_CHECK_ARG(dual_div != 0.f, "dual_div must be non-zero");
  EFFQ_CHECK_ARG(!next || (next->b0 && next->w0p && next->planes && ((uintptr_t)next->planes & 15) == 0),
                 "next right-hand side: null / misaligned pointer");
  effq_next_rhs nx;
  if (next) nx = *next; else { nx.b0 = nullptr; nx.w0p = nullptr; nx.rho = 0.f; nx.eta = 0.f; nx.planes = nullptr; }
  const int ldk = (int)effq_split3_ld((int64_t)c1 * taps + (has_bias ? 1 : 0));
  admm_project_kernel<<<grid_for((long long)c2 * c1 * taps), AD_THREADS, 0, (cudaStream_t)stream>>>(
      wstar, ldw, dual, wscale, xscale, nlvl_w, nlvl_a, c2, c1, taps, has_bias, dual_div, g_out, bstar_out,
      wcodes_out, tc_layout(c1, wcodes_out ? code_dtype : 0), st, nx, ldk, kp_, per_channel_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

// The two halves of effq_admm_track as launches of their own: the calibration loop scores an iterate with
// effq_admm_decide (or inside effq_quadform_delta) and lets the NEXT effq_admm_project save the best iterate
// (`keep`); effq_admm_keep flushes the last one after the loop.
extern "C" int effq_admm_decide(effq_admm_state* st, const double* sse, double numel, float* history,
                                const effq_peer_comm* comm, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(st && sse && numel > 0, "bad argument");
  EFFQ_CHECK_ARG(!comm || (comm->world >= 1 && comm->world <= EFFQ_PEER_MAX && comm->rank >= 0 &&
                           comm->rank < comm->world), "bad communicator");
  effq_peer_comm cm;
  if (comm) cm = *comm; else { cm.world = 1; cm.rank = 0; for (int i = 0; i < EFFQ_PEER_MAX; ++i) cm.slots[i] = nullptr; }
  admm_decide_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(st, sse, numel, history, &st->take_, cm);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_admm_keep(effq_admm_state* st, const float* g, const float* bstar, int64_t g_numel, int32_t c2,
                              float* best_g, float* best_b, const void* aux_src, void* aux_dst, int64_t aux_bytes,
                              const float* pc_src, float* pc_dst, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(st && g && best_g && g_numel > 0, "bad argument");
  EFFQ_CHECK_ARG(aux_bytes == 0 || (aux_src && aux_dst && aux_bytes % 16 == 0 &&
                                    ((uintptr_t)aux_src & 15) == 0 && ((uintptr_t)aux_dst & 15) == 0),
                 "aux buffers must be 16B aligned and sized");
  admm_keep_kernel<<<grid_for(g_numel), AD_THREADS, 0, (cudaStream_t)stream>>>(
      &st->take_, g, bstar, g_numel, c2, best_g, best_b, (const uint4*)aux_src, (uint4*)aux_dst, aux_bytes / 16, pc_src,
      pc_dst);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_admm_track(effq_admm_state* st, const double* sse, double numel, const float* g,
                               const float* bstar, int64_t g_numel, int32_t c2, float* best_g, float* best_b,
                               float* history, const void* aux_src, void* aux_dst, int64_t aux_bytes,
                               const effq_peer_comm* comm, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(st && sse && g && best_g && numel > 0, "bad argument");
  EFFQ_CHECK_ARG(!comm || (comm->world >= 1 && comm->world <= EFFQ_PEER_MAX && comm->rank >= 0 &&
                           comm->rank < comm->world), "bad communicator");
  effq_peer_comm cm;
  if (comm) cm = *comm; else { cm.world = 1; cm.rank = 0; for (int i = 0; i < EFFQ_PEER_MAX; ++i) cm.slots[i] = nullptr; }
  EFFQ_CHECK_ARG(aux_bytes == 0 || (aux_src && aux_dst && aux_bytes % 16 == 0 &&
                                    ((uintptr_t)aux_src & 15) == 0 && ((uintptr_t)aux_dst & 15) == 0),
                 "aux buffers must be 16B aligned and sized");
  cudaStream_t s = (cudaStream_t)stream;
  int* take = &st->take_;
  admm_decide_kernel<<<1, 1, 0, s>>>(st, sse, numel, history, take, cm);
  EFFQ_LAUNCH_CHECK();
  admm_keep_kernel<<<grid_for(g_numel), AD_THREADS, 0, s>>>(take, g, bstar, g_numel, c2, best_g, best_b,
                                                                           (const uint4*)aux_src, (uint4*)aux_dst, aux_bytes / 16);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
