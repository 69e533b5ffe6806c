// Conv-free scoring of an ADMM iterate from fp64 sufficient statistics.
//
// The reference scores every one of the 200 iterates of a layer with a full conv + MSE
// (src/models/EfficientQConv.py:118-122).  The conv input X^ (the quantised activations) is the same
// tensor in all of them, so with the UNWEIGHTED statistics  S = X^ X^T (K' x K', ones row for the bias)
// and, for a reference iterate (G_ref, b_ref) with residual R = Y - (G_ref X^ + b_ref),  T = R X^T:
//
//   sum (W^ X^ - Y)^2 = sum_r [ u_r S u_r^T - 2 u_r . T_r ] + sum R^2 ,   u = [G - G_ref | b* - b_ref]
//
// Every term is of the size of the loss itself (no cancellation against sum Y^2), S is exact (integer Gram
// of the codes x scale^2 in fp64), and one launch of C2 K'^2 fp64 FMAs replaces a 2 V C2 K flop conv that
// re-reads the V x C2 fp32 target: 0.40 ms -> a few microseconds on the 32-channel level of the BraTS net.
// With G_ref = NULL it is the plain form  w S w^T - 2 w.T + sum y^2  used for the two layers whose input
// is not quantised (conv0, final_cls).
//
// The kernel also does the best-iterate bookkeeping of the iterate it scored (admm_decide_dev), so
// the calibration loop needs no separate "track" launch.
#include "common.cuh"
#include "admm_decide.cuh"

namespace effq {

constexpr int QF_THREADS = 256;
constexpr int QF_TI = 64;        // rows of S (index i) per CTA
constexpr int QF_TR = 32;        // output channels per CTA
constexpr int QF_TJ = 16;        // j step staged in shared memory
constexpr int QF_MAX_CTAS = 4096;

struct QfArgs {
  const double* acc;     // [(kp + c2) x kp]: S then T
  const double* yy;      // device scalar: sum y^2 (plain form) or sum R^2 (delta form)
  const float* g;        // [c2][k]
  const float* bstar;    // [c2] or NULL
  const float* g_ref;    // [c2][k] or NULL
  const float* b_ref;    // [c2] or NULL
  int c2, k, kp;
  int n_it, n_rt, n_js, jchunk;
  double* partial;       // [n_ctas][QF_TR]
  unsigned int* done;
  double* sse;
  effq_admm_state* st;   // optional: decide in the tail
  double numel;
  float* history;
};

__device__ __forceinline__ double qf_u(const QfArgs& a, int r, int j) {
  if (r >= a.c2 || j >= a.kp) return 0.0;
  if (j < a.k) {
    const long long e = (long long)r * a.k + j;
    const double v = (double)a.g[e];
    return a.g_ref ? v - (double)a.g_ref[e] : v;
  }
  const double v = (double)a.bstar[r];
  return a.b_ref ? v - (double)a.b_ref[r] : v;
}

// CTA (it, rt, js): t[i][r] = sum_{j in chunk js} S'[i][j] u[r][j] for a 64 x 32 tile, then
// q_r += sum_i u[r][i] t[i][r]; the CTAs of row tile 0 also add -2 sum_{j in chunk} u[r][j] T[r][j].
// S is symmetric, so only its upper triangle is read: S'[i][j] = 2 S[i][j] (j > i), S[i][i], 0 (j < i) -- half the
// fp64 FMAs and half the L2 traffic; j steps entirely left of the row tile are not visited at all.
__global__ void __launch_bounds__(QF_THREADS)
quadform_delta_kernel(QfArgs a) {
  __shared__ double Ss[QF_TJ][QF_TI + 1];
  __shared__ double Us[QF_TJ][QF_TR + 1];
  __shared__ double red[QF_THREADS / 32][QF_TR];
  __shared__ bool last;
  const int cta = blockIdx.x;
  const int js = cta % a.n_js, rt = (cta / a.n_js) % a.n_rt, it = cta / (a.n_js * a.n_rt);
  const int i0 = it * QF_TI, r0 = rt * QF_TR;
  const int j_begin = js * a.jchunk, j_end = min(a.kp, j_begin + a.jchunk);
  const int t = threadIdx.x;
  const int ti = t % 16, tr = t / 16;            // thread tile: rows i0 + ti + 16 q (q < 4) x channels r0 + 2 tr, +1
  double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
  double lin = 0.0;                              // u.T share of channel r0 + t (threads t < 32 of row tile 0)
  // register prefetch of the next j step (4 S values + 2 u values per thread) hides the L2 latency behind the
  // 128 FMAs of the current one
  double s_pre[4], u_pre[2];
  auto fetch = [&](int j0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = t + q * QF_THREADS;
      const int ii = e / QF_TJ, jj = e % QF_TJ;
      const int i = i0 + ii, j = j0 + jj;
      double sv = (i < a.kp && j < j_end && j >= i) ? __ldg(a.acc + (long long)i * a.kp + j) : 0.0;
      if (j > i) sv += sv;
      s_pre[q] = sv;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = t + q * QF_THREADS;
      const int rr = e / QF_TJ, jj = e % QF_TJ;
      const int j = j0 + jj;
      u_pre[q] = j < j_end ? qf_u(a, r0 + rr, j) : 0.0;
    }
  };
  const int j_first = max(j_begin, i0);          // i0 and j_begin are multiples of QF_TJ: whole steps below the diagonal skipped
  if (j_first < j_end) fetch(j_first);
  for (int j0 = j_first; j0 < j_end; j0 += QF_TJ) {
    // stage S[i0..i0+63][j0..j0+15] (j contiguous in memory) and u[r0..r0+31][j0..j0+15]
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = t + q * QF_THREADS;
      Ss[e % QF_TJ][e / QF_TJ] = s_pre[q];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = t + q * QF_THREADS;
      Us[e % QF_TJ][e / QF_TJ] = u_pre[q];
    }
    __syncthreads();
    if (j0 + QF_TJ < j_end) fetch(j0 + QF_TJ);
#pragma unroll
    for (int jj = 0; jj < QF_TJ; ++jj) {
      const double u0 = Us[jj][tr * 2], u1 = Us[jj][tr * 2 + 1];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double sv = Ss[jj][ti + 16 * q];          // consecutive lanes -> consecutive doubles: conflict-free
        acc[q][0] = fma(sv, u0, acc[q][0]);
        acc[q][1] = fma(sv, u1, acc[q][1]);
      }
    }
    if (it == 0 && t < QF_TR) {                  // linear term of channel r0 + t over this j step
      const int r = r0 + t;
      if (r < a.c2) {
        const double* trow = a.acc + (long long)a.kp * a.kp + (long long)r * a.kp;
#pragma unroll 4
        for (int jj = 0; jj < QF_TJ; ++jj) {
          const int j = j0 + jj;
          if (j < j_end) lin = fma(Us[jj][t], __ldg(trow + j), lin);
        }
      }
    }
    __syncthreads();
  }
  // q_r share of this thread: sum over its 4 rows of u[r][i] * t[i][r]
  double q0 = 0.0, q1 = 0.0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + ti + 16 * q;
    q0 = fma(qf_u(a, r0 + tr * 2, i), acc[q][0], q0);
    q1 = fma(qf_u(a, r0 + tr * 2 + 1, i), acc[q][1], q1);
  }
  // reduce over the 16 threads (ti) that share a channel pair: they are the 16 lanes of a half warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    q0 += __shfl_xor_sync(0xffffffffu, q0, o);
    q1 += __shfl_xor_sync(0xffffffffu, q1, o);
  }
  // thread with ti == 0 holds channels tr*2, tr*2+1 (tr = 0..15 covers all 32)
  double* mine = a.partial + (long long)cta * QF_TR;
  if (ti == 0) { mine[tr * 2] = q0; mine[tr * 2 + 1] = q1; }
  __syncthreads();
  if (it == 0 && t < QF_TR) mine[t] = mine[t] - 2.0 * lin;      // same thread block wrote mine[]: ordered by the barrier
  __threadfence();
  __syncthreads();
  if (t == 0) last = (atomicAdd(a.done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // fixed-order fold of all partials: thread (w = t / 32, c = t % 32) sums the CTAs congruent to w mod 8
  // for channel slot c, then slots are added in a fixed tree -> bit-identical from run to run
  {
    const int c = t % 32, w = t / 32;
    double s = 0.0;
    for (int b = w; b < (int)gridDim.x; b += QF_THREADS / 32) s += ((volatile double*)a.partial)[(long long)b * QF_TR + c];
    red[w][c] = s;
  }
  __syncthreads();
  if (t < 32) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < QF_THREADS / 32; ++w) s += red[w][t];
    s = warp_sum(s);          // xor-butterfly: the same association in every run
    if (t == 0) {
      double total = s + *a.yy;
      if (!(total > 0.0)) total = total == total ? 0.0 : total;      // clamp tiny negatives, keep NaN
      *a.sse = total;
      *a.done = 0;
      if (a.st) admm_decide_dev(a.st, total, a.numel, a.history);
    }
  }
}

// fp64 copy of the tcgen05 Gram accumulator in real units and full (mirrored) form:
//   out[i][j] = s_i s_j acc[i][j]   (s = code_scale on the K code rows / columns, 1 on the bias and Y rows)
__global__ void gram_finalize_f64_kernel(const double* __restrict__ acc64, const float* __restrict__ x_scale,
                                         int k, int kp, int c2, int c1, double* __restrict__ out) {
  const double s = x_scale ? (double)__ldg(x_scale) : 1.0;
  const long long total = (long long)(kp + c2) * kp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / kp), j = (int)(e % kp);
    if (i < k && j < k) {
      // tiles below the diagonal (tap-major order, 128 x 256 tiles) were skipped by the tcgen05 kernel: mirror
      const int ti = (i % 27) * c1 + i / 27, tj = (j % 27) * c1 + j / 27;
      const bool skipped = (ti / 128) * 128 >= (tj / 256 + 1) * 256;
      out[e] = s * s * acc64[skipped ? (long long)j * kp + i : e];
      continue;
    }
    const double sj = j < k ? s : 1.0;
    const double si = i < k ? s : 1.0;
    out[e] = si * sj * acc64[e];
  }
}

}  // namespace effq

extern "C" int64_t effq_quadform_delta_workspace(int32_t c2, int32_t kp) {
  (void)c2; (void)kp;
  return 16 + (int64_t)effq::QF_MAX_CTAS * effq::QF_TR * 8;
}

extern "C" int effq_quadform_delta(const double* acc64, const double* sum_sq, const float* gw, const float* bstar,
                                   const float* g_ref, const float* b_ref, int32_t c2, int32_t k, int32_t has_bias,
                                   double* sse, void* workspace, effq_admm_state* st, double numel, float* history,
                                   void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(acc64 && sum_sq && gw && sse && workspace && c2 > 0 && k > 0, "bad argument");
  EFFQ_CHECK_ARG(!has_bias || bstar, "bias missing");
  EFFQ_CHECK_ARG(!g_ref || !has_bias || b_ref, "reference bias missing");
  EFFQ_CHECK_ARG(!st || numel > 0, "numel must be positive when the decision is fused");
  QfArgs a;
  a.acc = acc64; a.yy = sum_sq; a.g = gw; a.bstar = bstar; a.g_ref = g_ref; a.b_ref = g_ref ? b_ref : nullptr;
  a.c2 = c2; a.k = k; a.kp = k + (has_bias ? 1 : 0);
  a.n_it = (a.kp + QF_TI - 1) / QF_TI;
  a.n_rt = (c2 + QF_TR - 1) / QF_TR;
  // split the j range so that every SM holds ~8 CTAs (latency hiding: a CTA is a chain of dependent L2 loads),
  // in whole QF_TJ steps
  const int base = a.n_it * a.n_rt;
  int n_js = (8 * sm_count() + base - 1) / base;
  const int max_js = (a.kp + QF_TJ - 1) / QF_TJ;
  if (n_js > max_js) n_js = max_js;
  if (n_js < 1) n_js = 1;
  a.jchunk = ((a.kp + n_js - 1) / n_js + QF_TJ - 1) / QF_TJ * QF_TJ;
  a.n_js = (a.kp + a.jchunk - 1) / a.jchunk;
  const int ctas = a.n_it * a.n_rt * a.n_js;
  EFFQ_CHECK_ARG(ctas <= QF_MAX_CTAS, "system too large for the quadratic-form kernel");
  a.done = (unsigned int*)workspace;
  a.partial = (double*)((char*)workspace + 16);
  a.sse = sse; a.st = st; a.numel = numel; a.history = history;
  quadform_delta_kernel<<<ctas, QF_THREADS, 0, (cudaStream_t)stream>>>(a);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_gram_tc_f64(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y,
                                const float* att, const effq_geom* g, int32_t has_bias, int32_t att_exact,
                                double* acc64_out, void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(xcodes_ndhwc_bf16 && code_scale && y && g && acc64_out && workspace, "null pointer");
  EFFQ_CHECK_ARG(effq_gram_tc_supported(g), "geometry not supported by the tcgen05 Gram kernel");
  const int k = g->c1 * 27;
  const int kp = k + (has_bias ? 1 : 0);
  const int mrows = kp + g->c2;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t acc_bytes = (size_t)mrows * kp * 8;
  EFFQ_CUDA(cudaMemsetAsync(workspace, 0, acc_bytes + 16, s));
  if (int rc = effq_gram_tc_accumulate(xcodes_ndhwc_bf16, att, y, g, has_bias, att_exact, (double*)workspace, kp,
                                       (char*)workspace + acc_bytes, stream)) return rc;
  const long long total = (long long)mrows * kp;
  int fb = (int)((total + 255) / 256);
  if (fb > sm_count() * 16) fb = sm_count() * 16;
  gram_finalize_f64_kernel<<<fb, 256, 0, s>>>((const double*)workspace, code_scale, k, kp, g->c2, g->c1, acc64_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
