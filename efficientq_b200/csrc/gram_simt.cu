// Normal-equation statistics without ever materialising the im2col matrix.
// Replaces im2col_loop + QuadraSolver.getA0B0 (reference src/models/solver.py:86-111,
// :253-257, :282-314):
//
//     A0 = 2 * sum_v att_v * xhat_v xhat_v^T        (K' x K')
//     B0 = 2 * sum_v att_v * y_v    xhat_v^T        (C2 x K')
//
// xhat_v = the (c,kd,kh,kw)-ordered input patch of output voxel v, plus a trailing 1
// when the layer has a bias (solver.py:255-256).  B0 is exactly conv-wgrad with
// grad_out = att*Y; A0 is the im2col Gram matrix.
//
// Generic fp32 implementation (any geometry): one GEMM  [Xhat ; Y] (att.Xhat)^T with
// the reduction over voxels, 64x64 output tiles, 4x4 register micro-tiles, patches
// gathered on the fly, split over voxel ranges across CTAs.  fp32 FMA over short
// runs (512 voxels) that are folded into fp64 accumulators, fp64 atomics into the
// workspace, one finalize pass that applies 2 * x_scale^{1,2} and writes fp32.
// When x holds integer codes the short fp32 runs are exact.
#include "common.cuh"

namespace effq {

constexpr int GM_TILE = 64;
constexpr int GM_KC = 16;               // voxels per smem step
constexpr int GM_THREADS = 256;
constexpr int GM_FLUSH = 32;            // steps between fp32 -> fp64 folds (512 voxels)

struct RowDesc {                        // how to fetch one operand row
  int kind;                             // 0: patch row, 1: ones, 2: y row, 3: padding (zero)
  int c;                                // input channel or y channel
  int a, b, d;                          // tap offsets (kd,kh,kw)
};

__device__ __forceinline__ RowDesc decode_row(int r, int k, int kp, int mrows, const effq_geom& g) {
  RowDesc rd;
  rd.kind = 3; rd.c = 0; rd.a = rd.b = rd.d = 0;
  if (r < k) {
    const int taps = g.kd * g.kh * g.kw;
    rd.kind = 0;
    rd.c = r / taps;
    int t = r % taps;
    rd.d = t % g.kw; t /= g.kw;
    rd.b = t % g.kh; t /= g.kh;
    rd.a = t;
  } else if (r < kp) {
    rd.kind = 1;
  } else if (r < mrows) {
    rd.kind = 2;
    rd.c = r - kp;
  }
  return rd;
}

struct VoxPos {
  int n, od, oh, ow;
  bool live;
  long long sp;                         // voxel index inside the sample
};

__device__ __forceinline__ float fetch(const RowDesc& rd, const VoxPos& p, const float* __restrict__ x,
                                       const float* __restrict__ y, const effq_geom& g, const OutDims& o) {
  if (!p.live || rd.kind == 3) return 0.f;
  if (rd.kind == 1) return 1.f;
  if (rd.kind == 2) return __ldg(y + ((long long)p.n * g.c2 + rd.c) * o.vox_per_sample + p.sp);
  const int id = p.od * g.sd - g.pd + rd.a;
  const int ih = p.oh * g.sh - g.ph + rd.b;
  const int iw = p.ow * g.sw - g.pw + rd.d;
  if ((unsigned)id >= (unsigned)g.d || (unsigned)ih >= (unsigned)g.h || (unsigned)iw >= (unsigned)g.w) return 0.f;
  return __ldg(x + ((((long long)p.n * g.c1 + rd.c) * g.d + id) * g.h + ih) * g.w + iw);
}

__global__ void __launch_bounds__(GM_THREADS)
gram_f32_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ att,
                effq_geom g, OutDims o, int k, int kp, int mrows, int row_begin, long long vox_per_split,
                int flush_every, double* __restrict__ acc64) {
  __shared__ float Ls[GM_KC][GM_TILE + 4];
  __shared__ float Rs[GM_KC][GM_TILE + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int i0 = row_begin + blockIdx.y * GM_TILE, j0 = blockIdx.x * GM_TILE;
  const long long v_begin = (long long)blockIdx.z * vox_per_split;
  const long long v_end = min(o.vox, v_begin + vox_per_split);

  // loader role: this thread fetches voxel column `lk` of rows lr, lr+16, lr+32, lr+48
  const int lk = threadIdx.x % GM_KC, lr = threadIdx.x / GM_KC;
  RowDesc lrow[4], rrow[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    lrow[q] = decode_row(i0 + lr + 16 * q, k, kp, mrows, g);
    const int j = j0 + lr + 16 * q;
    rrow[q] = decode_row(j < kp ? j : mrows, k, kp, mrows, g);    // right operand has no y rows
  }

  float acc[4][4];
  double acc_d[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { acc[a][b] = 0.f; acc_d[a][b] = 0.0; }

  int step = 0;
  for (long long v0 = v_begin; v0 < v_end; v0 += GM_KC, ++step) {
    VoxPos p;
    const long long v = v0 + lk;
    p.live = v < v_end;
    {
      long long r = p.live ? v : 0;
      p.ow = (int)(r % o.ow); r /= o.ow;
      p.oh = (int)(r % o.oh); r /= o.oh;
      p.od = (int)(r % o.od); r /= o.od;
      p.n = (int)r;
      p.sp = (p.live ? v : 0) - (long long)p.n * o.vox_per_sample;
    }
    const float wv = (p.live && att) ? __ldg(att + v) : 1.f;
    float lv[4], rv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      lv[q] = fetch(lrow[q], p, x, y, g, o);
      rv[q] = __fmul_rn(fetch(rrow[q], p, x, y, g, o), wv);      // x_col * att, solver.py:293
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Ls[lk][lr + 16 * q] = lv[q];
      Rs[lk][lr + 16 * q] = rv[q];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GM_KC; ++kk) {
      const float4 l4 = *reinterpret_cast<const float4*>(&Ls[kk][ty * 4]);
      const float4 r4 = *reinterpret_cast<const float4*>(&Rs[kk][tx * 4]);
      const float l[4] = {l4.x, l4.y, l4.z, l4.w};
      const float r[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(l[a], r[b], acc[a][b]);
    }
    if ((step + 1) % flush_every == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc_d[a][b] += (double)acc[a][b]; acc[a][b] = 0.f; }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i >= mrows) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + tx * 4 + b;
      if (j >= kp) continue;
      const double t = acc_d[a][b] + (double)acc[a][b];
      if (t != 0.0) atomicAdd(acc64 + (long long)i * kp + j, t);
    }
  }
}

// tc_block != 0: the K x K block was accumulated on integer codes by the tcgen05 kernel and
// needs code_scale^2; the rows >= K (ones row, Y rows) were accumulated on real values; the
// bias column of the first K rows is mirrored from the ones row (A0 is symmetric).
__global__ void gram_finalize_kernel(const double* __restrict__ acc64, const float* __restrict__ x_scale,
                                     int k, int kp, int c2, int tc_block, float* __restrict__ a0,
                                     float* __restrict__ b0) {
  const double s = x_scale ? (double)__ldg(x_scale) : 1.0;
  const long long total = (long long)(kp + c2) * kp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / kp), j = (int)(e % kp);
    if (tc_block && i < k && j < k) {
      // tensor-core path: tiles below the diagonal (in the kernel's tap-major order, 128 x 256
      // tiles) were skipped -- mirror them from the transpose.  tc_block = C1.
      const int c1 = tc_block;
      const int ti = (i % 27) * c1 + i / 27, tj = (j % 27) * c1 + j / 27;
      const bool skipped = (ti / 128) * 128 >= (tj / 256 + 1) * 256;
      a0[e] = (float)(2.0 * s * s * acc64[skipped ? (long long)j * kp + i : e]);
      continue;
    }
    const double sj = j < k ? s : 1.0;
    if (i < kp) {
      const double si = i < k ? s : 1.0;
      a0[e] = (float)(2.0 * si * sj * acc64[e]);
    } else {
      b0[(long long)(i - kp) * kp + j] = (float)(2.0 * sj * acc64[e]);
    }
  }
}

}  // namespace effq

extern "C" int64_t effq_gram_workspace(const effq_geom* g, int32_t has_bias) {
  if (!g) return 0;
  const long long k = (long long)g->c1 * g->kd * g->kh * g->kw;
  const long long kp = k + (has_bias ? 1 : 0);
  return (int64_t)((kp + g->c2) * kp * 8 + 16);      // fp64 accumulator + abort flag
}

namespace effq {
// Launches the generic kernel for rows [row_begin, mrows) of [Xhat ; Y] against att.Xhat.
static int launch_gram_rows(const float* x, const float* y, const float* att, const effq_geom& g, const OutDims& o,
                            int k, int kp, int mrows, int row_begin, double* acc, cudaStream_t s,
                            int flush_every = GM_FLUSH) {
  const int tx = (kp + GM_TILE - 1) / GM_TILE, ty = (mrows - row_begin + GM_TILE - 1) / GM_TILE;
  long long splits = ((long long)sm_count() * 6 + (long long)tx * ty - 1) / ((long long)tx * ty);
  const long long max_splits = (o.vox + 511) / 512;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long per = (o.vox + splits - 1) / splits;
  per = (per + GM_KC - 1) / GM_KC * GM_KC;
  splits = (o.vox + per - 1) / per;
  dim3 grid(tx, ty, (unsigned)splits);
  gram_f32_kernel<<<grid, GM_THREADS, 0, s>>>(x, y, att, g, o, k, kp, mrows, row_begin, per, flush_every, acc);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
}  // namespace effq

extern "C" int effq_gram_f32(const float* x, const float* x_scale, const float* y, const float* att,
                             const effq_geom* g, int32_t has_bias, float* a0_out, float* b0_out,
                             void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && y && g && a0_out && b0_out && workspace, "null pointer");
  const OutDims o = out_dims(*g);
  EFFQ_CHECK_ARG(o.od > 0 && o.oh > 0 && o.ow > 0 && g->n > 0, "empty output");
  const int k = g->c1 * g->kd * g->kh * g->kw;
  const int kp = k + (has_bias ? 1 : 0);
  const int mrows = kp + g->c2;
  cudaStream_t s = (cudaStream_t)stream;
  EFFQ_CUDA(cudaMemsetAsync(workspace, 0, (size_t)mrows * kp * 8, s));
  if (int rc = launch_gram_rows(x, y, att, *g, o, k, kp, mrows, 0, (double*)workspace, s)) return rc;
  const long long total = (long long)mrows * kp;
  int fb = (int)((total + 255) / 256);
  if (fb > sm_count() * 16) fb = sm_count() * 16;
  gram_finalize_kernel<<<fb, 256, 0, s>>>((const double*)workspace, x_scale, k, kp, g->c2, 0, a0_out, b0_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}


// Tensor-core path: A0 (incl. bias row / column) and B0 in one tcgen05 kernel on the integer
// codes; code_scale (device fp32) turns codes into activations in the finalize pass.
extern "C" int effq_gram_tc(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y,
                            const float* att, const effq_geom* g, int32_t has_bias, int32_t att_exact,
                            float* a0_out, float* b0_out, void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(xcodes_ndhwc_bf16 && code_scale && y && g && a0_out && b0_out && workspace, "null pointer");
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
  gram_finalize_kernel<<<fb, 256, 0, s>>>((const double*)workspace, code_scale, k, kp, g->c2, g->c1, a0_out, b0_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

// ---- one pass for both the attention-weighted normal equations and the UNWEIGHTED Gram --------------------
namespace effq {
// rows [row_begin, row_end) of the tcgen05 accumulator as fp64 in real units, mirrored (see gram_finalize_kernel)
__global__ void gram_finalize_rows_f64_kernel(const double* __restrict__ acc64, const float* __restrict__ x_scale,
                                              int k, int kp, int c1, int row_begin, int row_end,
                                              double* __restrict__ out) {
  const double s = x_scale ? (double)__ldg(x_scale) : 1.0;
  const long long total = (long long)(row_end - row_begin) * kp;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int i = row_begin + (int)(e / kp), j = (int)(e % kp);
    const long long o = (long long)i * kp + j;
    if (i < k && j < k) {
      const int ti = (i % 27) * c1 + i / 27, tj = (j % 27) * c1 + j / 27;
      const bool skipped = (ti / 128) * 128 >= (tj / 256 + 1) * 256;
      out[o] = s * s * acc64[skipped ? (long long)j * kp + i : o];
      continue;
    }
    // bias row / column of the unweighted block: the column (rows < K) comes from the raw-code x ones-column
    // products, the row (i == K) from the unweighted ones row; Y rows: as accumulated
    out[o] = (i < k ? s : 1.0) * (j < k ? s : 1.0) * acc64[o];
  }
}
}  // namespace effq

// A0, B0 (fp32, as effq_gram_tc) AND the unweighted S = X^ X^T (K' x K', fp64, real units) from ONE pass over the
// codes: the conv-free scoring of the ADMM iterates needs S, and a second full Gram pass would double the cost.
// workspace: 2 * effq_gram_workspace bytes.
extern "C" int effq_gram_tc_dual(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y,
                                 const float* att, const effq_geom* g, int32_t att_exact, float* a0_out,
                                 float* b0_out, double* s_unweighted_out, void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(xcodes_ndhwc_bf16 && code_scale && y && g && a0_out && b0_out && s_unweighted_out && workspace,
                 "null pointer");
  EFFQ_CHECK_ARG(effq_gram_tc_supported(g), "geometry not supported by the tcgen05 Gram kernel");
  const int k = g->c1 * 27, kp = k + 1, mrows = kp + g->c2;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t acc_bytes = ((size_t)mrows * kp * 8 + 16 + 255) & ~(size_t)255;
  EFFQ_CUDA(cudaMemsetAsync(workspace, 0, 2 * acc_bytes, s));
  double* acc = (double*)workspace;
  double* acc2 = (double*)((char*)workspace + acc_bytes);
  if (int rc = effq_gram_tc_accumulate2(xcodes_ndhwc_bf16, att, y, g, 1, att_exact, acc, kp,
                                        (char*)workspace + (size_t)mrows * kp * 8, acc2, 0, stream)) return rc;
  const long long total = (long long)mrows * kp;
  int fb = (int)((total + 255) / 256);
  if (fb > sm_count() * 16) fb = sm_count() * 16;
  gram_finalize_kernel<<<fb, 256, 0, s>>>(acc, code_scale, k, kp, g->c2, g->c1, a0_out, b0_out);
  EFFQ_LAUNCH_CHECK();
  gram_finalize_rows_f64_kernel<<<fb, 256, 0, s>>>(acc2, code_scale, k, kp, g->c1, 0, kp, s_unweighted_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

// T = Y X^T (C2 x K', fp64, real units, UNWEIGHTED) for a new target y (the residual of a reference iterate): only
// the extra row block of the tcgen05 Gram kernel runs.  t_out: rows K'.. of the [(K'+C2) x K'] statistics matrix
// (pass its base pointer; rows [K', K'+C2) are written).  workspace: effq_gram_workspace bytes.
extern "C" int effq_gram_tc_rows_f64(const void* xcodes_ndhwc_bf16, const float* code_scale, const float* y,
                                     const effq_geom* g, double* stats_base, void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(xcodes_ndhwc_bf16 && code_scale && y && g && stats_base && workspace, "null pointer");
  EFFQ_CHECK_ARG(effq_gram_tc_supported(g), "geometry not supported by the tcgen05 Gram kernel");
  const int k = g->c1 * 27, kp = k + 1, mrows = kp + g->c2;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t acc_bytes = (size_t)mrows * kp * 8;
  EFFQ_CUDA(cudaMemsetAsync(workspace, 0, acc_bytes + 16, s));
  if (int rc = effq_gram_tc_accumulate2(xcodes_ndhwc_bf16, nullptr, y, g, 1, 1, (double*)workspace, kp,
                                        (char*)workspace + acc_bytes, nullptr, 1, stream)) return rc;
  const long long total = (long long)g->c2 * kp;
  int fb = (int)((total + 255) / 256);
  if (fb > sm_count() * 16) fb = sm_count() * 16;
  gram_finalize_rows_f64_kernel<<<fb, 256, 0, s>>>((const double*)workspace, code_scale, k, kp, g->c1, kp, mrows, stats_base);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

// ---- sufficient statistics for scoring WITHOUT re-running the conv ------------------------------
// For a layer whose input is not quantised (conv0, final_cls: q_first/q_last = 256,-1) the conv
// input is the same tensor in all 200 ADMM iterations, so
//     sum (W^ x^ - y)^2  =  sum_r [ w_r S w_r^T - 2 w_r . T_r ]  +  sum y^2
// with the UNWEIGHTED  S = X^ X^T (K' x K')  and  T = Y X^T (C2 x K')  accumulated once in fp64
// (fp32 products folded into fp64 every 16 voxels).  K' is 109 / 33 for these layers.
extern "C" int effq_gram_f64(const float* x, const float* y, const effq_geom* g, int32_t has_bias,
                             double* acc64_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && y && g && acc64_out, "null pointer");
  const OutDims o = out_dims(*g);
  EFFQ_CHECK_ARG(o.od > 0 && o.oh > 0 && o.ow > 0 && g->n > 0, "empty output");
  const int k = g->c1 * g->kd * g->kh * g->kw;
  const int kp = k + (has_bias ? 1 : 0);
  const int mrows = kp + g->c2;
  cudaStream_t s = (cudaStream_t)stream;
  EFFQ_CUDA(cudaMemsetAsync(acc64_out, 0, (size_t)mrows * kp * 8, s));
  return launch_gram_rows(x, y, nullptr, *g, o, k, kp, mrows, 0, acc64_out, s, 1);
}

namespace effq {
// one CTA per output channel r:  q_r = w_r S w_r^T - 2 w_r . T_r   (fp64), then a fixed-order sum
__global__ void __launch_bounds__(256)
quadform_kernel(const double* __restrict__ acc, double yy, const float* __restrict__ gw, const float* __restrict__ bstar,
                int c2, int k, int kp, double* __restrict__ per_row, unsigned int* __restrict__ done,
                double* __restrict__ sse) {
  __shared__ double scratch[32];
  __shared__ bool last;
  const int r = blockIdx.x;
  const double* S = acc;
  const double* T = acc + (long long)kp * kp + (long long)r * kp;
  auto wv = [&](int j) -> double { return j < k ? (double)gw[(long long)r * k + j] : (double)bstar[r]; };
  double part = 0.0;
  for (int i = threadIdx.x; i < kp; i += blockDim.x) {
    double t = 0.0;
    for (int j = 0; j < kp; ++j) t = fma(S[(long long)i * kp + j], wv(j), t);
    part += wv(i) * (t - 2.0 * T[i]);
  }
  part = block_sum(part, scratch);
  if (threadIdx.x == 0) {
    per_row[r] = part;
    __threadfence();
    last = (atomicAdd(done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = yy;
    for (int q = 0; q < c2; ++q) t += ((volatile double*)per_row)[q];
    *sse = t > 0.0 ? t : 0.0;
    *done = 0;
  }
}
}  // namespace effq

// sse = sum over all outputs of (conv(x, G) + b* - y)^2 from the statistics of effq_gram_f64.
// workspace: 16 B counter (zero on entry) + c2 doubles.
extern "C" int effq_quadform_sse(const double* acc64, double sum_y2, const float* gw, const float* bstar,
                                 int32_t c2, int32_t k, int32_t has_bias, double* sse, void* workspace,
                                 void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(acc64 && gw && sse && workspace && c2 > 0 && k > 0, "bad argument");
  EFFQ_CHECK_ARG(!has_bias || bstar, "bias missing");
  const int kp = k + (has_bias ? 1 : 0);
  quadform_kernel<<<c2, 256, 0, (cudaStream_t)stream>>>(acc64, sum_y2, gw, bstar, c2, k, kp,
                                                       (double*)((char*)workspace + 16), (unsigned int*)workspace, sse);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
