// Shared helpers for the effq_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/effq_b200.h"

namespace effq {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int  sm_count();
// Grid cap of a streaming (HBM-bound, read-once / write-once) kernel WITHOUT per-CTA set-up.  One work chunk per CTA
// and the hardware block scheduler beat a persistent grid of `per_sm` CTAs per SM with a grid stride on B200: glue ReLU
// 0.87 -> 1.04, add 0.97 -> 1.06, max-pool 0.67 -> 0.96 of the copy peak (profiles/r02_hbm_bench.md), so the cap is only
// the grid-dimension limit; EFFQ_STREAM_PERSIST=1 restores the persistent grids (A/B).  Kernels that prepare something
// per CTA (fp64 scale constants, level tables) measured the other way round and keep their persistent grids.
long long stream_cap(int per_sm);

#define EFFQ_CHECK_ARG(cond, msg)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      effq::set_error("%s: %s", __func__, msg);                     \
      return 1;                                                     \
    }                                                               \
  } while (0)

#define EFFQ_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      effq::set_error("%s: %s -> %s", __func__, #call, cudaGetErrorString(e_)); \
      return 2;                                                                \
    }                                                                          \
  } while (0)

#define EFFQ_LAUNCH_CHECK()                                                     \
  do {                                                                          \
    cudaError_t e_ = cudaGetLastError();                                        \
    if (e_ != cudaSuccess) {                                                    \
      effq::set_error("%s: launch failed -> %s", __func__, cudaGetErrorString(e_)); \
      return 2;                                                                 \
    }                                                                           \
    effq::count_launch();                                                       \
  } while (0)

struct OutDims {
  int od, oh, ow;
  long long vox_per_sample;   // od*oh*ow
  long long vox;              // n*od*oh*ow
};

__host__ __device__ inline OutDims out_dims(const effq_geom& g) {
  OutDims o;
  o.od = (g.d + 2 * g.pd - g.kd) / g.sd + 1;
  o.oh = (g.h + 2 * g.ph - g.kh) / g.sh + 1;
  o.ow = (g.w + 2 * g.pw - g.kw) / g.sw + 1;
  o.vox_per_sample = (long long)o.od * o.oh * o.ow;
  o.vox = o.vox_per_sample * g.n;
  return o;
}

// ---- reductions -------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of a double; result valid in thread 0. `scratch` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();                 // scratch may be reused by back-to-back calls
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  double r = 0.0;
  if (wid == 0) {
    r = lane < nw ? scratch[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// Block-wide sums of NV doubles in ONE round (one pair of barriers, the shuffle chains of the NV values interleaved):
// the same reduction tree per value as block_sum, hence bit-identical results; results valid in thread 0.
// `scratch` >= 32 * NV doubles.  A fixed-point pass of the scale search needs two sums (sum b*v, sum b*b) and a fold
// needs two or four; back-to-back block_sum calls cost ~600 cycles each (two barriers + two dependent shuffle trees).
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  __syncthreads();                 // scratch may be reused by back-to-back calls
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[32 * k + wid] = v[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = lane < nw ? scratch[32 * k + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < NV; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    }
  } else {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = 0.0;
  }
}
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
  double v[2] = {a, b};
  block_sum_n<2>(v, scratch);
  a = v[0];
  b = v[1];
}

// ---- the reference's discretize, op for op (layer_helper.py:25-37) -----------
// fp32: every operation individually rounded (no FMA contraction), IEEE division,
// clamp that propagates NaN like torch.clamp, rintf == round-half-to-even.
struct QParamF {
  float lo, hi, delta;
};
__host__ __device__ inline QParamF make_qparam_f(float lo, float hi, int nlvl) {
  QParamF q;
  q.lo = lo;
  q.hi = hi;
  q.delta = (float)(((double)hi - (double)lo) / (double)(nlvl - 1));
  return q;
}
__device__ __forceinline__ float clampf_nan(float t, float lo, float hi) {
  return t < lo ? lo : (t > hi ? hi : t);
}
// level index of t (already divided by the scale)
__device__ __forceinline__ float level_index_f(float t, const QParamF& q) {
  t = clampf_nan(t, q.lo, q.hi);
  return rintf(__fdiv_rn(__fsub_rn(t, q.lo), q.delta));
}
__device__ __forceinline__ float level_value_f(float idx, const QParamF& q) {
  return __fadd_rn(__fmul_rn(idx, q.delta), q.lo);
}

// Fast path with exact fallback.  q ~ (x/alpha - lo)/delta computed with two multiplies (a few
// ulp off); if q is safely away from a rounding tie the rounded index provably equals the exact
// op-for-op result, otherwise (about 0.2 % of elements) the exact division sequence decides.
// Bit-exactness is therefore unchanged; the common path costs ~1/3 of the instructions, which is
// what keeps the fake-quant kernels on the HBM roofline instead of the ALU one.
struct QFastF {
  float inv_alpha, inv_delta, lm1;
};
__host__ __device__ inline QFastF make_qfast_f(float alpha, const QParamF& q, int nlvl) {
  QFastF f;
  f.inv_alpha = 1.0f / alpha;
  f.inv_delta = 1.0f / q.delta;
  f.lm1 = (float)(nlvl - 1);
  return f;
}
__device__ __forceinline__ float level_index_fast_f(float x, float alpha, const QParamF& q, const QFastF& f) {
  const float qa = (x * f.inv_alpha - q.lo) * f.inv_delta;           // approximate, unclamped
  const float r = rintf(qa);
  // |qa - r| close to 0.5 inside the level range -> possible tie: take the exact path.
  // margin 2e-3 >> accumulated error (<= ~8 ulp of values <= 256 -> 1.3e-4)
  const bool risky = fabsf(fabsf(qa - r) - 0.5f) < 2e-3f && qa > -1.0f && qa < f.lm1 + 1.0f;
  if (risky || !(x == x)) return level_index_f(__fdiv_rn(x, alpha), q);
  return fminf(fmaxf(r, 0.f), f.lm1);
}

struct QParamD {
  double lo, hi, delta;
};
__host__ __device__ inline QParamD make_qparam_d(float lo, float hi, int nlvl) {
  QParamD q;
  q.lo = lo;
  q.hi = hi;
  q.delta = ((double)hi - (double)lo) / (double)(nlvl - 1);
  return q;
}
__device__ __forceinline__ double level_index_d(double t, const QParamD& q) {
  t = t < q.lo ? q.lo : (t > q.hi ? q.hi : t);
  return rint(__ddiv_rn(__dsub_rn(t, q.lo), q.delta));
}
__device__ __forceinline__ double level_value_d(double idx, const QParamD& q) {
  return __dadd_rn(__dmul_rn(idx, q.delta), q.lo);
}
struct QFastD {
  double c1, c0, lm1;      // q ~ v*c1 + c0 with c1 = 1/(a*delta), c0 = -lo/delta
};
__host__ __device__ inline QFastD make_qfast_d(double a, const QParamD& q, int nlvl) {
  QFastD f;
  f.c1 = 1.0 / (a * q.delta);
  f.c0 = -q.lo / q.delta;
  f.lm1 = (double)(nlvl - 1);
  return f;
}
// exact fp64 level index of v/a with the division-free fast path (fallback near ties)
__device__ __forceinline__ double level_index_fast_d(double v, double a, const QParamD& q, const QFastD& f) {
  const double qa = fma(v, f.c1, f.c0);
  const double r = rint(qa);
  const bool risky = fabs(fabs(qa - r) - 0.5) < 1e-9 && qa > -1.0 && qa < f.lm1 + 1.0;
  if (risky || !(v == v)) return level_index_d(__ddiv_rn(v, a), q);
  return fmin(fmax(r, 0.0), f.lm1);
}

}  // namespace effq
