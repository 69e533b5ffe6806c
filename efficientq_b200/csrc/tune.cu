// End-to-end refinement of the activation ranges (reference src/ptqer.py:238-272,
// tune_activation_range: Adam on every alpha_act, loss = MSE of the quantised net's output).
// Two kernels of that path live here; the convolutions between them are untouched layer maths.
//
//   effq_fakequant_ste_bwd : backward of  qact = discretize(x/alpha, L, lo, hi) * alpha  under the
//       reference's straight-through estimator (layer_helper.py:13-37: round() has identity gradient,
//       torch.clamp passes the gradient where lo <= u <= hi).  With u = x/alpha, m = 1[lo <= u <= hi]:
//           d qact / d x     = m          (evaluated as (((g*alpha)*delta)/delta)/alpha, autograd's op order)
//           d qact / d alpha = D(u) - m * u            (D(u) = the level value; = hi for u > hi, lo for u < lo)
//       One pass over (x, g): writes grad_x = g*m (optional) and reduces  sum g*(D(u) - m*u)  with
//       warp shuffles, fp64 across the CTA, per-CTA partials folded in a fixed order by the last
//       CTA (deterministic, so replicated ranks stay in lock-step).  HBM-bound: 12 B per element.
//   effq_adam_step : torch.optim.Adam's update (no weight decay, no amsgrad) of the <= 64 range
//       parameters in one launch, op for op in fp32 (lerp form of the first moment).
#include "common.cuh"

namespace effq {

constexpr int STE_THREADS = 256;
constexpr int STE_MAX_BLOCKS = 2048;

template <bool WRITE_GX>
__global__ void __launch_bounds__(STE_THREADS)
fakequant_ste_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long long numel,
                         const float* __restrict__ alpha_p, QParamF q, float* __restrict__ gx,
                         double* __restrict__ grad_alpha, double* __restrict__ partials, unsigned int* __restrict__ done) {
  __shared__ double scratch[32];
  __shared__ bool is_last;
  const float alpha = __ldg(alpha_p);
  const long long nvec = numel / 4;
  const long long stride = (long long)gridDim.x * STE_THREADS;
  double acc = 0.0;
  // Instruction budget (ncu, profiles/r01_ste_ncu.md): the op-for-op form needs four IEEE divisions per
  // element (x/alpha, the level index, autograd's (..)/delta and (..)/alpha) and ran at 33 % of the HBM
  // peak with 114 instructions per element, a third of them branches around the division subroutine.
  // Here ONE predicate decides per element between
  //  * the fast path: u and the level index from multiplies (as level_index_fast_f, common.cuh), the two
  //    divisions of grad_x by correctly rounded reciprocals + Markstein's correction
  //        q = t*r;  q' = fma(fma(-q, b, t), r, q)        (= RN(t/b) for operands in the normal range),
  //  * the exact op-for-op sequence, taken near a rounding tie, near a clamp boundary, for NaN and for
  //    gradients outside the safely-normal range.  Same bits either way (tests/test_gpu_tune.py).
  const int nlvl = (int)(rintf((q.hi - q.lo) / q.delta)) + 1;
  const QFastF qf = make_qfast_f(alpha, q, nlvl);
  const float r_alpha = __frcp_rn(alpha), r_delta = __frcp_rn(q.delta);
  const float eps_hi = 1e-5f * fabsf(q.hi) + 1e-30f, eps_lo = 1e-5f * fabsf(q.lo) + 1e-30f;
  auto one = [&](float xv, float gv, float& gxo) -> double {
    float u = xv * qf.inv_alpha;
    const float qa = (u - q.lo) * qf.inv_delta;
    const float r = rintf(qa);
    const float ag = fabsf(gv);
    const bool tie = fabsf(fabsf(qa - r) - 0.5f) < 2e-3f && qa > -1.0f && qa < qf.lm1 + 1.0f;
    const bool edge = fabsf(u - q.hi) < eps_hi || (fabsf(u - q.lo) < eps_lo && xv != 0.f);
    const bool wild = !(xv == xv) || !(ag < 1e25f && (ag > 1e-20f || ag == 0.f));
    float idx, gq;
    const float t = __fmul_rn(__fmul_rn(gv, alpha), q.delta);   // autograd: (g*alpha)*delta ...
    if (tie || edge || wild) {
      u = __fdiv_rn(xv, alpha);                                  // exactly the reference's x / alpha
      idx = level_index_f(u, q);
      gq = __fdiv_rn(__fdiv_rn(t, q.delta), alpha);              // ... /delta, /alpha
    } else {
      idx = fminf(fmaxf(r, 0.f), qf.lm1);
      const float q1 = t * r_delta;
      const float d1 = fmaf(fmaf(-q1, q.delta, t), r_delta, q1);
      const float q2 = d1 * r_alpha;
      gq = fmaf(fmaf(-q2, alpha, d1), r_alpha, q2);
    }
    const bool in = (u >= q.lo) && (u <= q.hi);                  // torch.clamp backward; NaN -> false
    const float d = level_value_f(idx, q);
    gxo = in ? gq : 0.f;
    return (double)gv * (double)(in ? __fsub_rn(d, u) : d);      // d - u: within 1 ulp(u) of the exact difference
  };
  constexpr int U = 4;                                       // independent 128-bit load pairs in flight per thread
  for (long long i = (long long)blockIdx.x * STE_THREADS + threadIdx.x; i < nvec; i += stride * U) {
    float4 xv[U], gv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long j = i + k * stride;
      if (j < nvec) {
        xv[k] = __ldcs(reinterpret_cast<const float4*>(x) + j);
        gv[k] = __ldcs(reinterpret_cast<const float4*>(g) + j);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long j = i + k * stride;
      if (j >= nvec) continue;
      float4 o;
      acc += (one(xv[k].x, gv[k].x, o.x) + one(xv[k].y, gv[k].y, o.y)) + (one(xv[k].z, gv[k].z, o.z) + one(xv[k].w, gv[k].w, o.w));
      if (WRITE_GX) __stcs(reinterpret_cast<float4*>(gx) + j, o);
    }
  }
  if (blockIdx.x == 0) {
    const long long t = nvec * 4 + threadIdx.x;
    if (t < numel) {
      float o;
      acc += one(x[t], g[t], o);
      if (WRITE_GX) gx[t] = o;
    }
  }
  const double tot = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = tot;
    __threadfence();
    is_last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += STE_THREADS) s += partials[b];   // fixed assignment
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
      *grad_alpha += s;          // accumulate: several uses of one alpha in a graph add up, like autograd
      *done = 0;                 // workspace ready for the next launch on this stream
    }
  }
}

__global__ void adam_step_kernel(float* __restrict__ p, const double* __restrict__ grad, float grad_scale,
                                 float* __restrict__ m, float* __restrict__ v, int n, float lr, float beta1,
                                 float beta2, float eps, int step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = (float)(grad[i] * (double)grad_scale);
  // torch/optim/adam.py (_single_tensor_adam): exp_avg.lerp_(grad, 1 - beta1);
  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float m1 = __fadd_rn(m[i], __fmul_rn(1.f - beta1, __fsub_rn(gr, m[i])));
  const float v1 = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(1.f - beta2, __fmul_rn(gr, gr)));
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v1), bc2_sqrt), eps);
  m[i] = m1;
  v[i] = v1;
  p[i] = __fadd_rn(p[i], __fmul_rn(-step_size, __fdiv_rn(m1, denom)));
}

// NCDHW fp32 -> three NDHWC bf16 planes with hi + mid + lo == x exactly (24 significant bits): the operand
// format of the tensor-core dgrad (tune.conv_dgrad).  Same data movement as quantize_act_ndhwc_v2 (fakequant.cu):
// a thread owns 4 channels x 4 consecutive voxels, transposes in registers, writes 8-byte 4-channel packs into
// shared-memory tiles that already have the output layout (store order rotated per lane -> conflict free), and
// the tiles leave as straight 16-byte copies.  Persistent, 10 B per element instead of ~38 B for the
// permute + five elementwise passes it replaces.
constexpr int SP_THREADS = 256;

__device__ __forceinline__ void split3(float v, float& hi, float& mid, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(v));
  const float r1 = __fsub_rn(v, hi);                 // exact (Sterbenz / 16 leftover bits)
  mid = __bfloat162float(__float2bfloat16_rn(r1));
  lo = __bfloat162float(__float2bfloat16_rn(__fsub_rn(r1, mid)));
}

// Fixed-point digits (the FP conv of the calibration's first pass, ops.conv3d_fp): with the channel's power-of-two
// scale 2^e > max|x| the integer X = rint(x * 2^(23 - e)), |X| <= 2^23, is cut into three balanced base-256 digits
// X = d0 * 2^16 + d1 * 2^8 + d2 (d0 in [-128, 128], d1, d2 in [-128, 127]): small integers, exact in bf16, whose products
// and sums the tensor core accumulates WITHOUT rounding (|sum| < 2^24) -- unlike the floating-point planes above,
// whose unaligned products lose bits at every accumulation step.
__device__ __forceinline__ void dig3(float v, float scale, float& d0, float& d1, float& d2) {
  const int X = __float2int_rn(v * scale);
  const int e2 = ((X + 128) & 255) - 128;
  const int X1 = (X - e2) >> 8;
  const int e1 = ((X1 + 128) & 255) - 128;
  d0 = (float)((X1 - e1) >> 8);
  d1 = (float)e1;
  d2 = (float)e2;
}

template <bool DIGITS>
__global__ void __launch_bounds__(SP_THREADS)
split3_ndhwc_kernel(const float* __restrict__ x, int c, long long dhw, int tile_v, long long n_tiles,
                    __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_mid, __nv_bfloat16* __restrict__ o_lo,
                    const int* __restrict__ ch_exp) {
  extern __shared__ __align__(16) uint8_t sp_raw[];
  uint2* tiles[3];
  for (int p = 0; p < 3; ++p) tiles[p] = reinterpret_cast<uint2*>(sp_raw + (size_t)p * tile_v * c * 2);
  __nv_bfloat16* outs[3] = {o_hi, o_mid, o_lo};
  const long long tiles_per_sample = dhw / tile_v;
  const int groups = c >> 2;
  const int gw = groups < 32 ? groups : 32;
  const int vqw = 32 / gw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lg = lane % gw, lq = lane / gw;
  const int quads = tile_v >> 2;
  const int gblocks = (groups + gw - 1) / gw;
  const int items = gblocks * ((quads + vqw - 1) / vqw);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n_idx = tile / tiles_per_sample;
    const long long v0 = (tile % tiles_per_sample) * tile_v;
    const float* xs = x + n_idx * (long long)c * dhw + v0;
    for (int it = warp; it < items; it += SP_THREADS / 32) {
      const int g4 = (it % gblocks) * gw + lg, vq = (it / gblocks) * vqw + lq;
      if (g4 >= groups || vq >= quads) continue;
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldcs(reinterpret_cast<const float4*>(xs + (long long)(4 * g4 + k) * dhw) + vq);
      uint2 pk[3][4];                                           // [plane][voxel]: 4 channels as bf16
      float sc[4] = {1.f, 1.f, 1.f, 1.f};
      if (DIGITS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int e = 23 - __ldg(ch_exp + 4 * g4 + k);
          e = e < -126 ? -126 : (e > 127 ? 127 : e);
          sc[k] = __int_as_float((127 + e) << 23);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float h[4], m[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float val = i == 0 ? v[k].x : i == 1 ? v[k].y : i == 2 ? v[k].z : v[k].w;
          if (DIGITS) dig3(val, sc[k], h[k], m[k], l[k]);
          else split3(val, h[k], m[k], l[k]);
        }
        const float* src[3] = {h, m, l};
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const __nv_bfloat162 a = __floats2bfloat162_rn(src[p][0], src[p][1]);
          const __nv_bfloat162 b = __floats2bfloat162_rn(src[p][2], src[p][3]);
          pk[p][i].x = *reinterpret_cast<const uint32_t*>(&a);
          pk[p][i].y = *reinterpret_cast<const uint32_t*>(&b);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int vi = (i + lq) & 3;
        const int row = 4 * vq + vi;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint2 w = vi == 0 ? pk[p][0] : vi == 1 ? pk[p][1] : vi == 2 ? pk[p][2] : pk[p][3];
          tiles[p][row * groups + g4] = w;
        }
      }
    }
    __syncthreads();
    const int n16 = tile_v * c / 8;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      uint4* dst = reinterpret_cast<uint4*>(outs[p] + (n_idx * dhw + v0) * c);
      const uint4* src = reinterpret_cast<const uint4*>(tiles[p]);
      for (int e = threadIdx.x; e < n16; e += SP_THREADS) dst[e] = src[e];
    }
    __syncthreads();
  }
}

// max|x| per channel of an NCDHW tensor (the channel scales of the fixed-point digits).  Non-negative floats order
// like their bit patterns, so the per-CTA maxima combine with an integer atomicMax; `out` must be zeroed by the caller.
__global__ void __launch_bounds__(256)
channel_absmax_kernel(const float* __restrict__ x, int c, long long dhw, long long chunks_per_row, float* __restrict__ out) {
  __shared__ float red[8];
  const long long row = blockIdx.x / chunks_per_row, chunk = blockIdx.x % chunks_per_row;     // row = n * c + channel
  const long long quads = dhw >> 2, per = (quads + chunks_per_row - 1) / chunks_per_row;
  const long long q0 = chunk * per, q1 = q0 + per < quads ? q0 + per : quads;
  const float4* src = reinterpret_cast<const float4*>(x + row * dhw);
  float m = 0.f;
  for (long long q = q0 + threadIdx.x; q < q1; q += 256) {
    const float4 v = __ldg(src + q);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  if (chunk == chunks_per_row - 1)
    for (long long e = (quads << 2) + threadIdx.x; e < dhw; e += 256) m = fmaxf(m, fabsf(x[row * dhw + e]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    atomicMax(reinterpret_cast<int*>(out + (row % c)), __float_as_int(m));
  }
}

}  // namespace effq

extern "C" int effq_channel_absmax(const float* x, int32_t n, int32_t c, int64_t dhw, float* out_zeroed, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && out_zeroed && n > 0 && c > 0 && dhw > 0, "bad argument");
  EFFQ_CHECK_ARG(((uintptr_t)x & 15) == 0 && dhw % 4 == 0, "x must be 16B aligned with dhw % 4 == 0");
  const long long rows = (long long)n * c;
  long long chunks = ((long long)sm_count() * 8 + rows - 1) / rows;
  const long long max_chunks = (dhw / 4 + 1023) / 1024;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  EFFQ_CHECK_ARG(rows * chunks < (1ll << 31), "tensor too large");
  channel_absmax_kernel<<<(unsigned)(rows * chunks), 256, 0, (cudaStream_t)stream>>>(x, c, dhw, chunks, out_zeroed);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_fixdigits_ndhwc(const float* x, int32_t n, int32_t c, int64_t dhw, const int32_t* ch_exp, void* d0_out,
                                    void* d1_out, void* d2_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && ch_exp && d0_out && d1_out && d2_out, "null pointer");
  const int tile_v = effq_split3_ndhwc_supported(c, dhw);
  EFFQ_CHECK_ARG(tile_v > 0, "shape not supported (c % 8, channel groups a power of two or a multiple of 32, dhw % 16)");
  EFFQ_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)d0_out & 15) == 0 && ((uintptr_t)d1_out & 15) == 0 &&
                     ((uintptr_t)d2_out & 15) == 0, "pointers must be 16B aligned");
  if (n <= 0) return 0;
  const long long n_tiles = (long long)n * (dhw / tile_v);
  const long long cap = (long long)sm_count() * 4;
  const size_t smem = (size_t)tile_v * c * 6;
  static bool configured = false;
  if (!configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(split3_ndhwc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = true;
  }
  split3_ndhwc_kernel<true><<<(unsigned)(n_tiles < cap ? n_tiles : cap), SP_THREADS, smem, (cudaStream_t)stream>>>(
      x, c, dhw, tile_v, n_tiles, (__nv_bfloat16*)d0_out, (__nv_bfloat16*)d1_out, (__nv_bfloat16*)d2_out, ch_exp);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_split3_ndhwc_supported(int32_t c, int64_t dhw) {
  const int groups = c / 4;
  const bool g_ok = c > 0 && c % 8 == 0 && (groups >= 32 ? groups % 32 == 0 : (groups & (groups - 1)) == 0);
  if (!g_ok) return 0;
  for (int cand = 256; cand >= 16; cand >>= 1)
    if (dhw % cand == 0 && (long long)cand * c <= 8192) return cand;
  return 0;
}

extern "C" int effq_split3_ndhwc(const float* x, int32_t n, int32_t c, int64_t dhw, void* hi_out, void* mid_out,
                                 void* lo_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && hi_out && mid_out && lo_out, "null pointer");
  const int tile_v = effq_split3_ndhwc_supported(c, dhw);
  EFFQ_CHECK_ARG(tile_v > 0, "shape not supported (c % 8, channel groups a power of two or a multiple of 32, dhw % 16)");
  EFFQ_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)hi_out & 15) == 0 && ((uintptr_t)mid_out & 15) == 0 &&
                     ((uintptr_t)lo_out & 15) == 0, "pointers must be 16B aligned");
  if (n <= 0) return 0;
  const long long n_tiles = (long long)n * (dhw / tile_v);
  const long long cap = (long long)sm_count() * 4;
  const size_t smem = (size_t)tile_v * c * 6;
  static bool configured = false;
  if (!configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(split3_ndhwc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = true;
  }
  split3_ndhwc_kernel<false><<<(unsigned)(n_tiles < cap ? n_tiles : cap), SP_THREADS, smem, (cudaStream_t)stream>>>(
      x, c, dhw, tile_v, n_tiles, (__nv_bfloat16*)hi_out, (__nv_bfloat16*)mid_out, (__nv_bfloat16*)lo_out, nullptr);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t effq_ste_bwd_workspace(void) { return 16 + 8 * (int64_t)effq::STE_MAX_BLOCKS; }

extern "C" int effq_fakequant_ste_bwd(const float* x, const float* grad_out, int64_t numel, const float* alpha,
                                      float lo, float hi, int32_t nlvl, float* grad_x_out, double* grad_alpha_acc,
                                      void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && grad_out && alpha && grad_alpha_acc && workspace, "null pointer");
  EFFQ_CHECK_ARG(nlvl >= 2, "nlvl must be >= 2");
  EFFQ_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)grad_out & 15) == 0 &&
                     (!grad_x_out || ((uintptr_t)grad_x_out & 15) == 0), "pointers must be 16B aligned");
  if (numel <= 0) return 0;
  const QParamF q = make_qparam_f(lo, hi, nlvl);
  long long blocks = (numel / 4 + STE_THREADS * 4 - 1) / (STE_THREADS * 4);
  const long long cap = (long long)sm_count() * 8 < STE_MAX_BLOCKS ? (long long)sm_count() * 8 : STE_MAX_BLOCKS;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  unsigned int* done = (unsigned int*)workspace;          // zero-initialised by the caller, self-resetting
  double* partials = (double*)((char*)workspace + 16);
  cudaStream_t s = (cudaStream_t)stream;
  if (grad_x_out)
    fakequant_ste_bwd_kernel<true><<<(unsigned)blocks, STE_THREADS, 0, s>>>(x, grad_out, numel, alpha, q, grad_x_out,
                                                                            grad_alpha_acc, partials, done);
  else
    fakequant_ste_bwd_kernel<false><<<(unsigned)blocks, STE_THREADS, 0, s>>>(x, grad_out, numel, alpha, q, nullptr,
                                                                             grad_alpha_acc, partials, done);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_adam_step(float* params, const double* grads, float grad_scale, float* exp_avg, float* exp_avg_sq,
                              int32_t n, float lr, float beta1, float beta2, float eps, int32_t step, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "null pointer");
  EFFQ_CHECK_ARG(n > 0 && step >= 1, "bad size / step");
  adam_step_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(params, grads, grad_scale, exp_avg, exp_avg_sq, n,
                                                                       lr, beta1, beta2, eps, step);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
