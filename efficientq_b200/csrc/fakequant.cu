// Fused fake-quant kernels (HBM-bound).
//   effq_fakequant_f32      : reference PTQConv._quantize_act/_quantize_w
//                             (src/models/PTQConv.py:110-116 -> layer_helper.py:25-37)
//   effq_quantize_act_ndhwc : same arithmetic, emits channels-last bf16 integer codes,
//                             the operand format of the tcgen05 conv.
#include "common.cuh"
#include <cuda_fp8.h>

namespace effq {

constexpr int FQ_THREADS = 256;
constexpr int FQ_VEC = 4;        // one 128-bit load per thread per step
constexpr int FQ_UNROLL = 4;     // independent 128-bit loads in flight per thread

template <bool WRITE_Y, bool WRITE_C>
__global__ void __launch_bounds__(FQ_THREADS)
fakequant_f32_kernel(const float* __restrict__ x, long long numel, const float* __restrict__ alpha_p,
                     QParamF q, float* __restrict__ y, uint8_t* __restrict__ code) {
  const float alpha = __ldg(alpha_p);
  const QFastF qf = make_qfast_f(alpha, q, (int)(rintf((q.hi - q.lo) / q.delta)) + 1);
  const long long nvec = numel / FQ_VEC;
  const long long stride = (long long)gridDim.x * FQ_THREADS;
  long long i = (long long)blockIdx.x * FQ_THREADS + threadIdx.x;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (; i < nvec; i += stride * FQ_UNROLL) {
    float4 v[FQ_UNROLL];
#pragma unroll
    for (int u = 0; u < FQ_UNROLL; ++u) {
      long long j = i + u * stride;
      if (j < nvec) v[u] = __ldcs(x4 + j);          // streamed: read once
    }
#pragma unroll
    for (int u = 0; u < FQ_UNROLL; ++u) {
      long long j = i + u * stride;
      if (j >= nvec) continue;
      float in[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      float out[4];
      uint32_t packed = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float idx = level_index_fast_f(in[e], alpha, q, qf);
        out[e] = __fmul_rn(level_value_f(idx, q), alpha);
        packed |= ((uint32_t)(int)idx & 0xffu) << (8 * e);
      }
      if (WRITE_Y) __stcs(reinterpret_cast<float4*>(y) + j, make_float4(out[0], out[1], out[2], out[3]));
      if (WRITE_C) reinterpret_cast<uint32_t*>(code)[j] = packed;
    }
  }
  // tail (numel % 4), handled by block 0
  if (blockIdx.x == 0) {
    long long t = nvec * FQ_VEC + threadIdx.x;
    if (t < numel) {
      float idx = level_index_fast_f(x[t], alpha, q, qf);
      if (WRITE_Y) y[t] = __fmul_rn(level_value_f(idx, q), alpha);
      if (WRITE_C) code[t] = (uint8_t)(int)idx;
    }
  }
}

// Qact = a_act * b_act with b from project_by_iter's final fp64 discretize
// (reference EfficientQConv.py:68-70, layer_helper.py:67): fp32(a) * fp32(level(x/a)).
__global__ void __launch_bounds__(FQ_THREADS)
fakequant_state_kernel(const float* __restrict__ x, long long numel, const effq_scale_state* __restrict__ st,
                       QParamD q, float* __restrict__ y) {
  const double a64 = st->a;
  const float a32 = (float)a64;
  const QFastD qf = make_qfast_d(a64, q, (int)(rint((q.hi - q.lo) / q.delta)) + 1);
  const long long nvec = numel / 4;
  const long long stride = (long long)gridDim.x * FQ_THREADS;
  for (long long i = (long long)blockIdx.x * FQ_THREADS + threadIdx.x; i < nvec; i += stride) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(x) + i);
    float4 o;
    o.x = __fmul_rn(a32, (float)level_value_d(level_index_fast_d((double)v.x, a64, q, qf), q));
    o.y = __fmul_rn(a32, (float)level_value_d(level_index_fast_d((double)v.y, a64, q, qf), q));
    o.z = __fmul_rn(a32, (float)level_value_d(level_index_fast_d((double)v.z, a64, q, qf), q));
    o.w = __fmul_rn(a32, (float)level_value_d(level_index_fast_d((double)v.w, a64, q, qf), q));
    __stcs(reinterpret_cast<float4*>(y) + i, o);
  }
  if (blockIdx.x == 0) {
    const long long t = nvec * 4 + threadIdx.x;
    if (t < numel) y[t] = __fmul_rn(a32, (float)level_value_d(level_index_fast_d((double)x[t], a64, q, qf), q));
  }
}

// NCDHW fp32 -> NDHWC bf16 codes.  One CTA handles TILE_V consecutive voxels of one sample for
// all C channels: 128-bit loads along the voxel axis (1 KB contiguous per channel at TILE_V = 256),
// transpose through shared memory, one contiguous TILE_V*C*2-byte store in 16-byte pieces.
constexpr int QA_THREADS = 256;

template <bool F64, int TILE_V>
__global__ void __launch_bounds__(QA_THREADS)
quantize_act_ndhwc_kernel(const float* __restrict__ x, int c, long long dhw, int nlvl,
                          const effq_scale_state* __restrict__ st, const float* __restrict__ alpha_f32,
                          __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ out8) {
  extern __shared__ __nv_bfloat16 tile[];      // channel-major [c][TILE_V + 4]: both phases conflict-free
  constexpr int VP = TILE_V + 4;
  const long long tiles_per_sample = (dhw + TILE_V - 1) / TILE_V;
  const long long n_idx = blockIdx.x / tiles_per_sample;
  const long long v0 = (blockIdx.x % tiles_per_sample) * TILE_V;
  const int nv = (int)min((long long)TILE_V, dhw - v0);
  const float* xs = x + n_idx * (long long)c * dhw + v0;

  const QParamF qf = make_qparam_f(0.f, 1.f, nlvl);
  const QParamD qd = make_qparam_d(0.f, 1.f, nlvl);
  const double a64 = F64 ? st->a : 1.0;
  const float a32 = F64 ? 1.f : __ldg(alpha_f32);
  const QFastD fd = make_qfast_d(a64, qd, nlvl);
  const QFastF ff = make_qfast_f(a32, qf, nlvl);
  auto code_of = [&](float val) -> float {
    return F64 ? (float)level_index_fast_d((double)val, a64, qd, fd) : level_index_fast_f(val, a32, qf, ff);
  };

  constexpr int VQ = TILE_V / 4;                // float4 groups per channel row
  const bool vec_ok = (dhw % 4 == 0) && (nv == TILE_V) && ((reinterpret_cast<uintptr_t>(xs) & 15) == 0);
  if (vec_ok) {
    for (int e = threadIdx.x; e < c * VQ; e += QA_THREADS) {
      const int ch = e / VQ, vq = e % VQ;        // a warp reads 512 contiguous bytes of one channel
      const float4 val = __ldcs(reinterpret_cast<const float4*>(xs + (long long)ch * dhw) + vq);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(code_of(val.x), code_of(val.y));
      const __nv_bfloat162 hi = __floats2bfloat162_rn(code_of(val.z), code_of(val.w));
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(&tile[ch * VP + vq * 4]) = pk;     // 8 B, consecutive lanes consecutive
    }
  } else {
    for (int e = threadIdx.x; e < c * TILE_V; e += QA_THREADS) {
      const int ch = e / TILE_V, v = e % TILE_V;
      float code = 0.f;
      if (v < nv) code = code_of(__ldcs(xs + (long long)ch * dhw + v));
      tile[ch * VP + v] = __float2bfloat16_rn(code);
    }
  }
  __syncthreads();
  if (out8) {
    // e4m3 copy of the same codes (exact for <= 16 levels): 16-byte pieces of 16 channels
    uint8_t* dst8 = out8 + (n_idx * dhw + v0) * c;
    const int chunks8 = c / 16;                 // c % 16 == 0 (host check)
    for (int e = threadIdx.x; e < nv * chunks8; e += QA_THREADS) {
      const int q = e / nv, v = e % nv;
      const unsigned short* col = reinterpret_cast<const unsigned short*>(tile) + (16 * q) * VP + v;
      uint32_t wd[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 lo2, hi2;
        lo2.x = __uint_as_float((uint32_t)col[(4 * k + 0) * VP] << 16);
        lo2.y = __uint_as_float((uint32_t)col[(4 * k + 1) * VP] << 16);
        hi2.x = __uint_as_float((uint32_t)col[(4 * k + 2) * VP] << 16);
        hi2.y = __uint_as_float((uint32_t)col[(4 * k + 3) * VP] << 16);
        wd[k] = (uint32_t)__nv_cvt_float2_to_fp8x2(lo2, __NV_SATFINITE, __NV_E4M3) |
                ((uint32_t)__nv_cvt_float2_to_fp8x2(hi2, __NV_SATFINITE, __NV_E4M3) << 16);
      }
      reinterpret_cast<uint4*>(dst8 + (long long)v * c)[q] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
  if (!out) return;
  __nv_bfloat16* dst = out + (n_idx * dhw + v0) * c;
  const int chunks = c / 8;                     // c % 8 == 0 (host check): 16-byte pieces of 8 channels
  for (int e = threadIdx.x; e < nv * chunks; e += QA_THREADS) {
    const int q = e / nv, v = e % nv;           // lanes walk v: 2-byte reads of one channel row, no conflicts
    const unsigned short* col = reinterpret_cast<const unsigned short*>(tile) + (8 * q) * VP + v;
    uint4 piece;
    piece.x = (uint32_t)col[0] | ((uint32_t)col[VP] << 16);
    piece.y = (uint32_t)col[2 * VP] | ((uint32_t)col[3 * VP] << 16);
    piece.z = (uint32_t)col[4 * VP] | ((uint32_t)col[5 * VP] << 16);
    piece.w = (uint32_t)col[6 * VP] | ((uint32_t)col[7 * VP] << 16);
    reinterpret_cast<uint4*>(dst + (long long)v * c)[q] = piece;
  }
}

}  // namespace effq

extern "C" int effq_fakequant_f32(const float* x, int64_t numel, const float* alpha, float lo, float hi,
                                  int32_t nlvl, float* y_out, uint8_t* code_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && alpha, "null input");
  EFFQ_CHECK_ARG(nlvl >= 2, "nlvl must be >= 2");
  EFFQ_CHECK_ARG(!code_out || nlvl <= 256, "uint8 codes need nlvl <= 256");
  EFFQ_CHECK_ARG(y_out || code_out, "nothing to write");
  EFFQ_CHECK_ARG(((uintptr_t)x & 15) == 0 && (!y_out || ((uintptr_t)y_out & 15) == 0) &&
                     (!code_out || ((uintptr_t)code_out & 3) == 0), "pointers must be 16B aligned");
  if (numel <= 0) return 0;
  const QParamF q = make_qparam_f(lo, hi, nlvl);
  const long long nvec = numel / FQ_VEC;
  long long blocks = (nvec + (long long)FQ_THREADS * FQ_UNROLL - 1) / ((long long)FQ_THREADS * FQ_UNROLL);
  const long long cap = (long long)sm_count() * 8;     // 8 resident CTAs/SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaStream_t s = (cudaStream_t)stream;
  if (y_out && code_out)
    fakequant_f32_kernel<true, true><<<(unsigned)blocks, FQ_THREADS, 0, s>>>(x, numel, alpha, q, y_out, code_out);
  else if (y_out)
    fakequant_f32_kernel<true, false><<<(unsigned)blocks, FQ_THREADS, 0, s>>>(x, numel, alpha, q, y_out, code_out);
  else
    fakequant_f32_kernel<false, true><<<(unsigned)blocks, FQ_THREADS, 0, s>>>(x, numel, alpha, q, y_out, code_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_fakequant_state(const float* x, int64_t numel, const effq_scale_state* state, float lo,
                                    float hi, int32_t nlvl, float* y_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && state && y_out, "null pointer");
  EFFQ_CHECK_ARG(nlvl >= 2, "nlvl must be >= 2");
  if (numel <= 0) return 0;
  long long blocks = (numel + FQ_THREADS - 1) / FQ_THREADS;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  fakequant_state_kernel<<<(unsigned)blocks, FQ_THREADS, 0, (cudaStream_t)stream>>>(
      x, numel, state, make_qparam_d(lo, hi, nlvl), y_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_quantize_act_ndhwc(const float* x, int32_t n, int32_t c, int64_t dhw, int32_t nlvl,
                                       const effq_scale_state* state, const float* alpha_f32,
                                       int32_t use_f64, void* codes_bf16_out, void* codes_e4m3_out,
                                       void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && (codes_bf16_out || codes_e4m3_out), "null pointer");
  EFFQ_CHECK_ARG(!codes_e4m3_out || (nlvl <= 16 && c % 16 == 0 && ((uintptr_t)codes_e4m3_out & 15) == 0),
                 "e4m3 codes need nlvl <= 16, c % 16 == 0 and a 16B aligned output");
  EFFQ_CHECK_ARG(use_f64 ? state != nullptr : alpha_f32 != nullptr, "missing scale");
  EFFQ_CHECK_ARG(nlvl >= 2 && nlvl <= 256, "nlvl out of range for bf16-exact codes");
  EFFQ_CHECK_ARG(c > 0 && c % 8 == 0 && c <= 512, "channel count must be a multiple of 8 and <= 512");
  EFFQ_CHECK_ARG(((uintptr_t)codes_bf16_out & 15) == 0, "output must be 16B aligned");
  if (n <= 0 || dhw <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  __nv_bfloat16* out = (__nv_bfloat16*)codes_bf16_out;
  uint8_t* out8 = (uint8_t*)codes_e4m3_out;
  const bool big = c <= 64;                     // 256-voxel tiles while the transpose tile stays <= 36 KB
  const int tile_v = big ? 256 : 64;
  const long long tiles = (long long)n * ((dhw + tile_v - 1) / tile_v);
  EFFQ_CHECK_ARG(tiles < (1ll << 31), "too many tiles");
  const size_t smem = (size_t)c * (tile_v + 4) * sizeof(__nv_bfloat16);
  static bool configured = false;
  if (!configured) {
    const int mx = 512 * (64 + 4) * 2;
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    configured = true;
  }
  if (big) {
    if (use_f64) quantize_act_ndhwc_kernel<true, 256><<<(unsigned)tiles, QA_THREADS, smem, s>>>(x, c, dhw, nlvl, state, alpha_f32, out, out8);
    else         quantize_act_ndhwc_kernel<false, 256><<<(unsigned)tiles, QA_THREADS, smem, s>>>(x, c, dhw, nlvl, state, alpha_f32, out, out8);
  } else {
    if (use_f64) quantize_act_ndhwc_kernel<true, 64><<<(unsigned)tiles, QA_THREADS, smem, s>>>(x, c, dhw, nlvl, state, alpha_f32, out, out8);
    else         quantize_act_ndhwc_kernel<false, 64><<<(unsigned)tiles, QA_THREADS, smem, s>>>(x, c, dhw, nlvl, state, alpha_f32, out, out8);
  }
  EFFQ_LAUNCH_CHECK();
  return 0;
}
