// Fused fake-quant kernels (HBM-bound).
//   effq_fakequant_f32      : reference PTQConv._quantize_act/_quantize_w
//                             (src/models/PTQConv.py:110-116 -> layer_helper.py:25-37)
//   effq_quantize_act_ndhwc : same arithmetic, emits channels-last bf16 integer codes,
//                             the operand format of the tcgen05 conv.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp8.h>
#include <stdlib.h>

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
// Generic form (any level count): fp64 per element.
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

// <= 256 levels (every configured layer): the fp64 kernel above spends ~40 issue slots per element on conversions and
// fp64 arithmetic and reaches 59 % of the HBM peak (BENCH_r01).  Here the level index comes from ONE fp32 FMA + rint;
// only elements within a safety margin of a rounding tie (or NaN) take the exact fp64 sequence, so the index is the
// fp64 kernel's bit for bit (same argument as quantize_act_ndhwc_v2), and the output value fp32(a) * fp32(level) is
// read from a per-CTA table of the nlvl possible results, built once with the fp64 expressions of the generic kernel.
// Four independent 128-bit loads in flight per thread, streaming loads and stores.
__global__ void __launch_bounds__(FQ_THREADS)
fakequant_state_lut_kernel(const float* __restrict__ x, long long numel, const effq_scale_state* __restrict__ st,
                           QParamD q, int nlvl, float* __restrict__ y) {
  __shared__ float lut[256];
  const double a64 = st->a;
  const float a32 = (float)a64;
  const QFastD qf = make_qfast_d(a64, q, nlvl);
  for (int i = threadIdx.x; i < nlvl; i += FQ_THREADS) lut[i] = __fmul_rn(a32, (float)level_value_d((double)i, q));
  __syncthreads();
  const float c1f = (float)qf.c1, c0f = (float)qf.c0, lm1 = (float)(nlvl - 1);
  const float tol = 1e-5f * (lm1 + 1.f) + 1e-5f;                   // >= 40x the fp32 evaluation error of qa
  auto value_of = [&](float val) -> float {
    const float qa = fmaf(val, c1f, c0f);
    const float r = rintf(qa);
    const bool risky = (fabsf(fabsf(qa - r) - 0.5f) < tol && qa > -1.0f && qa < lm1 + 1.0f) || !(val == val);
    if (!(val == val)) return __fmul_rn(a32, val);                 // NaN propagates like the generic kernel's
    int idx;
    if (risky) idx = (int)level_index_fast_d((double)val, a64, q, qf);
    else idx = (int)fminf(fmaxf(r, 0.f), lm1);
    return lut[idx];
  };
  const long long nvec = numel / 4;
  const long long stride = (long long)gridDim.x * FQ_THREADS;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * FQ_THREADS + threadIdx.x; i < nvec; i += stride * FQ_UNROLL) {
    float4 v[FQ_UNROLL];
#pragma unroll
    for (int u = 0; u < FQ_UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j < nvec) v[u] = __ldcs(x4 + j);
    }
#pragma unroll
    for (int u = 0; u < FQ_UNROLL; ++u) {
      const long long j = i + u * stride;
      if (j >= nvec) continue;
      __stcs(reinterpret_cast<float4*>(y) + j,
             make_float4(value_of(v[u].x), value_of(v[u].y), value_of(v[u].z), value_of(v[u].w)));
    }
  }
  if (blockIdx.x == 0) {
    const long long t = nvec * 4 + threadIdx.x;
    if (t < numel) y[t] = value_of(x[t]);
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

// ---- v2: register transpose ------------------------------------------------------------------
// The v1 kernel above is issue-bound (ncu: 106 instructions per element, 40-47 % of the HBM peak,
// profiles/r01_conv_sweep.md): fp64 arithmetic for every element plus one LDS.U16 per element in the
// transpose.  v2, used whenever dhw % 4 == 0 and the tile is full:
//  * a thread owns 4 channels x 4 consecutive voxels: four independent 128-bit loads (one per channel
//    row), a 4x4 transpose in registers, and per voxel ONE 32-bit (e4m3) / 64-bit (bf16) shared-memory
//    store of its 4-channel pack into a tile that already has the output layout [voxel][channel];
//    lanes rotate which voxel they store first, which makes the stores bank-conflict free;
//  * the tile leaves as a straight 16-byte LDS.128 -> STG.128 copy (the block is contiguous in NDHWC);
//  * the level index comes from one fp32 FMA + rint; only elements within a safety margin of a rounding
//    tie (or NaN) take the exact fp64 sequence, so the codes are identical to v1's.
template <bool F64, bool W16, bool W8>
__global__ void __launch_bounds__(QA_THREADS)
quantize_act_ndhwc_v2_kernel(const float* __restrict__ x, int c, long long dhw, int nlvl, int tile_v, long long n_tiles,
                             const effq_scale_state* __restrict__ st, const float* __restrict__ alpha_f32,
                             __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ out8) {
  extern __shared__ __align__(16) uint8_t tile_raw[];
  uint32_t* t16 = reinterpret_cast<uint32_t*>(tile_raw);                                    // [tile_v][c] bf16
  uint32_t* t8 = reinterpret_cast<uint32_t*>(tile_raw + (W16 ? (size_t)tile_v * c * 2 : 0)); // [tile_v][c] e4m3
  const long long tiles_per_sample = dhw / tile_v;                 // host guarantees dhw % tile_v == 0

  const QParamF qf = make_qparam_f(0.f, 1.f, nlvl);
  const QParamD qd = make_qparam_d(0.f, 1.f, nlvl);
  const double a64 = F64 ? st->a : 1.0;
  const float a32 = F64 ? 1.f : __ldg(alpha_f32);
  const QFastD fd = make_qfast_d(a64, qd, nlvl);
  const QFastF ff = make_qfast_f(a32, qf, nlvl);
  const float c1f = (float)fd.c1, lm1 = (float)(nlvl - 1);
  const float tol = 1e-5f * (lm1 + 1.f) + 1e-5f;                   // >= 40x the fp32 evaluation error of qa
  auto code_of = [&](float val) -> float {
    if (F64) {
      const float qa = val * c1f;                                  // lo = 0: c0 = 0
      const float r = rintf(qa);
      const bool risky = (fabsf(fabsf(qa - r) - 0.5f) < tol && qa > -1.0f && qa < lm1 + 1.0f) || !(val == val);
      if (risky) return (float)level_index_fast_d((double)val, a64, qd, fd);
      return fminf(fmaxf(r, 0.f), lm1);
    }
    return level_index_fast_f(val, a32, qf, ff);
  };

  const int groups = c >> 2;                                       // 4-channel groups
  const int gw = groups < 32 ? groups : 32;                        // groups covered by one warp
  const int vqw = 32 / gw;                                         // voxel quads covered by one warp (gw is a power of 2 <= 32
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;      //  or groups >= 32; host check)
  const int lg = lane % gw, lq = lane / gw;
  const int quads = tile_v >> 2;
  const int gblocks = (groups + gw - 1) / gw;
  const int items = gblocks * ((quads + vqw - 1) / vqw);           // warp-level work items
  // Persistent: a CTA walks tiles with a grid stride, so the scale set-up above (fp64 divisions) is paid once
  // per CTA: 41-47 % -> 53-61 % of the HBM peak.  Measured and rejected: issuing all of a thread's loads of a
  // tile first (80 registers + spills: 28 %), prefetching the next tile's loads across the transpose
  // (126 registers, 2 CTAs per SM: 33 %).  What is left is memory-latency stall on the first use of the loaded
  // data (ncu: 31 % of the samples); a TMA-staged ring is the next step.
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n_idx = tile / tiles_per_sample;
    const long long v0 = (tile % tiles_per_sample) * tile_v;
    const float* xs = x + n_idx * (long long)c * dhw + v0;
    for (int it = warp; it < items; it += QA_THREADS / 32) {
      const int gb = it % gblocks, qb = it / gblocks;
      const int g4 = gb * gw + lg, vq = qb * vqw + lq;
      if (g4 >= groups || vq >= quads) continue;
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldcs(reinterpret_cast<const float4*>(xs + (long long)(4 * g4 + k) * dhw) + vq);
      float cd[4][4];                                              // [voxel][channel]
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cd[0][k] = code_of(v[k].x); cd[1][k] = code_of(v[k].y); cd[2][k] = code_of(v[k].z); cd[3][k] = code_of(v[k].w);
      }
      uint32_t p8[4];
      uint2 p16[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (W8)
          p8[i] = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(cd[i][0], cd[i][1]), __NV_SATFINITE, __NV_E4M3) |
                  ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(cd[i][2], cd[i][3]), __NV_SATFINITE, __NV_E4M3) << 16);
        if (W16) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(cd[i][0], cd[i][1]);
          const __nv_bfloat162 hi = __floats2bfloat162_rn(cd[i][2], cd[i][3]);
          p16[i].x = *reinterpret_cast<const uint32_t*>(&lo);
          p16[i].y = *reinterpret_cast<const uint32_t*>(&hi);
        }
      }
      // store order rotated by the lane's quad index: in each store instruction the warp's quads write
      // different voxel rows -> different banks
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int vi = (i + lq) & 3;
        const int row = 4 * vq + vi;
        if (W8) {
          const uint32_t w = vi == 0 ? p8[0] : vi == 1 ? p8[1] : vi == 2 ? p8[2] : p8[3];
          t8[row * groups + g4] = w;
        }
        if (W16) {
          const uint2 w = vi == 0 ? p16[0] : vi == 1 ? p16[1] : vi == 2 ? p16[2] : p16[3];
          reinterpret_cast<uint2*>(t16)[row * groups + g4] = w;
        }
      }
    }
    __syncthreads();
    if (W8) {
      uint4* dst = reinterpret_cast<uint4*>(out8 + (n_idx * dhw + v0) * c);
      const uint4* src = reinterpret_cast<const uint4*>(t8);
      const int n16 = tile_v * c / 16;
      for (int e = threadIdx.x; e < n16; e += QA_THREADS) dst[e] = src[e];
    }
    if (W16) {
      uint4* dst = reinterpret_cast<uint4*>(out + (n_idx * dhw + v0) * c);
      const uint4* src = reinterpret_cast<const uint4*>(t16);
      const int n16 = tile_v * c / 8;
      for (int e = threadIdx.x; e < n16; e += QA_THREADS) dst[e] = src[e];
    }
    __syncthreads();                                               // the tile is reused by the next iteration
  }
}

// ---- v3: TMA-staged ring ---------------------------------------------------------------------
// v2 waits on its own global loads (ncu, profiles/r02_layer_ncu.md: DRAM 44 %, SM 50 %, warps active 49 %), and more
// loads per thread cost registers (measured and rejected above).  v3 takes the loads off the threads: a producer warp
// streams whole tiles (C rows of TV voxels, C * TV = 4096 elements = 16 KB) into a 3-stage shared-memory ring with one
// cp.async.bulk per channel row (mbarrier complete_tx); 8 consumer warps read a stage with conflict-free LDS.128 (row
// groups are skewed by 16 B so that the 8 lanes of a quarter-warp, which own 8 different channel groups, hit 8
// different bank quads), run v2's index arithmetic and register transpose, and assemble the NDHWC tile in one of two
// output buffers, which leaves as ONE cp.async.bulk shared -> global store per code type while the next tile is being
// computed.  One named barrier per tile; a stage goes back to the producer with one mbarrier arrive per warp; three
// CTAs (74 KB each) per SM overlap each other's barrier phases.
// Measured (tools/hbm_bench.py, 32 x 32 x 64^3, profiles/r02_hbm_bench.md): bf16 + e4m3 codes 4.29 TB/s = 65 % of the
// copy peak (v2: 62 %), e4m3 only 59 % (v2: 53 %); C = 64 with both code types 54 % (v2: 49 %).  Variants measured and
// rejected: one CTA per SM with 16 consumer warps, 32 KB tiles and four stages (64 % / 47 %: all warps of the SM move
// in lock-step through load -> convert -> store -> barrier); two CTAs per SM with 32 KB tiles and two items per thread
// (57 %).  More resident warps help and longer copies do not: what is left is the issue rate of the convert + transpose
// code (~50 instructions per element incl. the tie test), not the memory system.
constexpr int Q3_CONSUMERS = 256;
constexpr int Q3_THREADS = Q3_CONSUMERS + 32;
constexpr int Q3_STAGES = 3;
constexpr int Q3_CTAS_PER_SM = 3;                      // CTAs in different phases hide each other's per-tile barrier
constexpr int Q3_ELEMS = 4096;                         // elements per tile (C * tile_v)
constexpr int Q3_ITEMS = Q3_ELEMS / 16 / Q3_CONSUMERS; // a thread's item = 4 channels x 4 voxels
constexpr uint32_t Q3_IN_BYTES = Q3_ELEMS * 4 + 512;   // + 16 B skew per 4-row group (<= 32 groups)
constexpr uint32_t Q3_SMEM = Q3_STAGES * Q3_IN_BYTES + 2 * (Q3_ELEMS * 2) + 2 * Q3_ELEMS + 128;

__device__ __forceinline__ void q3_wait(uint32_t bar, uint32_t parity) {
  unsigned int spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > (1u << 26)) __trap();                // a lost transaction must fail the launch, not hang the GPU
  }
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

template <bool F64, bool W16, bool W8>
__global__ void __launch_bounds__(Q3_THREADS, Q3_CTAS_PER_SM)
quantize_act_ndhwc_v3_kernel(const float* __restrict__ x, int c, long long dhw, int nlvl, int tile_v, long long n_tiles,
                             const effq_scale_state* __restrict__ st, const float* __restrict__ alpha_f32,
                             __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ out8) {
  extern __shared__ __align__(128) uint8_t q3_raw[];
  __shared__ __align__(8) uint64_t q3_bars[2 * Q3_STAGES];
  uint8_t* base = q3_raw + ((128u - (smem_u32(q3_raw) & 127u)) & 127u);
  uint8_t* in_base = base;
  uint8_t* o16_base = base + Q3_STAGES * Q3_IN_BYTES;                  // 2 x [tile_v][c] bf16
  uint8_t* o8_base = o16_base + 2 * (Q3_ELEMS * 2);                    // 2 x [tile_v][c] e4m3
  const uint32_t bar0 = smem_u32(q3_bars);
  auto FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(Q3_STAGES + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Q3_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), Q3_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long tiles_per_sample = dhw / tile_v;                     // host guarantees dhw % tile_v == 0
  const uint32_t row_bytes = (uint32_t)tile_v * 4u;

  if (warp == Q3_CONSUMERS / 32) {
    // ===== producer: one bulk copy per channel row, rows spread over the lanes =====
    int stage = 0;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_idx = tile / tiles_per_sample;
      const long long v0 = (tile % tiles_per_sample) * tile_v;
      const float* xs = x + n_idx * (long long)c * dhw + v0;
      if (lane == 0) {
        q3_wait(EMPTY(stage), phase ^ 1u);
        mbar_expect_tx(FULL(stage), (uint32_t)c * row_bytes);
      }
      __syncwarp();
      const uint32_t dst0 = smem_u32(in_base) + (uint32_t)stage * Q3_IN_BYTES;
      for (int r = lane; r < c; r += 32)
        bulk_g2s(dst0 + (uint32_t)r * row_bytes + 16u * (uint32_t)(r >> 2), xs + (long long)r * dhw, row_bytes, FULL(stage));
      if (++stage == Q3_STAGES) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // ===== consumers =====
  const QParamF qf = make_qparam_f(0.f, 1.f, nlvl);
  const QParamD qd = make_qparam_d(0.f, 1.f, nlvl);
  const double a64 = F64 ? st->a : 1.0;
  const float a32 = F64 ? 1.f : __ldg(alpha_f32);
  const QFastD fd = make_qfast_d(a64, qd, nlvl);
  const QFastF ff = make_qfast_f(a32, qf, nlvl);
  const float c1f = (float)fd.c1, lm1 = (float)(nlvl - 1);
  const float tol = 1e-5f * (lm1 + 1.f) + 1e-5f;
  auto code_of = [&](float val) -> float {
    if (F64) {
      const float qa = val * c1f;
      const float r = rintf(qa);
      const bool risky = (fabsf(fabsf(qa - r) - 0.5f) < tol && qa > -1.0f && qa < lm1 + 1.0f) || !(val == val);
      if (risky) return (float)level_index_fast_d((double)val, a64, qd, fd);
      return fminf(fmaxf(r, 0.f), lm1);
    }
    return level_index_fast_f(val, a32, qf, ff);
  };
  const int groups = c >> 2;                       // 8, 16 or 32 (host check)
  const int gw = groups;                           // one warp spans all channel groups ...
  const int vqw = 32 / gw;                         // ... of vqw voxel quads; 8 warps x vqw quads = tile_v / 4
  const int g4 = lane % gw, lq = lane / gw;
  int stage = 0;
  uint32_t phase = 0;
  int ob = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long n_idx = tile / tiles_per_sample;
    const long long v0 = (tile % tiles_per_sample) * tile_v;
    q3_wait(FULL(stage), phase);
    uint32_t* t16 = reinterpret_cast<uint32_t*>(o16_base + (size_t)ob * (Q3_ELEMS * 2));
    uint32_t* t8 = reinterpret_cast<uint32_t*>(o8_base + (size_t)ob * Q3_ELEMS);
    float4 v[Q3_ITEMS][4];
#pragma unroll
    for (int it = 0; it < Q3_ITEMS; ++it) {
      const int vq = (warp + it * (Q3_CONSUMERS / 32)) * vqw + lq;
      const uint8_t* in = in_base + (size_t)stage * Q3_IN_BYTES + (size_t)(4 * g4) * row_bytes + 16u * (uint32_t)g4 + 16u * (uint32_t)vq;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[it][k] = *reinterpret_cast<const float4*>(in + (size_t)k * row_bytes);
    }
#pragma unroll
    for (int it = 0; it < Q3_ITEMS; ++it) {
      const int vq = (warp + it * (Q3_CONSUMERS / 32)) * vqw + lq;
      float cd[4][4];                                                   // [voxel][channel]
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cd[0][k] = code_of(v[it][k].x); cd[1][k] = code_of(v[it][k].y); cd[2][k] = code_of(v[it][k].z); cd[3][k] = code_of(v[it][k].w);
      }
      uint32_t p8[4];
      uint2 p16[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (W8)
          p8[i] = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(cd[i][0], cd[i][1]), __NV_SATFINITE, __NV_E4M3) |
                  ((uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(cd[i][2], cd[i][3]), __NV_SATFINITE, __NV_E4M3) << 16);
        if (W16) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(cd[i][0], cd[i][1]);
          const __nv_bfloat162 hi = __floats2bfloat162_rn(cd[i][2], cd[i][3]);
          p16[i].x = *reinterpret_cast<const uint32_t*>(&lo);
          p16[i].y = *reinterpret_cast<const uint32_t*>(&hi);
        }
      }
      if (it == Q3_ITEMS - 1) {
        // every loaded value has been consumed: the stage can go back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(EMPTY(stage));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {                                       // rotated store order: conflict-free (see v2)
        const int vi = (i + lq) & 3;
        const int row = 4 * vq + vi;
        if (W8) {
          const uint32_t w = vi == 0 ? p8[0] : vi == 1 ? p8[1] : vi == 2 ? p8[2] : p8[3];
          t8[row * groups + g4] = w;
        }
        if (W16) {
          const uint2 w = vi == 0 ? p16[0] : vi == 1 ? p16[1] : vi == 2 ? p16[2] : p16[3];
          reinterpret_cast<uint2*>(t16)[row * groups + g4] = w;
        }
      }
    }
    fence_proxy_async();                                                // generic writes -> visible to the bulk store
    if (threadIdx.x == 0) bulk_wait_read0();                            // the previous tile's store has left its buffer
    asm volatile("bar.sync 1, %0;" ::"n"(Q3_CONSUMERS) : "memory");
    if (threadIdx.x == 0) {
      if (W16) bulk_s2g(out + (n_idx * dhw + v0) * c, smem_u32(t16), (uint32_t)(Q3_ELEMS * 2));
      if (W8) bulk_s2g(out8 + (n_idx * dhw + v0) * c, smem_u32(t8), (uint32_t)Q3_ELEMS);
      bulk_commit();
    }
    ob ^= 1;
    if (++stage == Q3_STAGES) { stage = 0; phase ^= 1u; }
  }
  if (threadIdx.x == 0) bulk_wait_read0();
}

template <bool F64>
static int launch_quantize_v3(const float* x, int n, int c, long long dhw, int nlvl, const effq_scale_state* st,
                              const float* alpha, __nv_bfloat16* out, uint8_t* out8, cudaStream_t s) {
  const int tile_v = Q3_ELEMS / c;
  const long long n_tiles = (long long)n * (dhw / tile_v);
  const long long cap = (long long)sm_count() * Q3_CTAS_PER_SM;
  const long long ctas = n_tiles < cap ? n_tiles : cap;
  static bool configured = false;
  if (!configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_v3_kernel<F64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q3_SMEM));
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_v3_kernel<F64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q3_SMEM));
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_v3_kernel<F64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Q3_SMEM));
    configured = true;
  }
  if (out && out8)
    quantize_act_ndhwc_v3_kernel<F64, true, true><<<(unsigned)ctas, Q3_THREADS, Q3_SMEM, s>>>(x, c, dhw, nlvl, tile_v, n_tiles, st, alpha, out, out8);
  else if (out)
    quantize_act_ndhwc_v3_kernel<F64, true, false><<<(unsigned)ctas, Q3_THREADS, Q3_SMEM, s>>>(x, c, dhw, nlvl, tile_v, n_tiles, st, alpha, out, out8);
  else
    quantize_act_ndhwc_v3_kernel<F64, false, true><<<(unsigned)ctas, Q3_THREADS, Q3_SMEM, s>>>(x, c, dhw, nlvl, tile_v, n_tiles, st, alpha, out, out8);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

template <bool F64>
static int launch_quantize_v2(const float* x, int n, int c, long long dhw, int nlvl, int tile_v,
                              const effq_scale_state* st, const float* alpha, __nv_bfloat16* out, uint8_t* out8,
                              cudaStream_t s) {
  const long long n_tiles = (long long)n * (dhw / tile_v);
  const size_t smem = (size_t)tile_v * c * ((out ? 2 : 0) + (out8 ? 1 : 0));
  const long long cap = (long long)sm_count() * 8;                 // grid stride beyond 8 CTAs per SM (the scale set-up is per CTA)
  const long long tiles = n_tiles < cap ? n_tiles : cap;
  static bool configured = false;
  if (!configured) {
    const int mx = 96 * 1024;
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_v2_kernel<F64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_v2_kernel<F64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    EFFQ_CUDA(cudaFuncSetAttribute(quantize_act_ndhwc_v2_kernel<F64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    configured = true;
  }
  if (out && out8)
    quantize_act_ndhwc_v2_kernel<F64, true, true><<<(unsigned)tiles, QA_THREADS, smem, s>>>(x, c, dhw, nlvl, tile_v, n_tiles, st, alpha, out, out8);
  else if (out)
    quantize_act_ndhwc_v2_kernel<F64, true, false><<<(unsigned)tiles, QA_THREADS, smem, s>>>(x, c, dhw, nlvl, tile_v, n_tiles, st, alpha, out, out8);
  else
    quantize_act_ndhwc_v2_kernel<F64, false, true><<<(unsigned)tiles, QA_THREADS, smem, s>>>(x, c, dhw, nlvl, tile_v, n_tiles, st, alpha, out, out8);
  EFFQ_LAUNCH_CHECK();
  return 0;
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
  const long long cap = (long long)sm_count() * 8;     // 8 resident CTAs/SM, grid-stride beyond (measured against one
                                                       // chunk per CTA: 0.96 vs 0.92 of the copy peak; the set-up is per CTA)
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
  static const bool generic_only = [] { const char* v = getenv("EFFQ_FQ_STATE_F64"); return v && *v == '1'; }();
  if (nlvl <= 256 && !generic_only && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y_out & 15) == 0) {
    long long blocks = (numel / 4 + (long long)FQ_THREADS * FQ_UNROLL - 1) / ((long long)FQ_THREADS * FQ_UNROLL);
    const long long cap = (long long)sm_count() * 8;   // persistent: the table is built once per CTA (0.86 vs 0.73 of the peak)
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    fakequant_state_lut_kernel<<<(unsigned)blocks, FQ_THREADS, 0, (cudaStream_t)stream>>>(
        x, numel, state, make_qparam_d(lo, hi, nlvl), nlvl, y_out);
    EFFQ_LAUNCH_CHECK();
    return 0;
  }
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
  // v3 (TMA-staged ring) where it measured faster than v2: C = 32 (the level-1 tensors: three quarters of the
  // activation bytes of the BraTS net) and C = 64 with both code types; whole 4096-element tiles, 16-byte aligned rows
  {
    static const bool no_v3 = [] { const char* v = getenv("EFFQ_QA_V3"); return v && *v == '0'; }();
    if (!no_v3 && (c == 32 || (c == 64 && out && out8)) && dhw % (Q3_ELEMS / c) == 0 && ((uintptr_t)x & 15) == 0 &&
        (!out8 || ((uintptr_t)out8 & 15) == 0))
      return use_f64 ? launch_quantize_v3<true>(x, n, c, dhw, nlvl, state, alpha_f32, out, out8, s)
                     : launch_quantize_v3<false>(x, n, c, dhw, nlvl, state, alpha_f32, out, out8, s);
  }
  // v2 (register transpose): full tiles only, channel groups a power of two below 32 or a multiple of 32
  {
    const int groups = c / 4;
    const bool g_ok = (c % 4 == 0) && (groups >= 32 ? groups % 32 == 0 : (groups & (groups - 1)) == 0);
    static const bool v1_only = [] { const char* v = getenv("EFFQ_QA_V1"); return v && *v == '1'; }();
    int tv = 0;
    for (int cand = 256; cand >= 16; cand >>= 1)                  // largest tile that divides dhw and fits 48 KB
      if (dhw % cand == 0 && (size_t)cand * c * 3 <= 48 * 1024) { tv = cand; break; }
    // measured (profiles/r01_conv_sweep.md): v2 wins up to C = 128 (47-61 % vs 41-47 % of the HBM peak), v1 above
    if (g_ok && tv > 0 && !v1_only && c <= 128 && ((uintptr_t)x & 15) == 0 && (!out8 || c % 16 == 0) && (!out || c % 8 == 0))
      return use_f64 ? launch_quantize_v2<true>(x, n, c, dhw, nlvl, tv, state, alpha_f32, out, out8, s)
                     : launch_quantize_v2<false>(x, n, c, dhw, nlvl, tv, state, alpha_f32, out, out8, s);
  }
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
