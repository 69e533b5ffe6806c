// tcgen05 implicit-GEMM 3D convolution on integer codes with the reconstruction-error
// reduction fused into the epilogue.  Replaces, for quantised-activation layers,
//     out_q = F.conv3d(Qact, G, b*) ; loss = F.mse_loss(out_q, out_fp)
// (reference src/models/EfficientQConv.py:118-122, 200x per layer, and :161-165).
//
// Arithmetic: Qact = a_x * cx/(La-1) and G = a_w * cw/(Lw-1) with cx in 0..La-1 and cw an
// integer in [-(Lw-1), Lw-1].  Codes of <= 256 levels are exact in bf16 (kind::f16, K = 16 per
// instruction), codes of <= 16 levels also in e4m3 (kind::f8f6f4, K = 32 per instruction: half
// the MMA instructions and half the operand bytes).  The products are exact and the fp32 TMEM
// accumulation of integers is exact below 2^24, so either way the tensor-core result is
// conv_scale * (exact integer) + bias -- at least as accurate as the reference's fp32 conv, and
// bit-identical between the two operand types.
//
// GEMM view: M = output voxels (tile = 16 h-rows x 8 w, one d-plane -> 128 rows),
// N = C2, K = taps * C1.  No im2col: ONE 5-D TMA box (cp.async.bulk.tensor, tile mode) brings the
// halo block (kd x (16+kh-1) x (8+kw-1) voxels x one channel block of <= 128 bytes) of a tile
// into shared memory as K-major rows with the UMMA 128/64/32-byte swizzle (tc_layout.cuh);
// box coordinates outside the volume are zero-filled by the hardware, which is exactly the
// convolution's zero padding and the ragged tile edge.  The A descriptor
// of tap (a,b,c) is the same block with the start address shifted by ((a*HH + b)*WP + c) rows
// (the swizzle is a function of the absolute address, so any row may start a descriptor), so
// every activation byte is fetched from L2 once per tile and reused by all taps.  Weights
// [tap][block][C2][CG] (pre-swizzled in global memory) arrive by 1-D bulk TMA copies
// (cp.async.bulk + mbarrier complete_tx), resident for the whole kernel when they fit, else
// through a ring.  One elected thread issues tcgen05.mma (M=128, N=C2) into a double-buffered
// TMEM accumulator; four epilogue warps drain it with tcgen05.ld, apply scale+bias, read the
// fp32 target (NCDHW), and reduce att*(out-target)^2.
//
// The fp32 target tile of the epilogue ([32 channels][16 h][8 w], NCDHW) is also a TMA box into
// a small shared-memory ring when the shape allows (W % 4 == 0, C2 % 32 == 0, room for two
// 16 KB stages); otherwise the epilogue threads load it directly.
//
// Warp roles (256 threads): 0 weight TMA | 1 MMA issuer + TMEM owner | 2-5 epilogue |
// 6 halo TMA | 7 target TMA.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include "tc_layout.cuh"
#include "tc_ptx.cuh"

namespace effq {

constexpr int TC_THREADS = 256;
constexpr int TC_TILE_H = 16;
constexpr int TC_TILE_W = 8;
constexpr uint32_t TC_TGT_BYTES = 32u * TC_TILE_H * TC_TILE_W * 4u;     // one target stage: 32 channels of a tile
constexpr int TC_EPI = 128;

struct TcParams {
  const uint8_t* xq;           // NDHWC codes (bf16 or e4m3, eb bytes each)
  const uint8_t* wq;           // [tap][group][C2][CG] pre-swizzled (tc_layout.cuh)
  const float* bias;
  const float* conv_scale;
  const float* scale_vec;      // optional [c2_total]: per-output-channel scales (conv_scale is then ignored)
  float* out;                  // NCDHW or null
  const float* target;         // NCDHW or null
  const float* att;            // N,D,H,W or null
  double* sse;
  unsigned int* ws_done;       // workspace: [done, abort, pad, pad] then partials
  double* ws_partial;
  int n, c1, c2, d, h, w;
  int kd, kh, kw, pd, ph, pw, taps;
  int cg, n_groups, nch;       // channels per halo block, blocks per tile, 16-byte chunks per row
  int eb;                      // bytes per code: 2 = bf16 (kind::f16), 1 = e4m3 (kind::f8f6f4)
  int hh, wp, hv;              // halo rows per plane, halo cols, halo voxels
  int tiles_h, tiles_w;
  long long n_tiles;
  int n_halo_stages, n_w_stages, w_resident;
  int ctas_per_sm;             // 2 when two CTAs are meant to share an SM (grid = 2 x SMs)
  int n_tgt_stages;            // > 0: the target tile arrives by TMA through a ring of this many stages
  unsigned int off_tgt;
  unsigned int halo_bytes, halo_tx_bytes, wtile_bytes;   // stage stride, bytes one halo box delivers, weight tile
  unsigned int off_bias, off_halo, off_w;
  unsigned int tmem_cols;
  int swz, rp, debug;          // operand swizzle width (128/64/32 B), row pitch, bring-up debug bits
  unsigned int wstage_bytes;   // shared-memory stride between weight tiles (1024-aligned when swizzled)
  // C2 > 256 runs as launches over chunks of 256 output channels (c2 = the chunk): the chunk's rows of every
  // [C2_total][CG] weight tile are contiguous in global memory, its outputs / targets sit c2_off channels into
  // the NCDHW tensors, and launches after the first add their squared error to *sse.
  int c2_total, c2_off, sse_accumulate;
  int add_out;                 // 1: `target` is the previous content of `out` and is ADDED to the result (no squared error)
  unsigned int wsrc_tile_bytes, wsrc_off;       // global stride between weight tiles, byte offset of the chunk's rows
  unsigned long long* dbg;     // bring-up timeline buffer ([tile][8] clock stamps of CTA 0) or null
};

__device__ __forceinline__ void dbg_stamp(const TcParams& p, unsigned int tile, int slot) {
  if (p.dbg && blockIdx.x == 0) {
    const unsigned int i = tile / gridDim.x;
    if (i < 256) p.dbg[i * 8 + slot] = clock64();
  }
}


// ---- epilogue pieces (NC = 32 or 16 channels of one voxel per thread) -------------------------
template <int NC>
__device__ __forceinline__ void epi_load_targets(float (&tv)[32], const float* tp, long long chan, bool want) {
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    tv[j] = want ? __ldg(tp) : 0.f;
    tp += chan;
  }
}
// v: raw accumulators in, outputs (scale * acc + bias) out; returns sum (out - target)^2 of the chunk
template <int NC>
__device__ __forceinline__ float epi_chunk(uint32_t (&v)[32], const float (&tv)[32], const float* bias_c0, float scale,
                                           const float* svec_c0 = nullptr, bool add = false) {
  float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NC; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias_c0 + j);
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // per-output-channel scales (optional extension): a warp-uniform broadcast load from L1
      const float sc = svec_c0 ? __ldg(svec_c0 + j + k) : scale;
      float o = fmaf(__uint_as_float(v[j + k]), sc, bb[k]);
      if (add) o += tv[j + k];                               // accumulating launch: out = conv + previous out
      const float dlt = o - tv[j + k];
      e[k] = fmaf(dlt, dlt, e[k]);
      v[j + k] = __float_as_uint(o);
    }
  }
  return (e[0] + e[1]) + (e[2] + e[3]);
}
template <int NC>
__device__ __forceinline__ void epi_store(const uint32_t (&v)[32], float* op, long long chan) {
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    __stcs(op, __uint_as_float(v[j]));
    op += chan;
  }
}

// KS: kernel edge (3 or 1), KK: MMAs per (tap, channel block) = row bytes / 32, WRES: weights
// resident in shared memory and FP8: e4m3 operands are compile-time, so the single MMA-issuing
// thread runs straight-line code.
template <int KS, int KK, bool WRES, bool FP8>
__global__ void __launch_bounds__(TC_THREADS, (KK == 1) ? 2 : 1)
conv3d_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024 B alignment
  // barrier block (8 B each): halo_full[4] halo_empty[4] w_full[8] w_empty[8] wres tmem_full[2] tmem_empty[2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_HF = 0, B_HE = 4, B_WF = 8, B_WE = 16, B_WRES = 24, B_TF = 25, B_TE = 27, B_TMEMPTR = 30;
  constexpr int B_GF = 32, B_GE = 36;            // target ring: full (TMA complete_tx) / empty (128 epilogue threads)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + B_TMEMPTR);
  float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);
  const uint32_t halo0 = smem_u32(smem + p.off_halo);
  const uint32_t wsm0 = smem_u32(smem + p.off_w);
  __shared__ double red_scratch[32];
  __shared__ bool is_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile unsigned int* abort_flag = p.ws_done + 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(BAR(B_HF + i), 1); mbar_init(BAR(B_HE + i), 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(BAR(B_WF + i), 1); mbar_init(BAR(B_WE + i), 1); }
    mbar_init(BAR(B_WRES), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_TF + i), 1); mbar_init(BAR(B_TE + i), TC_EPI); }
    for (int i = 0; i < 4; ++i) { mbar_init(BAR(B_GF + i), 1); mbar_init(BAR(B_GE + i), TC_EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.c2; i += TC_THREADS) bias_s[i] = p.bias ? p.bias[p.c2_off + i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const unsigned int tile0 = blockIdx.x, tstep = gridDim.x, n_tiles = (unsigned int)p.n_tiles;   // < 2^31 (host check)
  double err_acc = 0.0;

  if (warp == 0) {
    // ===== weight loader: one elected lane issues the 1-D bulk TMA copies =====
    if (elect_one()) {
      if (WRES) {
        const uint32_t total = p.wtile_bytes * (uint32_t)(p.n_groups * p.taps);
        mbar_expect_tx(BAR(B_WRES), total);
        for (int g = 0; g < p.n_groups; ++g)
          for (int t = 0; t < p.taps; ++t) {
            const uint8_t* src = p.wq + ((long long)t * p.n_groups + g) * p.wsrc_tile_bytes + p.wsrc_off;
            bulk_g2s(wsm0 + (uint32_t)(g * p.taps + t) * p.wstage_bytes, src, p.wtile_bytes, BAR(B_WRES));
          }
      } else {
        Pipe wp{0, 0};
        bool ok = true;
        for (unsigned int tile = tile0; tile < n_tiles && ok; tile += tstep)
          for (int g = 0; g < p.n_groups && ok; ++g)
            for (int t = 0; t < p.taps; ++t) {
              if (!mbar_wait<32>(BAR(B_WE + wp.stage), wp.phase ^ 1u, abort_flag)) { ok = false; break; }
              mbar_expect_tx(BAR(B_WF + wp.stage), p.wtile_bytes);
              const uint8_t* src = p.wq + ((long long)t * p.n_groups + g) * p.wsrc_tile_bytes + p.wsrc_off;
              bulk_g2s(wsm0 + (uint32_t)wp.stage * p.wstage_bytes, src, p.wtile_bytes, BAR(B_WF + wp.stage));
              wp.advance(p.n_w_stages);
            }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: D[128 voxels][C2] += X(tap slice of the halo) * W(tap)^T.  One elected lane
    // runs the whole loop; per MMA it does two 32-bit adds on the descriptors' address fields. =====
    if (elect_one()) {
      const uint32_t idesc = umma_idesc(128, p.c2, FP8);
      const uint64_t h_tmpl = umma_desc_sw(0, (uint32_t)(p.wp * p.rp), p.swz, 0);
      const uint64_t w_tmpl = umma_desc_sw(0, (uint32_t)(8 * p.rp), p.swz, 0);
      const uint32_t h_hi = (uint32_t)(h_tmpl >> 32), h_lo0 = (uint32_t)h_tmpl;
      const uint32_t w_hi = (uint32_t)(w_tmpl >> 32), w_lo0 = (uint32_t)w_tmpl;
      const uint32_t row16 = (uint32_t)p.rp >> 4;                 // row pitch in 16 B units
      const uint32_t step_b = (uint32_t)p.wp * row16, step_a = (uint32_t)(p.hh * p.wp) * row16;
      const uint32_t wst16 = p.wstage_bytes >> 4;
      const uint32_t hst16 = p.halo_bytes >> 4;
      const uint32_t h_base = h_lo0 + (halo0 >> 4), w_base = w_lo0 + (wsm0 >> 4);
      Pipe hp{0, 0}, wp{0, 0}, ap{0, 0};
      bool ok = true;
      if (WRES) ok = mbar_wait(BAR(B_WRES), 0, abort_flag);
      for (unsigned int tile = tile0; tile < n_tiles && ok; tile += tstep) {
        if (!mbar_wait(BAR(B_TE + ap.stage), ap.phase ^ 1u, abort_flag)) { ok = false; break; }
        tc_fence_after();
        dbg_stamp(p, tile, 0);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ap.stage * p.c2);
        for (int g = 0; g < p.n_groups && ok; ++g) {
          if (!mbar_wait(BAR(B_HF + hp.stage), hp.phase, abort_flag)) { ok = false; break; }
          tc_fence_after();
          if (g == 0) dbg_stamp(p, tile, 1);
          const uint32_t h_lo = h_base + (uint32_t)hp.stage * hst16;
          uint32_t w_lo = w_base + (uint32_t)(g * p.taps) * wst16;              // resident weights
#pragma unroll
          for (int a = 0; a < KS; ++a) {
#pragma unroll
            for (int b = 0; b < KS; ++b) {
#pragma unroll
              for (int c = 0; c < KS; ++c) {
                if (!WRES) {
                  if (ok && !mbar_wait(BAR(B_WF + wp.stage), wp.phase, abort_flag)) ok = false;
                  tc_fence_after();
                  w_lo = w_base + (uint32_t)wp.stage * wst16;
                }
                if (ok) {
                  const uint32_t ha = h_lo + (uint32_t)a * step_a + (uint32_t)b * step_b + (uint32_t)c * row16;
#pragma unroll
                  for (int kk = 0; kk < KK; ++kk)
                    tc_mma<FP8>(d_tmem, ha + 2u * kk, h_hi, w_lo + 2u * kk, w_hi, idesc,
                                (g == 0 && a == 0 && b == 0 && c == 0 && kk == 0) ? 0u : 1u);
                  if (!WRES) { tc_commit(BAR(B_WE + wp.stage)); wp.advance(p.n_w_stages); }
                }
                if (WRES) w_lo += wst16;
              }
            }
          }
          if (ok) tc_commit(BAR(B_HE + hp.stage));
          hp.advance(p.n_halo_stages);
        }
        if (ok) tc_commit(BAR(B_TF + ap.stage));
        dbg_stamp(p, tile, 2);
        ap.advance(2);
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===== epilogue: TMEM -> registers -> scale+bias -> out / squared error =====
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;               // tile row = hy*8 + wx
    const int hy = row >> 3, wx = row & 7;
    const float scale = p.scale_vec ? 1.f : __ldg(p.conv_scale);
    const float* svec = p.scale_vec ? p.scale_vec + p.c2_off : nullptr;
    const bool add_out = p.add_out != 0;
    const long long plane = (long long)p.h * p.w;
    const long long chan = (long long)p.d * plane;
    Pipe ap{0, 0}, gp{0, 0};
    const float* tgt_row = reinterpret_cast<const float*>(smem + p.off_tgt) + row;
    bool ok = true;
    for (unsigned int tile = tile0; tile < n_tiles && ok; tile += tstep) {
      unsigned int r = tile;
      const int tw = (int)(r % (unsigned)p.tiles_w); r /= (unsigned)p.tiles_w;
      const int th = (int)(r % (unsigned)p.tiles_h); r /= (unsigned)p.tiles_h;
      const int dd = (int)(r % (unsigned)p.d); r /= (unsigned)p.d;
      const int nn = (int)r;
      const int oh = th * TC_TILE_H + hy, ow = tw * TC_TILE_W + wx;
      const bool live = oh < p.h && ow < p.w;
      const long long sp = (long long)dd * plane + (long long)oh * p.w + ow;
      const long long base = ((long long)nn * p.c2_total + p.c2_off) * chan + sp;
      if (p.n_tgt_stages > 0) {
        // ---- target tile staged in shared memory by the TMA warp: [32 channels][128 voxels] fp32 ----
        float* optr = p.out + base;
        const bool store = live && p.out != nullptr;
        float e32 = 0.f;
        for (int c0 = 0; c0 < p.c2 && ok; c0 += 32) {
          if (!mbar_wait<32>(BAR(B_GF + gp.stage), gp.phase, abort_flag)) { ok = false; break; }
          if (c0 == 0) {
            if (!mbar_wait<32>(BAR(B_TF + ap.stage), ap.phase, abort_flag)) { ok = false; break; }
            if (threadIdx.x == 64) dbg_stamp(p, tile, 3);
            tc_fence_after();
          }
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ap.stage * p.c2 + c0), v);
          float tv[32];
          const float* ts = tgt_row + gp.stage * (int)(TC_TGT_BYTES / 4);
#pragma unroll
          for (int j = 0; j < 32; ++j) tv[j] = ts[j * (TC_TILE_H * TC_TILE_W)];     // lanes read consecutive floats
          tc_wait_ld();
          e32 += epi_chunk<32>(v, tv, bias_s + c0, scale, svec ? svec + c0 : nullptr, add_out);
          mbar_arrive(BAR(B_GE + gp.stage));
          gp.advance(p.n_tgt_stages);
          if (store) epi_store<32>(v, optr, chan);
          optr += 32 * chan;
        }
        if (!ok) break;
        tc_fence_before();
        mbar_arrive(BAR(B_TE + ap.stage));
        if (threadIdx.x == 64) dbg_stamp(p, tile, 4);
        ap.advance(2);
        if (live) {
          const float wv = p.att ? __ldg(p.att + (long long)nn * chan + sp) : 1.f;
          err_acc += (double)e32 * (double)wv;
        }
        continue;
      }
      // Target values do not depend on the MMA: issue all loads of a 32-channel chunk before
      // anything consumes them (32 independent requests in flight per thread), the first chunk
      // even before waiting for the accumulator.  Addresses advance by one running pointer per
      // chunk (the epilogue is instruction-bound: a 64-bit multiply per load tripled its length).
      const bool want_t = live && p.target != nullptr && !(p.debug & 16);   // bit 16: bring-up, skip the target reads
      // (a second register buffer for the next chunk's loads was tried: 168 registers with spills,
      //  12-25 % slower -- profiles/r01_conv_layout.md)
      float tv[32];
      const float* tptr = p.target + base;          // only dereferenced when want_t
      float* optr = p.out + base;                   // only dereferenced when live && p.out
      const bool store = live && p.out != nullptr;
      if (p.c2 >= 32) epi_load_targets<32>(tv, tptr, chan, want_t); else epi_load_targets<16>(tv, tptr, chan, want_t);
      if (!mbar_wait<64>(BAR(B_TF + ap.stage), ap.phase, abort_flag)) { ok = false; break; }
      if (threadIdx.x == 64) dbg_stamp(p, tile, 3);
      tc_fence_after();
      float e32 = 0.f;
      for (int c0 = 0; c0 < p.c2; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ap.stage * p.c2 + c0);
        const int rest = p.c2 - c0 - 32;            // channels after this chunk (>= 16 or <= 0)
        if (rest >= 0) {
          tc_ld32(taddr, v);
          tc_wait_ld();
          e32 += epi_chunk<32>(v, tv, bias_s + c0, scale, svec ? svec + c0 : nullptr, add_out);
          if (rest >= 32) epi_load_targets<32>(tv, tptr + 32 * chan, chan, want_t);      // next chunk's loads overlap
          else if (rest > 0) epi_load_targets<16>(tv, tptr + 32 * chan, chan, want_t);   // this chunk's stores
          if (store) epi_store<32>(v, optr, chan);
        } else {
          tc_ld16(taddr, v);
          tc_wait_ld();
          e32 += epi_chunk<16>(v, tv, bias_s + c0, scale, svec ? svec + c0 : nullptr, add_out);
          if (store) epi_store<16>(v, optr, chan);
        }
        tptr += 32 * chan;
        optr += 32 * chan;
      }
      tc_fence_before();
      mbar_arrive(BAR(B_TE + ap.stage));
      if (threadIdx.x == 64) dbg_stamp(p, tile, 4);
      ap.advance(2);
      if (want_t) {
        const float wv = p.att ? __ldg(p.att + (long long)nn * chan + sp) : 1.f;
        err_acc += (double)e32 * (double)wv;
      }
    }
  } else if (warp == 6) {
    // ===== halo loader: one 5-D TMA box per (tile, channel block); padding and ragged edges are
    // the hardware's out-of-bounds zero fill =====
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&xmap)) : "memory");
      Pipe hp{0, 0};
      bool ok = true;
      for (unsigned int tile = tile0; tile < n_tiles && ok; tile += tstep) {
        unsigned int r = tile;
        const int tw = (int)(r % (unsigned)p.tiles_w); r /= (unsigned)p.tiles_w;
        const int th = (int)(r % (unsigned)p.tiles_h); r /= (unsigned)p.tiles_h;
        const int dd = (int)(r % (unsigned)p.d); r /= (unsigned)p.d;
        const int nn = (int)r;
        const int h0 = th * TC_TILE_H - p.ph, w0 = tw * TC_TILE_W - p.pw, d0 = dd - p.pd;
        for (int g = 0; g < p.n_groups; ++g) {
          if (!mbar_wait<32>(BAR(B_HE + hp.stage), hp.phase ^ 1u, abort_flag)) { ok = false; break; }
          if (g == 0) dbg_stamp(p, tile, 5);
          mbar_expect_tx(BAR(B_HF + hp.stage), p.halo_tx_bytes);
          tma_load_5d(halo0 + (uint32_t)hp.stage * p.halo_bytes, &xmap, g * p.cg, w0, h0, d0, nn, BAR(B_HF + hp.stage));
          if (g == 0) dbg_stamp(p, tile, 6);
          hp.advance(p.n_halo_stages);
        }
      }
    }
    __syncwarp();
  } else {
    // ===== target loader: the epilogue's fp32 target tile, 32 channels per box =====
    if (p.n_tgt_stages > 0 && elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      const uint32_t tgt0 = smem_u32(smem + p.off_tgt);
      Pipe gp{0, 0};
      bool ok = true;
      for (unsigned int tile = tile0; tile < n_tiles && ok; tile += tstep) {
        unsigned int r = tile;
        const int tw = (int)(r % (unsigned)p.tiles_w); r /= (unsigned)p.tiles_w;
        const int th = (int)(r % (unsigned)p.tiles_h); r /= (unsigned)p.tiles_h;
        const int dd = (int)(r % (unsigned)p.d); r /= (unsigned)p.d;
        const int nn = (int)r;
        for (int c0 = 0; c0 < p.c2; c0 += 32) {
          if (!mbar_wait<32>(BAR(B_GE + gp.stage), gp.phase ^ 1u, abort_flag)) { ok = false; break; }
          mbar_expect_tx(BAR(B_GF + gp.stage), TC_TGT_BYTES);
          tma_load_5d(tgt0 + (uint32_t)gp.stage * TC_TGT_BYTES, &tmap, tw * TC_TILE_W, th * TC_TILE_H, dd, p.c2_off + c0, nn,
                      BAR(B_GF + gp.stage));
          gp.advance(p.n_tgt_stages);
        }
      }
    }
    __syncwarp();
  }

  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
  if (!p.target || p.add_out) return;
  // deterministic reduction: block tree -> per-CTA partial -> last CTA folds in index order
  const double bsum = block_sum(err_acc, red_scratch);
  if (threadIdx.x == 0) {
    p.ws_partial[blockIdx.x] = bsum;
    __threadfence();
    is_last = (atomicAdd(p.ws_done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += TC_THREADS) t += ((volatile double*)p.ws_partial)[b];
    t = block_sum(t, red_scratch);
    if (threadIdx.x == 0) {
      const double prev = p.sse_accumulate ? *p.sse : 0.0;                                    // later channel chunks add up
      *p.sse = (*abort_flag != 0u) ? __longlong_as_double(0x7ff8000000000000ll) : prev + t;   // NaN on abort
      *p.ws_done = 0;
    }
  }
}

static bool tc_plan(const effq_geom& g, int code_dtype, TcParams& p) {
  if (code_dtype != CODE_BF16 && code_dtype != CODE_E4M3) return false;
  if (!((g.kd == 3 && g.kh == 3 && g.kw == 3 && g.pd == 1 && g.ph == 1 && g.pw == 1) ||
        (g.kd == 1 && g.kh == 1 && g.kw == 1 && g.pd == 0 && g.ph == 0 && g.pw == 0)))
    return false;
  if (g.sd != 1 || g.sh != 1 || g.sw != 1) return false;
  if (g.c2 % 16 != 0 || g.c2 < 16 || g.c2 > 256) return false;
  const TcLayout lay = tc_layout(g.c1, code_dtype);
  if (g.c1 <= 0 || g.c1 % lay.cg != 0) return false;
  if (lay.rp != 32 && lay.rp != 64 && lay.rp != 128) return false;   // 2, 4 or 8 chunks per row
  p.n = g.n; p.c1 = g.c1; p.c2 = g.c2; p.d = g.d; p.h = g.h; p.w = g.w;
  p.kd = g.kd; p.kh = g.kh; p.kw = g.kw; p.pd = g.pd; p.ph = g.ph; p.pw = g.pw;
  p.taps = g.kd * g.kh * g.kw;
  p.cg = lay.cg;
  p.n_groups = lay.groups;
  p.nch = lay.nch;
  p.eb = lay.eb;
  p.swz = lay.swz;
  p.rp = lay.rp;
  static const int dbg_env = [] { const char* v = getenv("EFFQ_TC_DEBUG"); return (v && *v) ? atoi(v) : 0; }();
  p.debug = dbg_env;
  p.hh = TC_TILE_H + g.kh - 1;
  p.wp = TC_TILE_W + g.kw - 1;
  // NB halo rows are NOT padded to the 8-row swizzle period: the hardware applies the XOR to the
  // absolute shared-memory address (verified on B200, profiles/r01_conv_layout.md), so a
  // descriptor may start at any row and 8-row groups may start at any phase; base_offset stays 0.
  p.hv = g.kd * p.hh * p.wp;
  p.tiles_h = (g.h + TC_TILE_H - 1) / TC_TILE_H;
  p.tiles_w = (g.w + TC_TILE_W - 1) / TC_TILE_W;
  p.n_tiles = (long long)g.n * g.d * p.tiles_h * p.tiles_w;
  p.halo_tx_bytes = (uint32_t)(p.hv * p.rp);
  p.halo_bytes = (p.halo_tx_bytes + 1023u) & ~1023u;
  p.wtile_bytes = (uint32_t)(p.rp * g.c2);
  p.c2_total = g.c2; p.c2_off = 0; p.sse_accumulate = 0; p.add_out = 0;
  p.wsrc_tile_bytes = p.wtile_bytes; p.wsrc_off = 0;
  p.wstage_bytes = (p.wtile_bytes + 1023u) & ~1023u;
  p.off_bias = 384;
  p.off_halo = (p.off_bias + (uint32_t)g.c2 * 4u + 1023u) & ~1023u;
  // 32-byte rows (C1 = 32 e4m3 / 16 bf16): everything is small, so TWO CTAs share an SM (each with
  // half the shared memory and <= 128 registers): their tile pipelines interleave and hide the
  // per-tile serialisation of the single MMA-issuing thread.
  static const bool one_cta = [] { const char* v = getenv("EFFQ_TC_ONE_CTA"); return v && *v == '1'; }();
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * g.c2)) cols <<= 1;
  p.tmem_cols = cols;
  if (cols > 512) return false;
  p.ctas_per_sm = (p.rp == 32 && !one_cta && 2 * cols <= 512) ? 2 : 1;   // both CTAs' accumulators must fit TMEM
  const uint32_t budget = p.ctas_per_sm == 2 ? 110u * 1024u : 224u * 1024u;
  const uint32_t w_all = p.wstage_bytes * (uint32_t)(p.n_groups * p.taps);
  // Shared-memory plan: halo ring (>= 2 stages), weights (resident, else a ring of >= 2 tiles), and --
  // when the target can be a TMA box -- a target ring of >= 2 stages; leftover space deepens the rings.
  const bool tgt_ok = (g.w % 4 == 0) && (g.c2 % 32 == 0);
  p.n_halo_stages = 2;
  p.n_tgt_stages = 0;
  if (p.off_halo + 2u * p.halo_bytes + w_all <= budget) {
    p.w_resident = 1;
    p.n_w_stages = 0;
    uint32_t left = budget - (p.off_halo + w_all + 2u * p.halo_bytes);
    if (tgt_ok && left >= 2u * TC_TGT_BYTES) { p.n_tgt_stages = 2; left -= 2u * TC_TGT_BYTES; }
    for (int round = 0; round < 2; ++round) {            // alternate: one more halo stage, one more target stage
      if (p.n_halo_stages < 4 && left >= p.halo_bytes) { ++p.n_halo_stages; left -= p.halo_bytes; }
      if (p.n_tgt_stages > 0 && p.n_tgt_stages < 4 && left >= TC_TGT_BYTES) { ++p.n_tgt_stages; left -= TC_TGT_BYTES; }
    }
    p.off_tgt = p.off_halo + (uint32_t)p.n_halo_stages * p.halo_bytes;
    p.off_w = p.off_tgt + (uint32_t)p.n_tgt_stages * TC_TGT_BYTES;
  } else {
    p.w_resident = 0;
    p.off_tgt = p.off_halo + 2u * p.halo_bytes;       // the loader pipeline needs >= 2 halo stages
    if (p.off_tgt + 2u * p.wstage_bytes > budget) return false;
    if (tgt_ok && p.off_tgt + 2u * TC_TGT_BYTES + 3u * p.wstage_bytes <= budget) p.n_tgt_stages = 2;
    p.off_w = p.off_tgt + (uint32_t)p.n_tgt_stages * TC_TGT_BYTES;
    int ws = (int)((budget - p.off_w) / p.wstage_bytes);
    p.n_w_stages = ws > 8 ? 8 : ws;
  }
  return true;
}

static uint32_t tc_smem_bytes(const TcParams& p) {
  const uint32_t w = p.w_resident ? p.wstage_bytes * (uint32_t)(p.n_groups * p.taps)
                                  : p.wstage_bytes * (uint32_t)p.n_w_stages;
  return p.off_w + w + 1024u;      // + slack for the manual 1024 B alignment of the base
}

template <int KS, int KK, bool WRES, bool FP8>
static int tc_launch(const TcParams& p, const CUtensorMap& xmap, const CUtensorMap& tmap, uint32_t smem, unsigned ctas, cudaStream_t s) {
  static uint32_t configured = 0;
  if (smem > configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<KS, KK, WRES, FP8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  conv3d_tc_kernel<KS, KK, WRES, FP8><<<ctas, TC_THREADS, smem, s>>>(p, xmap, tmap);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

template <int KS, int KK>
static int tc_dispatch_w(const TcParams& p, const CUtensorMap& m, const CUtensorMap& t, uint32_t smem, unsigned ctas, cudaStream_t s) {
  if (p.eb == 1)
    return p.w_resident ? tc_launch<KS, KK, true, true>(p, m, t, smem, ctas, s) : tc_launch<KS, KK, false, true>(p, m, t, smem, ctas, s);
  return p.w_resident ? tc_launch<KS, KK, true, false>(p, m, t, smem, ctas, s) : tc_launch<KS, KK, false, false>(p, m, t, smem, ctas, s);
}

static int tc_dispatch(const TcParams& p, const CUtensorMap& m, const CUtensorMap& t, int ks, int kk, uint32_t smem, unsigned ctas,
                       cudaStream_t s) {
  if (ks == 3) {
    if (kk == 1) return tc_dispatch_w<3, 1>(p, m, t, smem, ctas, s);
    if (kk == 2) return tc_dispatch_w<3, 2>(p, m, t, smem, ctas, s);
    return tc_dispatch_w<3, 4>(p, m, t, smem, ctas, s);
  }
  if (kk == 1) return tc_dispatch_w<1, 1>(p, m, t, smem, ctas, s);
  if (kk == 2) return tc_dispatch_w<1, 2>(p, m, t, smem, ctas, s);
  return tc_dispatch_w<1, 4>(p, m, t, smem, ctas, s);
}

// Tensor map of the NDHWC code tensor, dims innermost first (C, W, H, D, N), box = one halo block.
// Tensor map of the NCDHW fp32 target, dims innermost first (W, H, D, C2, N), box = 32 channels of
// one output tile, dense in shared memory as [channel][h][w].
static int tc_make_tmap(const TcParams& p, const float* target, CUtensorMap* map) {
  EncodeTiledFn encode = tc_encoder();
  if (!encode) return 2;
  const cuuint64_t dims[5] = {(cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.d, (cuuint64_t)p.c2_total, (cuuint64_t)p.n};
  const cuuint64_t strides[4] = {dims[0] * 4, dims[0] * dims[1] * 4, dims[0] * dims[1] * dims[2] * 4,
                                 dims[0] * dims[1] * dims[2] * dims[3] * 4};
  const cuuint32_t box[5] = {(cuuint32_t)TC_TILE_W, (cuuint32_t)TC_TILE_H, 1u, 32u, 1u};
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(target), dims, strides, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("effq_conv3d_tc: cuTensorMapEncodeTiled(target) failed (%d)", (int)rc); return 2; }
  return 0;
}

static int tc_make_xmap(const TcParams& p, const void* xcodes, CUtensorMap* map) {
  EncodeTiledFn encode = tc_encoder();
  if (!encode) return 2;
  const cuuint64_t eb = (cuuint64_t)p.eb;
  const cuuint64_t dims[5] = {(cuuint64_t)p.c1, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.d, (cuuint64_t)p.n};
  const cuuint64_t strides[4] = {dims[0] * eb, dims[0] * dims[1] * eb, dims[0] * dims[1] * dims[2] * eb,
                                 dims[0] * dims[1] * dims[2] * dims[3] * eb};
  const cuuint32_t box[5] = {(cuuint32_t)p.cg, (cuuint32_t)p.wp, (cuuint32_t)p.hh, (cuuint32_t)p.kd, 1u};
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  const CUtensorMapSwizzle swz = p.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : (p.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const CUresult rc = encode(map, p.eb == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                             const_cast<void*>(xcodes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("effq_conv3d_tc: cuTensorMapEncodeTiled failed (%d)", (int)rc); return 2; }
  return 0;
}

}  // namespace effq

// C2 > 256 (one tcgen05.mma covers N <= 256 and the accumulator is double-buffered in 512 TMEM columns):
// chunks of 256 output channels, one launch each.
static bool tc_chunked(const effq_geom& g) { return g.c2 > 256 && g.c2 % 256 == 0 && g.c2 <= 1024; }

extern "C" int effq_conv3d_tc_supported(const effq_geom* g, int32_t code_dtype) {
  effq::TcParams p;
  if (!g) return 0;
  effq_geom gg = *g;
  if (tc_chunked(gg)) gg.c2 = 256;
  return effq::tc_plan(gg, code_dtype, p) ? 1 : 0;
}

extern "C" int64_t effq_conv3d_tc_workspace(const effq_geom* g) {
  (void)g;
  return 16 + 8 * 1024;
}

static int conv3d_tc_impl(const void* xcodes, const void* wcodes, int32_t code_dtype, const float* bias,
                          const float* conv_scale, const float* scale_vec, const effq_geom* g, float* out,
                          const float* target, const float* att, double* sse, void* workspace, void* stream,
                          int add_out = 0);

extern "C" int effq_conv3d_tc(const void* xcodes, const void* wcodes, int32_t code_dtype, const float* bias,
                              const float* conv_scale, const effq_geom* g, float* out, const float* target, const float* att, double* sse,
                              void* workspace, void* stream) {
  return conv3d_tc_impl(xcodes, wcodes, code_dtype, bias, conv_scale, nullptr, g, out, target, att, sse, workspace, stream);
}

// The same conv with one scale per OUTPUT CHANNEL (scale_vec[c2], device): out[c] = scale_vec[c] * (integer) + bias[c].
extern "C" int effq_conv3d_tc_pc(const void* xcodes, const void* wcodes, int32_t code_dtype, const float* bias,
                                 const float* scale_vec, const effq_geom* g, float* out, const float* target,
                                 const float* att, double* sse, void* workspace, void* stream) {
  return conv3d_tc_impl(xcodes, wcodes, code_dtype, bias, scale_vec, scale_vec, g, out, target, att, sse, workspace, stream);
}

// out += scale * conv(xcodes, wcodes) (+ bias): the accumulating form behind the fp32-accurate FP conv, which sums six
// products of fixed-point digit planes (efficientq_b200/ops.py conv3d_fp); `out` must hold the partial result of the earlier launches.
extern "C" int effq_conv3d_tc_acc(const void* xcodes, const void* wcodes, int32_t code_dtype, const float* bias,
                                  const float* conv_scale, const float* scale_vec, const effq_geom* g, float* out_inout,
                                  void* workspace, void* stream) {
  EFFQ_CHECK_ARG(out_inout && (conv_scale || scale_vec), "null pointer");
  return conv3d_tc_impl(xcodes, wcodes, code_dtype, bias, conv_scale ? conv_scale : scale_vec, scale_vec, g, out_inout,
                        out_inout, nullptr, nullptr, workspace, stream, 1);
}

static int conv3d_tc_impl(const void* xcodes, const void* wcodes, int32_t code_dtype, const float* bias,
                          const float* conv_scale, const float* scale_vec, const effq_geom* g, float* out,
                          const float* target, const float* att, double* sse, void* workspace, void* stream, int add_out) {
  using namespace effq;
  EFFQ_CHECK_ARG(xcodes && wcodes && conv_scale && g && workspace, "null pointer");
  EFFQ_CHECK_ARG(out || target, "nothing to compute");
  EFFQ_CHECK_ARG(!target || sse || add_out, "sse required with target");
  EFFQ_CHECK_ARG(((uintptr_t)xcodes & 15) == 0 && ((uintptr_t)wcodes & 15) == 0, "operands must be 16B aligned");
  const int n_chunks = tc_chunked(*g) ? g->c2 / 256 : 1;
  effq_geom gg = *g;
  if (n_chunks > 1) gg.c2 = 256;
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
  TcParams p;
  EFFQ_CHECK_ARG(tc_plan(gg, code_dtype, p), "geometry / code type not supported by the tcgen05 path");
  p.c2_total = g->c2;
  p.c2_off = chunk * gg.c2;
  p.sse_accumulate = chunk > 0 ? 1 : 0;
  p.add_out = add_out;
  p.wsrc_tile_bytes = (uint32_t)(p.rp * g->c2);
  p.wsrc_off = (uint32_t)(p.rp * p.c2_off);
  EFFQ_CHECK_ARG(p.n_tiles < (1ll << 31), "too many tiles");
  p.xq = (const uint8_t*)xcodes;
  p.wq = (const uint8_t*)wcodes;
  p.bias = bias;
  p.conv_scale = conv_scale;
  p.scale_vec = scale_vec;
  p.out = out;
  p.target = target;
  p.att = att;
  p.sse = sse;
  p.ws_done = (unsigned int*)workspace;
  p.ws_partial = (double*)((char*)workspace + 16);
  // bring-up: EFFQ_TC_DEBUG bit 8 -> CTA 0 writes a per-tile clock timeline behind the partials
  // (the caller must then pass a workspace of at least 16 + 8 KB + 16 KB)
  p.dbg = (p.debug & 8) ? (unsigned long long*)((char*)workspace + 16 + 8 * 1024) : nullptr;
  const uint32_t smem = tc_smem_bytes(p);
  const long long max_ctas = (long long)sm_count() * p.ctas_per_sm;
  const long long ctas = p.n_tiles < max_ctas ? p.n_tiles : max_ctas;
  const int kk = p.rp / 32;
  alignas(64) CUtensorMap xmap, tmap;
  if (int rc = tc_make_xmap(p, xcodes, &xmap)) return rc;
  static const bool tgt_tma_off = [] { const char* v = getenv("EFFQ_TC_TARGET_TMA"); return v && *v == '0'; }();
  if (!target || ((uintptr_t)target & 15) != 0 || tgt_tma_off) p.n_tgt_stages = 0;     // epilogue loads the target itself
  if (p.n_tgt_stages > 0) {
    if (int rc = tc_make_tmap(p, target, &tmap)) return rc;
  } else {
    tmap = xmap;                                       // never dereferenced
  }
  if (int rc = tc_dispatch(p, xmap, tmap, g->kd, kk, smem, (unsigned)ctas, (cudaStream_t)stream)) return rc;
  }
  return 0;
}
