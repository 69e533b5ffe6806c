// tcgen05 implicit-GEMM for the im2col Gram matrix of a 3x3x3 / stride-1 / pad-1 layer with
// quantised activations:      S[i][j] = sum_v att_v * xhat_i(v) * xhat_j(v),   i, j < K = 27*C1
// (the K x K block of A0 = 2*S of reference src/models/solver.py:282-314; the bias row/column
// and B0 come from the generic kernel in gram_simt.cu).  The im2col matrix is never formed.
//
// GEMM view: M = 128 rows i, N = 256 rows j, reduction over output voxels (K dim = voxels).
// Row index order inside this kernel is tap-major, i = tap*C1 + c, so that 8 consecutive rows
// are the 8 channels of one 16-byte NDHWC vector; results are scattered to the reference's
// (c, tap) order in the epilogue.
//
//   * operand tiles are built by 8 "builder" warps straight from global memory (NDHWC bf16
//     codes, 128-bit loads through L1, zero for padding): each 16-byte vector is the codes of
//     8 consecutive rows at one voxel, i.e. one chunk of an MN-major, 128B-swizzled UMMA tile
//     [64-row block][voxel][64 rows].  L1 serves the 27-fold tap reuse of a voxel block.
//   * the left operand carries the attention weight: p = att_v * code (fp32), split into
//     bf16 hi + bf16 lo (two MMAs, error <= 2^-17 per term, exact for the reference's integer
//     masks); the rows of B0 (att_v * y, arbitrary fp32) and the bias row use a three-term
//     split (24 bits: fp32-exact); the right operand is the raw integer code (exact).
//   * when C1 is a multiple of 32 the right operand is not built by threads at all: a 64-row
//     (C1 % 64 == 0, SWIZZLE_128B) or 32-row (C1 = 32, SWIZZLE_64B) block of it is exactly one
//     tap's channel slice of an 8x8 voxel block, i.e. ONE 5-D TMA box of the NDHWC codes at
//     tap-shifted coordinates (out-of-range coordinates arrive as zeros = the padding); the
//     constant blocks behind row K (the ones column, zero fill) are written once per item.  The
//     builders then only gather, weight and split the left operand (4 chunks instead of 12).
//   * one thread issues tcgen05.mma M=128,N=256,K=16 (both operands MN-major) -- the only
//     shape at which the SS tensor pipe is not starved by the A-operand read (B200: one A row
//     per cycle, profiles/r01_conv_layout.md) -- into a 128x256 fp32 TMEM accumulator.
//   * a work item = (128-row block, 256-row block, voxel range of <= 32768 voxels): fp32
//     accumulation stays exact / short; the epilogue adds the tile into an fp64 workspace
//     with atomics (deterministic to fp32 after the final rounding).
#include "common.cuh"
#include "tc_layout.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace effq {

constexpr int GT_BUILDERS = 512;           // 16 warps
constexpr int GT_RCS = GT_BUILDERS / 64;   // builder threads per voxel (each takes every GT_RCS-th chunk)
constexpr int GT_ZS = 16 / GT_RCS;         // left-operand chunks per thread and stage (128 rows = 16 chunks per voxel)
constexpr int GT_PS = 32 / GT_RCS;         // right-operand chunks per thread and stage (thread-built path)
constexpr int GT_THREADS = GT_BUILDERS + 64;        // + MMA warp + right-operand TMA warp
constexpr int GT_STAGES = 2;
constexpr int GT_KV = 64;                  // voxels per stage: 8 h-rows x 8 w
constexpr int GT_BM = 128, GT_BN = 256;
constexpr uint32_t GT_ZBYTES = 2 * GT_KV * 128;      // 128 rows = 2 blocks of 64 rows: 16 KB
constexpr uint32_t GT_PBYTES = 4 * GT_KV * 128;      // 256 rows: 32 KB
constexpr uint32_t GT_STAGE_BYTES = 3 * GT_ZBYTES + GT_PBYTES;   // Zhi, Zlo, Zlo2, P = 80 KB
constexpr unsigned int GT_SPIN_LIMIT = 1u << 26;

struct GtParams {
  const __nv_bfloat16* xq;     // NDHWC codes
  const float* att;            // N,D,H,W or null
  const float* y;              // NCDHW fp32 target (rows of B0) or null
  double* acc;                 // fp64 workspace [(K'+C2) x K'], leading dimension ld (reference row order)
  unsigned int* flags;         // [0] abort
  int n, c1, d, h, w;
  int k;                       // 27 * c1
  int c2, has_bias;            // B0 rows / ones row+column
  int mx0;                     // first extra left row (ones row, then y channels from mx0 + 8), multiple of 128
  int ld;
  int mb_n, nb_n, splits;      // work decomposition
  int hb_h, hb_w;              // 8x8 voxel blocks per plane (ceil)
  long long hb_total;          // n*d*hb_h*hb_w
  long long hb_per_split;
  int single;                  // att * code exact in bf16: one term for the weighted codes (no lo plane / MMA)
  double* acc2;                // optional second workspace: the UNWEIGHTED Gram of the codes (and their column sums),
                               // accumulated by the same pass in a second TMEM accumulator (raw codes as left operand)
  int mb_begin;                // first row block to compute (rows-only passes start at the extra-row block)
  int p_tma;                   // right operand by TMA
  int pblk;                    // rows per right-operand block (64: SWIZZLE_128B, 32: SWIZZLE_64B)
};

__device__ __forceinline__ uint32_t gt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void gt_mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool gt_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
template <int SLEEP_NS = 0>
__device__ __forceinline__ bool gt_mbar_wait(uint32_t bar, uint32_t parity, volatile unsigned int* abort_flag) {
  unsigned int spins = 0;
  while (!gt_mbar_try(bar, parity)) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);       // do not steal issue slots from the MMA warp
    if ((++spins & 0x3ffu) == 0) {
      if (*abort_flag != 0u) return false;
      if (spins > GT_SPIN_LIMIT) { *abort_flag = 1u; __threadfence(); return false; }
    }
  }
  return true;
}
__device__ __forceinline__ void gt_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gt_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool gt_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void gt_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

// MN-major, 128B-swizzled UMMA descriptor: 64 rows contiguous (128 B), K (voxel) rows at
// 128 B, 8-voxel groups sbo apart, 64-row blocks lbo apart.
__device__ __forceinline__ uint64_t gt_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

// A slot = the 16-byte chunk (8 consecutive rows at this thread's voxel) a builder thread fills.
//   kind 1: 8 channels of the tap-shifted activation vector   (rows < K)
//   kind 2: the "ones" chunk: row 0 = 1, rows 1-7 = 0         (bias row / column)
//   kind 3: 8 channels of the fp32 target y                   (left operand only: rows of B0)
//   kind 0: zero padding
// The patch x patch part of S is symmetric: a (128-row, 256-column) tile that lies entirely
// below the diagonal is not computed; the finalize pass mirrors it from its transpose.
__device__ __host__ __forceinline__ bool gt_tile_skipped(int mb, int nb, int mx0, int mb_begin = 0) {
  return mb < mb_begin || (mb * GT_BM < mx0 && mb * GT_BM >= (nb + 1) * GT_BN);
}

// Coordinates of an 8x8 voxel block; `next` walks blocks in linear order without divisions (the
// loaders and builders advance one block per pipeline stage).
struct BlockPos {
  int bw, bh, dd, nn;
  __device__ __forceinline__ void set(long long hb, const GtParams& p) {
    long long q = hb;
    bw = (int)(q % p.hb_w); q /= p.hb_w;
    bh = (int)(q % p.hb_h); q /= p.hb_h;
    dd = (int)(q % p.d); q /= p.d;
    nn = (int)q;
  }
  __device__ __forceinline__ void next(const GtParams& p) {
    if (++bw == p.hb_w) {
      bw = 0;
      if (++bh == p.hb_h) {
        bh = 0;
        if (++dd == p.d) { dd = 0; ++nn; }
      }
    }
  }
};

// item -> (voxel range z, column block nb, row block mb), voxel range FASTEST: concurrently running
// CTAs work on different voxel ranges of the same tile pair.  (The voxel-range-slowest order, which
// lets the ~20 tile pairs of a range share it through L2, cut the DRAM reads -- 23 GB per launch
// against 1.6 GB algorithmic, ncu -- but ran 16 % slower: DRAM is at 10 % of its peak here, and the
// mixed order puts the slow three-term row blocks and the fp64 epilogue atomics of all tiles in
// flight at once.  profiles/r01_layer_ncu.md)
__device__ __forceinline__ void gt_decode(long long item, const GtParams& p, int& z, int& nb, int& mb) {
  z = (int)(item % p.splits);
  const long long t = item / p.splits;
  nb = (int)(t % p.nb_n);
  mb = (int)(t / p.nb_n);
}

struct Slot {
  int rel;          // kind 1: element offset relative to the voxel's own vector; kind 3: first y channel
  int tap;          // (a) | (b << 2) | (c << 4) | kind << 6      (a,b,c in 0..2)
  uint32_t dst;     // byte offset inside the stage (Zhi / P region), swizzle applied
};

__device__ __forceinline__ Slot make_slot(int row0, int rc, int kvox, const GtParams& p, bool left) {
  Slot s;
  const int r = row0 + rc * 8;
  s.tap = 0;
  s.rel = 0;
  if (r < p.k) {
    const int tap = r / p.c1, c0 = r % p.c1;
    const int a = tap / 9, b = (tap / 3) % 3, c = tap % 3;
    s.rel = (((a - 1) * p.h + (b - 1)) * p.w + (c - 1)) * p.c1 + c0;
    s.tap = a | (b << 2) | (c << 4) | (1 << 6);
  } else if (left) {
    if (r == p.mx0 && p.has_bias) s.tap = 2 << 6;
    else if (p.y && r >= p.mx0 + 8 && r < p.mx0 + 8 + p.c2) { s.tap = 3 << 6; s.rel = r - p.mx0 - 8; }
  } else if (r == p.k && p.has_bias) {
    s.tap = 2 << 6;
  }
  const uint32_t blk = (uint32_t)(rc >> 3), ch = (uint32_t)(rc & 7);
  s.dst = blk * (GT_KV * 128u) + (uint32_t)kvox * 128u + ((ch ^ (uint32_t)(kvox & 7)) << 4);
  return s;
}

// two fp32 -> packed bf16x2 (first argument in the low half), round to nearest even
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}

// PF (right operand by TMA only): the builders load the codes and the attention weight of voxel block hb + 1 before
// they wait for, weight, split and store block hb -- without it every builder thread serialises the latency of its own
// L1/L2 gather with its store phase in every pipeline stage (ncu: tensor pipe 39 % active with the builders at a
// third of their issue slots, profiles/r02_layer_ncu.md).
template <bool PF>
__global__ void __launch_bounds__(GT_THREADS, 1)
gram_tc_kernel(const GtParams p, const __grid_constant__ CUtensorMap pmap) {
  extern __shared__ __align__(1024) uint8_t gsm_raw[];
  uint8_t* gsm = gsm_raw + ((1024u - (gt_smem_u32(gsm_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bars[2 * GT_STAGES + 2];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar0 = gt_smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(GT_STAGES + s); };
  const uint32_t TFULL = bar0 + 8u * (2 * GT_STAGES), TEMPTY = bar0 + 8u * (2 * GT_STAGES + 1);
  const uint32_t stage0 = gt_smem_u32(gsm);
  volatile unsigned int* abort_flag = p.flags;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // FULL: every builder thread + (TMA path) the arrive.expect_tx of the right-operand loader
    for (int s = 0; s < GT_STAGES; ++s) { gt_mbar_init(FULL(s), GT_BUILDERS + (p.p_tma ? 1 : 0)); gt_mbar_init(EMPTY(s), 1); }
    gt_mbar_init(TFULL, 1);
    gt_mbar_init(TEMPTY, GT_BUILDERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == GT_BUILDERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_smem_u32(&tmem_slot)),
                 "r"(p.acc2 ? 512u : 256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  const long long n_items = (long long)p.mb_n * p.nb_n * p.splits;

  if (warp == GT_BUILDERS / 32) {
    // ===== MMA issuer: warp-uniform control flow, one elected lane issues =====
    {
      uint32_t idesc = 0;
      idesc |= 1u << 4;                       // D = f32
      idesc |= 1u << 7;                       // A = bf16
      idesc |= 1u << 10;                      // B = bf16
      idesc |= 1u << 15;                      // A MN-major
      idesc |= 1u << 16;                      // B MN-major
      idesc |= (uint32_t)(GT_BN >> 3) << 17;
      idesc |= (uint32_t)(GT_BM >> 4) << 24;
      const uint64_t tmpl = gt_desc(0, GT_KV * 128u, 1024u);
      // right operand: 64-row blocks (128 B per voxel) or, for the TMA path at C1 = 32, 32-row blocks
      // (64 B per voxel, SWIZZLE_64B: 8-voxel groups 512 B apart, blocks GT_KV * 64 B apart)
      const bool p32 = p.p_tma && p.pblk == 32;
      uint64_t ptmpl = tmpl;
      if (p32) {
        ptmpl = gt_desc(0, GT_KV * 64u, 512u);
        ptmpl = (ptmpl & ~((uint64_t)7 << 61)) | ((uint64_t)4 << 61);      // SWIZZLE_64B
      }
      const uint32_t pk16 = p32 ? (16u * 64u >> 4) : (16u * 128u >> 4);     // 16 voxels of the right operand, in 16 B units
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      bool ok = true;
      for (long long item = blockIdx.x; item < n_items && ok; item += gridDim.x) {
        int z, nb_, mb_;
        gt_decode(item, p, z, nb_, mb_);
        if (gt_tile_skipped(mb_, nb_, p.mx0, p.mb_begin)) continue;
        const bool three = mb_ * GT_BM >= p.mx0;                                            // y / ones rows: 3-term split
        const bool two = three || !p.single;                                                // weighted codes: hi + lo unless exact
        const bool dual = p.acc2 != nullptr && !three;                                      // patch rows: + the unweighted Gram
        long long hb0 = (long long)z * p.hb_per_split;
        long long hb1 = hb0 + p.hb_per_split < p.hb_total ? hb0 + p.hb_per_split : p.hb_total;
        if (!gt_mbar_wait(TEMPTY, tphase ^ 1u, abort_flag)) { ok = false; break; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t accum = 0;
        for (long long hb = hb0; hb < hb1; ++hb) {
          if (!gt_mbar_wait(FULL(stage), phase, abort_flag)) { ok = false; break; }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t zhi = (stage0 + (uint32_t)stage * GT_STAGE_BYTES) >> 4;
          const uint32_t zlo = zhi + (GT_ZBYTES >> 4), zl2 = zhi + ((2 * GT_ZBYTES) >> 4), pp = zhi + ((3 * GT_ZBYTES) >> 4);
          if (gt_elect_one()) {
#pragma unroll
            for (int ks = 0; ks < GT_KV / 16; ++ks) {
              const uint32_t koff = (uint32_t)ks * (16u * 128u >> 4);
              const uint64_t bd = ptmpl | (uint64_t)((pp + (uint32_t)ks * pk16) & 0x3fffu);
              gt_mma(tmem_base, tmpl | (uint64_t)((zhi + koff) & 0x3fffu), bd, idesc, ks == 0 ? accum : 1u);
              if (two) gt_mma(tmem_base, tmpl | (uint64_t)((zlo + koff) & 0x3fffu), bd, idesc, 1u);
              if (three) gt_mma(tmem_base, tmpl | (uint64_t)((zl2 + koff) & 0x3fffu), bd, idesc, 1u);
              // raw codes (in the plane the weighted split does not use: lo when single, else the third) x raw codes
              if (dual) gt_mma(tmem_base + 256u, tmpl | (uint64_t)(((p.single ? zlo : zl2) + koff) & 0x3fffu), bd, idesc,
                               ks == 0 ? accum : 1u);
            }
            gt_commit(EMPTY(stage));
            if (hb == hb1 - 1) gt_commit(TFULL);
          }
          __syncwarp();
          accum = 1;
          if (++stage == GT_STAGES) { stage = 0; phase ^= 1u; }
        }
        tphase ^= 1u;
      }
    }
  } else if (warp == GT_BUILDERS / 32 + 1) {
    // ===== right-operand loader: one TMA box per (tap, channel slice) block of the 256-row tile =====
    if (p.p_tma && elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&pmap)) : "memory");
      const uint32_t blk_bytes = (uint32_t)(GT_KV * p.pblk * 2);
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (long long item = blockIdx.x; item < n_items && ok; item += gridDim.x) {
        int z, nb, mb_;
        gt_decode(item, p, z, nb, mb_);
        if (gt_tile_skipped(mb_, nb, p.mx0, p.mb_begin)) continue;
        int n_live = (p.k - nb * GT_BN) / p.pblk;                    // blocks of this tile below row K
        n_live = n_live < 0 ? 0 : (n_live > GT_BN / p.pblk ? GT_BN / p.pblk : n_live);
        long long hb0 = (long long)z * p.hb_per_split;
        long long hb1 = hb0 + p.hb_per_split < p.hb_total ? hb0 + p.hb_per_split : p.hb_total;
        // per-item box table: channel offset and tap shift of every live block (no divisions per stage:
        // this single thread is on the latency path of every stage)
        int bc0[8], bsh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r0 = nb * GT_BN + j * p.pblk;
          const int tap = r0 / p.c1;
          bc0[j] = r0 % p.c1;
          bsh[j] = (tap / 9) | (((tap / 3) % 3) << 2) | ((tap % 3) << 4);
        }
        BlockPos bp;
        bp.set(hb0, p);
        for (long long hb = hb0; hb < hb1; ++hb) {
          if (!gt_mbar_wait<32>(EMPTY(stage), phase ^ 1u, abort_flag)) { ok = false; break; }
          const uint32_t pdst = stage0 + (uint32_t)stage * GT_STAGE_BYTES + 3u * GT_ZBYTES;
          mbar_expect_tx(FULL(stage), (uint32_t)n_live * blk_bytes);
          const int w0 = bp.bw * 8 - 1, h0 = bp.bh * 8 - 1, d0 = bp.dd - 1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < n_live)
              tma_load_5d(pdst + (uint32_t)j * blk_bytes, &pmap, bc0[j], w0 + ((bsh[j] >> 4) & 3), h0 + ((bsh[j] >> 2) & 3),
                          d0 + (bsh[j] & 3), bp.nn, FULL(stage));
          }
          bp.next(p);
          if (++stage == GT_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===== builders (also the epilogue) =====
    const int t = threadIdx.x;                       // 0..255
    const int kvox = t & (GT_KV - 1);                // this thread's voxel inside every stage
    const int vy = kvox >> 3, vx = kvox & 7;
    const int rc_base = t >> 6;                      // 0..GT_RCS-1
    int stage = 0;
    uint32_t phase = 0, tphase = 0;
    bool ok = true;
    const long long plane = (long long)p.h * p.w;
    for (long long item = blockIdx.x; item < n_items && ok; item += gridDim.x) {
      int z, nb, mb;
      gt_decode(item, p, z, nb, mb);
      if (gt_tile_skipped(mb, nb, p.mx0, p.mb_begin)) continue;
      Slot zs[GT_ZS], ps[GT_PS];
#pragma unroll
      for (int i = 0; i < GT_ZS; ++i) zs[i] = make_slot(mb * GT_BM, rc_base + GT_RCS * i, kvox, p, true);
#pragma unroll
      for (int i = 0; i < GT_PS; ++i) ps[i] = make_slot(nb * GT_BN, rc_base + GT_RCS * i, kvox, p, false);
      const bool p_tma = p.p_tma != 0;
      const bool single = p.single != 0;
      const bool mb_patch = mb * GT_BM < p.mx0;       // a row block of weighted codes only (no y / ones rows): lo plane unused when exact
      long long hb0 = (long long)z * p.hb_per_split;
      long long hb1 = hb0 + p.hb_per_split < p.hb_total ? hb0 + p.hb_per_split : p.hb_total;
      BlockPos bp;
      bp.set(hb0, p);
      const long long chan = (long long)p.d * plane;
      // this thread's voxel of a block: position, liveness, attention weight
      struct VoxPos { int dd, nn, vh, vw; bool vlive; long long vidx; float aw; };
      auto make_pos = [&](const BlockPos& b) {
        VoxPos q;
        q.dd = b.dd; q.nn = b.nn;
        q.vh = b.bh * 8 + vy; q.vw = b.bw * 8 + vx;
        q.vlive = q.vh < p.h && q.vw < p.w;
        q.vidx = ((long long)q.nn * p.d + q.dd) * plane + (long long)q.vh * p.w + q.vw;
        q.aw = q.vlive ? (p.att ? __ldg(p.att + q.vidx) : 1.f) : 0.f;
        return q;
      };
      // the left operand's code chunks (kind 1) of that voxel: raw bf16 codes, zero for padding
      auto gather_z = [&](const VoxPos& q, uint4 (&zq)[GT_ZS]) {
        const __nv_bfloat16* vptr = p.xq + q.vidx * p.c1;
#pragma unroll
        for (int i = 0; i < GT_ZS; ++i) {
          const int tp = zs[i].tap;
          zq[i] = make_uint4(0, 0, 0, 0);
          if ((tp >> 6) == 1) {
            const int gd = q.dd + (tp & 3) - 1, gh = q.vh + ((tp >> 2) & 3) - 1, gw = q.vw + ((tp >> 4) & 3) - 1;
            if (q.vlive && (unsigned)gd < (unsigned)p.d && (unsigned)gh < (unsigned)p.h && (unsigned)gw < (unsigned)p.w)
              zq[i] = __ldg(reinterpret_cast<const uint4*>(vptr + zs[i].rel));
          }
        }
      };
      VoxPos nxt;
      uint4 zn[GT_ZS];
      if (PF) {
        nxt = make_pos(bp);
        bp.next(p);
        gather_z(nxt, zn);
      }
      for (long long hb = hb0; hb < hb1; ++hb) {
        VoxPos cur;
        uint4 zv[GT_ZS], pv[GT_PS];           // kind 1: raw bf16 codes; kind 3 (left only): handled below
        if (PF) {
          cur = nxt;
#pragma unroll
          for (int i = 0; i < GT_ZS; ++i) zv[i] = zn[i];
          if (hb + 1 < hb1) {                  // next block's loads go out before this block's wait + store phase
            nxt = make_pos(bp);
            bp.next(p);
            gather_z(nxt, zn);
          }
        } else {
          cur = make_pos(bp);
          bp.next(p);
          gather_z(cur, zv);
        }
        const int dd = cur.dd, nn = cur.nn, vh = cur.vh, vw = cur.vw;
        const bool vlive = cur.vlive;
        const long long vidx = cur.vidx;
        const __nv_bfloat16* vptr = p.xq + vidx * p.c1;
        const float aw = cur.aw;
        float yv[GT_ZS][8];               // kind 3: the 8 target channels of this voxel
#pragma unroll
        for (int i = 0; i < GT_ZS; ++i) {
          if ((zs[i].tap >> 6) == 3) {
            const float* yp = p.y + ((long long)nn * p.c2 + zs[i].rel) * chan + (vidx - (long long)nn * chan);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              yv[i][e] = (vlive && zs[i].rel + e < p.c2) ? __ldg(yp + (long long)e * chan) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < GT_PS; ++i) {
          const int tp = ps[i].tap, kind = tp >> 6;
          pv[i] = make_uint4(0, 0, 0, 0);
          if (PF || p_tma) continue;                  // the TMA warp delivers the right operand
          if (kind == 1) {
            const int gd = dd + (tp & 3) - 1, gh = vh + ((tp >> 2) & 3) - 1, gw = vw + ((tp >> 4) & 3) - 1;
            if (vlive && (unsigned)gd < (unsigned)p.d && (unsigned)gh < (unsigned)p.h && (unsigned)gw < (unsigned)p.w)
              pv[i] = __ldg(reinterpret_cast<const uint4*>(vptr + ps[i].rel));
          } else if (kind == 2 && vlive) {
            pv[i].x = 0x3f80u;        // bf16 1.0 in row 0 of the chunk
          }
        }
        if (!gt_mbar_wait<64>(EMPTY(stage), phase ^ 1u, abort_flag)) { ok = false; break; }
        uint8_t* sbase = gsm + (size_t)stage * GT_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < GT_ZS; ++i) {
          const int kind = zs[i].tap >> 6;
          float pr[8];                                   // att-weighted left-operand values
          if (kind == 3) {
#pragma unroll
            for (int e = 0; e < 8; ++e) pr[e] = yv[i][e] * aw;
          } else {
            const uint32_t wds[4] = {zv[i].x, zv[i].y, zv[i].z, zv[i].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              pr[2 * e] = __uint_as_float(wds[e] << 16) * aw;
              pr[2 * e + 1] = __uint_as_float(wds[e] & 0xffff0000u) * aw;
            }
            if (kind == 2) {
              pr[0] = aw;                                // ones row: att_v * 1 (aw is 0 for dead voxels)
              if (p.acc2) pr[1] = vlive ? 1.f : 0.f;     // + an UNWEIGHTED ones row: column sums of the codes, voxel count
            }
          }
          uint32_t hi[4], lo[4], l2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = pr[2 * e], p1 = pr[2 * e + 1];
            hi[e] = pack2(p0, p1);                                        // one cvt.rn.bf16x2.f32
            const float r0 = p0 - __uint_as_float(hi[e] << 16), r1 = p1 - __uint_as_float(hi[e] & 0xffff0000u);   // exact
            lo[e] = 0;
            l2[e] = 0;
            if (kind == 1 && single) continue;                            // att * code is exact in bf16: hi is the value
            lo[e] = pack2(r0, r1);
            if (kind >= 2)                                                // third term: 24 bits in total
              l2[e] = pack2(r0 - __uint_as_float(lo[e] << 16), r1 - __uint_as_float(lo[e] & 0xffff0000u));
          }
          *reinterpret_cast<uint4*>(sbase + zs[i].dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (!(single && mb_patch)) *reinterpret_cast<uint4*>(sbase + GT_ZBYTES + zs[i].dst) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          if (kind >= 2) *reinterpret_cast<uint4*>(sbase + 2 * GT_ZBYTES + zs[i].dst) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
          // unweighted pass riding along: the raw codes of a patch row block (zero for padding / rows behind K)
          if (p.acc2 && mb_patch)
            *reinterpret_cast<uint4*>(sbase + (single ? GT_ZBYTES : 2 * GT_ZBYTES) + zs[i].dst) = kind == 1 ? zv[i] : make_uint4(0, 0, 0, 0);
        }
        if (!p_tma) {
#pragma unroll
          for (int i = 0; i < GT_PS; ++i) *reinterpret_cast<uint4*>(sbase + 3 * GT_ZBYTES + ps[i].dst) = pv[i];
        } else if (hb - hb0 < GT_STAGES) {
          // first use of this stage buffer in the item: the constant blocks behind row K (TMA never
          // writes them): the ones column (row K; dead voxels are zeroed by the left operand) and zeros
          const int rowb = p.pblk * 2;                               // bytes per voxel row of a block
          const int cpb = p.pblk / 8;                                // 16-byte chunks per voxel row
          const int first = (p.k - nb * GT_BN + p.pblk - 1) / p.pblk;
          for (int j = first < 0 ? 0 : first; j < GT_BN / p.pblk; ++j) {
            for (int ch = rc_base; ch < cpb; ch += GT_RCS) {
              const uint32_t x = p.pblk == 64 ? (uint32_t)(kvox & 7) : (uint32_t)((kvox >> 1) & 3);
              const uint32_t dst = (uint32_t)j * (uint32_t)(GT_KV * rowb) + (uint32_t)kvox * rowb + (((uint32_t)ch ^ x) << 4);
              uint4 val = make_uint4(0, 0, 0, 0);
              if (p.has_bias && nb * GT_BN + j * p.pblk == p.k && ch == 0) val.x = 0x3f80u;   // bf16 1.0 in row K
              *reinterpret_cast<uint4*>(sbase + 3 * GT_ZBYTES + dst) = val;
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        gt_mbar_arrive(FULL(stage));
        if (++stage == GT_STAGES) { stage = 0; phase ^= 1u; }
      }
      if (!ok) break;
      // ---- epilogue of the item: TMEM tile -> fp64 workspace (reference row order) ----
      if (!gt_mbar_wait<64>(TFULL, tphase, abort_flag)) { ok = false; break; }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tphase ^= 1u;
      const int qd = warp & 3, half = warp >> 2;        // TMEM lane quadrant, column half (builder warps 0..7)
      const int ri = mb * GT_BM + qd * 32 + lane;        // left row index (tap-major patch rows, then extras)
      const int kp = p.k + p.has_bias;
      int ref_i = -1;                                     // destination row in the workspace
      if (ri < p.k) ref_i = (ri % p.c1) * 27 + ri / p.c1;
      else if (ri == p.mx0 && p.has_bias) ref_i = p.k;                                   // bias row of A0
      else if (p.y && ri >= p.mx0 + 8 && ri < p.mx0 + 8 + p.c2) ref_i = kp + (ri - p.mx0 - 8);   // B0 rows
      const bool row_ok = ref_i >= 0;
      // unweighted statistics riding along (p.acc2): the patch rows of the second accumulator, and -- from the FIRST
      // accumulator -- the unweighted ones row (row mx0 + 1 of the extra block) as the bias row
      const int ref_i2 = !p.acc2 ? -1 : (ri < p.k ? ref_i : ((ri == p.mx0 + 1 && p.has_bias) ? p.k : -1));
      constexpr int GT_EPI_COLS = GT_BN / (GT_BUILDERS / 128);      // columns per warp group of four quadrant warps
      const int n_acc = (p.acc2 && mb * GT_BM < p.mx0) ? 2 : 1;
      for (int a_i = 0; a_i < n_acc; ++a_i) {
        for (int c0 = half * GT_EPI_COLS; c0 < half * GT_EPI_COLS + GT_EPI_COLS; c0 += 32) {
          uint32_t v[32];
          gt_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(a_i * 256 + c0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // accumulator 0 -> weighted workspace (all rows) and, for the unweighted ones row, the second workspace;
          // accumulator 1 -> second workspace (patch rows)
          double* dst = a_i == 0 ? p.acc : p.acc2;
          const int dst_row = a_i == 0 ? ref_i : ref_i2;
          const bool extra_unw = a_i == 0 && ri == p.mx0 + 1 && ref_i2 >= 0;
          if (extra_unw) dst = p.acc2;
          const int row = extra_unw ? ref_i2 : dst_row;
          if (row >= 0 && (a_i == 1 || row_ok || extra_unw)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int cj = nb * GT_BN + c0 + j;
              const float val = __uint_as_float(v[j]);
              if (val != 0.f && (cj < p.k || (cj == p.k && p.has_bias))) {
                const int ref_j = cj < p.k ? (cj % p.c1) * 27 + cj / p.c1 : p.k;
                atomicAdd(dst + (long long)row * p.ld + ref_j, (double)val);
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      gt_mbar_arrive(TEMPTY);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == GT_BUILDERS / 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.acc2 ? 512u : 256u) : "memory");
  }
}

}  // namespace effq

extern "C" int effq_gram_tc_supported(const effq_geom* g) {
  if (!g) return 0;
  const bool k3 = g->kd == 3 && g->kh == 3 && g->kw == 3 && g->pd == 1 && g->ph == 1 && g->pw == 1 &&
                  g->sd == 1 && g->sh == 1 && g->sw == 1;
  return (k3 && g->c1 % 8 == 0 && g->c1 >= 8 && g->c1 <= 512) ? 1 : 0;
}

// Accumulates (+=) into acc64[ld * i + j], reference row order, NOT yet scaled by 2 or the code
// scale: rows [0,K) x cols [0,K'] the attention-weighted Gram of the codes (incl. the bias
// column), row K the bias row, rows K'.. the B0 rows (att*y against the codes) when y != NULL.
// acc64 must have been zeroed by the caller.
static int gram_tc_accumulate_impl(const void* xcodes_ndhwc_bf16, const float* att, const float* y,
                                   const effq_geom* g, int32_t has_bias, int32_t att_exact, double* acc64,
                                   int32_t ld, void* flags, double* acc64_unweighted, int32_t rows_only, void* stream);

extern "C" int effq_gram_tc_accumulate(const void* xcodes_ndhwc_bf16, const float* att, const float* y,
                                       const effq_geom* g, int32_t has_bias, int32_t att_exact, double* acc64,
                                       int32_t ld, void* flags, void* stream) {
  return gram_tc_accumulate_impl(xcodes_ndhwc_bf16, att, y, g, has_bias, att_exact, acc64, ld, flags, nullptr, 0, stream);
}

// The same pass with the UNWEIGHTED Gram of the codes accumulated beside the weighted one (acc64_unweighted, same
// layout; only its K' x K' block is filled), or -- rows_only -- restricted to the extra row block (ones row and the
// rows of y against the codes): the cheap second pass of the conv-free scoring (csrc/quadform.cu).
extern "C" int effq_gram_tc_accumulate2(const void* xcodes_ndhwc_bf16, const float* att, const float* y,
                                        const effq_geom* g, int32_t has_bias, int32_t att_exact, double* acc64,
                                        int32_t ld, void* flags, double* acc64_unweighted, int32_t rows_only,
                                        void* stream) {
  return gram_tc_accumulate_impl(xcodes_ndhwc_bf16, att, y, g, has_bias, att_exact, acc64, ld, flags, acc64_unweighted,
                                 rows_only, stream);
}

static int gram_tc_accumulate_impl(const void* xcodes_ndhwc_bf16, const float* att, const float* y,
                                   const effq_geom* g, int32_t has_bias, int32_t att_exact, double* acc64,
                                   int32_t ld, void* flags, double* acc64_unweighted, int32_t rows_only, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(!acc64_unweighted || has_bias, "the unweighted statistics need the bias (ones) row");
  EFFQ_CHECK_ARG(xcodes_ndhwc_bf16 && g && acc64 && flags, "null pointer");
  EFFQ_CHECK_ARG(effq_gram_tc_supported(g), "geometry not supported by the tcgen05 Gram kernel");
  EFFQ_CHECK_ARG(((uintptr_t)xcodes_ndhwc_bf16 & 15) == 0, "codes must be 16B aligned");
  GtParams p;
  p.xq = (const __nv_bfloat16*)xcodes_ndhwc_bf16;
  p.att = att;
  p.y = y;
  p.acc = acc64;
  p.flags = (unsigned int*)flags;
  p.n = g->n; p.c1 = g->c1; p.d = g->d; p.h = g->h; p.w = g->w;
  p.k = 27 * g->c1;
  p.c2 = g->c2;
  p.has_bias = has_bias ? 1 : 0;
  p.single = (att_exact || !att) ? 1 : 0;
  p.ld = ld;
  p.acc2 = acc64_unweighted;
  p.mx0 = (p.k + GT_BM - 1) / GT_BM * GT_BM;                      // extras start on a row-block boundary
  p.mb_begin = rows_only ? p.mx0 / GT_BM : 0;
  const int extra_rows = (has_bias || y) ? 8 + (y ? g->c2 : 0) : 0;
  p.mb_n = (p.mx0 + extra_rows + GT_BM - 1) / GT_BM;
  p.nb_n = (p.k + (has_bias ? 8 : 0) + GT_BN - 1) / GT_BN;
  p.hb_h = (g->h + 7) / 8;
  p.hb_w = (g->w + 7) / 8;
  p.hb_total = (long long)g->n * g->d * p.hb_h * p.hb_w;
  // <= 512 voxel blocks (32768 voxels) per item keeps the fp32 TMEM accumulation short; at
  // least ~4 waves of items so the persistent grid balances
  long long splits = (p.hb_total + 511) / 512;
  const long long tiles = (long long)p.mb_n * p.nb_n;
  const long long want = ((long long)sm_count() * 4 + tiles - 1) / tiles;
  if (splits < want) splits = want;
  if (splits > p.hb_total) splits = p.hb_total;
  if (splits < 1) splits = 1;
  p.hb_per_split = (p.hb_total + splits - 1) / splits;
  splits = (p.hb_total + p.hb_per_split - 1) / p.hb_per_split;
  p.splits = (int)splits;
  const size_t smem = (size_t)GT_STAGES * GT_STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(gram_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EFFQ_CUDA(cudaFuncSetAttribute(gram_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const long long items = tiles * splits;
  const int ctas = (int)(items < sm_count() ? items : sm_count());
  // right operand by TMA: a block = one tap's slice of 64 (or, at C1 = 32, all 32) channels
  static const bool no_ptma = [] { const char* v = getenv("EFFQ_GRAM_PTMA"); return v && *v == '0'; }();
  p.p_tma = (!no_ptma && (g->c1 == 32 || g->c1 % 64 == 0)) ? 1 : 0;
  p.pblk = g->c1 == 32 ? 32 : 64;
  alignas(64) CUtensorMap pmap;
  memset(&pmap, 0, sizeof(pmap));
  if (p.p_tma) {
    EncodeTiledFn encode = tc_encoder();
    if (!encode) return 2;
    const cuuint64_t dims[5] = {(cuuint64_t)p.c1, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.d, (cuuint64_t)p.n};
    const cuuint64_t strides[4] = {dims[0] * 2, dims[0] * dims[1] * 2, dims[0] * dims[1] * dims[2] * 2,
                                   dims[0] * dims[1] * dims[2] * dims[3] * 2};
    const cuuint32_t box[5] = {(cuuint32_t)p.pblk, 8u, 8u, 1u, 1u};
    const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    const CUresult rc = encode(&pmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xcodes_ndhwc_bf16), dims,
                               strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               p.pblk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("effq_gram_tc: cuTensorMapEncodeTiled failed (%d)", (int)rc); return 2; }
  }
  static const bool no_pf = [] { const char* v = getenv("EFFQ_GRAM_PREFETCH"); return v && *v == '0'; }();
  if (p.p_tma && !no_pf) gram_tc_kernel<true><<<ctas, GT_THREADS, smem, (cudaStream_t)stream>>>(p, pmap);
  else                   gram_tc_kernel<false><<<ctas, GT_THREADS, smem, (cudaStream_t)stream>>>(p, pmap);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
