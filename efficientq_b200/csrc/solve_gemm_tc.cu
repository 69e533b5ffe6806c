// Proximal-step GEMM on the tensor cores with fp32-class accuracy:   w* = B A^-1
// (reference src/models/solver.py:331-342 solves A x = B^T with an fp32 LU in every one of the
// 200 iterations; A takes 5 values per layer, so A^-1 is formed once per value -- layer_engine.py --
// and the per-iteration work is this C2 x K' x K' product).
//
// fp32 operands are split into three bf16 terms  x = x0 + x1 + x2  (each residual is exact in
// fp32, 3 x 8 = 24 significand bits), and the six products of combined order <= 2
//     x0 y0 | x0 y1, x1 y0 | x0 y2, x1 y1, x2 y0
// are accumulated in the fp32 TMEM accumulator: the dropped terms are <= 2^-24 relative, i.e. the
// result is as accurate as an fp32 FMA chain (tests compare against fp64).  Six bf16 MMAs at the
// tensor rate replace one SIMT fp32 GEMM.
//
// D[m x n] = sum_{(i,j)} A_i[m x k] * B_j[n x k]^T, all planes bf16 K-major with a row pitch of
// ldk elements (multiple of 64 -> TMA-legal, K tail zero-padded by the producers).  A^-1 is
// symmetric, so its rows serve as the K-major "B" operand directly.
//   CTA tile 128 x 128, K step 64 (128-byte rows, SWIZZLE_128B), 2 stages x 96 KB, split-K over
//   gridDim.z with a deterministic second pass.  Warps: 0 TMA | 1 MMA | 2-5 epilogue.
//
// Accumulation chains are kept short on purpose.  The TMEM accumulator truncates (it does not round
// to nearest), so its error grows linearly with the number of MMAs chained into one accumulator; on
// the real, heavily cancelling systems (sum|b||a| / |result| ~ 10^2) one long chain was 5-25x less
// accurate than an fp32 SGEMM (tools/solve_accuracy.py).  Therefore: the leading products x0 y0
// rotate over THREE TMEM accumulators, the five correction products (2^-8 .. 2^-16 of the leading
// one, so their truncation is harmless) go to a fourth, the epilogue adds the four in registers
// (round to nearest), and split-K keeps every leading chain at <= 4 K steps (16 MMAs); the
// partials are folded in fp32 in a fixed order.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace effq {

constexpr int SG_THREADS = 192;
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 64;
constexpr uint32_t SG_PLANE_BYTES = SG_BM * SG_BK * 2;          // 16 KB: one operand plane of one stage
constexpr uint32_t SG_STAGE_BYTES = 6 * SG_PLANE_BYTES;         // A0 A1 A2 B0 B1 B2
constexpr int SG_STAGES = 2;
constexpr int SG_ACCS = 4;                                       // TMEM accumulators: 3 leading (rotated) + 1 corrections
constexpr int SG_MAIN = 3;
constexpr int SG_CHAIN = 4;                                      // K steps per leading accumulator before split-K takes over

struct SgParams {
  float* out;                  // [splits][m][ldo] when splits > 1 (partials), else the result [m][ldo]
  int m, n, k, ldo;
  int ksteps, splits;
  unsigned int* abort_flag;
  // general form (effq_gemm_tc_ex): result = alpha * A B^T + beta * c_in, applied here when splits == 1 and in the
  // fold pass otherwise; lower_only skips the tiles strictly above the diagonal (symmetric updates)
  const float* c_in;
  long long ldc;
  float alpha, beta;
  int lower_only;
  // triangular B operand: skip the K steps where its rows are structurally zero.  1: B = rows of a LOWER triangular
  // matrix (row j is zero beyond column j); 2: B = rows of its transpose (row j is zero before column j)
  int tri;
};

__global__ void __launch_bounds__(SG_THREADS, 1)
solve_gemm_tc_kernel(const SgParams p, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // full[2] empty[2] acc_full tmem_ptr
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_FULL = 0, B_EMPTY = 2, B_ACC = 4, B_TMEMPTR = 6;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + B_TMEMPTR);
  const uint32_t stage0 = smem_u32(smem + 1024);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile unsigned int* abort_flag = p.abort_flag;
  const int n0 = blockIdx.x * SG_BN, m0 = blockIdx.y * SG_BM;
  if (p.lower_only && n0 > m0 + SG_BM - 1) return;               // whole CTA, before any barrier / TMEM allocation
  int ks_begin = (int)(((long long)p.ksteps * blockIdx.z) / p.splits);
  int ks_end = (int)(((long long)p.ksteps * (blockIdx.z + 1)) / p.splits);
  if (p.tri == 1) ks_end = min(ks_end, (n0 + SG_BN - 1) / SG_BK + 1);      // columns of this tile's B rows: k <= n0 + 127
  if (p.tri == 2) ks_begin = max(ks_begin, n0 / SG_BK);                    // k >= n0
  if (ks_end < ks_begin) ks_end = ks_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SG_STAGES; ++i) { mbar_init(BAR(B_FULL + i), 1); mbar_init(BAR(B_EMPTY + i), 1); }
    mbar_init(BAR(B_ACC), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)(SG_ACCS * SG_BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&amap)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&bmap)) : "memory");
      Pipe sp{0, 0};
      for (int ks = ks_begin; ks < ks_end; ++ks) {
        if (!mbar_wait<32>(BAR(B_EMPTY + sp.stage), sp.phase ^ 1u, abort_flag)) break;
        const uint32_t base = stage0 + (uint32_t)sp.stage * SG_STAGE_BYTES;
        mbar_expect_tx(BAR(B_FULL + sp.stage), SG_STAGE_BYTES);
#pragma unroll
        for (int pl = 0; pl < 3; ++pl) {
          tma_load_3d(base + (uint32_t)pl * SG_PLANE_BYTES, &amap, ks * SG_BK, m0, pl, BAR(B_FULL + sp.stage));
          tma_load_3d(base + (uint32_t)(3 + pl) * SG_PLANE_BYTES, &bmap, ks * SG_BK, n0, pl, BAR(B_FULL + sp.stage));
        }
        sp.advance(SG_STAGES);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc(SG_BM, SG_BN, false);
      const uint64_t tmpl = umma_desc_sw(0, 8u * 128u, 128, 0);
      const uint32_t hi = (uint32_t)(tmpl >> 32), lo0 = (uint32_t)tmpl;
      Pipe sp{0, 0};
      bool ok = true;
      for (int ks = ks_begin; ks < ks_end && ok; ++ks) {
        const int rel = ks - ks_begin;
        const uint32_t d_main = tmem_base + (uint32_t)((rel % SG_MAIN) * SG_BN);
        const uint32_t d_corr = tmem_base + (uint32_t)(SG_MAIN * SG_BN);
        uint32_t acc_main = rel < SG_MAIN ? 0u : 1u;                 // first visit of an accumulator overwrites
        uint32_t acc_corr = rel == 0 ? 0u : 1u;
        if (!mbar_wait(BAR(B_FULL + sp.stage), sp.phase, abort_flag)) { ok = false; break; }
        tc_fence_after();
        const uint32_t base16 = lo0 + ((stage0 + (uint32_t)sp.stage * SG_STAGE_BYTES) >> 4);
        constexpr uint32_t PL16 = SG_PLANE_BYTES >> 4;
        // smallest terms first: (2,0) (0,2) (1,1) (1,0) (0,1) (0,0)
        constexpr int IA[6] = {2, 0, 1, 1, 0, 0};
        constexpr int IB[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
        for (int t = 0; t < 6; ++t) {
          const uint32_t a_lo = base16 + (uint32_t)IA[t] * PL16, b_lo = base16 + (uint32_t)(3 + IB[t]) * PL16;
#pragma unroll
          for (int kk = 0; kk < SG_BK / 16; ++kk) {
            if (t == 5) { tc_mma<false>(d_main, a_lo + 2u * kk, hi, b_lo + 2u * kk, hi, idesc, acc_main); acc_main = 1u; }
            else        { tc_mma<false>(d_corr, a_lo + 2u * kk, hi, b_lo + 2u * kk, hi, idesc, acc_corr); acc_corr = 1u; }
          }
        }
        tc_commit(BAR(B_EMPTY + sp.stage));
        sp.advance(SG_STAGES);
      }
      if (ok) tc_commit(BAR(B_ACC));
    }
    __syncwarp();
  } else {
    // epilogue: TMEM -> registers -> fp32 rows (each thread owns one row of the tile)
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    float* orow = p.out + ((long long)blockIdx.z * p.m + row) * p.ldo + n0;
    const bool vec = (p.ldo & 3) == 0 && ((uintptr_t)p.out & 15) == 0;
    if (ks_end > ks_begin) {
      if (mbar_wait<64>(BAR(B_ACC), 0, abort_flag)) {
        tc_fence_after();
        const int n_main = min(SG_MAIN, ks_end - ks_begin);
        for (int c0 = 0; c0 < SG_BN; c0 += 32) {
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
          tc_wait_ld();
          for (int a = 1; a <= n_main; ++a) {                        // leading accumulators, then the corrections (fp32, RN)
            uint32_t u[32];
            const int slot = a < n_main ? a : SG_MAIN;
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * SG_BN + c0), u);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
          }
          if (row < p.m && p.splits == 1 && (p.c_in != nullptr || p.alpha != 1.f)) {
            const float* crow = p.c_in ? p.c_in + (long long)row * p.ldc + n0 + c0 : nullptr;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float r = p.alpha * __uint_as_float(v[j]);
              if (crow && n0 + c0 + j < p.n) r = fmaf(p.beta, crow[j], r);
              v[j] = __float_as_uint(r);
            }
          }
          if (row < p.m) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int col = n0 + c0 + j;
              if (vec && col + 3 < p.n) {
                *reinterpret_cast<float4*>(orow + c0 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3]));
              } else {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                  if (col + t < p.n) orow[c0 + j + t] = __uint_as_float(v[j + t]);
              }
            }
          }
        }
      }
    } else if (row < p.m) {
      for (int c = 0; c < SG_BN; ++c)
        if (n0 + c < p.n) orow[c] = 0.f;           // empty K range (cannot happen with splits <= ksteps)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(SG_ACCS * SG_BN)) : "memory");
  }
}

// deterministic split-K fold: out[r][c] = sum_s partial[s][r][c] in index order
__global__ void __launch_bounds__(256)
solve_gemm_fold_kernel(const float* __restrict__ partial, int splits, int m, int n, int ldp, float* out, long long ldo,
                       float alpha, float beta, const float* c_in, long long ldc) {
  const long long total = (long long)m * n;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int r = (int)(e / n), c = (int)(e % n);
    const long long o = (long long)r * ldo + c, q = (long long)r * ldp + c;
    float acc = partial[q];
    for (int s = 1; s < splits; ++s) acc += partial[(long long)s * m * ldp + q];
    if (c_in != nullptr || alpha != 1.f) {
      acc *= alpha;
      if (c_in) acc = fmaf(beta, c_in[(long long)r * ldc + c], acc);
    }
    out[o] = acc;
  }
}

// fp32 [rows][cols] (leading dimension ld) -> three bf16 planes [3][rows][ldk], K tail zeroed
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ src, int rows, int cols, long long ld, __nv_bfloat16* __restrict__ planes,
              int ldk) {
  const long long total = (long long)rows * ldk;
  const long long plane = total;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int r = (int)(e / ldk), c = (int)(e % ldk);
    const float x = c < cols ? src[(long long)r * ld + c] : 0.f;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x);
    const float r1 = __fsub_rn(x, __bfloat162float(h0));
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const float r2 = __fsub_rn(r1, __bfloat162float(h1));
    planes[e] = h0;
    planes[plane + e] = h1;
    planes[2 * plane + e] = __float2bfloat16_rn(r2);
  }
}

static int sg_make_map(const void* planes, int rows, int k, long long ldk, CUtensorMap* map, long long plane_stride = 0) {
  EncodeTiledFn encode = tc_encoder();
  if (!encode) return 2;
  if (plane_stride == 0) plane_stride = (long long)rows * ldk;
  const cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, 3};
  const cuuint64_t strides[2] = {(cuuint64_t)ldk * 2, (cuuint64_t)plane_stride * 2};
  const cuuint32_t box[3] = {(cuuint32_t)SG_BK, (cuuint32_t)SG_BM, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(planes), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("effq_solve_gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)rc); return 2; }
  return 0;
}

static int sg_splits(int m, int n, int k) {
  const int tiles = ((m + SG_BM - 1) / SG_BM) * ((n + SG_BN - 1) / SG_BN);
  const int ksteps = (k + SG_BK - 1) / SG_BK;
  int s = sm_count() / tiles;                                   // fill the machine ...
  if (s > ksteps / 4) s = ksteps / 4;
  const int chain = (ksteps + SG_MAIN * SG_CHAIN - 1) / (SG_MAIN * SG_CHAIN);   // ... and keep the chains short
  if (s < chain) s = chain;
  if (s > 16) s = 16;
  return s < 1 ? 1 : s;
}

}  // namespace effq

extern "C" int64_t effq_split3_ld(int64_t cols) { return (cols + 63) / 64 * 64; }

extern "C" int effq_split3_bf16(const float* src, int32_t rows, int32_t cols, int64_t ld, void* planes_out,
                                void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(src && planes_out && rows > 0 && cols > 0 && ld >= cols, "bad argument");
  EFFQ_CHECK_ARG(((uintptr_t)planes_out & 15) == 0, "planes must be 16B aligned");
  const int ldk = (int)effq_split3_ld(cols);
  long long blocks = ((long long)rows * ldk + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  split3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld, (__nv_bfloat16*)planes_out, ldk);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t effq_solve_gemm_tc_workspace(int32_t m, int32_t n, int32_t k, int64_t ldo) {
  (void)ldo;                                                    // partials are compact, independent of the output pitch
  const int s = effq::sg_splits(m, n, k);
  return 16 + (s > 1 ? (int64_t)s * m * ((n + 3) / 4 * 4) * 4 : 0);
}

namespace effq {
static int sg_launch(const CUtensorMap& amap, const CUtensorMap& bmap, int m, int n, int k, float* out, long long ldo,
                     void* workspace, float alpha, float beta, const float* c_in, long long ldc, int lower_only,
                     cudaStream_t s, int tri = 0) {
  SgParams p;
  p.m = m; p.n = n; p.k = k; p.ldo = (int)ldo;
  p.ksteps = (k + SG_BK - 1) / SG_BK;
  p.splits = sg_splits(m, n, k);
  p.abort_flag = (unsigned int*)workspace;                     // word 0: abort flag (zero-initialised by the caller)
  p.c_in = c_in; p.ldc = ldc; p.alpha = alpha; p.beta = beta; p.lower_only = lower_only & 1;
  p.tri = tri;
  float* partial = (float*)((char*)workspace + 16);
  const int ldp = (n + 3) / 4 * 4;                              // split-K partials are compact: [splits][m][ldp]
  if (p.splits > 1) { p.out = partial; p.ldo = ldp; } else { p.out = out; }
  const uint32_t smem = 1024 + SG_STAGES * SG_STAGE_BYTES + 1024;
  static bool configured = false;
  if (!configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(solve_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const dim3 grid((n + SG_BN - 1) / SG_BN, (m + SG_BM - 1) / SG_BM, p.splits);
  solve_gemm_tc_kernel<<<grid, SG_THREADS, smem, s>>>(p, amap, bmap);
  EFFQ_LAUNCH_CHECK();
  if (p.splits > 1) {
    long long blocks = ((long long)m * n + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    solve_gemm_fold_kernel<<<(unsigned)blocks, 256, 0, s>>>(partial, p.splits, m, n, ldp, out, ldo, alpha, beta, c_in, ldc);
    EFFQ_LAUNCH_CHECK();
  }
  return 0;
}
}  // namespace effq

extern "C" int effq_solve_gemm_tc(const void* a_planes, const void* b_planes, int32_t m, int32_t n, int32_t k,
                                  float* out, int64_t ldo, void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(a_planes && b_planes && out && workspace, "null pointer");
  EFFQ_CHECK_ARG(m > 0 && n > 0 && k > 0 && ldo >= n, "bad shape");
  EFFQ_CHECK_ARG(((uintptr_t)a_planes & 15) == 0 && ((uintptr_t)b_planes & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                     ((uintptr_t)workspace & 15) == 0, "pointers must be 16B aligned");
  const int ldk = (int)effq_split3_ld(k);
  alignas(64) CUtensorMap amap, bmap;
  if (int rc = sg_make_map(a_planes, m, k, ldk, &amap)) return rc;
  if (int rc = sg_make_map(b_planes, n, k, ldk, &bmap)) return rc;
  return sg_launch(amap, bmap, m, n, k, out, ldo, workspace, 1.f, 0.f, nullptr, 0, 0, (cudaStream_t)stream);
}

// General form on SUB-BLOCKS of split-plane matrices:  out = alpha * A B^T + beta * c_in  with
// A = rows x k block starting at a_planes (row pitch a_ld elements, planes a_plane_stride elements apart), B likewise
// (n rows).  The building block of the blocked Cholesky factorisation and triangular inverse (csrc/chol_tc.cu).
extern "C" int64_t effq_gemm_tc_ex_workspace(int32_t m, int32_t n, int32_t k, int64_t ldo) {
  return effq_solve_gemm_tc_workspace(m, n, k, ldo);
}
extern "C" int effq_gemm_tc_ex(const void* a_planes, int64_t a_ld, int64_t a_plane_stride, const void* b_planes,
                               int64_t b_ld, int64_t b_plane_stride, int32_t m, int32_t n, int32_t k, float alpha,
                               float beta, const float* c_in, int64_t ldc, float* out, int64_t ldo,
                               int32_t lower_only, void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(a_planes && b_planes && out && workspace, "null pointer");
  EFFQ_CHECK_ARG(m > 0 && n > 0 && k > 0 && ldo >= n && (!c_in || ldc >= n), "bad shape");
  EFFQ_CHECK_ARG(a_ld % 8 == 0 && b_ld % 8 == 0 && a_plane_stride % 8 == 0 && b_plane_stride % 8 == 0 && a_ld >= k &&
                     b_ld >= k, "row pitch / plane stride must be multiples of 8 elements and cover k");
  EFFQ_CHECK_ARG(((uintptr_t)a_planes & 15) == 0 && ((uintptr_t)b_planes & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                 "operands must be 16B aligned");
  alignas(64) CUtensorMap amap, bmap;
  if (int rc = sg_make_map(a_planes, m, k, a_ld, &amap, a_plane_stride)) return rc;
  if (int rc = sg_make_map(b_planes, n, k, b_ld, &bmap, b_plane_stride)) return rc;
  // lower_only: bit 0 = skip the tiles above the diagonal; bits 1-2 = triangular B operand (SgParams::tri)
  return sg_launch(amap, bmap, m, n, k, out, ldo, workspace, alpha, beta, c_in, ldc, lower_only, (cudaStream_t)stream,
                   (lower_only >> 1) & 3);
}
