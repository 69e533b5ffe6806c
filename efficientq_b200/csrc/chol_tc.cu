// Blocked Cholesky factorisation and triangular inverse of the normal matrix  A = A0 + rho*quasi_eye + eta*I
// (reference src/models/solver.py:316-331 solves with an fp32 LU of A in every iteration; here A is factorised
// and inverted once per distinct rho, efficientq_b200/spd_inverse.py).
//
// The O(n^3) work runs on the tensor cores through effq_gemm_tc_ex (split-bf16, fp32-class accuracy); this file
// holds what is left: the 128 x 128 diagonal blocks.
//   effq_potrf_tile: one CTA, block in shared memory: L11 = chol(A11) written over the block's lower triangle,
//                    W11 = L11^-1 (lower triangular) and its transpose as fp32 -- the right-looking panel step
//                    L21 = A21 L11^-T and the block triangular inverse both multiply with W11.
// plus two small layout kernels: fp32 block -> three bf16 planes written INTO a larger plane matrix (optionally
// transposed), so that factor panels become tensor-core operands without a copy of the whole matrix.
#include "common.cuh"

namespace effq {

constexpr int PT_NB = 128;
constexpr int PT_THREADS = 256;

// a: [nb][lda] fp32 (in: SPD block, lower triangle read; out: L11 in the lower triangle, upper untouched)
// w: [nb][ldw] block of the (lower triangular) inverse; wt: [PT_NB][PT_NB] dense transpose, zero padded
//
// Both phases are organised so that the O(n^3) work has no long dependent chains and few block barriers:
//   Cholesky in panels of 8 columns: warp 0 factors the 8 x 8 diagonal block, every row below solves against it in
//   registers (8 values per thread), then one rank-8 update of the trailing triangle -- 4 barriers per panel;
//   W = L^-1 by recursive doubling: the 8 x 8 diagonal blocks by substitution, then for s = 8, 16, 32, 64 the
//   off-diagonal blocks  W21 = -W22 (L21 W11)  of all block pairs of size s as two small dense products.
constexpr int PT_PW = 8;

// C[r][c] (r < rows, c < cols) = sign * sum_k A[r][k] B[k][c], k < kk; all in shared memory with pitch LD; every
// thread of the block takes outputs t, t + 256, ... of ALL `pairs` sub-problems (problem q: A, B, C advanced by
// a_step, b_step, c_step floats)
__device__ __forceinline__ void pt_small_gemm(const float* A, int a_step, const float* B, int b_step, float* C, int c_step,
                                              int pairs, int rows, int cols, int kk, int ld, float sign) {
  // a thread takes 4 consecutive columns of one row: one A value feeds four independent FMA chains
  const int c4 = cols >> 2, per = rows * c4, total = pairs * per;
  for (int e = threadIdx.x; e < total; e += PT_THREADS) {
    const int q = e / per, rc = e - q * per, r = rc / c4, c = (rc - r * c4) << 2;
    const float* ar = A + q * a_step + r * ld;
    const float* bc = B + q * b_step + c;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = 0; k < kk; ++k) {
      const float av = ar[k];
      const float* bk = bc + k * ld;
      acc[0] = fmaf(av, bk[0], acc[0]);
      acc[1] = fmaf(av, bk[1], acc[1]);
      acc[2] = fmaf(av, bk[2], acc[2]);
      acc[3] = fmaf(av, bk[3], acc[3]);
    }
    float* cr = C + q * c_step + r * ld + c;
    cr[0] = sign * acc[0];
    cr[1] = sign * acc[1];
    cr[2] = sign * acc[2];
    cr[3] = sign * acc[3];
  }
}

__global__ void __launch_bounds__(PT_THREADS, 1)
potrf_tile_kernel(float* __restrict__ a, long long lda, int nb, float* __restrict__ w, long long ldw,
                  float* __restrict__ wt, int* __restrict__ info, int block_index) {
  extern __shared__ float sm[];
  constexpr int LD = PT_NB + 1;
  float* L = sm;                         // [PT_NB][LD]
  float* W = sm + PT_NB * LD;            // [PT_NB][LD]
  float* T = W + PT_NB * LD;             // [PT_NB / 2][LD]: scratch of the doubling steps (64 * s floats, pitch LD)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // load: 8 independent global loads per thread in flight (one CTA streams the 64 KB block from L2)
  for (int e0 = t; e0 < PT_NB * PT_NB; e0 += 8 * PT_THREADS) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * PT_THREADS, i = e / PT_NB, j = e % PT_NB;
      // rows / columns beyond nb: identity (keeps every step below well defined for a partial last block)
      v[u] = (i < nb && j <= i) ? a[(long long)i * lda + j] : (i == j ? 1.f : 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * PT_THREADS, i = e / PT_NB, j = e % PT_NB;
      L[i * LD + j] = v[u];
      W[i * LD + j] = 0.f;
    }
  }
  __shared__ int bad;
  if (t == 0) bad = 0;
  __syncthreads();
  const int tx = t & 15, ty = t >> 4;
  for (int p = 0; p < PT_NB; p += PT_PW) {
    // (1) 8 x 8 diagonal block, warp 0: lane r < 8 owns row p + r of the block
    if (warp == 0) {
      for (int c = 0; c < PT_PW; ++c) {
        const float akk = L[(p + c) * LD + p + c];
        const bool ok = akk > 0.f;
        if (!ok) { if (lane == 0 && bad == 0) bad = p + c + 1; }
        const float d = ok ? sqrtf(akk) : 1.f;
        __syncwarp();
        if (lane == c) L[(p + c) * LD + p + c] = d;
        if (lane > c && lane < PT_PW) L[(p + lane) * LD + p + c] /= d;
        __syncwarp();
        // rank-1 update of the rest of the 8 x 8 block: a[r][c2] -= l[r][c] l[c2][c], c < c2 <= r
        if (lane > c && lane < PT_PW)
          for (int c2 = c + 1; c2 <= lane; ++c2)
            L[(p + lane) * LD + p + c2] = fmaf(-L[(p + lane) * LD + p + c], L[(p + c2) * LD + p + c], L[(p + lane) * LD + p + c2]);
        __syncwarp();
      }
    }
    __syncthreads();
    if (bad) break;
    // (2) rows below the diagonal block: row r solves  x L11^T = a[r][p..p+7]  in registers
    for (int r = p + PT_PW + t; r < PT_NB; r += PT_THREADS) {
      float v[PT_PW];
#pragma unroll
      for (int c = 0; c < PT_PW; ++c) v[c] = L[r * LD + p + c];
#pragma unroll
      for (int c = 0; c < PT_PW; ++c) {
#pragma unroll
        for (int m = 0; m < c; ++m) v[c] = fmaf(-v[m], L[(p + c) * LD + p + m], v[c]);
        v[c] /= L[(p + c) * LD + p + c];
      }
#pragma unroll
      for (int c = 0; c < PT_PW; ++c) L[r * LD + p + c] = v[c];
    }
    __syncthreads();
    // (3) rank-8 update of the trailing lower triangle
    for (int i = p + PT_PW + ty; i < PT_NB; i += 16) {
      float li[PT_PW];
#pragma unroll
      for (int c = 0; c < PT_PW; ++c) li[c] = L[i * LD + p + c];
      for (int j = p + PT_PW + tx; j <= i; j += 16) {
        float acc = L[i * LD + j];
#pragma unroll
        for (int c = 0; c < PT_PW; ++c) acc = fmaf(-li[c], L[j * LD + p + c], acc);
        L[i * LD + j] = acc;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad) {
    if (t == 0 && bad <= nb) atomicCAS(info, 0, block_index * PT_NB + bad);   // first failing pivot (1-based), like LAPACK's info
    if (bad <= nb) return;                                                    // (block-uniform: bad is shared)
  }
  // ---- W = L^-1.  8 x 8 diagonal blocks: thread (blk, col) substitutes one column of its block
  if (t < PT_NB) {
    const int b0 = (t / PT_PW) * PT_PW, j = t % PT_PW;
    float x[PT_PW];
#pragma unroll
    for (int i = 0; i < PT_PW; ++i) {
      float acc = (i == j) ? 1.f : 0.f;
#pragma unroll
      for (int m = 0; m < i; ++m) acc = fmaf(-L[(b0 + i) * LD + b0 + m], x[m], acc);
      x[i] = (i < j) ? 0.f : acc / L[(b0 + i) * LD + b0 + i];
    }
#pragma unroll
    for (int i = 0; i < PT_PW; ++i) W[(b0 + i) * LD + b0 + j] = x[i];
  }
  __syncthreads();
  // doubling: blocks of size s -> 2s;  W21 = -W22 (L21 W11) for every pair
  for (int s2 = PT_PW; s2 < PT_NB; s2 *= 2) {
    const int pairs = PT_NB / (2 * s2);
    const int step = 2 * s2 * LD + 2 * s2;                        // next pair along the diagonal
    // T_q = L21_q W11_q
    pt_small_gemm(L + s2 * LD, step, W, step, T, s2 * LD, pairs, s2, s2, s2, LD, 1.f);
    __syncthreads();
    // W21_q = -W22_q T_q
    pt_small_gemm(W + s2 * LD + s2, step, T, s2 * LD, W + s2 * LD, step, pairs, s2, s2, s2, LD, -1.f);
    __syncthreads();
  }
  for (int e = t; e < PT_NB * PT_NB; e += PT_THREADS) {
    const int i = e / PT_NB, j = e % PT_NB;
    if (i < nb && j <= i) a[(long long)i * lda + j] = L[i * LD + j];
    const float wv = (i < nb && j <= i) ? W[i * LD + j] : 0.f;
    if (i < nb && j < nb) w[(long long)i * ldw + j] = wv;
    wt[(long long)j * PT_NB + i] = wv;
  }
}

// src [rows][cols] fp32 (pitch ld) -> three bf16 terms written at dst (a position inside a plane matrix with row
// pitch dst_ld and plane stride dst_plane, both in elements); transpose != 0 writes src^T.  Columns from `cols`
// up to pad_cols (in the destination's K direction) are zero-filled.
__global__ void __launch_bounds__(256)
split3_block_kernel(const float* __restrict__ src, int rows, int cols, long long ld, __nv_bfloat16* __restrict__ dst,
                    long long dst_ld, long long dst_plane, int transpose, int pad_k) {
  const int out_rows = transpose ? cols : rows, out_k = transpose ? rows : cols;
  const long long total = (long long)out_rows * pad_k;
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    const int r = (int)(e / pad_k), c = (int)(e % pad_k);
    float x = 0.f;
    if (c < out_k) x = transpose ? src[(long long)c * ld + r] : src[(long long)r * ld + c];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x);
    const float r1 = __fsub_rn(x, __bfloat162float(h0));
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const float r2 = __fsub_rn(r1, __bfloat162float(h1));
    const long long o = (long long)r * dst_ld + c;
    dst[o] = h0;
    dst[dst_plane + o] = h1;
    dst[2 * dst_plane + o] = __float2bfloat16_rn(r2);
  }
}

}  // namespace effq

extern "C" int effq_potrf_tile(float* a, int64_t lda, int32_t nb, float* w_out, int64_t ldw, float* wt_out,
                               int32_t* info, int32_t block_index, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(a && w_out && wt_out && info, "null pointer");
  EFFQ_CHECK_ARG(nb >= 1 && nb <= PT_NB && lda >= nb && ldw >= nb, "block must be 1..128 wide");
  const size_t smem = (2 * (size_t)PT_NB + PT_NB / 2) * (PT_NB + 1) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  potrf_tile_kernel<<<1, PT_THREADS, smem, (cudaStream_t)stream>>>(a, lda, nb, w_out, ldw, wt_out, info, block_index);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_split3_block(const float* src, int32_t rows, int32_t cols, int64_t ld, void* dst, int64_t dst_ld,
                                 int64_t dst_plane, int32_t transpose, int32_t pad_k, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(src && dst && rows > 0 && cols > 0 && ld >= cols, "bad argument");
  const int out_k = transpose ? rows : cols;
  EFFQ_CHECK_ARG(pad_k >= out_k && dst_ld >= pad_k, "destination pitch / padding too small");
  const long long total = (long long)(transpose ? cols : rows) * pad_k;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  split3_block_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld, (__nv_bfloat16*)dst,
                                                                          dst_ld, dst_plane, transpose, pad_k);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
