// Micro-benchmark: cycles per tcgen05.mma (cta_group::1) on B200 as a function of the kind
// (f16/bf16 K=16, f8f6f4/e4m3 K=32, i8 K=32), M, N, swizzle width and operand placement.  Operands are K-major
// SWIZZLE_128B tiles in shared memory (contents irrelevant).  One CTA per SM is launched on
// every SM so the numbers include any chip-level effects.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t sbo, int rp) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(rp == 128 ? 2 : (rp == 64 ? 4 : 6)) << 61;      // SWIZZLE_128B / 64B / 32B
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

// a_mode 2 = the conv kernel's real descriptor walk: 27 taps x (rp/32) K-steps, A = tap-shifted
//            view of a 3x18x10-row halo, B = a distinct N x rp weight tile per tap
// a_mode / b_mode: 0 = canonical tile (8-row groups contiguous, start aligned),
//                  1 = "halo" tile (groups 10 rows apart, start shifted by one row) -> every 8-row
//                      group straddles two swizzle atoms, as the conv kernel's tap-shifted operand
// rp: row pitch / swizzle width in bytes (128 or 64)
template <int KIND>
__device__ __forceinline__ void mma_lohi(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
  else if (KIND == 1)
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
  else
    asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
}

// The conv kernel's descriptor walk with its issue code: 27 unrolled taps x KK K-steps per "tile",
// A = tap-shifted view of a 3x18x10-row halo, B = a distinct N x rp weight tile per tap (wrapping
// inside the B region), 32-bit arithmetic on the descriptors' low words only.
template <int KIND, int KK>
__device__ __forceinline__ void conv_walk(uint32_t tmem, uint32_t idesc, uint32_t a0, uint32_t b0, int n, int rp, int tiles) {
  const uint64_t at = desc_sw(0, (uint32_t)(10 * rp), rp), bt = desc_sw(0, (uint32_t)(8 * rp), rp);
  const uint32_t a_hi = (uint32_t)(at >> 32), b_hi = (uint32_t)(bt >> 32);
  const uint32_t a_base = (uint32_t)at + (a0 >> 4), b_base = (uint32_t)bt + (b0 >> 4);
  const uint32_t row16 = (uint32_t)rp >> 4, step_b = 10u * row16, step_a = 180u * row16;
  const uint32_t b_tile16 = (uint32_t)(n * rp) >> 4;
  const uint32_t b_wrap = (118u * 1024u) / (uint32_t)(n * rp);
  for (int t = 0; t < tiles; ++t) {
    uint32_t w_lo = b_base, wi = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint32_t ha = a_base + (uint32_t)a * step_a + (uint32_t)b * step_b + (uint32_t)c * row16;
#pragma unroll
          for (int kk = 0; kk < KK; ++kk) mma_lohi<KIND>(tmem, ha + 2u * kk, a_hi, w_lo + 2u * kk, b_hi, idesc);
          w_lo += b_tile16;
          if (++wi == b_wrap) { wi = 0; w_lo = b_base; }
        }
  }
}

template <int KIND>      // 0 = kind::f16 (bf16), 1 = kind::f8f6f4 (e4m3), 2 = kind::i8 (u8 x s8)
__global__ void __launch_bounds__(128, 1)
mma_bench(int m, int n, int a_mode, int b_mode, int rp, int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  // zero the operands (finite values)
  for (int i = threadIdx.x; i < (190 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    uint32_t idesc = ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    if (KIND == 0) idesc |= (1u << 4) | (1u << 7) | (1u << 10);          // D=f32, A=B=bf16
    if (KIND == 1) idesc |= (1u << 4);                                    // D=f32, A=B=e4m3 (format 0)
    if (KIND == 2) idesc |= (2u << 4) | (0u << 7) | (1u << 10);          // D=s32, A=u8, B=s8
    const uint32_t a0 = smem_u32(sm), b0 = a0 + 72 * 1024;           // A region 72 KB, B region 118 KB
    const uint64_t at = desc_sw(0, (uint32_t)((a_mode ? 10 : 8) * rp), rp), bt = desc_sw(0, (uint32_t)((b_mode ? 10 : 8) * rp), rp);
    const uint32_t a_sh = a_mode ? (uint32_t)rp : 0u, b_sh = b_mode ? (uint32_t)rp : 0u;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      if (a_mode == 2) {
        const int kk_n = rp / 32, tiles = iters / (27 * kk_n);
        if (kk_n == 1) conv_walk<KIND, 1>(tmem, idesc, a0, b0, n, rp, tiles);
        else if (kk_n == 2) conv_walk<KIND, 2>(tmem, idesc, a0, b0, n, rp, tiles);
        else conv_walk<KIND, 4>(tmem, idesc, a0, b0, n, rp, tiles);
      } else
      for (int i = 0; i < iters; ++i) {
        // K advance inside the row: 32 B per MMA, wrapping inside the row pitch
        const uint32_t koff = (uint32_t)((i * 32) & (rp - 1));
        uint32_t aoff = a_sh + koff, boff = b_sh + koff;

        const uint64_t ad = at | (uint64_t)(((a0 + aoff) >> 4) & 0x3fffu);
        const uint64_t bd = bt | (uint64_t)(((b0 + boff) >> 4) & 0x3fffu);
        const uint32_t d = tmem;
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        else if (KIND == 1)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      t1 = clock64();
    }
    __syncwarp();
    // wait for completion
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    const long long t2 = clock64();
    if (t0 != 0 && blockIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int KIND>
static int sweep(const char* kind, int sms, unsigned long long* out) {
  const size_t smem_use = 200 * 1024;
  cudaFuncSetAttribute(mma_bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_use);
  const int iters = 4320;            // multiple of 27 * 4
  const int ns[] = {32, 64, 128, 256};
  const int rps[] = {128, 64, 32};
  for (int ni = 0; ni < 4; ++ni)
    for (int am = 0; am < 3; ++am)
      for (int ri = 0; ri < 3; ++ri) {
        const int m = 128, n = ns[ni], rp = rps[ri];
        mma_bench<KIND><<<sms, 128, smem_use>>>(m, n, am, 0, rp, iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s error %s\n", kind, cudaGetErrorString(e)); return 1; }
        unsigned long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%-7s %-3d %-3d %-5s %-4d %-14.1f %-12.1f\n", kind, m, n, am == 2 ? "conv" : (am ? "halo" : "canon"), rp,
               (double)h[0] / iters, (double)h[1] / iters);
      }
  return 0;
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 16);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("kind    M   N   A     rp   issue_cyc/mma  done_cyc/mma\n");
  if (sweep<0>("bf16", sms, out)) return 1;
  if (sweep<1>("e4m3", sms, out)) return 1;
  if (sweep<2>("i8", sms, out)) return 1;
  return 0;
}
