// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16, K=16, cta_group::1) on B200 as a
// function of M, N, operand reuse and accumulator dependence.  Operands are K-major
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
  d |= (uint64_t)(rp == 128 ? 2 : 4) << 61;      // SWIZZLE_128B / SWIZZLE_64B
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}

// a_mode / b_mode: 0 = canonical tile (8-row groups contiguous, start aligned),
//                  1 = "halo" tile (groups 10 rows apart, start shifted by one row) -> every 8-row
//                      group straddles two swizzle atoms, as the conv kernel's tap-shifted operand
// rp: row pitch / swizzle width in bytes (128 or 64)
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
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    const uint32_t a0 = smem_u32(sm), b0 = a0 + 64 * 1024;           // A region 64 KB, B region 128 KB
    const uint64_t at = desc_sw(0, (uint32_t)((a_mode ? 10 : 8) * rp), rp), bt = desc_sw(0, (uint32_t)((b_mode ? 10 : 8) * rp), rp);
    const uint32_t a_sh = a_mode ? (uint32_t)rp : 0u, b_sh = b_mode ? (uint32_t)rp : 0u;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        // K advance inside the row: 32 B per MMA, wrapping inside the row pitch
        const uint32_t koff = (uint32_t)((i * 32) & (rp - 1));
        const uint64_t ad = at | (uint64_t)(((a0 + a_sh + koff) >> 4) & 0x3fffu);
        const uint64_t bd = bt | (uint64_t)(((b0 + b_sh + koff) >> 4) & 0x3fffu);
        const uint32_t d = tmem;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 16);
  const size_t smem = 8 * 16384 + 8 * 32768 + 1024;      // 385 KB? too big -> use 4 slices each
  (void)smem;
  const size_t smem_use = 200 * 1024;
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_use);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("M   N   A     B     rp   issue_cyc/mma  done_cyc/mma\n");
  const int iters = 4096;
  const int ms[] = {128, 64};
  const int ns[] = {32, 64, 128, 256};
  for (int mi = 0; mi < 2; ++mi)
    for (int ni = 0; ni < 4; ++ni)
      for (int am = 0; am < 2; ++am)
        for (int bm = 0; bm < 2; ++bm)
          for (int acc = 128; acc >= 64; acc /= 2) {
            const int m = ms[mi], n = ns[ni];
            // extents: A rows m -> (m/8 groups) * 10 rows * rp + rp <= 64 KB; B rows n -> <= 128 KB   (checked: max 41 KB)
            mma_bench<<<sms, 128, smem_use>>>(m, n, am, bm, acc, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            unsigned long long h[2];
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("%-3d %-3d %-5s %-5s %-4d %-14.1f %-12.1f\n", m, n, am ? "halo" : "canon", bm ? "halo" : "canon", acc,
                   (double)h[0] / iters, (double)h[1] / iters);
          }
  return 0;
}
