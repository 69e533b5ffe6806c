// Generic fp32 direct 3D convolution with the reconstruction-error reduction fused
// into its epilogue.  Replaces  F.conv3d(Qact, G, b*) + F.mse_loss(., out_fp)
// (reference src/models/EfficientQConv.py:118-122 and :161-165) for the layers the
// tensor-core path does not take: un-quantised activations (conv0, final_cls with
// q_first/q_last = 256,-1), strided or odd geometries.  Also the on-device
// cross-check of the tcgen05 kernel in the tests.
//
// One thread = one output voxel x C2T output channels; weights for a slice of input
// channels are staged in shared memory and broadcast; activations are read through L1
// (neighbouring taps overlap).  fp32 FMA accumulation; squared error in fp32 per
// thread, then fp64 across the block; per-CTA partials are folded in a fixed order by
// the last CTA to finish, so the result is deterministic.
#include "common.cuh"

namespace effq {

constexpr int CS_THREADS = 128;
constexpr int CS_CCH = 4;               // input channels staged per step

struct ConvWs {
  unsigned int done;
  unsigned int pad[3];
  double partial[1];
};

template <int C2T>
__global__ void __launch_bounds__(CS_THREADS)
conv3d_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  effq_geom g, OutDims o, float* __restrict__ out, const float* __restrict__ target,
                  const float* __restrict__ att, double* __restrict__ sse, ConvWs* ws) {
  extern __shared__ float wsm[];        // [CS_CCH][taps][C2T]
  __shared__ double scratch[32];
  __shared__ bool is_last;
  const int taps = g.kd * g.kh * g.kw;
  const int c2_0 = blockIdx.y * C2T;
  const long long vox = (long long)blockIdx.x * CS_THREADS + threadIdx.x;
  const bool live = vox < o.vox;

  int n = 0, od = 0, oh = 0, ow = 0;
  if (live) {
    long long r = vox;
    ow = (int)(r % o.ow); r /= o.ow;
    oh = (int)(r % o.oh); r /= o.oh;
    od = (int)(r % o.od); r /= o.od;
    n = (int)r;
  }
  const int id0 = od * g.sd - g.pd, ih0 = oh * g.sh - g.ph, iw0 = ow * g.sw - g.pw;
  const long long chan_stride = (long long)g.d * g.h * g.w;
  const float* xn = x + (long long)n * g.c1 * chan_stride;

  float acc[C2T];
#pragma unroll
  for (int j = 0; j < C2T; ++j) acc[j] = 0.f;

  for (int c0 = 0; c0 < g.c1; c0 += CS_CCH) {
    const int cc = min(CS_CCH, g.c1 - c0);
    __syncthreads();
    for (int e = threadIdx.x; e < cc * taps * C2T; e += CS_THREADS) {
      const int j = e % C2T;
      const int t = (e / C2T) % taps;
      const int cl = e / (C2T * taps);
      const int c2 = c2_0 + j;
      wsm[e] = c2 < g.c2 ? __ldg(w + ((long long)c2 * g.c1 + (c0 + cl)) * taps + t) : 0.f;
    }
    __syncthreads();
    if (live) {
      for (int cl = 0; cl < cc; ++cl) {
        const float* xc = xn + (long long)(c0 + cl) * chan_stride;
        const float* wc = wsm + cl * taps * C2T;
        int t = 0;
        for (int a = 0; a < g.kd; ++a) {
          const int id = id0 + a;
          const bool okd = (unsigned)id < (unsigned)g.d;
          for (int b = 0; b < g.kh; ++b) {
            const int ih = ih0 + b;
            const bool okh = okd && (unsigned)ih < (unsigned)g.h;
            const float* row = xc + ((long long)id * g.h + ih) * g.w;
            for (int c = 0; c < g.kw; ++c, ++t) {
              const int iw = iw0 + c;
              float xv = 0.f;
              if (okh && (unsigned)iw < (unsigned)g.w) xv = __ldg(row + iw);
              const float* wt = wc + t * C2T;
#pragma unroll
              for (int j = 0; j < C2T; j += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(wt + j);
                acc[j + 0] = fmaf(xv, w4.x, acc[j + 0]);
                acc[j + 1] = fmaf(xv, w4.y, acc[j + 1]);
                acc[j + 2] = fmaf(xv, w4.z, acc[j + 2]);
                acc[j + 3] = fmaf(xv, w4.w, acc[j + 3]);
              }
            }
          }
        }
      }
    }
  }

  double err = 0.0;
  if (live) {
    const long long sp = vox - (long long)n * o.vox_per_sample;     // voxel inside the sample
    float e32 = 0.f;
#pragma unroll
    for (int j = 0; j < C2T; ++j) {
      const int c2 = c2_0 + j;
      if (c2 < g.c2) {
        const float v = acc[j] + (bias ? __ldg(bias + c2) : 0.f);
        const long long oi = ((long long)n * g.c2 + c2) * o.vox_per_sample + sp;
        if (out) out[oi] = v;
        if (target) {
          const float dlt = v - __ldg(target + oi);
          e32 = fmaf(dlt, dlt, e32);
        }
      }
    }
    if (target) err = (double)e32 * (att ? (double)__ldg(att + vox) : 1.0);
  }
  if (!target) return;

  err = block_sum(err, scratch);
  const unsigned int nblk = gridDim.x * gridDim.y;
  if (threadIdx.x == 0) {
    ws->partial[blockIdx.y * gridDim.x + blockIdx.x] = err;
    __threadfence();
    is_last = (atomicAdd(&ws->done, 1u) == nblk - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < nblk; b += CS_THREADS) t += ((volatile double*)ws->partial)[b];
    t = block_sum(t, scratch);
    if (threadIdx.x == 0) {
      *sse = t;
      ws->done = 0;
    }
  }
}

static inline int pick_c2t(int c2) { return c2 <= 4 ? 4 : (c2 <= 16 ? 16 : 32); }

}  // namespace effq

extern "C" int64_t effq_conv3d_f32_workspace(const effq_geom* g) {
  using namespace effq;
  if (!g) return 0;
  const OutDims o = out_dims(*g);
  const int c2t = pick_c2t(g->c2);
  const long long bx = (o.vox + CS_THREADS - 1) / CS_THREADS;
  const long long by = (g->c2 + c2t - 1) / c2t;
  return (int64_t)(16 + 8 * bx * by + 64);
}

extern "C" int effq_conv3d_f32(const float* x, const float* w, const float* bias, const effq_geom* g,
                               float* out, const float* target, const float* att, double* sse,
                               void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && w && g, "null pointer");
  EFFQ_CHECK_ARG(out || target, "nothing to compute");
  EFFQ_CHECK_ARG(!target || (sse && workspace), "sse/workspace required with target");
  EFFQ_CHECK_ARG(g->n > 0 && g->c1 > 0 && g->c2 > 0 && g->kd > 0 && g->kh > 0 && g->kw > 0 &&
                     g->sd > 0 && g->sh > 0 && g->sw > 0, "bad geometry");
  const OutDims o = out_dims(*g);
  EFFQ_CHECK_ARG(o.od > 0 && o.oh > 0 && o.ow > 0, "empty output");
  const int c2t = pick_c2t(g->c2);
  const long long bx = (o.vox + CS_THREADS - 1) / CS_THREADS;
  const int by = (g->c2 + c2t - 1) / c2t;
  EFFQ_CHECK_ARG(bx < (1ll << 31) && by < 65536, "grid too large");
  const int taps = g->kd * g->kh * g->kw;
  const size_t smem = (size_t)CS_CCH * taps * c2t * sizeof(float);
  EFFQ_CHECK_ARG(smem <= 48 * 1024, "kernel too large for the generic conv");
  dim3 grid((unsigned)bx, (unsigned)by);
  cudaStream_t s = (cudaStream_t)stream;
  ConvWs* ws = (ConvWs*)workspace;
  // ws->done is zero on entry: zeroed by the caller at allocation, re-zeroed by the last CTA.
  if (c2t == 4)       conv3d_f32_kernel<4><<<grid, CS_THREADS, smem, s>>>(x, w, bias, *g, o, out, target, att, sse, ws);
  else if (c2t == 16) conv3d_f32_kernel<16><<<grid, CS_THREADS, smem, s>>>(x, w, bias, *g, o, out, target, att, sse, ws);
  else                conv3d_f32_kernel<32><<<grid, CS_THREADS, smem, s>>>(x, w, bias, *g, o, out, target, att, sse, ws);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
