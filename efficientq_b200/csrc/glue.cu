// Glue ops between the quantizer layers (reference src/models/factory_blk.py:18-93,147-166 and
// factoryQ.py:66-81): ReLU, MaxPool3d(k, k) (+ the ReLU that follows it in a "mid" unit), trilinear upsampling by an
// integer factor (+ the skip connection that is added to it), residual add.  All of them are HBM-bound passes over
// NCDHW fp32 tensors; each kernel moves every byte once with 128-bit accesses and does the neighbouring elementwise op
// in the same pass (pool+ReLU: one pass instead of two; upsample+add: 12 B per output element instead of 20).
// Arithmetic: ReLU / max / add are exact operations; the interpolation follows the library's op order (source index
// scale*(dst+0.5)-0.5 clamped at 0, lambda products nested w -> h -> d) in fp32.
#include "common.cuh"

namespace effq {

constexpr int GL_THREADS = 256;

__device__ __forceinline__ float relu_nan(float v) { return v < 0.f ? 0.f : v; }     // NaN passes through like torch.relu

// y = relu(x)  or  y = a + b  or  y = relu(a + b); in place allowed (y == a)
template <bool ADD, bool RELU>
__global__ void __launch_bounds__(GL_THREADS)
glue_elementwise_kernel(const float* a, const float* __restrict__ b, long long numel, float* y) {   // y may alias a
  const long long nvec = numel >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  // a CTA walks contiguous 16 KB chunks (256 threads x 4 x 16 B): four coalesced 4 KB rows in flight per warp group
  constexpr long long CHUNK = GL_THREADS * 4;
  for (long long c0 = (long long)blockIdx.x * CHUNK; c0 < nvec; c0 += (long long)gridDim.x * CHUNK) {
    float4 va[4], vb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long j = c0 + u * GL_THREADS + threadIdx.x;
      if (j < nvec) {
        va[u] = __ldcs(a4 + j);
        if (ADD) vb[u] = __ldcs(b4 + j);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long j = c0 + u * GL_THREADS + threadIdx.x;
      if (j >= nvec) continue;
      float4 o = va[u];
      if (ADD) { o.x = __fadd_rn(o.x, vb[u].x); o.y = __fadd_rn(o.y, vb[u].y); o.z = __fadd_rn(o.z, vb[u].z); o.w = __fadd_rn(o.w, vb[u].w); }
      if (RELU) { o.x = relu_nan(o.x); o.y = relu_nan(o.y); o.z = relu_nan(o.z); o.w = relu_nan(o.w); }
      __stcs(reinterpret_cast<float4*>(y) + j, o);
    }
  }
  if (blockIdx.x == 0) {
    const long long t = (nvec << 2) + threadIdx.x;
    if (t < numel) {
      float o = a[t];
      if (ADD) o = __fadd_rn(o, b[t]);
      if (RELU) o = relu_nan(o);
      y[t] = o;
    }
  }
}

// max that propagates NaN like the library's pooling (a NaN anywhere in the window wins)
__device__ __forceinline__ float max_nan(float m, float v) { return (v > m || v != v) ? v : m; }

// MaxPool3d(kernel = stride = (kd, kh, kw)), no padding, floor mode; optional ReLU on the result (relu(max) == max(relu)).
// A thread produces VW consecutive outputs along w from kw*VW consecutive inputs of each of the kd*kh rows.
template <int KW, bool RELU>
__global__ void __launch_bounds__(GL_THREADS)
glue_maxpool_kernel(const float* __restrict__ x, long long nc, int d, int h, int w, int kd, int kh, int od, int oh, int ow,
                    float* __restrict__ y) {
  constexpr int VW = 4;                                         // outputs per thread
  const int owq = (ow + VW - 1) / VW;
  const long long total = nc * od * oh * (long long)owq;
  const bool vec = (w % 4 == 0) && (ow % VW == 0);
  for (long long e = (long long)blockIdx.x * GL_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * GL_THREADS) {
    const int q = (int)(e % owq);
    long long r = e / owq;
    const int yh = (int)(r % oh); r /= oh;
    const int yd = (int)(r % od);
    const long long ch = r / od;
    const float* src = x + ((ch * d + (long long)yd * kd) * h + (long long)yh * kh) * w + (long long)q * VW * KW;
    float m[VW];
#pragma unroll
    for (int i = 0; i < VW; ++i) m[i] = -INFINITY;
    const int nout = min(VW, ow - q * VW);
    for (int a = 0; a < kd; ++a)
      for (int b = 0; b < kh; ++b) {
        const float* row = src + ((long long)a * h + b) * w;
        if (vec) {
          float v[VW * KW];
#pragma unroll
          for (int i = 0; i < VW * KW / 4; ++i) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(row) + i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
          }
#pragma unroll
          for (int i = 0; i < VW; ++i)
#pragma unroll
            for (int k = 0; k < KW; ++k) m[i] = max_nan(m[i], v[i * KW + k]);
        } else {
          for (int i = 0; i < nout; ++i)
            for (int k = 0; k < KW; ++k) m[i] = max_nan(m[i], row[i * KW + k]);
        }
      }
    float* dst = y + ((ch * od + yd) * oh + yh) * (long long)ow + (long long)q * VW;
    if (RELU) {
#pragma unroll
      for (int i = 0; i < VW; ++i) m[i] = relu_nan(m[i]);
    }
    if (vec) {
      __stcs(reinterpret_cast<float4*>(dst), make_float4(m[0], m[1], m[2], m[3]));
    } else {
      for (int i = 0; i < nout; ++i) dst[i] = m[i];
    }
  }
}

// source position of output index o for an upsampling by the integer factor f (align_corners = False):
// src = (1/f) * (o + 0.5) - 0.5, clamped at 0; i0 = floor(src), i1 = i0 + (i0 < n - 1), lambda1 = src - i0.
struct Tap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Tap make_tap(int o, float rscale, int n) {
  float s = __fsub_rn(__fmul_rn(rscale, __fadd_rn((float)o, 0.5f)), 0.5f);
  s = s < 0.f ? 0.f : s;
  Tap t;
  t.i0 = (int)s;
  if (t.i0 > n - 1) t.i0 = n - 1;
  t.i1 = t.i0 + (t.i0 < n - 1 ? 1 : 0);
  t.l1 = __fsub_rn(s, (float)t.i0);
  t.l0 = __fsub_rn(1.f, t.l1);
  return t;
}

// y = trilinear_upsample(x, factors (fd, fh, fw)) [+ skip].  A thread produces 4 consecutive outputs along w.
template <bool ADD>
__global__ void __launch_bounds__(GL_THREADS)
glue_upsample_kernel(const float* __restrict__ x, const float* __restrict__ skip, long long nc, int d, int h, int w, int fd,
                     int fh, int fw, float* __restrict__ y) {
  const int od = d * fd, oh = h * fh, ow = w * fw;
  const int owq = (ow + 3) >> 2;
  const long long total = nc * od * oh * (long long)owq;
  const float rd = 1.0f / (float)fd, rh = 1.0f / (float)fh, rw = 1.0f / (float)fw;
  const bool vec = (ow % 4 == 0);
  for (long long e = (long long)blockIdx.x * GL_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * GL_THREADS) {
    const int q = (int)(e % owq);
    long long r = e / owq;
    const int yh = (int)(r % oh); r /= oh;
    const int yd = (int)(r % od);
    const long long ch = r / od;
    const Tap td = make_tap(yd, rd, d), th = make_tap(yh, rh, h);
    const float* p00 = x + ((ch * d + td.i0) * h + th.i0) * (long long)w;
    const float* p01 = x + ((ch * d + td.i0) * h + th.i1) * (long long)w;
    const float* p10 = x + ((ch * d + td.i1) * h + th.i0) * (long long)w;
    const float* p11 = x + ((ch * d + td.i1) * h + th.i1) * (long long)w;
    const long long o0 = ((ch * od + yd) * oh + yh) * (long long)ow + (long long)q * 4;
    float out[4];
    const int nout = min(4, ow - q * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= nout) { out[i] = 0.f; continue; }
      const Tap tw = make_tap(q * 4 + i, rw, w);
      // the library's expression (UpSampleTrilinear3d): nested w -> h -> d, left to right
      const float val =
          td.l0 * (th.l0 * (tw.l0 * __ldg(p00 + tw.i0) + tw.l1 * __ldg(p00 + tw.i1)) +
                   th.l1 * (tw.l0 * __ldg(p01 + tw.i0) + tw.l1 * __ldg(p01 + tw.i1))) +
          td.l1 * (th.l0 * (tw.l0 * __ldg(p10 + tw.i0) + tw.l1 * __ldg(p10 + tw.i1)) +
                   th.l1 * (tw.l0 * __ldg(p11 + tw.i0) + tw.l1 * __ldg(p11 + tw.i1)));
      out[i] = val;
    }
    if (vec) {
      float4 o = make_float4(out[0], out[1], out[2], out[3]);
      if (ADD) {
        const float4 s = __ldcs(reinterpret_cast<const float4*>(skip + o0));
        o.x = __fadd_rn(o.x, s.x); o.y = __fadd_rn(o.y, s.y); o.z = __fadd_rn(o.z, s.z); o.w = __fadd_rn(o.w, s.w);
      }
      __stcs(reinterpret_cast<float4*>(y + o0), o);
    } else {
      for (int i = 0; i < nout; ++i) y[o0 + i] = ADD ? __fadd_rn(out[i], skip[o0 + i]) : out[i];
    }
  }
}

static inline unsigned glue_grid(long long work_items) {
  long long blocks = (work_items + GL_THREADS - 1) / GL_THREADS;
  const long long cap = stream_cap(8);                         // one chunk per CTA (common.cuh)
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace effq

extern "C" int effq_glue_elementwise(const float* a, const float* b, int64_t numel, int32_t relu, float* y_out,
                                     void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(a && y_out, "null pointer");
  EFFQ_CHECK_ARG(b || relu, "nothing to do: neither an addend nor a ReLU");
  EFFQ_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)y_out & 15) == 0 && (!b || ((uintptr_t)b & 15) == 0),
                 "pointers must be 16B aligned");
  if (numel <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  // one 16 KB chunk per CTA (stream_cap, common.cuh)
  const unsigned grid = glue_grid((numel >> 2) / 4 + GL_THREADS);
  if (b && relu) glue_elementwise_kernel<true, true><<<grid, GL_THREADS, 0, s>>>(a, b, numel, y_out);
  else if (b)    glue_elementwise_kernel<true, false><<<grid, GL_THREADS, 0, s>>>(a, b, numel, y_out);
  else           glue_elementwise_kernel<false, true><<<grid, GL_THREADS, 0, s>>>(a, b, numel, y_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_glue_maxpool3d(const float* x, int64_t nc, int32_t d, int32_t h, int32_t w, int32_t kd, int32_t kh,
                                   int32_t kw, int32_t relu, float* y_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && y_out, "null pointer");
  EFFQ_CHECK_ARG(kd >= 1 && kh >= 1 && kw >= 1 && kw <= 2, "window: kd, kh >= 1, kw in {1, 2}");
  EFFQ_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)y_out & 15) == 0, "pointers must be 16B aligned");
  const int od = d / kd, oh = h / kh, ow = w / kw;
  if (nc <= 0 || od <= 0 || oh <= 0 || ow <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = glue_grid(nc * od * oh * (long long)((ow + 3) / 4));
  if (kw == 2) {
    if (relu) glue_maxpool_kernel<2, true><<<grid, GL_THREADS, 0, s>>>(x, nc, d, h, w, kd, kh, od, oh, ow, y_out);
    else      glue_maxpool_kernel<2, false><<<grid, GL_THREADS, 0, s>>>(x, nc, d, h, w, kd, kh, od, oh, ow, y_out);
  } else {
    if (relu) glue_maxpool_kernel<1, true><<<grid, GL_THREADS, 0, s>>>(x, nc, d, h, w, kd, kh, od, oh, ow, y_out);
    else      glue_maxpool_kernel<1, false><<<grid, GL_THREADS, 0, s>>>(x, nc, d, h, w, kd, kh, od, oh, ow, y_out);
  }
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_glue_upsample_trilinear(const float* x, const float* skip, int64_t nc, int32_t d, int32_t h, int32_t w,
                                            int32_t fd, int32_t fh, int32_t fw, float* y_out, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && y_out, "null pointer");
  EFFQ_CHECK_ARG(fd >= 1 && fh >= 1 && fw >= 1, "integer factors >= 1");
  EFFQ_CHECK_ARG(((uintptr_t)y_out & 15) == 0 && (!skip || ((uintptr_t)skip & 15) == 0), "pointers must be 16B aligned");
  if (nc <= 0 || d <= 0 || h <= 0 || w <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = glue_grid(nc * d * fd * h * fh * (long long)((w * fw + 3) / 4));
  if (skip) glue_upsample_kernel<true><<<grid, GL_THREADS, 0, s>>>(x, skip, nc, d, h, w, fd, fh, fw, y_out);
  else      glue_upsample_kernel<false><<<grid, GL_THREADS, 0, s>>>(x, skip, nc, d, h, w, fd, fh, fw, y_out);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
