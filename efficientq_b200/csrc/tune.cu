// End-to-end refinement of the activation ranges (reference src/ptqer.py:238-272,
// tune_activation_range: Adam on every alpha_act, loss = MSE of the quantised net's output).
// Two kernels of that path live here; the convolutions between them are untouched layer maths.
//
//   effq_fakequant_ste_bwd : backward of  qact = discretize(x/alpha, L, lo, hi) * alpha  under the
//       reference's straight-through estimator (layer_helper.py:13-37: round() has identity gradient,
//       torch.clamp passes the gradient where lo <= u <= hi).  With u = x/alpha, m = 1[lo <= u <= hi]:
//           d qact / d x     = m          (evaluated as (((g*alpha)*delta)/delta)/alpha, autograd's op order)
//           d qact / d alpha = D(u) - m * u            (D(u) = the level value; = hi for u > hi, lo for u < lo)
//       One pass over (x, g): writes grad_x = g*m (optional) and reduces  sum g*(D(u) - m*u)  with
//       warp shuffles, fp64 across the CTA, per-CTA partials folded in a fixed order by the last
//       CTA (deterministic, so replicated ranks stay in lock-step).  HBM-bound: 12 B per element.
//   effq_adam_step : torch.optim.Adam's update (no weight decay, no amsgrad) of the <= 64 range
//       parameters in one launch, op for op in fp32 (lerp form of the first moment).
#include "common.cuh"

namespace effq {

constexpr int STE_THREADS = 256;
constexpr int STE_MAX_BLOCKS = 2048;

template <bool WRITE_GX>
__global__ void __launch_bounds__(STE_THREADS)
fakequant_ste_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long long numel,
                         const float* __restrict__ alpha_p, QParamF q, float* __restrict__ gx,
                         double* __restrict__ grad_alpha, double* __restrict__ partials, unsigned int* __restrict__ done) {
  __shared__ double scratch[32];
  __shared__ bool is_last;
  const float alpha = __ldg(alpha_p);
  const long long nvec = numel / 4;
  const long long stride = (long long)gridDim.x * STE_THREADS;
  double acc = 0.0;
  auto one = [&](float xv, float gv, float& gxo) -> double {
    const float u = __fdiv_rn(xv, alpha);                    // exactly the reference's x / alpha
    const bool in = (u >= q.lo) && (u <= q.hi);              // torch.clamp backward; NaN -> false
    const float d = level_value_f(level_index_f(u, q), q);
    // autograd's op order through Qvar*alpha, t*delta+lo, round, (var-lo)/delta, clamp, x/alpha
    gxo = in ? __fdiv_rn(__fdiv_rn(__fmul_rn(__fmul_rn(gv, alpha), q.delta), q.delta), alpha) : 0.f;
    return (double)gv * (in ? (double)d - (double)u : (double)d);      // fp64: exact products, exact sum order below
  };
  for (long long i = (long long)blockIdx.x * STE_THREADS + threadIdx.x; i < nvec; i += stride) {
    const float4 xv = __ldcs(reinterpret_cast<const float4*>(x) + i);
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 o;
    acc += (one(xv.x, gv.x, o.x) + one(xv.y, gv.y, o.y)) + (one(xv.z, gv.z, o.z) + one(xv.w, gv.w, o.w));
    if (WRITE_GX) __stcs(reinterpret_cast<float4*>(gx) + i, o);
  }
  if (blockIdx.x == 0) {
    const long long t = nvec * 4 + threadIdx.x;
    if (t < numel) {
      float o;
      acc += one(x[t], g[t], o);
      if (WRITE_GX) gx[t] = o;
    }
  }
  const double tot = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = tot;
    __threadfence();
    is_last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double s = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += STE_THREADS) s += partials[b];   // fixed assignment
    s = block_sum(s, scratch);
    if (threadIdx.x == 0) {
      *grad_alpha += s;          // accumulate: several uses of one alpha in a graph add up, like autograd
      *done = 0;                 // workspace ready for the next launch on this stream
    }
  }
}

__global__ void adam_step_kernel(float* __restrict__ p, const double* __restrict__ grad, float grad_scale,
                                 float* __restrict__ m, float* __restrict__ v, int n, float lr, float beta1,
                                 float beta2, float eps, int step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = (float)(grad[i] * (double)grad_scale);
  // torch/optim/adam.py (_single_tensor_adam): exp_avg.lerp_(grad, 1 - beta1);
  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float m1 = __fadd_rn(m[i], __fmul_rn(1.f - beta1, __fsub_rn(gr, m[i])));
  const float v1 = __fadd_rn(__fmul_rn(v[i], beta2), __fmul_rn(1.f - beta2, __fmul_rn(gr, gr)));
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v1), bc2_sqrt), eps);
  m[i] = m1;
  v[i] = v1;
  p[i] = __fadd_rn(p[i], __fmul_rn(-step_size, __fdiv_rn(m1, denom)));
}

}  // namespace effq

extern "C" int64_t effq_ste_bwd_workspace(void) { return 16 + 8 * (int64_t)effq::STE_MAX_BLOCKS; }

extern "C" int effq_fakequant_ste_bwd(const float* x, const float* grad_out, int64_t numel, const float* alpha,
                                      float lo, float hi, int32_t nlvl, float* grad_x_out, double* grad_alpha_acc,
                                      void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(x && grad_out && alpha && grad_alpha_acc && workspace, "null pointer");
  EFFQ_CHECK_ARG(nlvl >= 2, "nlvl must be >= 2");
  EFFQ_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)grad_out & 15) == 0 &&
                     (!grad_x_out || ((uintptr_t)grad_x_out & 15) == 0), "pointers must be 16B aligned");
  if (numel <= 0) return 0;
  const QParamF q = make_qparam_f(lo, hi, nlvl);
  long long blocks = (numel / 4 + STE_THREADS - 1) / STE_THREADS;
  const long long cap = (long long)sm_count() * 8 < STE_MAX_BLOCKS ? (long long)sm_count() * 8 : STE_MAX_BLOCKS;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  unsigned int* done = (unsigned int*)workspace;          // zero-initialised by the caller, self-resetting
  double* partials = (double*)((char*)workspace + 16);
  cudaStream_t s = (cudaStream_t)stream;
  if (grad_x_out)
    fakequant_ste_bwd_kernel<true><<<(unsigned)blocks, STE_THREADS, 0, s>>>(x, grad_out, numel, alpha, q, grad_x_out,
                                                                            grad_alpha_acc, partials, done);
  else
    fakequant_ste_bwd_kernel<false><<<(unsigned)blocks, STE_THREADS, 0, s>>>(x, grad_out, numel, alpha, q, nullptr,
                                                                             grad_alpha_acc, partials, done);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_adam_step(float* params, const double* grads, float grad_scale, float* exp_avg, float* exp_avg_sq,
                              int32_t n, float lr, float beta1, float beta2, float eps, int32_t step, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "null pointer");
  EFFQ_CHECK_ARG(n > 0 && step >= 1, "bad size / step");
  adam_step_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(params, grads, grad_scale, exp_avg, exp_avg_sq, n,
                                                                       lr, beta1, beta2, eps, step);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
