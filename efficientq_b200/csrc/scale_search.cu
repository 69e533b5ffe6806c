// Device-resident scale search: reference project_by_iter
// (src/models/layer_helper.py:40-70).
//
//   a0 = mean|v| ; repeat { b = Q(v/a) ; a = sum(b*v)/sum(b*b) } until |da| <= 1e-5
//
// The reference does one host round trip (.item()) per pass.  Here the whole fp64
// fixed point runs inside ONE cooperative launch: every pass streams v (fp32) once
// -- from HBM for activations, from L2 for the <= 7 MB weight tensors --, reduces the
// two sums in fp64 (warp shuffle -> block -> per-CTA partial), crosses a grid
// barrier, and every CTA then folds the partials in the same fixed order, so all
// CTAs see bit-identical scales and the loop exit is uniform.  Algorithmic bytes
// per pass: numel*4 (+numel*4 when v = v1 + v2).
//
// Large tensors (activations) additionally use INTERVAL-STABLE partial sums: late in the search
// the scale moves by tiny steps, so for an interval I = [a-w, a+w] around the current scale most
// elements have the same level index for EVERY scale in I (the index is monotone in the scale, so
// comparing the two end points decides it).  One classifying pass adds those elements into
// "stable" sums once and copies the few ambiguous ones (near a rounding boundary) to a list; while
// the scale stays inside I a pass reads only the list.  The level index of every element is the
// one the plain pass would compute, so the scales differ from the plain passes only through the
// fp64 summation order (~1e-16 relative); the pass count is the same.
#include "common.cuh"
#include "peer.cuh"
#include <stdlib.h>
#include <cooperative_groups.h>

namespace effq {

// scale_search_bucket.cu: 0 launched, 1 / 2 error, -1 tensor not for that kernel
int scale_search_bucket_launch(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows, int64_t cols,
                               int nlvl, float lo, float hi, effq_scale_state* state, float* sorted_g,
                               int64_t sorted_floats, cudaStream_t s);

constexpr int SS_THREADS = 512;
constexpr int SS_MAX_CTAS = 1024;           // partial slots per parity
constexpr unsigned long long SS_SPIN_LIMIT = 1ull << 28;

struct SSWorkspace {
  unsigned int barrier;                      // monotonically increasing arrival counter
  unsigned int abort_flag;
  unsigned int list_count[3];                // ambiguous-list fill counters (rotate per classifying pass)
  unsigned int pad[3];
  double diag[4];                            // streamed search: pass-type counts of the last launch
  double gsum[4];                            // sharded search: sums over all ranks of the current pass
  double partial[2][SS_MAX_CTAS][4];
  // followed by the ambiguous list (floats) when the caller's workspace is larger
};

struct VecView {
  const float* v1;
  const float* v2;
  long long ld1, ld2, rows, cols;
};

__device__ __forceinline__ float load_v(const VecView& vv, long long r, long long c) {
  float a = __ldg(vv.v1 + r * vv.ld1 + c);
  if (vv.v2) a = __fadd_rn(a, __ldg(vv.v2 + r * vv.ld2 + c));   // fp32 add first, as `w_star + dual`
  return a;
}

// Per-element work of a search pass.  The passes use reciprocal multiplies (v * (1/a),
// (t-lo) * (1/delta)) instead of the reference's two fp64 divisions: the level index can
// differ only when v/a sits within ~1e-16 (relative) of a rounding boundary, which changes
// a sum by one level step on one element out of millions -- far below the 1e-5 stopping
// rule.  The FINAL discretize (effq_fakequant_state / effq_quantize_act_ndhwc /
// effq_admm_project) keeps the exact divisions and is bit-exact.
// With b = idx*delta + lo:  sum b*v = delta*S_iv + lo*S_v ;  sum b*b = delta^2*S_ii + 2*delta*lo*S_i
// + lo^2*n, so a pass only accumulates S_iv = sum idx*v (fp64 FMA) and the INTEGER sums S_ii, S_i
// (exact); S_v and n are pass-invariant.  Per element: 1 FMA (q), 2 min/max, 1 rint, 1 FMA on the
// fp64 pipe plus integer work -- about half of the straightforward formulation.
struct PassQ {
  double c1, c0, lm1;
};
__device__ __forceinline__ PassQ make_passq(double a, const QParamD& q) {
  PassQ p;
  p.c1 = 1.0 / (a * q.delta);
  p.c0 = -q.lo / q.delta;
  p.lm1 = (q.hi - q.lo) / q.delta;
  return p;
}
struct PassAcc {
  double s_iv, s_ii, s_i;       // idx <= 255: s_ii, s_i stay exact integers in fp64 (< 2^53)
};
__device__ __forceinline__ void accum_idx(double v, const PassQ& p, PassAcc& acc) {
  const double idx = rint(fmin(fmax(fma(v, p.c1, p.c0), 0.0), p.lm1));
  acc.s_iv = fma(idx, v, acc.s_iv);
  acc.s_ii = fma(idx, idx, acc.s_ii);
  acc.s_i += idx;
}
// (sum b*v, sum b*b) of this thread's share from the index sums; s_v / cnt are the thread's
// pass-invariant sum of v and element count.
__device__ __forceinline__ void finish_bv(const PassAcc& acc, const QParamD& q, double s_v, double cnt, double& s0,
                                          double& s1) {
  s0 = q.delta * acc.s_iv + q.lo * s_v;
  s1 = q.delta * q.delta * acc.s_ii + 2.0 * q.delta * q.lo * acc.s_i + q.lo * q.lo * cnt;
}
// compatibility wrapper (one element at a time)
__device__ __forceinline__ void accum_bv(double v, const PassQ& p, const QParamD& q, double& s0, double& s1) {
  const double idx = rint(fmin(fmax(fma(v, p.c1, p.c0), 0.0), p.lm1));
  const double b = __dadd_rn(__dmul_rn(idx, q.delta), q.lo);
  s0 = fma(b, v, s0);
  s1 = fma(b, b, s1);
}

// One pass over this CTA's share.  MODE 0: {sum|v|, 0}. MODE 1: {sum b*v, sum b*b}.
template <int MODE>
__device__ __forceinline__ void pass_sums(const VecView& vv, double a, const QParamD& q,
                                          long long cta, long long nctas, double& s0, double& s1) {
  s0 = 0.0;
  s1 = 0.0;
  const PassQ pq = make_passq(MODE == 1 ? a : 1.0, q);
  PassAcc pa{0.0, 0.0, 0.0};
  double s_v = 0.0, cnt = 0.0;
  const long long numel = vv.rows * vv.cols;
  const bool flat = (vv.ld1 == vv.cols) && (!vv.v2 || vv.ld2 == vv.cols);
  const long long stride = nctas * SS_THREADS;
  long long i = cta * SS_THREADS + threadIdx.x;
  if (flat && (numel % 4 == 0) && (((uintptr_t)vv.v1 & 15) == 0) && (!vv.v2 || ((uintptr_t)vv.v2 & 15) == 0)) {
    const long long nvec = numel / 4;
    const float4* p1 = reinterpret_cast<const float4*>(vv.v1);
    const float4* p2 = reinterpret_cast<const float4*>(vv.v2);
    for (; i < nvec; i += stride) {
      float4 t = __ldg(p1 + i);
      if (p2) {
        const float4 u = __ldg(p2 + i);
        t.x = __fadd_rn(t.x, u.x); t.y = __fadd_rn(t.y, u.y);
        t.z = __fadd_rn(t.z, u.z); t.w = __fadd_rn(t.w, u.w);
      }
      const float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double v = (double)e[k];
        if (MODE == 0) s0 += fabs(v);
        else { accum_idx(v, pq, pa); s_v += v; cnt += 1.0; }
      }
    }
  } else {
    // strided rows (e.g. w* carries the bias column: ld = K+1): work item = (row, 512-col chunk)
    const long long chunks = (vv.cols + SS_THREADS - 1) / SS_THREADS;
    const long long items = vv.rows * chunks;
    for (long long it = cta; it < items; it += nctas) {
      const long long r = it / chunks;
      const long long c = (it % chunks) * SS_THREADS + threadIdx.x;
      if (c >= vv.cols) continue;
      const double v = (double)load_v(vv, r, c);
      if (MODE == 0) s0 += fabs(v);
      else { accum_idx(v, pq, pa); s_v += v; cnt += 1.0; }
    }
  }
  if (MODE == 1) finish_bv(pa, q, s_v, cnt, s0, s1);
}

// Every CTA folds the per-CTA partials in the same order (thread t takes slots
// t, t+512, ... then the fixed block tree), so all CTAs get bit-identical sums.
__device__ __forceinline__ void fold_partials(const double (*part)[4], unsigned int nctas,
                                              double* scratch, double* bc) {
  double t0 = 0.0, t1 = 0.0;
  for (unsigned int b = threadIdx.x; b < nctas; b += SS_THREADS) {
    t0 += ((const volatile double*)part[b])[0];
    t1 += ((const volatile double*)part[b])[1];
  }
  block_sum2(t0, t1, scratch);
  if (threadIdx.x == 0) { bc[0] = t0; bc[1] = t1; }
  __syncthreads();
}
__device__ __forceinline__ void fold_partials4(const double (*part)[4], unsigned int nctas,
                                               double* scratch, double* bc) {
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (unsigned int b = threadIdx.x; b < nctas; b += SS_THREADS)
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] += ((const volatile double*)part[b])[k];
  block_sum_n<4>(t, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) bc[k] = t[k];
  }
  __syncthreads();
}

// Grid barrier for a cooperative (co-resident) launch.  Bounded spin: on timeout the
// abort flag makes every CTA leave instead of hanging the device.
__device__ __forceinline__ bool grid_barrier(SSWorkspace* ws, unsigned int& target, unsigned int nctas) {
  __shared__ int ok_s;
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nctas;
    __threadfence();
    atomicAdd(&ws->barrier, 1u);
    unsigned long long spins = 0;
    int ok = 1;
    while (*((volatile unsigned int*)&ws->barrier) < target) {
      if (*((volatile unsigned int*)&ws->abort_flag) != 0u || ++spins > SS_SPIN_LIMIT) {
        atomicExch(&ws->abort_flag, 1u);
        ok = 0;
        break;
      }
    }
    __threadfence();
    ok_s = ok;
  }
  __syncthreads();
  return ok_s != 0;
}

// REG_ITEMS > 0: the tensor is small (weights: <= 3.5 M elements) -> every thread keeps its
// REG_ITEMS elements in registers for the whole search and a pass touches no memory at all.
// REG_ITEMS == 0: stream v from HBM/L2 every pass (activations).
template <int REG_ITEMS>
__global__ void __launch_bounds__(SS_THREADS)
scale_search_kernel(VecView vv, int nlvl, float lo, float hi, effq_scale_state* state, SSWorkspace* ws) {
  __shared__ double scratch[128];
  __shared__ double bc[2];
  const QParamD q = make_qparam_d(lo, hi, nlvl);
  const unsigned int nctas = gridDim.x;
  const long long numel = vv.rows * vv.cols;
  const int max_pass = nlvl * 100;
  unsigned int target = 0;
  int parity = 0;

  double vreg[REG_ITEMS > 0 ? REG_ITEMS : 1];
  int nvalid = 0;                      // this thread owns elements k = 0 .. nvalid-1 (strided, so a prefix)
  if (REG_ITEMS > 0) {
    const long long stride = (long long)nctas * SS_THREADS;
    long long i = (long long)blockIdx.x * SS_THREADS + threadIdx.x;
#pragma unroll
    for (int k = 0; k < REG_ITEMS; ++k, i += stride) {
      vreg[k] = 0.0;
      if (i < numel) { vreg[k] = (double)load_v(vv, i / vv.cols, i % vv.cols); nvalid = k + 1; }
    }
  }

  // pass "-1": a0 = mean|v|
  double s0, s1;
  if (REG_ITEMS > 0) {
    s0 = 0.0;
#pragma unroll
    for (int k = 0; k < REG_ITEMS; ++k) s0 += fabs(vreg[k]);
  } else {
    pass_sums<0>(vv, 0.0, q, blockIdx.x, nctas, s0, s1);
  }
  s0 = block_sum(s0, scratch);
  if (threadIdx.x == 0) { ws->partial[parity][blockIdx.x][0] = s0; ws->partial[parity][blockIdx.x][1] = 0.0; }
  if (!grid_barrier(ws, target, nctas)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
  fold_partials(ws->partial[parity], nctas, scratch, bc);
  double a = bc[0] / (double)numel;
  double a_prev = -999.0;
  int passes = 0;
  double last0 = 0.0, last1 = 0.0;

  while (fabs(a - a_prev) > 1e-5 && passes < max_pass) {
    parity ^= 1;
    if (REG_ITEMS > 0) {
      const PassQ pq = make_passq(a, q);
      s0 = 0.0;
      s1 = 0.0;
#pragma unroll
      for (int k = 0; k < REG_ITEMS; ++k)
        if (k < nvalid) accum_bv(vreg[k], pq, q, s0, s1);  // padding must not count: Q(0) != 0 on symmetric grids
    } else {
      pass_sums<1>(vv, a, q, blockIdx.x, nctas, s0, s1);
    }
    block_sum2(s0, s1, scratch);
    if (threadIdx.x == 0) { ws->partial[parity][blockIdx.x][0] = s0; ws->partial[parity][blockIdx.x][1] = s1; }
    if (!grid_barrier(ws, target, nctas)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
    fold_partials(ws->partial[parity], nctas, scratch, bc);
    last0 = bc[0];
    last1 = bc[1];
    __syncthreads();
    a_prev = a;
    a = last0 / last1;
    ++passes;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state->a = a;
    state->a_prev = a_prev;
    state->s_bv = last0;
    state->s_bb = last1;
    state->passes = passes;
    state->converged = fabs(a - a_prev) <= 1e-5 ? 1 : 0;
    state->failed = (passes == max_pass) ? 1 : 0;
  }
}

// ---- streaming variant with interval-stable partial sums (activations) ---------------------
struct ClassAcc {
  PassAcc st, am;                 // stable / ambiguous index sums at the current scale
  double sv_st, n_st, sv_am, n_am;
};
__device__ __forceinline__ double idx_of(double v, const PassQ& p) {
  return rint(fmin(fmax(fma(v, p.c1, p.c0), 0.0), p.lm1));
}
// One element against the interval [a_lo, a_hi] (plo = pass constants of a_hi, phi of a_lo): stable ->
// index sums; ambiguous -> index sums at the current scale + appended to `list`.
// Ambiguous elements are staged in shared memory and flushed to the global list with ONE global
// atomic per ~2 K elements (a global atomic per warp made the classifying pass ~18x a plain pass).
constexpr unsigned int SS_STAGE_FLUSH = 2048;                 // flush when at least this many are staged
constexpr unsigned int SS_STAGE_CAP = SS_STAGE_FLUSH + 4 * SS_THREADS;   // + one loop iteration's worst case
__device__ __forceinline__ void classify(double v, const PassQ& pa, const PassQ& plo, const PassQ& phi, ClassAcc& c,
                                         float* sbuf, unsigned int* scnt) {
  namespace cg = cooperative_groups;
  const double idx = idx_of(v, pa);
  const bool amb = idx_of(v, plo) != idx_of(v, phi);        // index monotone in the scale: end points decide
  if (!amb) {
    c.st.s_iv = fma(idx, v, c.st.s_iv); c.st.s_ii = fma(idx, idx, c.st.s_ii); c.st.s_i += idx;
    c.sv_st += v; c.n_st += 1.0;
  } else {
    c.am.s_iv = fma(idx, v, c.am.s_iv); c.am.s_ii = fma(idx, idx, c.am.s_ii); c.am.s_i += idx;
    c.sv_am += v; c.n_am += 1.0;
    cg::coalesced_group grp = cg::coalesced_threads();
    unsigned int base = 0;
    if (grp.thread_rank() == 0) base = atomicAdd(scnt, grp.size());
    base = grp.shfl(base, 0);
    sbuf[base + grp.thread_rank()] = (float)v;
  }
}
// All threads of the CTA (after a __syncthreads): move the staged elements to the global list.
__device__ __forceinline__ void stage_flush(float* sbuf, unsigned int* scnt, unsigned int* sbase, float* list,
                                            unsigned int* counter, unsigned int cap) {
  const unsigned int n = *scnt;
  if (threadIdx.x == 0) *sbase = atomicAdd(counter, n);
  __syncthreads();
  const unsigned int g0 = *sbase;
  for (unsigned int i = threadIdx.x; i < n; i += SS_THREADS)
    if (g0 + i < cap) list[g0 + i] = sbuf[i];
  __syncthreads();
  if (threadIdx.x == 0) *scnt = 0;
  __syncthreads();
}
// block-reduce the four class sums into this CTA's partial slot
__device__ __forceinline__ void publish4(const ClassAcc& c, const QParamD& q, double (*slot)[4], double* scratch) {
  double v4[4];
  finish_bv(c.st, q, c.sv_st, c.n_st, v4[0], v4[1]);
  finish_bv(c.am, q, c.sv_am, c.n_am, v4[2], v4[3]);
  block_sum_n<4>(v4, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) slot[blockIdx.x][k] = v4[k];
  }
}

constexpr unsigned int SS_RECLASS_MIN = 1u << 16;      // lists shorter than this are not worth re-classifying

__global__ void __launch_bounds__(SS_THREADS, 2)
scale_search_stream_kernel(VecView vv, int nlvl, float lo, float hi, effq_scale_state* state, SSWorkspace* ws,
                           float* list_mem, unsigned int cap, float wmax0, float dthr, effq_peer_comm comm) {
  __shared__ double scratch[128];
  __shared__ double bc[4];
  __shared__ float sbuf[SS_STAGE_CAP];
  __shared__ unsigned int scnt, sbase;
  const QParamD q = make_qparam_d(lo, hi, nlvl);
  const unsigned int nctas = gridDim.x;
  const long long numel = vv.rows * vv.cols;
  const int max_pass = nlvl * 100;
  unsigned int target = 0;
  int parity = 0;
  if (threadIdx.x == 0) scnt = 0;
  __syncthreads();

  // Sharded volumes: every pass ends with an in-kernel all-reduce of the two local sums over NVLink peer memory.
  // Every CTA holds the same local sums after the fold; CTA 0 publishes them to all ranks, and EVERY CTA then waits
  // for the ranks' contributions in the local slot buffer and adds them in rank order (peer_recv3) -- no second grid
  // barrier to hand the global sums round.  The exchange number is tracked identically by all CTAs (base read at
  // kernel start, one increment per exchange; CTA 0 stores the final count for the next launch), and the slot
  // parities are safe because CTA 0 can only send exchange s + 1 after this rank's grid barrier of pass s + 1, which
  // every CTA reaches after it has consumed exchange s.  Which elements are stable / listed is a local matter; the
  // scale -- hence every decision below -- is global and identical on all ranks.
  __shared__ double xbc[3];
  const unsigned long long xbase = comm.world > 1 ? *(volatile unsigned long long*)peer_counter(comm, 0) : 0ull;
  unsigned long long xcount = 0;
  auto exchange = [&](double& x0, double& x1) -> bool {
    if (comm.world <= 1) return true;
    ++xcount;
    if (threadIdx.x == 0) {
      double v[3] = {x0, x1, 0.0};
      if (blockIdx.x == 0) peer_send3(comm, 0, xbase + xcount, v);
      const bool ok = peer_recv3(comm, 0, xbase + xcount, v);
      xbc[0] = v[0];
      xbc[1] = v[1];
      xbc[2] = ok ? 0.0 : 1.0;
    }
    __syncthreads();
    x0 = xbc[0];
    x1 = xbc[1];
    const bool ok = xbc[2] == 0.0;
    __syncthreads();
    return ok;
  };
  auto exchange_done = [&]() {              // CTA 0: leave the exchange counter where the next launch expects it
    if (comm.world > 1 && blockIdx.x == 0 && threadIdx.x == 0) *peer_counter(comm, 0) = xbase + xcount;
  };

  double s0, s1;
  pass_sums<0>(vv, 0.0, q, blockIdx.x, nctas, s0, s1);
  s0 = block_sum(s0, scratch);
  if (threadIdx.x == 0) { ws->partial[parity][blockIdx.x][0] = s0; ws->partial[parity][blockIdx.x][1] = 0.0; }
  if (!grid_barrier(ws, target, nctas)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
  fold_partials(ws->partial[parity], nctas, scratch, bc);
  double g_abs = bc[0], g_cnt = (double)numel;
  if (!exchange(g_abs, g_cnt)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
  double a = g_abs / g_cnt;
  double a_prev = -999.0;
  int passes = 0;
  double last0 = 0.0, last1 = 0.0;

  // Interval state: identical in every CTA (derived from the folded sums and one global counter).
  // Two nested intervals.  OUTER [olo, ohi] (as wide as the list capacity allows) is classified from
  // the tensor: stable sums ost*, ambiguous elements in list L1.  INNER [ilo, ihi] (what the next few
  // steps need) is classified from L1 -- a fraction of a pass --: sums ist* of the L1 elements that are
  // stable for it, the rest in L2 (two buffers, so an L2 pass can narrow the inner interval further).
  // Leaving the inner interval costs a pass over L1, leaving the outer one a pass over the tensor.
  bool have_outer = false, have_inner = false;
  double olo = 0.0, ohi = 0.0, ost0 = 0.0, ost1 = 0.0;
  double ilo = 0.0, ihi = 0.0, ist0 = 0.0, ist1 = 0.0;
  double d = 1e300, d_prev = 0.0, wmax = wmax0, d_retry = 1e300;
  unsigned int n1 = 0, n2 = 0;
  float* l1 = list_mem;
  float* l2cur = list_mem + cap;
  float* l2nxt = list_mem + 2 * (size_t)cap;
  int build_id = 0, n_builds = 0, n_list = 0, n_reclass = 0, n_l1 = 0;

  const bool flat = (vv.ld1 == vv.cols) && (!vv.v2 || vv.ld2 == vv.cols) && (numel % 4 == 0) &&
                    (((uintptr_t)vv.v1 & 15) == 0) && (!vv.v2 || ((uintptr_t)vv.v2 & 15) == 0);

  while (fabs(a - a_prev) > 1e-5 && passes < max_pass) {
    parity ^= 1;
    double p0, p1;
    double r = d_prev > 0.0 ? d / d_prev : 0.9;               // observed contraction of the step
    r = fmin(fmax(r, 0.0), 0.98);
    const double need = fmax(4.0 * d, 2.0 * r / (1.0 - r) * d);   // half-width that should hold until convergence
    if (have_outer && a >= olo && a <= ohi) {
      // ---- list pass ----
      const bool in_inner = have_inner && a >= ilo && a <= ihi;
      const float* src;
      unsigned int src_n;
      bool reclass;
      double nlo = ilo, nhi = ihi;
      if (!in_inner) {
        // (re)build the inner interval from L1 -- unless it would be (nearly) the whole outer interval
        src = l1;
        src_n = n1;
        nlo = fmax(olo, a - need);
        nhi = fmin(ohi, a + need);
        reclass = n1 >= SS_RECLASS_MIN && (nhi - nlo) <= 0.7 * (ohi - olo);
        have_inner = false;
        ist0 = 0.0;
        ist1 = 0.0;
        ++n_l1;
      } else {
        src = l2cur;
        src_n = n2;
        reclass = n2 >= SS_RECLASS_MIN && 3.0 * need <= 0.5 * (ihi - ilo);
        if (reclass) { nlo = fmax(ilo, a - need); nhi = fmin(ihi, a + need); }
      }
      const PassQ pq = make_passq(a, q);
      // counters rotate over three slots: slot k is re-armed two barriers after its last reader
      unsigned int* counter = &ws->list_count[build_id % 3];
      if (reclass && blockIdx.x == 0 && threadIdx.x == 0) ws->list_count[(build_id + 1) % 3] = 0;
      const PassQ plo = make_passq(nhi, q), phi = make_passq(nlo, q);
      ClassAcc c{{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, 0.0, 0.0, 0.0, 0.0};
      const unsigned int nvec = (src_n + 3u) / 4u;
      const float4* lv = reinterpret_cast<const float4*>(src);
      // block-uniform trip count (the staging flush synchronises the CTA)
      for (unsigned int i0 = blockIdx.x * SS_THREADS; i0 < nvec; i0 += nctas * SS_THREADS) {
        const unsigned int i = i0 + threadIdx.x;
        if (i < nvec) {
          const float4 t = __ldcg(lv + i);                     // written earlier in this launch: no .nc path
          const float e[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (4u * i + (unsigned)k < src_n) {
              const double v = (double)e[k];
              if (reclass) classify(v, pq, plo, phi, c, sbuf, &scnt);
              else { accum_idx(v, pq, c.am); c.sv_am += v; c.n_am += 1.0; }
            }
          }
        }
        if (reclass) {
          __syncthreads();
          if (scnt >= SS_STAGE_FLUSH) stage_flush(sbuf, &scnt, &sbase, l2nxt, counter, cap);
        }
      }
      if (reclass) {
        __syncthreads();
        if (scnt > 0) stage_flush(sbuf, &scnt, &sbase, l2nxt, counter, cap);
      }
      publish4(c, q, ws->partial[parity], scratch);
      if (!grid_barrier(ws, target, nctas)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
      fold_partials4(ws->partial[parity], nctas, scratch, bc);
      ist0 += bc[0];                                           // newly stable elements (zero without re-classification)
      ist1 += bc[1];
      p0 = ost0 + ist0 + bc[2];
      p1 = ost1 + ist1 + bc[3];
      if (reclass) {
        n2 = *((volatile unsigned int*)counter);
        float* t = l2cur; l2cur = l2nxt; l2nxt = t;
        ilo = nlo;
        ihi = nhi;
        have_inner = true;
        ++build_id;
        ++n_reclass;
      }
      ++n_list;
    } else if (flat && cap > 0 && passes >= 2 && d <= (double)dthr * fabs(a) && d <= d_retry) {
      // ---- classifying pass over the tensor: stable sums + ambiguous list L1 for the outer interval ----
      const double w = fmax(4.0 * d, fmin(need, wmax * fabs(a)));
      olo = a - w;
      ohi = a + w;
      unsigned int* counter = &ws->list_count[build_id % 3];
      if (blockIdx.x == 0 && threadIdx.x == 0) ws->list_count[(build_id + 1) % 3] = 0;
      const PassQ pa_ = make_passq(a, q), plo = make_passq(ohi, q), phi = make_passq(olo, q);
      ClassAcc c{{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, 0.0, 0.0, 0.0, 0.0};
      const long long nvec = numel / 4;
      const float4* p1v = reinterpret_cast<const float4*>(vv.v1);
      const float4* p2v = reinterpret_cast<const float4*>(vv.v2);
      for (long long i0 = (long long)blockIdx.x * SS_THREADS; i0 < nvec; i0 += (long long)nctas * SS_THREADS) {
        const long long i = i0 + threadIdx.x;
        if (i < nvec) {
          float4 t = __ldg(p1v + i);
          if (p2v) {
            const float4 u = __ldg(p2v + i);
            t.x = __fadd_rn(t.x, u.x); t.y = __fadd_rn(t.y, u.y);
            t.z = __fadd_rn(t.z, u.z); t.w = __fadd_rn(t.w, u.w);
          }
          classify((double)t.x, pa_, plo, phi, c, sbuf, &scnt);
          classify((double)t.y, pa_, plo, phi, c, sbuf, &scnt);
          classify((double)t.z, pa_, plo, phi, c, sbuf, &scnt);
          classify((double)t.w, pa_, plo, phi, c, sbuf, &scnt);
        }
        __syncthreads();
        if (scnt >= SS_STAGE_FLUSH) stage_flush(sbuf, &scnt, &sbase, l1, counter, cap);
      }
      __syncthreads();
      if (scnt > 0) stage_flush(sbuf, &scnt, &sbase, l1, counter, cap);
      publish4(c, q, ws->partial[parity], scratch);
      if (!grid_barrier(ws, target, nctas)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
      fold_partials4(ws->partial[parity], nctas, scratch, bc);
      ost0 = bc[0];
      ost1 = bc[1];
      p0 = ost0 + bc[2];
      p1 = ost1 + bc[3];
      n1 = *((volatile unsigned int*)counter);
      have_outer = n1 <= cap;
      have_inner = false;
      if (!have_outer) { wmax *= 0.5; d_retry = 0.5 * d; }     // list overflow: plain passes until the step has halved
      ++build_id;
      ++n_builds;
    } else {
      // ---- plain pass ----
      have_outer = false;
      have_inner = false;
      pass_sums<1>(vv, a, q, blockIdx.x, nctas, s0, s1);
      block_sum2(s0, s1, scratch);
      if (threadIdx.x == 0) { ws->partial[parity][blockIdx.x][0] = s0; ws->partial[parity][blockIdx.x][1] = s1; }
      if (!grid_barrier(ws, target, nctas)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
      fold_partials(ws->partial[parity], nctas, scratch, bc);
      p0 = bc[0];
      p1 = bc[1];
    }
    __syncthreads();
    if (!exchange(p0, p1)) { if (blockIdx.x == 0 && threadIdx.x == 0) state->failed = 2; return; }
    last0 = p0;
    last1 = p1;
    a_prev = a;
    a = last0 / last1;
    d_prev = d;
    d = fabs(a - a_prev);
    ++passes;
  }
  exchange_done();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    state->a = a;
    state->a_prev = a_prev;
    state->s_bv = last0;
    state->s_bb = last1;
    state->passes = passes;
    state->converged = fabs(a - a_prev) <= 1e-5 ? 1 : 0;
    state->failed = (passes == max_pass) ? 1 : 0;
    // diagnostics: classifying passes over the tensor | list passes | of which re-classifying | last list length
    ws->diag[0] = (double)n_builds;
    ws->diag[1] = (double)n_list;
    ws->diag[2] = (double)n_reclass;
    ws->diag[3] = (double)n_l1;
  }
}

// ---- cluster variant: tensors up to 128 K elements (measured crossover against the 148-CTA
// grid variant, profiles/r01_scale_search.md).  One thread-block cluster; each CTA keeps its slice of v in
// shared memory for the whole search; a pass ends with ONE hardware cluster barrier and every
// CTA folds the per-CTA partial sums straight out of its peers' shared memory (DSMEM) in rank
// order, so all CTAs see bit-identical scales.  No global-memory round trip per pass.
constexpr int SC_THREADS = 512;
constexpr int SC_MAX_ELEMS = 28 * 1024;          // doubles of shared memory per CTA (224 KB)

__global__ void __launch_bounds__(SC_THREADS, 1)
scale_search_cluster_kernel(VecView vv, int nlvl, float lo, float hi, effq_scale_state* state, int per_cta) {
  namespace cg = cooperative_groups;
  // this CTA's slice of v, widened ONCE: a pass is bound by the conversion unit (F2F.F64.F32 + FRND.F64 per element at a
  // quarter of the fp64 rate, profiles/r02_scale_search_units.md), and re-converting the slice in each of ~50 passes was
  // half of that
  extern __shared__ double sv[];
  __shared__ double scratch[128];
  __shared__ double slot[2][2];                              // [parity][sum index], read by peers
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int rank = cluster.block_rank(), nranks = cluster.num_blocks();
  const QParamD q = make_qparam_d(lo, hi, nlvl);
  const long long numel = vv.rows * vv.cols;
  const long long begin = (long long)rank * per_cta;
  const int mine = (int)max(0ll, min((long long)per_cta, numel - begin));
  for (int i = threadIdx.x; i < mine; i += SC_THREADS) {
    const long long e = begin + i;
    sv[i] = (double)load_v(vv, e / vv.cols, e % vv.cols);
  }
  __syncthreads();

  __shared__ double folded[2][2];                            // [parity]: the next fold writes the other half, so one
                                                             // barrier per fold suffices (a warp can be at most one
                                                             // fold ahead of the slowest reader: cluster.sync between)
  // warp 0: lane r fetches rank r's partials over DSMEM, then a serial shuffle sum in rank order
  // (the same order in every CTA -> bit-identical scales); result broadcast through smem.
  auto fold = [&](int parity, double& t0, double& t1) {
    if (threadIdx.x < 32) {
      double x0 = 0.0, x1 = 0.0;
      if (threadIdx.x < nranks) {
        const double* peer = cluster.map_shared_rank(&slot[parity][0], threadIdx.x);
        x0 = peer[0];
        x1 = peer[1];
      }
      double a0 = 0.0, a1 = 0.0;
      for (unsigned int r = 0; r < nranks; ++r) {
        a0 += __shfl_sync(0xffffffffu, x0, (int)r);
        a1 += __shfl_sync(0xffffffffu, x1, (int)r);
      }
      if (threadIdx.x == 0) { folded[parity][0] = a0; folded[parity][1] = a1; }
    }
    __syncthreads();
    t0 = folded[parity][0];
    t1 = folded[parity][1];
  };

  int parity = 0;
  double s0 = 0.0, s1 = 0.0;
  double my_sv = 0.0, my_cnt = 0.0;                  // pass-invariant: this thread's sum of v and count
  for (int i = threadIdx.x; i < mine; i += SC_THREADS) { my_sv += sv[i]; my_cnt += 1.0; }
  for (int i = threadIdx.x; i < mine; i += SC_THREADS) s0 += fabs(sv[i]);
  s0 = block_sum(s0, scratch);
  if (threadIdx.x == 0) { slot[parity][0] = s0; slot[parity][1] = 0.0; }
  cluster.sync();
  double t0, t1;
  fold(parity, t0, t1);
  double a = t0 / (double)numel, a_prev = -999.0;
  int passes = 0;
  const int max_pass = nlvl * 100;
  double last0 = 0.0, last1 = 0.0;
  while (fabs(a - a_prev) > 1e-5 && passes < max_pass) {
    parity ^= 1;
    const PassQ pq = make_passq(a, q);
    s0 = 0.0;
    s1 = 0.0;
    {
      PassAcc pa{0.0, 0.0, 0.0};
      for (int i = threadIdx.x; i < mine; i += SC_THREADS) accum_idx(sv[i], pq, pa);
      finish_bv(pa, q, my_sv, my_cnt, s0, s1);
    }
    block_sum2(s0, s1, scratch);
    if (threadIdx.x == 0) { slot[parity][0] = s0; slot[parity][1] = s1; }
    cluster.sync();        // release/acquire: peers' slots are visible; the other parity is free again
    fold(parity, last0, last1);
    a_prev = a;
    a = last0 / last1;
    ++passes;
  }
  cluster.sync();          // nobody exits while a peer may still read its shared memory
  if (rank == 0 && threadIdx.x == 0) {
    state->a = a;
    state->a_prev = a_prev;
    state->s_bv = last0;
    state->s_bb = last1;
    state->passes = passes;
    state->converged = fabs(a - a_prev) <= 1e-5 ? 1 : 0;
    state->failed = (passes == max_pass) ? 1 : 0;
  }
}

// ---- per-row variant: one independent search per output channel (per-output-channel weight scales, the
// optional `lwq_channel_wise` extension; the reference's live path is per-tensor, PTQConv.py:26-27).  One CTA per
// row, the row in shared memory, no inter-CTA barrier: C2 searches run concurrently.
constexpr int SR_THREADS = 256;
__global__ void __launch_bounds__(SR_THREADS)
scale_search_rows_kernel(VecView vv, int nlvl, float lo, float hi, effq_scale_state* states) {
  extern __shared__ float srow[];
  __shared__ double scratch[128];
  __shared__ double bc[2];
  const QParamD q = make_qparam_d(lo, hi, nlvl);
  const long long r = blockIdx.x;
  const int cols = (int)vv.cols;
  for (int c = threadIdx.x; c < cols; c += SR_THREADS) srow[c] = load_v(vv, r, c);
  __syncthreads();
  double s0 = 0.0, my_sv = 0.0, my_cnt = 0.0;
  for (int c = threadIdx.x; c < cols; c += SR_THREADS) { const double v = (double)srow[c]; s0 += fabs(v); my_sv += v; my_cnt += 1.0; }
  s0 = block_sum(s0, scratch);
  if (threadIdx.x == 0) bc[0] = s0 / (double)cols;
  __syncthreads();
  double a = bc[0], a_prev = -999.0, last0 = 0.0, last1 = 0.0;
  int passes = 0;
  const int max_pass = nlvl * 100;
  const PassQ pq_base = make_passq(1.0, q);
  while (fabs(a - a_prev) > 1e-5 && passes < max_pass) {
    PassQ pq = pq_base;
    pq.c1 = 1.0 / (a * q.delta);
    PassAcc pa{0.0, 0.0, 0.0};
    for (int c = threadIdx.x; c < cols; c += SR_THREADS) accum_idx((double)srow[c], pq, pa);
    double t0, t1;
    finish_bv(pa, q, my_sv, my_cnt, t0, t1);
    block_sum2(t0, t1, scratch);
    __syncthreads();                         // everyone has read bc of the previous pass
    if (threadIdx.x == 0) { bc[0] = t0; bc[1] = t1; }
    __syncthreads();
    last0 = bc[0];
    last1 = bc[1];
    a_prev = a;
    a = last0 / last1;
    ++passes;
  }
  if (threadIdx.x == 0) {
    effq_scale_state* st = states + r;
    st->a = a;
    st->a_prev = a_prev;
    st->s_bv = last0;
    st->s_bb = last1;
    st->passes = passes;
    st->converged = fabs(a - a_prev) <= 1e-5 ? 1 : 0;
    st->failed = (passes == max_pass) ? 1 : 0;
  }
}

// ---- multi-GPU building blocks (one pass, no grid barrier) -----------------------
struct SPWorkspace {
  unsigned int done;
  unsigned int pad[3];
  double partial[SS_MAX_CTAS][4];
};

template <int MODE>
__global__ void __launch_bounds__(SS_THREADS)
scale_partial_kernel(VecView vv, int nlvl, float lo, float hi, const effq_scale_state* state,
                     double* sums, SPWorkspace* ws) {
  __shared__ double scratch[128];
  __shared__ bool last;
  const QParamD q = make_qparam_d(lo, hi, nlvl);
  double s0, s1;
  // a converged search keeps its scale: later passes are no-ops that re-emit the sums
  const double a = MODE == 1 ? state->a : 0.0;
  pass_sums<MODE>(vv, a, q, blockIdx.x, gridDim.x, s0, s1);
  block_sum2(s0, s1, scratch);
  if (threadIdx.x == 0) {
    ws->partial[blockIdx.x][0] = s0;
    ws->partial[blockIdx.x][1] = MODE == 0 ? 0.0 : s1;
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->done, 1u);
    last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {                       // block-uniform
    __shared__ double bc[2];
    __threadfence();
    fold_partials(ws->partial, gridDim.x, scratch, bc);
    if (threadIdx.x == 0) {
      sums[0] = bc[0];
      sums[1] = MODE == 0 ? (double)(vv.rows * vv.cols) : bc[1];
      ws->done = 0;
    }
  }
}

__global__ void scale_step_kernel(effq_scale_state* st, const double* sums, int mode, int nlvl) {
  if (mode == 0) {
    st->a = sums[0] / sums[1];
    st->a_prev = -999.0;
    st->passes = 0;
    st->converged = 0;
    st->failed = 0;
    return;
  }
  if (st->converged || st->failed) return;
  const double a_new = sums[0] / sums[1];
  st->a_prev = st->a;
  st->a = a_new;
  st->s_bv = sums[0];
  st->s_bb = sums[1];
  st->passes += 1;
  if (fabs(st->a - st->a_prev) <= 1e-5) st->converged = 1;
  else if (st->passes >= nlvl * 100) st->failed = 1;
}

static int pick_ctas(long long numel, int per_sm) {
  long long want = (numel + (long long)SS_THREADS * 16 - 1) / ((long long)SS_THREADS * 16);
  long long cap = (long long)sm_count() * per_sm;
  if (cap > SS_MAX_CTAS) cap = SS_MAX_CTAS;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace effq

static int64_t ss_base_bytes() {
  size_t a = sizeof(effq::SSWorkspace), b = sizeof(effq::SPWorkspace);
  return (int64_t)(((a > b ? a : b) + 255) & ~(size_t)255);
}
static bool ss_streams(long long numel) { return numel > (long long)effq::sm_count() * effq::SS_THREADS * 48; }

extern "C" int64_t effq_scale_search_workspace(int64_t numel) {
  // fixed part + (for tensors that are streamed from memory every pass) room for the ambiguous list:
  // three list buffers (L1 + two for L2) of a quarter of the elements each; smaller tensors: room for the bucket-sorted
  // copy of the bucketed search (scale_search_bucket.cu)
  return ss_base_bytes() + (ss_streams(numel) ? 3 * ((numel / 4 + 63) / 64 * 64) * 4 : (numel + 63) / 64 * 64 * 4);
}

template <int REG_ITEMS>
static int launch_search(effq::VecView vv, int nlvl, float lo, float hi, effq_scale_state* state,
                         effq::SSWorkspace* ws, int ctas, cudaStream_t s) {
  using namespace effq;
  void* args[] = {&vv, &nlvl, &lo, &hi, &state, &ws};
  EFFQ_CUDA(cudaLaunchCooperativeKernel((void*)scale_search_kernel<REG_ITEMS>, dim3(ctas), dim3(SS_THREADS), args, 0, s));
  count_launch();
  return 0;
}

extern "C" int effq_scale_search(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows,
                                 int64_t cols, int32_t nlvl, float lo, float hi, effq_scale_state* state,
                                 void* workspace, int64_t workspace_bytes, const effq_peer_comm* comm,
                                 void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(v1 && state && workspace, "null pointer");
  EFFQ_CHECK_ARG(!comm || (comm->world >= 1 && comm->world <= EFFQ_PEER_MAX && comm->rank >= 0 &&
                           comm->rank < comm->world), "bad communicator");
  const bool sharded = comm && comm->world > 1;
  effq_peer_comm cm;
  if (comm) cm = *comm; else { cm.world = 1; cm.rank = 0; for (int i = 0; i < EFFQ_PEER_MAX; ++i) cm.slots[i] = nullptr; }
  EFFQ_CHECK_ARG(workspace_bytes >= ss_base_bytes(), "workspace smaller than effq_scale_search_workspace(0)");
  EFFQ_CHECK_ARG(rows > 0 && cols > 0 && ld1 >= cols && (!v2 || ld2 >= cols), "bad shape");
  EFFQ_CHECK_ARG(nlvl >= 2, "nlvl must be >= 2");
  cudaStream_t s = (cudaStream_t)stream;
  const long long numel = rows * cols;
  const int sms = sm_count();                         // one CTA per SM: cheapest grid barrier
  VecView vv{v1, v2, ld1, ld2, rows, cols};
  if (!sharded) {
    // weights (and any tensor up to 4 M elements): bucket-sorted slices, O(levels) work per pass
    const int64_t room = (workspace_bytes - ss_base_bytes()) / 4;
    const int rc = scale_search_bucket_launch(v1, ld1, v2, ld2, rows, cols, nlvl, lo, hi, state,
                                              (float*)((char*)workspace + ss_base_bytes()), room > 0 ? room : 0, s);
    if (rc >= 0) return rc;
  }
  if (!sharded && numel <= 131072) {   // measured: cluster wins up to ~128 K elements, the 148-CTA grid beyond
    // small tensors (weights): one thread-block cluster, data resident in shared memory
    int nranks = (int)((numel + SC_MAX_ELEMS - 1) / SC_MAX_ELEMS);
    if (nranks < 8 && numel > 8192) {                 // spread the per-pass arithmetic a little
      const int want = (int)((numel + 8191) / 8192);
      nranks = want > 8 ? 8 : want;
    }
    int per_cta = (int)((numel + nranks - 1) / nranks);
    per_cta = (per_cta + 3) & ~3;
    const size_t smem = (size_t)per_cta * sizeof(double);
    static bool configured = false;
    if (!configured) {
      EFFQ_CUDA(cudaFuncSetAttribute(scale_search_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     SC_MAX_ELEMS * (int)sizeof(double)));
      configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nranks);
    cfg.blockDim = dim3(SC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nranks;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    EFFQ_CUDA(cudaLaunchKernelEx(&cfg, scale_search_cluster_kernel, vv, (int)nlvl, lo, hi, state, per_cta));
    count_launch();
    return 0;
  }
  EFFQ_CUDA(cudaMemsetAsync(workspace, 0, 32, s));    // barrier counter, abort flag, list counters
  SSWorkspace* ws = (SSWorkspace*)workspace;
  const long long per_cta_8 = (long long)SS_THREADS * 8;
  if (sharded) {
    // the in-kernel exchange lives in the streaming kernel: use it for every size
  } else if (numel <= (long long)sms * SS_THREADS * 8) {
    int ctas = (int)((numel + per_cta_8 - 1) / per_cta_8);
    return launch_search<8>(vv, nlvl, lo, hi, state, ws, ctas < 1 ? 1 : ctas, s);
  }
  else if (numel <= (long long)sms * SS_THREADS * 24) return launch_search<24>(vv, nlvl, lo, hi, state, ws, sms, s);
  else if (numel <= (long long)sms * SS_THREADS * 48) return launch_search<48>(vv, nlvl, lo, hi, state, ws, sms, s);
  {
    float* list = (float*)((char*)workspace + ss_base_bytes());
    long long room = (workspace_bytes - ss_base_bytes()) / 12 / 64 * 64;      // floats per list buffer
    static const bool plain = [] { const char* v = getenv("EFFQ_SCALE_PLAIN"); return v && *v == '1'; }();
    if (room < 0 || plain) room = 0;
    unsigned int cap = room > 0x7fffffffll ? 0x7fffffffu : (unsigned int)room;
    static int per_sm = 0;                             // co-resident CTAs per SM (register bound)
    if (per_sm == 0) {
      int occ = 0;
      EFFQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scale_search_stream_kernel, SS_THREADS, 0));
      per_sm = occ < 1 ? 1 : (occ > 2 ? 2 : occ);
    }
    const int ctas = sms * per_sm > SS_MAX_CTAS ? SS_MAX_CTAS : sms * per_sm;
    // interval policy: start once the step is below dthr * a (so that 4 steps fit the widest interval),
    // half-width at most wmax * a (ambiguous fraction ~ the relative half-width; the list holds 25 %)
    static const float wmax0 = [] { const char* v = getenv("EFFQ_SS_WMAX"); return v && *v ? (float)atof(v) : 0.04f; }();
    static const float dthr = [] { const char* v = getenv("EFFQ_SS_DTHR"); return v && *v ? (float)atof(v) : 0.01f; }();
    float wm = wmax0, dt = dthr;
    void* args[] = {&vv, &nlvl, &lo, &hi, &state, &ws, &list, &cap, &wm, &dt, &cm};
    EFFQ_CUDA(cudaLaunchCooperativeKernel((void*)scale_search_stream_kernel, dim3(ctas), dim3(SS_THREADS), args, 0, s));
    count_launch();
    return 0;
  }
}

extern "C" int effq_scale_partial(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows,
                                  int64_t cols, int32_t nlvl, float lo, float hi,
                                  const effq_scale_state* state, int32_t mode, double* sums,
                                  void* workspace, void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(v1 && state && workspace && sums, "null pointer");
  EFFQ_CHECK_ARG(rows > 0 && cols > 0 && ld1 >= cols && (!v2 || ld2 >= cols), "bad shape");
  EFFQ_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
  cudaStream_t s = (cudaStream_t)stream;
  const int ctas = pick_ctas(rows * cols, 2);
  VecView vv{v1, v2, ld1, ld2, rows, cols};
  SPWorkspace* ws = (SPWorkspace*)workspace;
  // ws->done must be zero on entry: zeroed by the caller once at allocation, and
  // re-zeroed by the last CTA of every launch.
  if (mode == 0) scale_partial_kernel<0><<<ctas, SS_THREADS, 0, s>>>(vv, nlvl, lo, hi, state, sums, ws);
  else           scale_partial_kernel<1><<<ctas, SS_THREADS, 0, s>>>(vv, nlvl, lo, hi, state, sums, ws);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_scale_step(effq_scale_state* state, const double* sums, int32_t mode, int32_t nlvl,
                               void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(state && sums, "null pointer");
  scale_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, sums, mode, nlvl);
  EFFQ_LAUNCH_CHECK();
  return 0;
}

extern "C" int effq_scale_search_rows(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows,
                                      int64_t cols, int32_t nlvl, float lo, float hi, effq_scale_state* states,
                                      void* stream) {
  using namespace effq;
  EFFQ_CHECK_ARG(v1 && states, "null pointer");
  EFFQ_CHECK_ARG(rows > 0 && cols > 0 && ld1 >= cols && (!v2 || ld2 >= cols) && nlvl >= 2, "bad shape");
  EFFQ_CHECK_ARG(cols <= 56 * 1024, "row too long for the per-row search (56 K elements)");
  const size_t smem = (size_t)cols * sizeof(float);
  static bool configured = false;
  if (!configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(scale_search_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024 * 4));
    configured = true;
  }
  VecView vv{v1, v2, ld1, ld2, rows, cols};
  scale_search_rows_kernel<<<(unsigned)rows, SR_THREADS, smem, (cudaStream_t)stream>>>(vv, (int)nlvl, lo, hi, states);
  EFFQ_LAUNCH_CHECK();
  return 0;
}
