// Weight scale search (reference project_by_iter, src/models/layer_helper.py:40-70) with O(levels) work per
// fixed-point pass instead of O(elements).
//
//   a0 = mean|v| ; repeat { b = Q(v/a) ; a = sum(b*v)/sum(b*b) } until |da| <= 1e-5
//
// The 200 ADMM iterations of a layer each run one search over the K'xC2 tensor w* + dual, ~50-65 passes at 16
// levels, and every pass of the kernels in scale_search.cu evaluates every element and then crosses a cluster / grid
// barrier: 1-3 us per pass, 0.39 s of the 1.84 s calibration step (profiles/r02_bench_n1.json).  Here each CTA first
// SORTS its slice into 4096 uniform value buckets (counting sort in shared memory) and keeps prefix counts and
// prefix sums per bucket.  For a scale a the level index idx(v) = rint(clamp(v*c1 + c0)) is monotone in v, so the
// elements of a bucket that lies strictly between two rounding thresholds all share one index and enter a pass
// through the prefix tables; only the (levels - 1) buckets that contain a threshold are evaluated element by element,
// with the very same expression as the plain pass.  The level index of every element is therefore the one the plain
// pass computes, and a pass costs ~70 elements per threshold instead of the whole slice.
//
// Sums are INTEGERS: v is taken as the fixed-point number v_fix = rint(v * 2^e) (e from mean|v| and the element count,
// so that no sum can overflow 63 bits), sum idx*v_fix, sum idx^2 and sum idx are exact and independent of the
// order of addition, so they are combined with native 32-bit shared-memory atomics (no block-level reduction tree;
// one block barrier per pass, plus one hardware cluster barrier when the tensor spans several CTAs) and the result is
// bit-reproducible.  The only deviation from fp64 sums is the
// fixed-point rounding of v: <= 2^-38 mean|v| per element at 2 M elements, ~1e-12 relative on the scale (the
// reference's own fp64 summation order moves it by ~1e-15; the stopping rule is 1e-5).
//
// STATUS: experimental, off by default (EFFQ_SS_BUCKET=1 enables it); see sb_plan for the measurement.
//
// One launch: a thread-block cluster of 1..16 CTAs x 1024 threads.  Slices up to 40 K elements stay in shared memory,
// larger ones are sorted into the caller's workspace and read through L1.  Tensors of at most 64 K elements with more
// than 64 levels (conv0 / final_cls: 256 levels, 3456 and 96 elements) skip the sort and evaluate their few elements
// per thread directly -- same integer sums, same single barrier per pass.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace effq {

constexpr int SB_THREADS = 1024;
constexpr int SB_NBK = 4096;                       // interior buckets; bucket 0 / SB_NBK + 1 take what lies outside +-Rg
constexpr int SB_NB = SB_NBK + 2;
constexpr int SB_TAB = SB_NB + 2;                  // table length (prefix tables have SB_NB + 1 entries)
constexpr int SB_MAX_RANKS = 16;
constexpr int SB_WORKERS = 16;                     // bucketed mode: cells are dealt to this many warps
constexpr int SB_WORKERS_DIRECT = 16;              // direct mode: ~8 elements per lane
constexpr int SB_MAX_THR = 64;                    // thresholds (levels - 1) of the bucketed mode
constexpr int SB_CHUNK = (SB_TAB + SB_THREADS - 1) / SB_THREADS;     // table entries per thread in the block scans
constexpr double SB_EDGE_EPS = 2e-3;               // bucket units: >> the fp32 rounding of the bucket index (2.5e-4)

struct SBView {
  const float* v1;
  const float* v2;
  long long ld1, ld2;
  unsigned int rows, cols, numel;
};

__device__ __forceinline__ float sb_load(const SBView& vv, unsigned int e) {
  const unsigned int r = e / vv.cols, c = e - r * vv.cols;
  float a = __ldg(vv.v1 + (long long)r * vv.ld1 + c);
  if (vv.v2) a = __fadd_rn(a, __ldg(vv.v2 + (long long)r * vv.ld2 + c));     // fp32 add first, as `w_star + dual`
  return a;
}

struct SBShared {
  unsigned long long bar_tot[2];                   // mbarriers: all ranks' totals of a pass have arrived
  long long wslot[32][4];                          // [worker warp][sum idx*v_fix, sum idx^2, sum idx, -] of the current pass
  long long slot[2][SB_MAX_RANKS][4];              // [pass parity][source rank][sum idx*v_fix, sum idx^2, sum idx, -], written by peers
  long long setup_slot[SB_MAX_RANKS];              // set-up: every rank's sum of v_fix
  unsigned long long setup_acc;                    // set-up scratch of the direct mode
  double pub_c1;                                   // published by warp 0 for the pass: 1 / (a * delta) ...
  int pub_go, abort_flag;                          // ... whether there is a pass at all; a bounded wait gave up
  int zlo[SB_MAX_THR], zhi[SB_MAX_THR];            // ... and the zone of every threshold
  double cl_sum[SB_MAX_RANKS];                     // set-up exchange: every rank's sum|v| ...
  float cl_max[SB_MAX_RANKS];                      // ... and max|v|
  double red_d[32];
  float red_f[32];
  long long scan_ll[32];
  int scan_i[32];
  int start[SB_TAB];                               // start[k] = elements in buckets < k (exclusive prefix counts)
  int cursor[SB_TAB];
  long long pv[SB_TAB];                            // pv[k] = sum of v_fix over buckets < k
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Producer / consumer hand-over inside the CTA on hardware named barriers (ids 1 and 2; 0 is __syncthreads): the
// waiting side is parked by the barrier unit and released ~100 cycles after the last arrival (an mbarrier try_wait
// loop took ~300: profiles/r02_scale_search_bucket.md).  The fence orders the data written before the arrival.
__device__ __forceinline__ void sb_bar_arrive(int id, int threads) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void sb_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Bounded mbarrier waits (a lost arrival must not hang the device): false after ~2^22 polls or when a peer gave up.
__device__ __forceinline__ bool sb_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  unsigned int spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 0xffu) == 0) {
      if (*abort_flag) return false;
      if (spins > (1u << 22)) { *abort_flag = 1; return false; }
    }
  }
  return true;
}
__device__ __forceinline__ bool sb_wait_cluster(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
  unsigned int spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
    if ((++spins & 0xffu) == 0) {
      if (*abort_flag) return false;
      if (spins > (1u << 22)) { *abort_flag = 1; return false; }
    }
  }
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster; release.cluster orders this
// thread's earlier (remote) stores before the arrival
__device__ __forceinline__ void sb_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// exclusive scan of one value per thread over the block (1024 threads)
template <typename T>
__device__ __forceinline__ T sb_block_excl_scan(T x, T* warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const T y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();                                 // warp_tot may still be read from a previous scan
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    warp_tot[lane] = wi - w;                       // exclusive prefix of the warp totals
  }
  __syncthreads();
  return incl - x + warp_tot[warp];
}

__device__ __forceinline__ int sb_bucket(float v, float inv_w) {
  const float kf = floorf(fmaf(v, inv_w, (float)(SB_NBK / 2 + 1)));
  return kf < 0.f ? 0 : (kf > (float)(SB_NBK + 1) ? SB_NBK + 1 : (int)kf);   // NaN -> 0 (never reached: see `degenerate`)
}

// Conversions to and from 64-bit types are the slow instructions of this kernel (F2F.F64.F32, F2I.F64, F2I.S64.F64 issue
// at a few lanes per clock: with three of them per element the evaluation of 3.5 elements per thread took ~2000 cycles
// per pass, profiles/r02_scale_search_bucket.md), so the per-element paths build their doubles and fixed-point integers
// with integer instructions.
// exact float -> double (normal numbers and zeros by bit layout; denormals through the conversion unit)
__device__ __forceinline__ double sb_f2d(float v) {
  const unsigned int b = __float_as_uint(v);
  const unsigned int ex = (b >> 23) & 0xffu;
  if ((b << 1) == 0u) return __hiloint2double((int)(b & 0x80000000u), 0);
  if (ex == 0u || ex == 0xffu) return (double)v;
  return __hiloint2double((int)((b & 0x80000000u) | ((ex + 896u) << 20) | ((b & 0x7fffffu) >> 3)), (int)(b << 29));
}
// v_fix = v * 2^e rounded to the nearest integer (halves away from zero; any fixed rule serves, it is the only one
// used); the caller guarantees |v| * 2^e < 2^62
__device__ __forceinline__ long long sb_fix(float v, int e) {
  const unsigned int b = __float_as_uint(v);
  const int ex = (int)((b >> 23) & 0xffu);
  const long long m = (long long)((b & 0x7fffffu) | (ex ? 0x800000u : 0u));
  const int sh = (ex ? ex : 1) - 150 + e;                    // v = m * 2^(ex - 150)
  long long r;
  if (sh >= 0) r = m << (sh > 40 ? 40 : sh);
  else if (sh > -26) r = (m + (1ll << (-sh - 1))) >> (-sh);
  else r = 0;
  return (b >> 31) ? -r : r;
}
// exact level index of the plain pass (scale_search.cu accum_idx): the same expression, rint by the 1.5 * 2^52 trick
// (round-half-even like rint; the clamped argument lies in [0, levels - 1])
__device__ __forceinline__ int sb_idx(double v, double c1, double c0, double lm1) {
  return __double2loint(fmin(fmax(fma(v, c1, c0), 0.0), lm1) + 6755399441055744.0);
}
// The same index from fp32 arithmetic when that is provably safe: t32 = v*c1f + c0f differs from the fp64 value by at
// most ~4 * 6e-8 * 256 = 6e-5 inside the level range (<= 256 levels), so unless the clamped t32 lies within 1e-3 of a
// rounding boundary both round to the same integer; otherwise (and for NaN) the fp64 expression decides.
__device__ __forceinline__ int sb_idx_fast(float vf, float c1f, float c0f, float lm1f, double c1, double c0, double lm1) {
  const float tt = fminf(fmaxf(fmaf(vf, c1f, c0f), 0.f), lm1f);
  const float r = (tt + 12582912.f) - 12582912.f;                       // rint (half to even), |tt| < 2^22
  if (fabsf(tt - r) < 0.499f) return (int)__float_as_int(tt + 12582912.f) - 0x4b400000;
  return sb_idx(sb_f2d(vf), c1, c0, lm1);
}

// Zone of the threshold t-value T (= n + 0.5): the bucket that holds v_thr = (T - c0) * a * delta, widened by one
// bucket when v_thr sits within SB_EDGE_EPS of an edge (the fp32 bucket index of an element next to an edge may have
// rounded either way).  Every element whose index could differ from its side's index is inside [zlo, zhi].
__device__ __forceinline__ void sb_zone(double T, double c0, double ad, double inv_w_d, int& zlo, int& zhi) {
  const double kt = fma((T - c0) * ad, inv_w_d, (double)(SB_NBK / 2 + 1));
  const double kb = floor(kt), fr = kt - kb;
  double lo = kb - (fr < SB_EDGE_EPS ? 1.0 : 0.0), hi = kb + (fr > 1.0 - SB_EDGE_EPS ? 1.0 : 0.0);
  lo = fmin(fmax(lo, 0.0), (double)(SB_NBK + 1));
  hi = fmin(fmax(hi, 0.0), (double)(SB_NBK + 1));
  zlo = (int)lo;
  zhi = (int)hi;
}

__global__ void __launch_bounds__(SB_THREADS, 1)
scale_search_bucket_kernel(SBView vv, int nlvl, float lo, float hi, effq_scale_state* state, float* sorted_g,
                           unsigned int per_cta, int bucketed, int elems_in_smem, int debug) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char sb_raw[];
  SBShared& sh = *reinterpret_cast<SBShared*>(sb_raw);
  float* sv_s = reinterpret_cast<float*>(sb_raw + ((sizeof(SBShared) + 15) & ~(size_t)15));
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int rank = cluster.block_rank(), nranks = cluster.num_blocks();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const QParamD q = make_qparam_d(lo, hi, nlvl);
  const unsigned int begin = rank * per_cta;
  const unsigned int mine = begin < vv.numel ? min(per_cta, vv.numel - begin) : 0u;
  float* sv = elems_in_smem ? sv_s : sorted_g + begin;       // this CTA's (sorted) elements

  // ---- set-up 1: sum|v|, max|v| of the slice; all ranks' values through distributed shared memory ----
  // worker warps 1 .. n_workers: few, so that the per-warp reduction code is not issued 31 times over (with 8 warps per
  // scheduler doing three 64-bit shuffle trees each the pass was issue-bound: 950 cycles for the trees alone)
  const int n_workers = bucketed ? min(nlvl, SB_WORKERS) : max(1, min(SB_WORKERS_DIRECT, (int)((mine + 255u) / 256u)));
  if (t == 0) {
    sh.setup_acc = 0ull;
    sh.abort_flag = 0;
    mbar_init(smem_u32(&sh.bar_tot[0]), nranks);
    mbar_init(smem_u32(&sh.bar_tot[1]), nranks);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  double s_abs = 0.0;
  float m_abs = 0.f;
  bool nan_seen = false;
  for (unsigned int i = t; i < mine; i += SB_THREADS) {
    const float v = sb_load(vv, begin + i);
    if (!bucketed) sv_s[i] = v;                              // direct evaluation keeps the slice unsorted
    s_abs += fabs(sb_f2d(v));
    m_abs = fmaxf(m_abs, fabsf(v));
    nan_seen |= !(v == v);
  }
  s_abs = warp_sum(s_abs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m_abs = fmaxf(m_abs, __shfl_xor_sync(0xffffffffu, m_abs, o));
  nan_seen = __any_sync(0xffffffffu, nan_seen);
  if (lane == 0) { sh.red_d[warp] = s_abs; sh.red_f[warp] = nan_seen ? __int_as_float(0x7fc00000) : m_abs; }
  __syncthreads();
  if (warp == 0) {
    double s = sh.red_d[lane];                               // fixed tree: the same bits in every run
    float m = sh.red_f[lane];
    const bool bad = __any_sync(0xffffffffu, !(m == m));
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (bad) m = __int_as_float(0x7fc00000);
    if (lane < (int)nranks) {                                // lane r publishes this rank's pair in rank r's memory
      double* ps = cluster.map_shared_rank(&sh.cl_sum[rank], lane);
      float* pm = cluster.map_shared_rank(&sh.cl_max[rank], lane);
      *ps = s;
      *pm = m;
    }
  }
  cluster.sync();
  double tot_abs = 0.0;
  float vmax = 0.f;
  bool bad_in = false;
  for (unsigned int r = 0; r < nranks; ++r) {                // rank order: bit-identical in every CTA
    tot_abs += sh.cl_sum[r];
    const float m = sh.cl_max[r];
    bad_in |= !(m == m);
    vmax = fmaxf(vmax, m);
  }
  const double numel_d = (double)vv.numel;
  double a = tot_abs / numel_d;
  const int max_pass = nlvl * 100;
  // degenerate inputs: NaN -> the reference's a0 is NaN and its loop never runs; all zeros -> the plain kernel's
  // answer (a = 0 after one pass with every index 0); infinities -> a0 = inf, reported as is
  if (bad_in || !(vmax > 0.f) || !(vmax < 3e38f)) {
    if (rank == 0 && t == 0) {
      const bool zeros = !bad_in && vmax == 0.f;
      state->a = bad_in ? __longlong_as_double(0x7ff8000000000000ll) : (zeros ? 0.0 : a);
      state->a_prev = zeros ? 0.0 : -999.0;
      state->s_bv = 0.0;
      state->s_bb = zeros ? q.lo * q.lo * numel_d : 0.0;
      state->passes = zeros ? 1 : 0;
      state->converged = zeros ? 1 : 0;
      state->failed = 0;
    }
    cluster.sync();
    return;
  }
  // fixed point v_fix = rint(v * 2^e):  |sum idx*v_fix| <= (nlvl - 1) * 2^e * sum|v| = (nlvl - 1) * numel * 2^e * a0, so
  // with a0 < 2^xa and e = 62 - ceil(log2 numel) - ceil(log2 nlvl) - xa no sum (and no single v_fix: |v| <= numel * a0)
  // can leave 63 bits -- whatever the outliers -- and the resolution is ~2^-37 of the MEAN magnitude at 2 M elements
  int bits_n = 0, bits_l = 0;
  while ((1ull << bits_n) < (unsigned long long)vv.numel) ++bits_n;
  while ((1 << bits_l) < nlvl) ++bits_l;
  int E = 62 - bits_n - bits_l;
  if (E > 52) E = 52;
  int xa;
  frexp(a, &xa);                                             // a0 < 2^xa
  const int e_fix = E - xa;
  const double inv_S = ldexp(1.0, xa - E);

  // ---- set-up 2 (bucketed): counting sort of the slice + prefix tables ----
  long long sv_fix_local = 0;
  float inv_w = 0.f;
  double inv_w_d = 0.0;
  if (bucketed) {
    // buckets cover +-Rg: thresholds live inside (lo*a, hi*a) and a stays near a few mean|v|; what lies further out
    // lands in the two end buckets (evaluated element-wise only if a threshold ever gets there)
    const float rg = fminf(vmax * 1.0001f, 8.0f * (float)a);
    inv_w = (float)(SB_NBK / 2) / rg;
    inv_w_d = (double)inv_w;
    for (int k = t; k < SB_TAB; k += SB_THREADS) { sh.start[k] = 0; sh.pv[k] = 0; }
    __syncthreads();
    for (unsigned int i = t; i < mine; i += SB_THREADS) atomicAdd(&sh.start[sb_bucket(sb_load(vv, begin + i), inv_w)], 1);
    __syncthreads();
    {                                                        // exclusive scan of the counts, SB_CHUNK entries per thread
      int c[SB_CHUNK], s = 0;
#pragma unroll
      for (int j = 0; j < SB_CHUNK; ++j) { const int k = t * SB_CHUNK + j; c[j] = k < SB_NB ? sh.start[k] : 0; s += c[j]; }
      int base = sb_block_excl_scan<int>(s, sh.scan_i);
#pragma unroll
      for (int j = 0; j < SB_CHUNK; ++j) {
        const int k = t * SB_CHUNK + j;
        if (k < SB_TAB) { sh.start[k] = base; sh.cursor[k] = base; }
        base += c[j];
      }
    }
    __syncthreads();
    for (unsigned int i = t; i < mine; i += SB_THREADS) {
      const float v = sb_load(vv, begin + i);
      const int k = sb_bucket(v, inv_w);
      sv[atomicAdd(&sh.cursor[k], 1)] = v;
      atomicAdd(reinterpret_cast<unsigned long long*>(&sh.pv[k]), (unsigned long long)sb_fix(v, e_fix));
    }
    __syncthreads();
    {                                                        // exclusive scan of the bucket sums
      long long c[SB_CHUNK], s = 0;
#pragma unroll
      for (int j = 0; j < SB_CHUNK; ++j) { const int k = t * SB_CHUNK + j; c[j] = k < SB_NB ? sh.pv[k] : 0; s += c[j]; }
      long long base = sb_block_excl_scan<long long>(s, sh.scan_ll);
#pragma unroll
      for (int j = 0; j < SB_CHUNK; ++j) {
        const int k = t * SB_CHUNK + j;
        if (k < SB_TAB) sh.pv[k] = base;
        base += c[j];
      }
    }
    __syncthreads();
    sv_fix_local = sh.pv[SB_NB];
  } else {
    // direct mode: the slice as floats (written in set-up 1) and, behind them, as fixed-point integers, converted once
    long long* fx = reinterpret_cast<long long*>(sv_s + 2 * per_cta);
    long long s = 0;
    for (unsigned int i = t; i < mine; i += SB_THREADS) { const long long f = sb_fix(sv_s[i], e_fix); fx[i] = f; s += f; }
    s = warp_sum_ll(s);
    if (lane == 0) atomicAdd(&sh.setup_acc, (unsigned long long)s);
    __syncthreads();
    sv_fix_local = (long long)sh.setup_acc;
  }
  // sum of v_fix over all ranks: every rank leaves its share in every rank's slot table (buffer 0)
  if (warp == 0 && lane < (int)nranks) *cluster.map_shared_rank(&sh.setup_slot[rank], lane) = sv_fix_local;
  cluster.sync();
  long long sv_fix_all = 0;
  for (unsigned int r = 0; r < nranks; ++r) sv_fix_all += sh.setup_slot[r];
  const double s_v = (double)sv_fix_all * inv_S;

  // ---- the fixed point ----
  // Roles.  Warp 0 owns the scalar chain of a pass: totals -> a -> c1 and, one lane per threshold, the zones; it
  // publishes them and arrives on the "go" barrier.  Warps 1 .. n_workers wait there, sum their cells (or, in direct
  // mode, their elements), leave the warp's three integer sums in wslot[warp] and arrive on the "done" barrier, where
  // warp 0 waits; its lane w then takes wslot[w] and a shuffle tree gives the CTA's totals.  With several CTAs the
  // warp 0 of every rank stores its CTA's totals into every rank's slot table and arrives (release.cluster) on that
  // rank's bar_tot; it then waits for its own bar_tot and adds the slots in rank order.  The other warps never wake up
  // during the search: the fp64 divisions and the zone arithmetic of a pass are issued by ONE warp instead of 32, and
  // only the warps with work cross a barrier.  (Measured alternatives, profiles/r02_scale_search_bucket.md: every
  // warp doing the scalar chain redundantly, 64-bit shared atomics, redux.sync + 32-bit atomics, mbarrier hand-over.)
  const double c0 = -q.lo / q.delta, lm1 = (q.hi - q.lo) / q.delta;
  double a_prev = -999.0, last0 = 0.0, last1 = 0.0;
  int passes = 0;
  bool timed_out = false;
  const int bar_threads = 32 * (n_workers + 1);
  const float c0f = (float)c0, lm1f = (float)lm1;
  if (warp == 0) {
    double c1 = 1.0 / (a * q.delta);
    for (;;) {
      const bool cont = fabs(a - a_prev) > 1e-5 && passes < max_pass && !timed_out;
      if (cont && bucketed) {
        const double ad = a * q.delta;
        for (int n = lane; n < nlvl - 1; n += 32) sb_zone((double)n + 0.5, c0, ad, inv_w_d, sh.zlo[n], sh.zhi[n]);
      }
      if (lane == 0) { sh.pub_go = cont ? 1 : 0; sh.pub_c1 = c1; }
      __syncwarp();
      if (debug && rank == 0 && lane == 0 && passes < 16) ((long long*)sorted_g)[passes * 8 + 0] = clock64();
      sb_bar_arrive(1, bar_threads);                         // go
      if (!cont) break;
      sb_bar_sync(2, bar_threads);                           // done: every worker's sums are in wslot
      if (debug && rank == 0 && lane == 0 && passes < 16) ((long long*)sorted_g)[passes * 8 + 1] = clock64();
      long long g_iv = 0, g_ii = 0, g_i = 0;
      if (lane >= 1 && lane <= n_workers) { g_iv = sh.wslot[lane][0]; g_ii = sh.wslot[lane][1]; g_i = sh.wslot[lane][2]; }
      g_iv = warp_sum_ll(g_iv);
      g_ii = warp_sum_ll(g_ii);
      g_i = warp_sum_ll(g_i);
      if (nranks > 1) {
        const int par = passes & 1;
        if (lane < (int)nranks) {
          long long* dst = cluster.map_shared_rank(&sh.slot[par][rank][0], lane);
          dst[0] = g_iv;
          dst[1] = g_ii;
          dst[2] = g_i;
          sb_arrive_remote(smem_u32(&sh.bar_tot[par]), (uint32_t)lane);
        }
        if (!sb_wait_cluster(smem_u32(&sh.bar_tot[par]), (uint32_t)((passes >> 1) & 1), &sh.abort_flag)) {
          timed_out = true;
          continue;
        }
        g_iv = g_ii = g_i = 0;
        for (unsigned int r = 0; r < nranks; ++r) {
          g_iv += sh.slot[par][r][0];
          g_ii += sh.slot[par][r][1];
          g_i += sh.slot[par][r][2];
        }
      }
      const double t_iv = (double)g_iv * inv_S, t_ii = (double)g_ii, t_i = (double)g_i;
      last0 = q.delta * t_iv + q.lo * s_v;
      last1 = q.delta * q.delta * t_ii + 2.0 * q.delta * q.lo * t_i + q.lo * q.lo * numel_d;
      a_prev = a;
      a = last0 / last1;
      c1 = last1 / (last0 * q.delta);                        // = 1 / (a * delta) up to an ulp, without waiting for a
      if (debug && rank == 0 && lane == 0 && passes < 16) ((long long*)sorted_g)[passes * 8 + 2] = clock64() + (a > 1e300 ? 1 : 0);
      ++passes;
    }
  } else if (warp <= n_workers) {
    for (int p = 0;; ++p) {
      sb_bar_sync(1, bar_threads);                           // go
      if (!sh.pub_go) break;
      if (debug && rank == 0 && warp == 1 && lane == 0 && p < 16) ((long long*)sorted_g)[p * 8 + 3] = clock64();
      const double c1 = sh.pub_c1;
      const float c1f = (float)c1;
      long long s_iv = 0, s_ii = 0, s_i = 0;
      if (bucketed) {
        for (int cell = warp - 1; cell < nlvl; cell += n_workers) {
          // territory of the cell: buckets after the zone of its lower threshold up to the end of the zone of its upper
          // one; zone ends are non-decreasing in the threshold, so the territories partition 0 .. SB_NBK + 1
          int run_lo = 0, zlo = SB_NB, zhi = SB_NBK + 1;
          if (cell > 0) run_lo = sh.zhi[cell - 1] + 1;
          if (cell < nlvl - 1) { zlo = sh.zlo[cell]; zhi = sh.zhi[cell]; }
          if (zlo < run_lo) zlo = run_lo;
          if (zhi + 1 < zlo) zhi = zlo - 1;                  // zone swallowed by the previous one: nothing left
          if (lane == 0 && zlo > run_lo) {                   // uniform run: every element has index `cell`
            const long long cnt = sh.start[zlo] - sh.start[run_lo];
            s_iv += (long long)cell * (sh.pv[zlo] - sh.pv[run_lo]);
            s_ii += (long long)cell * cell * cnt;
            s_i += (long long)cell * cnt;
          }
          if (cell < nlvl - 1) {
            const int p1 = sh.start[zhi + 1];
            for (int e = sh.start[zlo] + lane; e < p1; e += 32) {
              const float vf = sv[e];
              const int idx = sb_idx(sb_f2d(vf), c1, c0, lm1);   // next to a rounding boundary: the fp32 shortcut rarely applies
              s_iv += (long long)idx * sb_fix(vf, e_fix);
              s_ii += idx * idx;
              s_i += idx;
            }
          }
        }
      } else {
        const float* fv = sv_s;
        const long long* fx = reinterpret_cast<const long long*>(sv_s + 2 * per_cta);
        const unsigned int stride = 32u * n_workers;
        for (unsigned int i0 = t - 32; i0 < mine; i0 += 4u * stride) {       // four independent chains in flight
          float vf[4];
          long long f[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const unsigned int i = i0 + j * stride;
            vf[j] = i < mine ? fv[i] : 0.f;
            f[j] = i < mine ? fx[i] : 0ll;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int idx = (i0 + j * stride < mine) ? sb_idx_fast(vf[j], c1f, c0f, lm1f, c1, c0, lm1) : 0;
            s_iv += (long long)idx * f[j];
            s_ii += idx * idx;
            s_i += idx;
          }
        }
      }
      if (debug && rank == 0 && warp == 1 && lane == 0 && p < 16) ((long long*)sorted_g)[p * 8 + 4] = clock64() + (s_iv == 0x7fffffffffffffffll ? 1 : 0);
      s_iv = warp_sum_ll(s_iv);
      s_ii = warp_sum_ll(s_ii);
      s_i = warp_sum_ll(s_i);
      if (lane == 0) { sh.wslot[warp][0] = s_iv; sh.wslot[warp][1] = s_ii; sh.wslot[warp][2] = s_i; }
      if (debug && rank == 0 && warp == 1 && lane == 0 && p < 16) ((long long*)sorted_g)[p * 8 + 5] = clock64();
      sb_bar_arrive(2, bar_threads);                         // done
    }
  }
  cluster.sync();                                            // nobody exits while a peer may still add to its memory
  if (rank == 0 && t == 0) {                                 // thread 0 is in warp 0: it holds the scalar chain
    state->a = a;
    state->a_prev = a_prev;
    state->s_bv = last0;
    state->s_bb = last1;
    state->passes = passes;
    state->converged = fabs(a - a_prev) <= 1e-5 ? 1 : 0;
    state->failed = (timed_out || sh.abort_flag) ? 2 : ((passes == max_pass) ? 1 : 0);
  }
}

// Launch plan; returns false when the tensor is not one for this kernel (the caller keeps its other variants).
struct SBPlan {
  int nranks, bucketed, in_smem;
  unsigned int per_cta;
  size_t smem;
};

static int sb_max_cluster() {
  // 16-CTA clusters are "non-portable": allowed on B200 once the attribute is set; ask the occupancy calculator
  static int cached = 0;
  if (cached) return cached;
  cached = 8;
  if (cudaFuncSetAttribute(scale_search_bucket_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16);
    cfg.blockDim = dim3(SB_THREADS);
    cfg.dynamicSmemBytes = sizeof(SBShared) + 64;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 16;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, scale_search_bucket_kernel, &cfg) == cudaSuccess && n >= 1) cached = 16;
  }
  (void)cudaGetLastError();
  return cached;
}

static bool sb_plan(long long numel, int nlvl, SBPlan& pl) {
  // Opt-in (EFFQ_SS_BUCKET=1, read at every call so that tests can switch it): measured on B200 the variant ties with
  // the plain cluster kernels up to 128 K elements and loses beyond (profiles/r02_scale_search_bucket.md) -- a pass of
  // either is a ~2400-cycle chain of dependent instructions (hand-over, a few elements, reduction, two fp64 divisions),
  // not a throughput problem, and the counting sort costs as much as twenty passes.
  const char* on_env = getenv("EFFQ_SS_BUCKET");
  const bool off = !(on_env && *on_env == '1');
  if (off || numel < 1 || numel > (1ll << 22) || nlvl < 2 || nlvl > 4096) return false;
  const size_t tab = (sizeof(SBShared) + 15) & ~(size_t)15;
  const size_t cap = (227u * 1024u - tab) / 4u;              // floats of shared memory left for the slice
  const int max_ranks = sb_max_cluster();
  pl.bucketed = (nlvl <= 64 && numel >= 8192) ? 1 : 0;
  if (!pl.bucketed && numel > 65536) return false;           // many levels AND many elements: the plain kernels
  const long long target = pl.bucketed ? 32768 : 8192;       // elements per CTA
  int r = 1;
  while (r < max_ranks && (numel + r - 1) / r > target) r <<= 1;
  pl.nranks = r;
  pl.per_cta = (unsigned int)((numel + r - 1) / r);
  pl.in_smem = (!pl.bucketed || pl.per_cta <= cap) ? 1 : 0;            // direct mode: doubles + fixed point, 16 B per element
  pl.smem = tab + (pl.bucketed ? (pl.in_smem ? (size_t)pl.per_cta * 4u : 0u) : (size_t)pl.per_cta * 16u) + 16u;
  return true;
}

// 0 launched, 1/2 error, -1 not applicable
int scale_search_bucket_launch(const float* v1, int64_t ld1, const float* v2, int64_t ld2, int64_t rows, int64_t cols,
                               int nlvl, float lo, float hi, effq_scale_state* state, float* sorted_g,
                               int64_t sorted_floats, cudaStream_t s) {
  SBPlan pl;
  const long long numel = rows * cols;
  if (!sb_plan(numel, nlvl, pl)) return -1;
  if (!pl.in_smem && (!sorted_g || sorted_floats < numel)) return -1;
  static size_t configured = 0;
  if (pl.smem > configured) {
    EFFQ_CUDA(cudaFuncSetAttribute(scale_search_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    configured = pl.smem;
  }
  SBView vv{v1, v2, ld1, ld2, (unsigned int)rows, (unsigned int)cols, (unsigned int)numel};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.nranks);
  cfg.blockDim = dim3(SB_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.nranks;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static const int dbg = [] { const char* v = getenv("EFFQ_SS_DEBUG"); return (v && *v == '1') ? 1 : 0; }();
  EFFQ_CUDA(cudaLaunchKernelEx(&cfg, scale_search_bucket_kernel, vv, nlvl, lo, hi, state, sorted_g, pl.per_cta, pl.bucketed,
                               pl.in_smem, (dbg && pl.in_smem && sorted_g && sorted_floats >= 512) ? 1 : 0));
  count_launch();
  return 0;
}

}  // namespace effq
