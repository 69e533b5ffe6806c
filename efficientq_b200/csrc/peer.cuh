// Tiny all-reduce of a few doubles between the GPUs of one node, done INSIDE a kernel over
// NVLink peer memory (no NCCL launch, no host round trip).  The sharded calibration exchanges
// 2 doubles per activation-search pass and 1 double per ADMM iteration -- thousands of
// latency-bound exchanges per layer; as NCCL calls they cost more than the kernels around them.
//
// Every rank owns one small buffer (effq_peer_alloc) that all peers map (CUDA IPC):
//   PeerSlot slot[EFFQ_PEER_CHANNELS][EFFQ_PEER_MAX][2];   // [channel][source rank][parity], 64 B each
//   unsigned long long next_seq[EFFQ_PEER_CHANNELS];       // local only, at EFFQ_PEER_CTR_OFFSET
//   unsigned long long aborted;                            // local only: sticky after a timeout
// An exchange with sequence number s: write my values into slot[ch][my rank][s & 1] of every rank
// (8-byte words that carry data and the sequence number together), spin until the local slots of all ranks carry s,
// add them in rank order (bit-identical result on every rank).  Two parities suffice: a rank can
// only be one exchange ahead of the slowest.  Waits are bounded (30 s of wall time); on timeout the abort flag is
// raised (sticky: later exchanges fail at once) and the caller leaves with a failure code instead
// of hanging the node.
#pragma once
#include "common.cuh"

namespace effq {

// One rank's contribution to one exchange: three doubles as six (data32, flag32) words.  A naturally aligned 8-byte
// store is single-copy atomic, so a word whose flag half carries the sequence number also carries that exchange's data
// half: no fence between "data" and "flag", no second round trip (the protocol of NCCL's LL path).  The first version
// (3 doubles, __threadfence_system, then a sequence word) paid a system fence -- a full NVLink round trip -- per exchange.
struct PeerSlot {
  unsigned long long w[8];          // w[2 i] = lo32(v[i]), w[2 i + 1] = hi32(v[i]), each | (flag32 << 32); w[6..7] pad
};
constexpr int EFFQ_PEER_CTR_OFFSET = EFFQ_PEER_CHANNELS * EFFQ_PEER_MAX * 2 * (int)sizeof(PeerSlot);
// A missing peer must end in an error, not in a hung node: the wait for a rank's sequence word is bounded by WALL
// TIME (%globaltimer, nanoseconds), not by a spin count whose duration depends on clocks and contention.  30 s covers
// the largest legitimate skew between ranks inside a layer (first-call module load, an fp64 factorisation fallback
// on one rank's side stream); the abort flag is sticky for the rest of the process (every later exchange fails at
// once, the layer engine raises), because the sequence counters of the ranks are no longer in step.
constexpr unsigned long long PEER_TIMEOUT_NS = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned long long* peer_counter(const effq_peer_comm& c, int ch) {
  return (unsigned long long*)((char*)c.slots[c.rank] + EFFQ_PEER_CTR_OFFSET) + ch;
}
__device__ __forceinline__ unsigned long long* peer_aborted(const effq_peer_comm& c) {
  return (unsigned long long*)((char*)c.slots[c.rank] + EFFQ_PEER_CTR_OFFSET) + EFFQ_PEER_CHANNELS;
}

// ONE thread of the rank: publish vals[0..2] as exchange `seq` of channel `ch` in every rank's buffer (own included).
__device__ __forceinline__ void peer_send3(const effq_peer_comm& c, int ch, unsigned long long seq, const double* vals) {
  const int par = (int)(seq & 1ull);
  const unsigned long long flag = (seq & 0xffffffffull) << 32;
  unsigned long long w[6];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[i]);
    w[2 * i] = flag | (bits & 0xffffffffull);
    w[2 * i + 1] = flag | (bits >> 32);
  }
  for (int p = 0; p < c.world; ++p) {
    volatile PeerSlot* s = (volatile PeerSlot*)c.slots[p] + ((ch * EFFQ_PEER_MAX + c.rank) * 2 + par);
#pragma unroll
    for (int k = 0; k < 6; ++k) s->w[k] = w[k];
  }
}

// Any thread (several CTAs may call it for the same exchange): wait for exchange `seq` of every rank in the LOCAL
// buffer and add the contributions in rank order -- the same bits in every caller and on every rank.
__device__ __forceinline__ bool peer_recv3(const effq_peer_comm& c, int ch, unsigned long long seq, double* out) {
  char* mine = (char*)c.slots[c.rank];
  unsigned long long* aborted = peer_aborted(c);
  const int par = (int)(seq & 1ull);
  const unsigned long long flag = seq & 0xffffffffull;
  double acc[3] = {0.0, 0.0, 0.0};
  bool ok = true;
  for (int r = 0; r < c.world && ok; ++r) {
    volatile PeerSlot* s = (volatile PeerSlot*)mine + ((ch * EFFQ_PEER_MAX + r) * 2 + par);
    unsigned long long w[6];
    unsigned long long spins = 0, t0 = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      while (((w[k] = s->w[k]) >> 32) != flag) {
        if ((++spins & 0xfffull) == 0) {                     // look at the clock every 4096 polls
          if (*(volatile unsigned long long*)aborted != 0ull) { ok = false; break; }
          const unsigned long long now = peer_now_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > PEER_TIMEOUT_NS) { ok = false; *aborted = 1ull; break; }
        }
      }
      if (!ok) break;
    }
    if (!ok) break;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      acc[i] += __longlong_as_double((long long)((w[2 * i] & 0xffffffffull) | (w[2 * i + 1] << 32)));
  }
  out[0] = acc[0];
  out[1] = acc[1];
  out[2] = acc[2];
  return ok;
}

// Called by ONE thread.  vals[0..2] in: local contribution; out: sum over ranks (rank order).
__device__ __forceinline__ bool peer_allreduce3(const effq_peer_comm& c, int ch, double* vals) {
  unsigned long long* ctr = peer_counter(c, ch);
  if (*(volatile unsigned long long*)peer_aborted(c) != 0ull) return false;
  const unsigned long long seq = *ctr + 1ull;
  *ctr = seq;
  peer_send3(c, ch, seq, vals);
  return peer_recv3(c, ch, seq, vals);
}

}  // namespace effq
