// Tiny all-reduce of a few doubles between the GPUs of one node, done INSIDE a kernel over
// NVLink peer memory (no NCCL launch, no host round trip).  The sharded calibration exchanges
// 2 doubles per activation-search pass and 1 double per ADMM iteration -- thousands of
// latency-bound exchanges per layer; as NCCL calls they cost more than the kernels around them.
//
// Every rank owns one small buffer (effq_peer_alloc) that all peers map (CUDA IPC):
//   PeerSlot slot[EFFQ_PEER_CHANNELS][EFFQ_PEER_MAX][2];   // [channel][source rank][parity]
//   unsigned long long next_seq[EFFQ_PEER_CHANNELS];       // local only, at EFFQ_PEER_CTR_OFFSET
//   unsigned long long aborted;                            // local only: sticky after a timeout
// An exchange with sequence number s: write my values into slot[ch][my rank][s & 1] of every rank
// (data, system fence, then the sequence word), spin until the local slots of all ranks carry s,
// add them in rank order (bit-identical result on every rank).  Two parities suffice: a rank can
// only be one exchange ahead of the slowest.  Waits are bounded (30 s of wall time); on timeout the abort flag is
// raised (sticky: later exchanges fail at once) and the caller leaves with a failure code instead
// of hanging the node.
#pragma once
#include "common.cuh"

namespace effq {

struct PeerSlot {
  double v[3];
  unsigned long long seq;
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

// Called by ONE thread.  vals[0..2] in: local contribution; out: sum over ranks (rank order).
__device__ __forceinline__ bool peer_allreduce3(const effq_peer_comm& c, int ch, double* vals) {
  char* mine = (char*)c.slots[c.rank];
  unsigned long long* ctr = (unsigned long long*)(mine + EFFQ_PEER_CTR_OFFSET) + ch;
  unsigned long long* aborted = (unsigned long long*)(mine + EFFQ_PEER_CTR_OFFSET) + EFFQ_PEER_CHANNELS;
  if (*aborted != 0ull) return false;
  const unsigned long long seq = *ctr + 1ull;
  *ctr = seq;
  const int par = (int)(seq & 1ull);
  for (int p = 0; p < c.world; ++p) {
    volatile PeerSlot* s = (volatile PeerSlot*)c.slots[p] + ((ch * EFFQ_PEER_MAX + c.rank) * 2 + par);
    s->v[0] = vals[0];
    s->v[1] = vals[1];
    s->v[2] = vals[2];
  }
  __threadfence_system();
  for (int p = 0; p < c.world; ++p) {
    volatile PeerSlot* s = (volatile PeerSlot*)c.slots[p] + ((ch * EFFQ_PEER_MAX + c.rank) * 2 + par);
    s->seq = seq;
  }
  double acc[3] = {0.0, 0.0, 0.0};
  bool ok = true;
  for (int r = 0; r < c.world; ++r) {
    volatile PeerSlot* s = (volatile PeerSlot*)mine + ((ch * EFFQ_PEER_MAX + r) * 2 + par);
    unsigned long long spins = 0, t0 = 0;
    while (s->seq != seq) {
      if ((++spins & 0xfffull) == 0) {                       // look at the clock every 4096 polls
        const unsigned long long now = peer_now_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > PEER_TIMEOUT_NS) { ok = false; *aborted = 1ull; break; }
      }
    }
    if (!ok) break;
    __threadfence_system();
    acc[0] += s->v[0];
    acc[1] += s->v[1];
    acc[2] += s->v[2];
  }
  vals[0] = acc[0];
  vals[1] = acc[1];
  vals[2] = acc[2];
  return ok;
}

}  // namespace effq
