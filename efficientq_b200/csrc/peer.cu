// Peer-memory plumbing for the in-kernel exchanges of peer.cuh: allocation of the per-rank slot
// buffer, CUDA IPC export / import.  One process per GPU (torchrun); the 64-byte handles travel
// through torch.distributed (efficientq_b200/dist.py).
#include "peer.cuh"
#include <string.h>

extern "C" int64_t effq_peer_bytes(void) { return 4096; }

extern "C" int effq_peer_alloc(void** dev_ptr, uint8_t* handle64) {
  using namespace effq;
  EFFQ_CHECK_ARG(dev_ptr && handle64, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  static_assert(EFFQ_PEER_CTR_OFFSET + (EFFQ_PEER_CHANNELS + 1) * 8 <= 4096, "slot buffer size");
  void* p = nullptr;
  EFFQ_CUDA(cudaMalloc(&p, (size_t)effq_peer_bytes()));
  EFFQ_CUDA(cudaMemset(p, 0, (size_t)effq_peer_bytes()));
  EFFQ_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  EFFQ_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}

extern "C" int effq_peer_open(const uint8_t* handle64, void** dev_ptr) {
  using namespace effq;
  EFFQ_CHECK_ARG(dev_ptr && handle64, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  EFFQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return 0;
}

extern "C" int effq_peer_close(void* dev_ptr) {
  using namespace effq;
  if (dev_ptr) EFFQ_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}

extern "C" int effq_peer_free(void* dev_ptr) {
  using namespace effq;
  if (dev_ptr) EFFQ_CUDA(cudaFree(dev_ptr));
  return 0;
}
