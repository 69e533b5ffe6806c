// Operand layouts shared by the tensor-core kernels and the kernels that produce their
// operands (effq_admm_project, effq_pack_wcodes).
//
// Channels are processed in blocks of CG = min(C1, 64).  A "row" is the CG channels of one
// voxel (activations) or of one output channel (weights): RP = 2*CG bytes = 128 / 64 / 32,
// stored with the UMMA shared-memory swizzle of the same width (SWIZZLE_128B / 64B / 32B):
// the 16-byte chunk index of a row is XOR-ed with a function of the row index.  Measured on
// B200: the
// hardware applies the XOR to the absolute shared-memory address, so descriptors may start at
// any row of a block whose base is 1024-byte aligned (profiles/r01_conv_layout.md).
#pragma once
#include <stdint.h>

namespace effq {

struct TcLayout {
  int cg;        // channels per block (16, 32, 64)
  int nch;       // 16-byte chunks per row (cg / 8)
  int rp;        // row pitch in bytes (2 * cg)
  int groups;    // channel blocks (c1 / cg)
  int swz;       // swizzle width in bytes: 32 / 64 / 128 (== rp)
};

__host__ __device__ inline TcLayout tc_layout(int c1) {
  TcLayout l;
  l.cg = c1 < 64 ? c1 : 64;
  l.nch = l.cg / 8;
  l.rp = l.cg * 2;
  l.groups = c1 / l.cg;
  l.swz = l.rp;
  return l;
}

// XOR applied to the 16-byte chunk index of row `row` (rows stacked at pitch rp from a
// 1024-byte aligned base): address bits [4,7) ^= bits [7,10) restricted to the swizzle width.
__host__ __device__ inline int tc_chunk_xor(int row, int swz) {
  return swz == 128 ? (row & 7) : (swz == 64 ? ((row >> 1) & 3) : (swz == 32 ? ((row >> 2) & 1) : 0));
}

// Element offset (in bf16 elements) of weight code (out channel r, in channel c, tap t).
//   [tap][group][C2 rows][nch chunks (swizzled)][8]
__host__ __device__ inline long long tc_wcode_index(int r, int c, int t, int c1, int c2, const TcLayout& l) {
  const int g = c / l.cg, cl = c % l.cg;
  const int chunk = (cl >> 3) ^ tc_chunk_xor(r, l.swz);
  return (((long long)t * l.groups + g) * c2 + r) * l.cg + chunk * 8 + (cl & 7);
}

}  // namespace effq
