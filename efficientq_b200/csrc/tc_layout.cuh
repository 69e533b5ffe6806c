// Operand layouts shared by the tensor-core kernels and the kernels that produce their
// operands (effq_quantize_act_ndhwc, effq_admm_project, effq_pack_wcodes).
//
// Codes are stored as bf16 (2 bytes, exact for <= 256 levels) or as e4m3 bytes (exact for <= 16
// levels: 4 significant bits; K = 32 per tcgen05.mma instead of 16).  Channels are processed in
// blocks of CG = min(C1, 128 bytes / element size).  A "row" is the CG channels of one voxel
// (activations) or of one output channel (weights): RP = CG * element size = 128 / 64 / 32 bytes,
// stored with the UMMA shared-memory swizzle of the same width (SWIZZLE_128B / 64B / 32B): the
// 16-byte chunk index of a row is XOR-ed with a function of the row index.  Measured on B200: the
// hardware applies the XOR to the absolute shared-memory address, so descriptors may start at any
// row of a block whose base is 1024-byte aligned (profiles/r01_conv_layout.md).
#pragma once
#include <stdint.h>

namespace effq {

enum CodeDtype { CODE_BF16 = 0, CODE_E4M3 = 1 };

struct TcLayout {
  int eb;        // bytes per code (2 = bf16, 1 = e4m3)
  int epc;       // codes per 16-byte chunk (8 / 16)
  int cg;        // channels per block
  int nch;       // 16-byte chunks per row
  int rp;        // row pitch in bytes (cg * eb): 32 / 64 / 128
  int groups;    // channel blocks (c1 / cg)
  int swz;       // swizzle width in bytes (== rp)
};

__host__ __device__ inline TcLayout tc_layout(int c1, int code_dtype) {
  TcLayout l;
  l.eb = code_dtype == CODE_E4M3 ? 1 : 2;
  l.epc = 16 / l.eb;
  const int cg_max = 128 / l.eb;
  l.cg = c1 < cg_max ? c1 : cg_max;
  l.nch = l.cg / l.epc;
  l.rp = l.cg * l.eb;
  l.groups = c1 / l.cg;
  l.swz = l.rp;
  return l;
}

// XOR applied to the 16-byte chunk index of row `row` (rows stacked at pitch rp from a
// 1024-byte aligned base): address bits [4,7) ^= bits [7,10) restricted to the swizzle width.
__host__ __device__ inline int tc_chunk_xor(int row, int swz) {
  return swz == 128 ? (row & 7) : (swz == 64 ? ((row >> 1) & 3) : (swz == 32 ? ((row >> 2) & 1) : 0));
}

// Element offset of weight code (out channel r, in channel c, tap t):
//   [tap][group][C2 rows][nch chunks (swizzled)][epc]
__host__ __device__ inline long long tc_wcode_index(int r, int c, int t, int c1, int c2, const TcLayout& l) {
  const int g = c / l.cg, cl = c % l.cg;
  const int chunk = (cl / l.epc) ^ tc_chunk_xor(r, l.swz);
  return (((long long)t * l.groups + g) * c2 + r) * l.cg + chunk * l.epc + (cl % l.epc);
}

}  // namespace effq
