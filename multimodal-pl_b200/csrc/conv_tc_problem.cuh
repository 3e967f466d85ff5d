// Shared between the host API (conv_tc.cu) and the kernel instantiation units (conv_tc_inst.cu): the launch description
// and the dispatch entry point of the tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace mmpl {

enum : int {
  MODE_S1K3 = 0,   // 3x3x3 stride 1 (fprop, or dgrad with the flipped/transposed packing)
  MODE_S1K1 = 1,   // 1x1x1 stride 1
  MODE_S2F = 2,    // 3x3x3 stride 2 fprop from the parity-split input
  MODE_S2D = 3,    // 3x3x3 stride 2 dgrad (one parity class of dX per item, strided store)
  MODE_S2K1F = 4,  // 1x1x1 stride 2 fprop (element-strided TMA straight from NDHWC)
  MODE_S2K1D = 5   // 1x1x1 stride 2 dgrad (only the even parity class is non-zero; dX is pre-zeroed)
};

// Epilogue variants, compiled separately so that no launch drags the others' code through the instruction cache (ncu:
// with one kernel holding all of it, 58 % of the epilogue warps' stall samples of the 64-channel kernels were
// instruction fetches -- 8.8 k of 11.7 k SASS instructions never executed in a given launch):
enum : int {
  EPI_PLAIN = 0,   // tcgen05.ld -> bf16 -> store
  EPI_FWD = 1,     // + residual add and / or GroupNorm statistics of the stored output (forward launches)
  EPI_GN = 2       // + first pass of the GroupNorm+ReLU backward of the producer of this launch's input (dgrad launches)
};

// Problem description for one launch.
//   a        : TMA source (X, dY or the parity-split P), channels = kred (reduction channels)
//   aN,aD..  : extents of the TMA source tensor
//   y        : output [N][D][H][W][nout]
struct TcProblem {
  const void* a;
  int64_t aN;
  int aD, aH, aW;
  const void* wp;
  const void* residual;
  void* y;
  int N, D, H, W;
  int kred, nout;
  double* stats = nullptr;
  const mmpl_gn_bwd_fuse* gn = nullptr;   // fused GroupNorm-backward reduction over the OUTPUT (dgrad launches)
  int* stats_fused = nullptr;             // out: 1 if the launch computed `stats` (one tile spans all output channels)
};


// Picks the tile configuration for (MODE, channels, problem size) and launches; one explicit instantiation per
// (MODE, EPI) pair lives in its own translation unit (conv_tc_inst.cu, see the Makefile).
template <int MODE, int EPI>
int dispatch_tc(const TcProblem& q, cudaStream_t s);

}  // namespace mmpl
