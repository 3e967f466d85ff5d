// Host side of the tcgen05 convolutions: problem set-up per public entry point, epilogue-variant selection, and the
// parity-split helper kernel.  The kernels themselves are instantiated in conv_tc_inst.cu (conv_tc_impl.cuh).
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc_problem.cuh"

namespace mmpl {
namespace {

// parity split: P[p*N + n][d'][h'][w'][c] = X[n][2d'+pd][2h'+ph][2w'+pw][c]  (zero where the source is out of range)
// Grid (ceil(Wp*C/8 / 256), Hp, 8*N*Dp): rows are block-uniform, all index arithmetic is 32-bit.
__global__ void __launch_bounds__(256)
parity_split_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ pout, int N, int D, int H, int W,
                    int C, int Dp, int Hp, int Wp) {
  const int vpv = C / 8;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= Wp * vpv) return;
  const int w = e / vpv, cv = e - w * vpv;
  const int h = blockIdx.y;
  int r = blockIdx.z;
  const int d = r % Dp;
  r /= Dp;
  const int n = r % N, pc = r / N;
  const int sd = 2 * d + (pc >> 2), sh = 2 * h + ((pc >> 1) & 1), sw = 2 * w + (pc & 1);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (sd < D && sh < H && sw < W)
    v = *reinterpret_cast<const uint4*>(x + ((((static_cast<int64_t>(n) * D + sd) * H + sh) * W + sw) * C) + cv * 8);
  *reinterpret_cast<uint4*>(pout + (((static_cast<int64_t>(blockIdx.z) * Hp + h) * Wp + w) * C) + cv * 8) = v;
}

}  // namespace

static int check_align(const void* a, const void* b, const void* c, const void* d) {
  MMPL_REQUIRE((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                reinterpret_cast<uintptr_t>(d)) % 16 == 0,
               MMPL_E_ALIGN, "conv_tc: pointers must be 16-byte aligned");
  return MMPL_OK;
}

// ---- stride 1 (k = 3 or 1): x [N,D,H,W,cin] -> y [N,D,H,W,cout]; also stride-1 dgrad with swapped channel roles
// GroupNorm statistics can be fused into the epilogue when one CTA tile spans all output channels
// The GroupNorm-backward reduction rides on every dgrad launch whose output has <= 512 channels in 16 groups.  Measured
// on B200 (cfg2) with the register-lean epilogue: fusing all layers 14.34 ms/step, only the <= 64-channel layers 14.66,
// none 15.5.  MMPL_GN_FUSE_MAXC narrows it for experiments.
bool conv_tc_can_fuse_gn_bwd(int nout) {
  static const int maxc = [] { const char* e = getenv("MMPL_GN_FUSE_MAXC"); return e ? atoi(e) : 512; }();
  return nout % 32 == 0 && nout <= maxc && nout <= 512;
}

int conv_tc_s1(const void* x, const void* wp, const void* residual, void* y, int N, int D, int H, int W, int cin,
               int cout, int ksize, double* stats, int* stats_fused, const mmpl_gn_bwd_fuse* gn, cudaStream_t s) {
  if (int e = check_align(x, wp, y, residual)) return e;
  TcProblem q{x, N, D, H, W, wp, residual, y, N, D, H, W, cin, cout, stats, gn, stats_fused};
  const int epi = gn != nullptr ? EPI_GN : ((stats != nullptr || residual != nullptr) ? EPI_FWD : EPI_PLAIN);
  if (ksize == 3) {
    if (epi == EPI_GN) return dispatch_tc<MODE_S1K3, EPI_GN>(q, s);
    return epi == EPI_FWD ? dispatch_tc<MODE_S1K3, EPI_FWD>(q, s) : dispatch_tc<MODE_S1K3, EPI_PLAIN>(q, s);
  }
  if (epi == EPI_GN) return dispatch_tc<MODE_S1K1, EPI_GN>(q, s);
  return epi == EPI_FWD ? dispatch_tc<MODE_S1K1, EPI_FWD>(q, s) : dispatch_tc<MODE_S1K1, EPI_PLAIN>(q, s);
}

// ---- stride 2 fprop.  k=3: `src` is the parity-split tensor P [8N][Dp][Hp][Wp][cin]; k=1: `src` is x itself.
int conv_tc_s2_fprop(const void* src, const void* wp, const void* residual, void* y, int N, int D, int H, int W, int cin,
                     int cout, int ksize, double* stats, int* stats_fused, cudaStream_t s) {
  if (int e = check_align(src, wp, y, residual)) return e;
  const int Do = (D + 1) / 2, Ho = (H + 1) / 2, Wo = (W + 1) / 2;   // == (in + 2*pad - k)/2 + 1 for k in {1,3}
  if (ksize == 3) {
    TcProblem q{src, static_cast<int64_t>(8) * N, Do, Ho, Wo, wp, residual, y, N, Do, Ho, Wo, cin, cout, stats, nullptr,
                stats_fused};
    return (stats != nullptr || residual != nullptr) ? dispatch_tc<MODE_S2F, EPI_FWD>(q, s) : dispatch_tc<MODE_S2F, EPI_PLAIN>(q, s);
  }
  TcProblem q{src, N, D, H, W, wp, residual, y, N, Do, Ho, Wo, cin, cout, stats, nullptr, stats_fused};
  return (stats != nullptr || residual != nullptr) ? dispatch_tc<MODE_S2K1F, EPI_FWD>(q, s) : dispatch_tc<MODE_S2K1F, EPI_PLAIN>(q, s);
}

// ---- stride 2 dgrad: dy [N,Do,Ho,Wo,cout] -> dx [N,D,H,W,cin] (k=1: dx must be zero-filled by the caller)
int conv_tc_s2_dgrad(const void* dy, const void* wp_dgrad, void* dx, int N, int D, int H, int W, int cin, int cout,
                     int ksize, const mmpl_gn_bwd_fuse* gn, cudaStream_t s) {
  if (int e = check_align(dy, wp_dgrad, dx, nullptr)) return e;
  const int Do = (D + 1) / 2, Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  TcProblem q{dy, N, Do, Ho, Wo, wp_dgrad, nullptr, dx, N, D, H, W, cout, cin, nullptr, gn};
  if (ksize == 3) return gn != nullptr ? dispatch_tc<MODE_S2D, EPI_GN>(q, s) : dispatch_tc<MODE_S2D, EPI_PLAIN>(q, s);
  return gn != nullptr ? dispatch_tc<MODE_S2K1D, EPI_GN>(q, s) : dispatch_tc<MODE_S2K1D, EPI_PLAIN>(q, s);
}

int parity_split(const void* x, void* pout, int N, int D, int H, int W, int C, cudaStream_t s) {
  MMPL_REQUIRE(C % 8 == 0, MMPL_E_SHAPE, "parity_split: C=%d", C);
  const int Dp = (D + 1) / 2, Hp = (H + 1) / 2, Wp = (W + 1) / 2;
  MMPL_REQUIRE(Hp <= 65535 && static_cast<int64_t>(8) * N * Dp <= 65535, MMPL_E_SHAPE, "parity_split: grid limits");
  const dim3 blocks((Wp * (C / 8) + 255) / 256, Hp, 8 * N * Dp);
  parity_split_kernel<<<blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(pout), N, D,
                                            H, W, C, Dp, Hp, Wp);
  MMPL_CHECK_LAUNCH("parity_split");
  return MMPL_OK;
}

}  // namespace mmpl
