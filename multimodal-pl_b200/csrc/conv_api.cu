// C-ABI dispatch for the convolution entry points (see include/mmpl_b200.h).
#include "common.cuh"

namespace mmpl {
int conv_direct_fprop(const void*, const void*, const void*, void*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv_direct_dgrad(const void*, const void*, const void*, void*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv_direct_wgrad(const void*, const void*, float*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int conv_tc_s1(const void*, const void*, const void*, void*, int, int, int, int, int, int, int, double*, int*,
               const mmpl_gn_bwd_fuse*, cudaStream_t);
bool conv_tc_can_fuse_gn_bwd(int nout);
int conv_tc_s2_fprop(const void*, const void*, const void*, void*, int, int, int, int, int, int, int, double*, int*,
                     cudaStream_t);
int conv_out_dim(int in, int k, int stride);
int conv_tc_s2_dgrad(const void*, const void*, void*, int, int, int, int, int, int, int, const mmpl_gn_bwd_fuse*, cudaStream_t);
int parity_split(const void*, void*, int, int, int, int, int, cudaStream_t);
int conv_tc_wgrad(const void*, const void*, float*, int, int, int, int, int, int, int, int, cudaStream_t);
size_t conv_tc_wgrad_workspace(int, int, int, int, int, int);

static int check_common(int n, int d, int h, int w, int cin, int cout, int k, int stride) {
  MMPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "conv3d: empty tensor [%d,%d,%d,%d]", n, d, h, w);
  MMPL_REQUIRE(k == 1 || k == 3, MMPL_E_SHAPE, "conv3d: kernel size %d (1 or 3)", k);
  MMPL_REQUIRE(stride == 1 || stride == 2, MMPL_E_SHAPE, "conv3d: stride %d (1 or 2)", stride);
  MMPL_REQUIRE(cin % 32 == 0 && cout % 32 == 0, MMPL_E_SHAPE, "conv3d: cin=%d cout=%d must be multiples of 32", cin, cout);
  return MMPL_OK;
}
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_parity_split(const void* x, void* p_out, int n, int d, int h, int w, int c, int dtype,
                                 mmpl_stream_t stream) {
  MMPL_REQUIRE(dtype == MMPL_BF16, MMPL_E_DTYPE, "parity_split: bf16 only");
  MMPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "parity_split: empty tensor");
  return parity_split(x, p_out, n, d, h, w, c, static_cast<cudaStream_t>(stream));
}

extern "C" int mmpl_conv3d_fprop(const void* x, const void* w_fprop, const void* residual, void* y, int n, int d, int h,
                                 int w, int cin, int cout, int ksize, int stride, int dtype, int algo,
                                 double* gn_stats_out, mmpl_stream_t stream) {
  if (int e = check_common(n, d, h, w, cin, cout, ksize, stride)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t out_spatial = static_cast<int64_t>(conv_out_dim(d, ksize, stride)) * conv_out_dim(h, ksize, stride) *
                              conv_out_dim(w, ksize, stride);
  if (algo == MMPL_ALGO_TCGEN05 || algo == MMPL_ALGO_TCGEN05_PSPLIT) {
    MMPL_REQUIRE(dtype == MMPL_BF16, MMPL_E_UNSUPPORTED, "conv3d_fprop: tcgen05 path needs bf16 (got dtype=%d)", dtype);
    int fused = 0;   // set by the launch when its epilogue produced the statistics (one tile spans all output channels)
    int rc;
    if (stride == 1) {
      rc = conv_tc_s1(x, w_fprop, residual, y, n, d, h, w, cin, cout, ksize, gn_stats_out, &fused, nullptr, s);
    } else {
      MMPL_REQUIRE((ksize == 3) == (algo == MMPL_ALGO_TCGEN05_PSPLIT), MMPL_E_UNSUPPORTED,
                   "conv3d_fprop: stride-2 3x3x3 takes the parity-split input (MMPL_ALGO_TCGEN05_PSPLIT), 1x1x1 takes x");
      rc = conv_tc_s2_fprop(x, w_fprop, residual, y, n, d, h, w, cin, cout, ksize, gn_stats_out, &fused, s);
    }
    if (rc || !gn_stats_out || fused) return rc;
    return mmpl_gn_stats(y, gn_stats_out, n, out_spatial, cout, 16, dtype, stream);
  }
  if (int e = conv_direct_fprop(x, w_fprop, residual, y, n, d, h, w, cin, cout, ksize, stride, dtype, s)) return e;
  if (gn_stats_out) return mmpl_gn_stats(y, gn_stats_out, n, out_spatial, cout, 16, dtype, stream);
  return MMPL_OK;
}

extern "C" int mmpl_conv3d_dgrad(const void* dy, const void* w_dgrad, const void* addend, void* dx, int n, int d, int h,
                                 int w, int cin, int cout, int ksize, int stride, int dtype, int algo,
                                 const mmpl_gn_bwd_fuse* gn, int* gn_fused_out, mmpl_stream_t stream) {
  if (gn_fused_out) *gn_fused_out = 0;
  if (int e = check_common(n, d, h, w, cin, cout, ksize, stride)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (algo == MMPL_ALGO_TCGEN05) {
    MMPL_REQUIRE(dtype == MMPL_BF16, MMPL_E_UNSUPPORTED, "conv3d_dgrad: tcgen05 path needs bf16 (got dtype=%d)", dtype);
    const mmpl_gn_bwd_fuse* fuse = nullptr;
    if (gn != nullptr && gn->ws != nullptr && conv_tc_can_fuse_gn_bwd(cin)) {
      MMPL_REQUIRE(gn->a && gn->beta && (gn->head == 0 || gn->head == 1), MMPL_E_SHAPE,
                   "conv3d_dgrad: incomplete GroupNorm-backward fusion request");
      // the parity-split copy only exists for (and is only indexed by) the stride-2 3x3x3 dgrad
      if (!gn->a_is_parity_split || (stride == 2 && ksize == 3)) fuse = gn;
    }
    int rc;
    // stride-1 dgrad is a correlation of dy with the flipped/transposed packing: channels swap roles
    if (stride == 1) {
      rc = conv_tc_s1(dy, w_dgrad, addend, dx, n, d, h, w, cout, cin, ksize, nullptr, nullptr, fuse, s);
    } else {
      MMPL_REQUIRE(addend == nullptr, MMPL_E_UNSUPPORTED, "conv3d_dgrad: stride-2 tcgen05 path has no addend input");
      if (ksize == 1)  // only the even parity class receives gradient; the rest of dx is zero
        MMPL_CUDA(cudaMemsetAsync(dx, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(n) * d * h * w * cin, s));
      rc = conv_tc_s2_dgrad(dy, w_dgrad, dx, n, d, h, w, cin, cout, ksize, fuse, s);
    }
    if (rc == MMPL_OK && fuse && gn_fused_out) *gn_fused_out = 1;
    return rc;
  }
  return conv_direct_dgrad(dy, w_dgrad, addend, dx, n, d, h, w, cin, cout, ksize, stride, dtype, s);
}

extern "C" size_t mmpl_conv3d_wgrad_workspace(int n, int d, int h, int w, int cin, int cout, int ksize, int stride,
                                              int algo) {
  (void)n, (void)d, (void)h, (void)w, (void)cin, (void)cout, (void)ksize, (void)stride, (void)algo;
  return 0;  // split-K partials live in TMEM; no global workspace is needed by any algorithm
}

extern "C" int mmpl_conv3d_wgrad(const void* x, const void* dy, float* dw_tapmajor, int n, int d, int h, int w, int cin,
                                 int cout, int ksize, int stride, int dtype, int algo, void* workspace,
                                 size_t workspace_bytes, mmpl_stream_t stream) {
  if (int e = check_common(n, d, h, w, cin, cout, ksize, stride)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (algo == MMPL_ALGO_TCGEN05 || algo == MMPL_ALGO_TCGEN05_PSPLIT) {
    MMPL_REQUIRE(dtype == MMPL_BF16, MMPL_E_UNSUPPORTED, "conv3d_wgrad: tcgen05 path needs bf16 (got dtype=%d)", dtype);
    MMPL_REQUIRE((stride == 2 && ksize == 3) == (algo == MMPL_ALGO_TCGEN05_PSPLIT), MMPL_E_UNSUPPORTED,
                 "conv3d_wgrad: stride-2 3x3x3 takes the parity-split input (MMPL_ALGO_TCGEN05_PSPLIT)");
    MMPL_REQUIRE(!(stride == 2 && ksize == 3) || cout % 64 == 0, MMPL_E_UNSUPPORTED, "conv3d_wgrad: stride-2 cout=%d", cout);
    (void)workspace, (void)workspace_bytes;
    return conv_tc_wgrad(x, dy, dw_tapmajor, n, d, h, w, cin, cout, ksize, stride, s);
  }
  return conv_direct_wgrad(x, dy, dw_tapmajor, n, d, h, w, cin, cout, ksize, stride, dtype, s);
}
