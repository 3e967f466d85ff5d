// Weight standardisation (forward + backward) fused with the tap-major packing the conv kernels consume.
// Reference: Conv3d.forward, unet3D.py:22-26 -- per out-channel: c = w - mean(w); w_hat = c / sqrt(var_unbiased(c) + 1e-12),
// recomputed every forward, gradients flow through it.  The reference spends ~8 ATen launches per conv per step on
// this; here it is one launch per direction, one block per out-channel, fp64 block reductions.
#include "common.cuh"

namespace mmpl {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* scratch) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0;
  for (int i = 0; i < kThreads / 32; ++i) t += scratch[i];
  return t;
}

// One out-channel (co) of one convolution.  stem_kch > 0 selects the stem packing (Cin = 1): pf is [cout][stem_kch]
// with the 27 taps in columns 0..26 and, for stem_kch = 64, again in columns 32..58 (hi/lo image split, see
// mmpl_stem_im2col); the padding columns are never written and must be zero.
template <typename T>
__device__ __forceinline__ void ws_fwd_one(const float* __restrict__ w, int cout, int cin, int taps, int standardise,
                                           float* __restrict__ w_hat, float* __restrict__ inv_std, T* __restrict__ pf,
                                           T* __restrict__ pd, int co, int stem_kch, double* scratch) {
  const int n = cin * taps;
  const float* wr = w + static_cast<int64_t>(co) * n;
  float mean = 0.f, istd = 1.f;
  if (standardise) {
    double s = 0;
    for (int i = threadIdx.x; i < n; i += kThreads) s += wr[i];
    const double mu = block_sum(s, scratch) / n;
    mean = static_cast<float>(mu);
    // variance of the centred fp32 weight (unbiased), as torch.var(weight.view(O,-1), dim=1) sees it
    double s1 = 0, s2 = 0;
    for (int i = threadIdx.x; i < n; i += kThreads) {
      const double c = static_cast<double>(wr[i] - mean);
      s1 += c;
      s2 += c * c;
    }
    const double t1 = block_sum(s1, scratch), t2 = block_sum(s2, scratch);
    const double var = (t2 - t1 * t1 / n) / (n > 1 ? n - 1 : 1);
    istd = static_cast<float>(1.0 / sqrt(var + 1e-12));
  }
  if (threadIdx.x == 0 && inv_std) inv_std[co] = istd;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const int ci = i / taps, t = i - ci * taps;
    const float wh = standardise ? (wr[i] - mean) * istd : wr[i];
    if (w_hat) w_hat[static_cast<int64_t>(co) * n + i] = wh;
    if (stem_kch > 0) {
      if (pf) {
        pf[static_cast<int64_t>(co) * stem_kch + t] = from_f32<T>(wh);
        if (stem_kch >= 64) pf[static_cast<int64_t>(co) * stem_kch + 32 + t] = from_f32<T>(wh);
      }
    } else {
      if (pf) pf[(static_cast<int64_t>(t) * cout + co) * cin + ci] = from_f32<T>(wh);
      if (pd) pd[(static_cast<int64_t>(taps - 1 - t) * cin + ci) * cout + co] = from_f32<T>(wh);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
ws_fwd_kernel(const float* __restrict__ w, int cout, int cin, int taps, int standardise, float* __restrict__ w_hat,
              float* __restrict__ inv_std, T* __restrict__ pf, T* __restrict__ pd) {
  __shared__ double scratch[kThreads / 32];
  ws_fwd_one<T>(w, cout, cin, taps, standardise, w_hat, inv_std, pf, pd, blockIdx.x, 0, scratch);
}

// All convolutions of a network in ONE launch: block b serves out-channel (b - first_block) of the table entry that
// contains b.  The table lives in device memory (mmpl_ws_entry, include/mmpl_b200.h).
template <typename T>
__global__ void __launch_bounds__(kThreads)
ws_fwd_batched_kernel(const mmpl_ws_entry* __restrict__ table, int count) {
  __shared__ double scratch[kThreads / 32];
  __shared__ mmpl_ws_entry e;
  if (threadIdx.x == 0) {
    int lo = 0;
    const int b = blockIdx.x;
    for (int i = 1; i < count; ++i)
      if (table[i].first_block <= b) lo = i;
    e = table[lo];
  }
  __syncthreads();
  ws_fwd_one<T>(e.w, e.cout, e.cin, e.taps, e.standardise, e.w_hat, e.inv_std, static_cast<T*>(e.packed_fprop),
                static_cast<T*>(e.packed_dgrad), blockIdx.x - e.first_block, e.stem_kch, scratch);
}

__global__ void __launch_bounds__(kThreads)
ws_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w_hat, const float* __restrict__ inv_std, int cout,
              int cin, int taps, int standardise, float* __restrict__ dw) {
  __shared__ double scratch[kThreads / 32];
  const int co = blockIdx.x, n = cin * taps;
  double s1 = 0, s2 = 0;
  if (standardise) {
    for (int i = threadIdx.x; i < n; i += kThreads) {
      const int ci = i / taps, t = i - ci * taps;
      const double gv = g[(static_cast<int64_t>(t) * cout + co) * cin + ci];
      s1 += gv;
      s2 += gv * static_cast<double>(w_hat[static_cast<int64_t>(co) * n + i]);
    }
    s1 = block_sum(s1, scratch) / n;
    s2 = block_sum(s2, scratch) / (n > 1 ? n - 1 : 1);
  }
  const float gm = static_cast<float>(s1), gw = static_cast<float>(s2), istd = standardise ? inv_std[co] : 1.f;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const int ci = i / taps, t = i - ci * taps;
    const float gv = g[(static_cast<int64_t>(t) * cout + co) * cin + ci];
    dw[static_cast<int64_t>(co) * n + i] =
        standardise ? (gv - gm - w_hat[static_cast<int64_t>(co) * n + i] * gw) * istd : gv;
  }
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_ws_weight_fwd(const float* w, int cout, int cin, int taps, int standardise, float* w_hat,
                                  float* inv_std, void* packed_fprop, void* packed_dgrad, int dtype,
                                  mmpl_stream_t stream) {
  MMPL_REQUIRE(cout > 0 && cin > 0 && (taps == 1 || taps == 27), MMPL_E_SHAPE, "ws_weight: cout=%d cin=%d taps=%d", cout,
               cin, taps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (ws_fwd_kernel<T><<<cout, kThreads, 0, s>>>(w, cout, cin, taps, standardise, w_hat, inv_std,
                                                                        static_cast<T*>(packed_fprop),
                                                                        static_cast<T*>(packed_dgrad))));
  MMPL_CHECK_LAUNCH("ws_weight_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_ws_weight_fwd_batched(const mmpl_ws_entry* table_dev, int count, int total_blocks, int dtype,
                                          mmpl_stream_t stream) {
  MMPL_REQUIRE(table_dev != nullptr && count > 0 && total_blocks > 0, MMPL_E_SHAPE, "ws_weight_fwd_batched: empty table");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (ws_fwd_batched_kernel<T><<<total_blocks, kThreads, 0, s>>>(table_dev, count)));
  MMPL_CHECK_LAUNCH("ws_weight_fwd_batched");
  return MMPL_OK;
}

extern "C" int mmpl_ws_weight_bwd(const float* g_hat_tapmajor, const float* w_hat, const float* inv_std, int cout,
                                  int cin, int taps, int standardise, float* dw, mmpl_stream_t stream) {
  MMPL_REQUIRE(cout > 0 && cin > 0 && (taps == 1 || taps == 27), MMPL_E_SHAPE, "ws_weight: cout=%d cin=%d taps=%d", cout,
               cin, taps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ws_bwd_kernel<<<cout, kThreads, 0, s>>>(g_hat_tapmajor, w_hat, inv_std, cout, cin, taps, standardise, dw);
  MMPL_CHECK_LAUNCH("ws_weight_bwd");
  return MMPL_OK;
}
