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

// All convolutions of a network in ONE launch.  A block serves EIGHT consecutive out-channels of the table entry that
// contains it (entry i owns blocks [first_block, first_block + ceil(cout/8))), so that both packings are written with
// coalesced stores: warp q computes mean / inv_std of out-channel co0+q, then the block walks the input channels in
// chunks of 32, stages the standardised values of the 8 x 32 x taps sub-filter in shared memory and writes
//   pf [tap][co][ci]   as 32 consecutive ci per (tap, co)            (64-byte segments)
//   pd [tap'][ci][co]  as 8 consecutive co per (tap', ci)            (16-byte vectors).
constexpr int kCoPerBlock = 8, kCiChunk = 32;

template <typename T>
__global__ void __launch_bounds__(kThreads)
ws_fwd_batched_kernel(const mmpl_ws_entry* __restrict__ table, int count) {
  __shared__ mmpl_ws_entry e;
  __shared__ float s_stat[kCoPerBlock][2];
  __shared__ float s_w[kCoPerBlock][kCiChunk * 27 + 1];
  if (threadIdx.x == 0) {
    int lo = 0;
    const int b = blockIdx.x;
    for (int i = 1; i < count; ++i)
      if (table[i].first_block <= b) lo = i;
    e = table[lo];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = (blockIdx.x - e.first_block) * kCoPerBlock;
  const int cout = e.cout, cin = e.cin, taps = e.taps, n = cin * taps;
  const int nco = min(kCoPerBlock, cout - co0);
  T* pf = static_cast<T*>(e.packed_fprop);
  T* pd = static_cast<T*>(e.packed_dgrad);
  // ---- statistics: warp q <-> out-channel co0 + q
  if (warp < nco) {
    const float* wr = e.w + static_cast<int64_t>(co0 + warp) * n;
    float mean = 0.f, istd = 1.f;
    if (e.standardise) {
      double sum = 0;
      for (int i = lane; i < n; i += 32) sum += wr[i];
      mean = static_cast<float>(warp_sum(sum) / n);
      double s1 = 0, s2 = 0;      // variance of the centred fp32 weight (unbiased), as torch.var(w.view(O,-1), dim=1) sees it
      for (int i = lane; i < n; i += 32) {
        const double c = static_cast<double>(wr[i] - mean);
        s1 += c;
        s2 += c * c;
      }
      const double t1 = warp_sum(s1), t2 = warp_sum(s2);
      const double var = (t2 - t1 * t1 / n) / (n > 1 ? n - 1 : 1);
      istd = static_cast<float>(1.0 / sqrt(var + 1e-12));
    }
    if (lane == 0) {
      s_stat[warp][0] = mean, s_stat[warp][1] = istd;
      if (e.inv_std) e.inv_std[co0 + warp] = istd;
    }
  }
  __syncthreads();
  if (e.stem_kch > 0) {   // Cin = 1 stem: [cout][stem_kch] with the taps in columns 0..26 (and 32..58)
    if (warp < nco && lane < taps) {
      const int co = co0 + warp;
      const float wv = e.w[static_cast<int64_t>(co) * n + lane];
      const float wh = e.standardise ? (wv - s_stat[warp][0]) * s_stat[warp][1] : wv;
      if (e.w_hat) e.w_hat[static_cast<int64_t>(co) * n + lane] = wh;
      if (pf) {
        pf[static_cast<int64_t>(co) * e.stem_kch + lane] = from_f32<T>(wh);
        if (e.stem_kch >= 64) pf[static_cast<int64_t>(co) * e.stem_kch + 32 + lane] = from_f32<T>(wh);
      }
    }
    return;
  }
  for (int ci0 = 0; ci0 < cin; ci0 += kCiChunk) {
    const int cic = min(kCiChunk, cin - ci0), m = cic * taps;     // sub-filter: nco x cic x taps
    __syncthreads();
    for (int idx = threadIdx.x; idx < nco * m; idx += kThreads) {
      const int q = idx / m, r = idx - q * m;
      const int64_t src = static_cast<int64_t>(co0 + q) * n + static_cast<int64_t>(ci0) * taps + r;
      const float wv = e.w[src];
      const float wh = e.standardise ? (wv - s_stat[q][0]) * s_stat[q][1] : wv;
      if (e.w_hat) e.w_hat[src] = wh;
      s_w[q][r] = wh;
    }
    __syncthreads();
    if (pf) {
      for (int idx = threadIdx.x; idx < taps * nco * cic; idx += kThreads) {
        const int cil = idx % cic, q = (idx / cic) % nco, t = idx / (cic * nco);
        pf[(static_cast<int64_t>(t) * cout + co0 + q) * cin + ci0 + cil] = from_f32<T>(s_w[q][cil * taps + t]);
      }
    }
    if (pd) {
      for (int idx = threadIdx.x; idx < taps * cic; idx += kThreads) {
        const int cil = idx % cic, t = idx / cic;
        T* dst = pd + (static_cast<int64_t>(taps - 1 - t) * cin + ci0 + cil) * cout + co0;
        if (nco == kCoPerBlock) {
          constexpr int VN = Vec<T>::N;     // 8 bf16 = one 16-byte store, 2 x 4 fp32 = two
#pragma unroll
          for (int h = 0; h < kCoPerBlock / VN; ++h) {
            Vec<T> v;
#pragma unroll
            for (int k = 0; k < VN; ++k) v.v[k] = s_w[h * VN + k][cil * taps + t];
            v.store(dst + h * VN);
          }
        } else {
          for (int q = 0; q < nco; ++q) dst[q] = from_f32<T>(s_w[q][cil * taps + t]);
        }
      }
    }
  }
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
  // total_blocks = sum over the entries of ceil(cout / 8) (a block serves 8 out-channels)
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
