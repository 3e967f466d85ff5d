// Sliding-window blending and the argmax/Dice reduction, on the device.
// Reference: predict_sliding, evaluate_amos.py:261-279 (prediction *= gaussian; full += prediction; count += gaussian;
// full /= count, all in float64 on the CPU with a D2H copy per tile) and get_dice/dice_score, :92-102, :128-141
// (argmax of softmax == argmax of logits; per class 2|P&T|/(|P|+|T|+1)).
#include "common.cuh"

namespace mmpl {
namespace {

template <typename A>
__global__ void __launch_bounds__(256)
sw_blend_kernel(A* __restrict__ acc, A* __restrict__ wsum, const float* __restrict__ tile, const float* __restrict__ g,
                int C, int D, int H, int W, int td, int th, int tw, int d0, int h0, int w0) {
  const int64_t tvox = static_cast<int64_t>(td) * th * tw;
  const int64_t vol = static_cast<int64_t>(D) * H * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < tvox;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % tw);
    const int y = static_cast<int>((i / tw) % th);
    const int z = static_cast<int>(i / (static_cast<int64_t>(tw) * th));
    const int64_t o = (static_cast<int64_t>(d0 + z) * H + (h0 + y)) * W + (w0 + x);
    const float gv = g[i];
    wsum[o] += static_cast<A>(gv);
    for (int c = 0; c < C; ++c) {
      // the reference multiplies in fp32 (prediction *= map) and accumulates in the accumulator type
      acc[c * vol + o] += static_cast<A>(tile[c * tvox + i] * gv);
    }
  }
}

template <typename A>
__global__ void __launch_bounds__(256)
sw_finalize_kernel(const A* __restrict__ acc, const A* __restrict__ wsum, const float* __restrict__ label,
                   float* __restrict__ out_logits, uint8_t* __restrict__ argmax, unsigned long long* __restrict__ counts,
                   int C, int64_t vol) {
  __shared__ unsigned long long s_cnt[3][32];
  if (threadIdx.x < 96) (&s_cnt[0][0])[threadIdx.x] = 0ull;
  __syncthreads();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < vol;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const A ws = wsum ? wsum[i] : static_cast<A>(1);
    int best = 0;
    A bv = acc[i] / ws;
    if (out_logits) out_logits[i] = static_cast<float>(bv);
    for (int c = 1; c < C; ++c) {
      const A v = acc[c * vol + i] / ws;
      if (out_logits) out_logits[c * vol + i] = static_cast<float>(v);
      if (v > bv) bv = v, best = c;  // first maximum wins, like torch.argmax
    }
    if (argmax) argmax[i] = static_cast<uint8_t>(best);
    if (label) {
      const float lv = label[i];
      const int li = static_cast<int>(lv);
      const bool lvalid = static_cast<float>(li) == lv && li >= 0 && li < C;
      atomicAdd(&s_cnt[1][best], 1ull);
      if (lvalid) {
        atomicAdd(&s_cnt[2][li], 1ull);
        if (li == best) atomicAdd(&s_cnt[0][best], 1ull);
      }
    }
  }
  __syncthreads();
  if (label && threadIdx.x < 96) {
    const int k = threadIdx.x / 32, c = threadIdx.x % 32;
    if (c < C && s_cnt[k][c]) atomicAdd(&counts[k * C + c], s_cnt[k][c]);
  }
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_sw_blend(void* acc, void* wsum, const float* tile_logits, const float* gauss, int c, int d, int h,
                             int w, int td, int th, int tw, int d0, int h0, int w0, int acc_bytes,
                             mmpl_stream_t stream) {
  MMPL_REQUIRE(d0 >= 0 && h0 >= 0 && w0 >= 0 && d0 + td <= d && h0 + th <= h && w0 + tw <= w, MMPL_E_SHAPE,
               "sw_blend: tile (%d,%d,%d)+(%d,%d,%d) outside volume (%d,%d,%d)", d0, h0, w0, td, th, tw, d, h, w);
  MMPL_REQUIRE(acc_bytes == 4 || acc_bytes == 8, MMPL_E_DTYPE, "sw_blend: acc_bytes=%d", acc_bytes);
  const int64_t tvox = static_cast<int64_t>(td) * th * tw;
  const int blocks = static_cast<int>(std::min<int64_t>((tvox + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (acc_bytes == 8)
    sw_blend_kernel<double><<<blocks, 256, 0, s>>>(static_cast<double*>(acc), static_cast<double*>(wsum), tile_logits,
                                                  gauss, c, d, h, w, td, th, tw, d0, h0, w0);
  else
    sw_blend_kernel<float><<<blocks, 256, 0, s>>>(static_cast<float*>(acc), static_cast<float*>(wsum), tile_logits, gauss,
                                                 c, d, h, w, td, th, tw, d0, h0, w0);
  MMPL_CHECK_LAUNCH("sw_blend");
  return MMPL_OK;
}

extern "C" int mmpl_sw_finalize(const void* acc, const void* wsum, const float* label, float* out_logits,
                                uint8_t* argmax, long long* counts, int c, int64_t voxels, int acc_bytes,
                                mmpl_stream_t stream) {
  MMPL_REQUIRE(c >= 1 && c <= 32, MMPL_E_SHAPE, "sw_finalize: classes=%d", c);
  MMPL_REQUIRE(acc_bytes == 4 || acc_bytes == 8, MMPL_E_DTYPE, "sw_finalize: acc_bytes=%d", acc_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (counts) MMPL_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * 3 * c, s));
  const int blocks = static_cast<int>(std::min<int64_t>((voxels + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  const float* lab = counts ? label : nullptr;
  if (acc_bytes == 8)
    sw_finalize_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(acc), static_cast<const double*>(wsum), lab,
                                                     out_logits, argmax, reinterpret_cast<unsigned long long*>(counts), c,
                                                     voxels);
  else
    sw_finalize_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(acc), static_cast<const float*>(wsum), lab,
                                                    out_logits, argmax, reinterpret_cast<unsigned long long*>(counts), c,
                                                    voxels);
  MMPL_CHECK_LAUNCH("sw_finalize");
  return MMPL_OK;
}
