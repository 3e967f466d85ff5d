// Sliding-window blending and the argmax/Dice reduction, on the device.
// Reference: predict_sliding, evaluate_amos.py:261-279 (prediction *= gaussian; full += prediction; count += gaussian;
// full /= count, all in float64 on the CPU with a D2H copy per tile) and get_dice/dice_score, :92-102, :128-141
// (argmax of softmax == argmax of logits; per class 2|P&T|/(|P|+|T|+1)).
#include "common.cuh"

namespace mmpl {
namespace {

// Two accumulator layouts: class-major [C][D][H][W] (what predict_sliding returns, like the reference) and depth-major
// [D][C][H][W] (`plane` = H*W > 0): a depth slab of the latter is ONE contiguous block, which is what the multi-GPU path
// reduce-scatters along D.
__device__ __forceinline__ int64_t acc_index(int c, int z, int64_t o, int C, int D, int64_t plane, bool d_outer) {
  return d_outer ? (static_cast<int64_t>(z) * C + c) * plane + o : (static_cast<int64_t>(c) * D + z) * plane + o;
}

template <typename A>
__global__ void __launch_bounds__(256)
sw_blend_kernel(A* __restrict__ acc, A* __restrict__ wsum, const float* __restrict__ tile, const float* __restrict__ g,
                int C, int D, int H, int W, int td, int th, int tw, int d0, int h0, int w0, int d_outer) {
  const int64_t tvox = static_cast<int64_t>(td) * th * tw;
  const int64_t plane = static_cast<int64_t>(H) * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < tvox;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % tw);
    const int y = static_cast<int>((i / tw) % th);
    const int z = static_cast<int>(i / (static_cast<int64_t>(tw) * th));
    const int64_t o = static_cast<int64_t>(h0 + y) * W + (w0 + x);
    const float gv = g[i];
    if (wsum) wsum[(d0 + z) * plane + o] += static_cast<A>(gv);
    for (int c = 0; c < C; ++c) {
      // the reference multiplies in fp32 (prediction *= map) and accumulates in the accumulator type
      acc[acc_index(c, d0 + z, o, C, D, plane, d_outer != 0)] += static_cast<A>(tile[c * tvox + i] * gv);
    }
  }
}

__device__ __forceinline__ bool label_class(float lv, int C, int& li) {
  li = static_cast<int>(lv);
  return static_cast<float>(li) == lv && li >= 0 && li < C;
}

// `vol` voxels = Dl depth planes x plane; acc is [C][Dl][plane] or, with d_outer, [Dl][C][plane]
template <typename A>
__global__ void __launch_bounds__(256)
sw_finalize_kernel(const A* __restrict__ acc, const A* __restrict__ wsum, const void* __restrict__ label, int label_u8,
                   float* __restrict__ out_logits, uint8_t* __restrict__ argmax, unsigned long long* __restrict__ counts,
                   int C, int64_t vol, int64_t plane, int d_outer) {
  __shared__ unsigned long long s_cnt[3][32];
  if (threadIdx.x < 96) (&s_cnt[0][0])[threadIdx.x] = 0ull;
  __syncthreads();
  const int Dl = static_cast<int>(vol / plane);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < vol;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int z = static_cast<int>(i / plane);
    const int64_t o = i - z * plane;
    const A ws = wsum ? wsum[i] : static_cast<A>(1);
    int best = 0;
    A bv = acc[acc_index(0, z, o, C, Dl, plane, d_outer != 0)] / ws;
    if (out_logits) out_logits[i] = static_cast<float>(bv);
    for (int c = 1; c < C; ++c) {
      const A v = acc[acc_index(c, z, o, C, Dl, plane, d_outer != 0)] / ws;
      if (out_logits) out_logits[c * vol + i] = static_cast<float>(v);
      if (v > bv) bv = v, best = c;  // first maximum wins, like torch.argmax
    }
    if (argmax) argmax[i] = static_cast<uint8_t>(best);
    if (label) {
      const float lv = label_u8 ? static_cast<float>(static_cast<const uint8_t*>(label)[i]) : static_cast<const float*>(label)[i];
      int li;
      const bool lvalid = label_class(lv, C, li);
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

// fp32 accumulators, 4 voxels per thread (16-byte loads per class plane, one 4-byte argmax store, 4-byte label load):
// the production path of predict_sliding_dice.  No division: argmax(acc_c / w) == argmax(acc_c) for w > 0, and w > 0
// everywhere the windows cover (the Gaussian map is floored at its smallest non-zero value, evaluate_amos.py:195).
template <bool D_OUTER>
__global__ void __launch_bounds__(256)
sw_finalize_f32x4_kernel(const float* __restrict__ acc, const void* __restrict__ label, int label_u8,
                         uint8_t* __restrict__ argmax, unsigned long long* __restrict__ counts, int C, int64_t vol,
                         int64_t plane) {
  __shared__ unsigned int s_cnt[3][32];
  if (threadIdx.x < 96) (&s_cnt[0][0])[threadIdx.x] = 0u;
  __syncthreads();
  const int Dl = static_cast<int>(vol / plane);
  const int64_t nv = vol / 4;
  for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < nv;
       q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = q * 4;
    const int z = static_cast<int>(i / plane);
    const int64_t o = i - z * plane;
    float4 bv = __ldg(reinterpret_cast<const float4*>(acc + acc_index(0, z, o, C, Dl, plane, D_OUTER)));
    int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
    for (int c = 1; c < C; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(acc + acc_index(c, z, o, C, Dl, plane, D_OUTER)));
      if (v.x > bv.x) bv.x = v.x, b0 = c;
      if (v.y > bv.y) bv.y = v.y, b1 = c;
      if (v.z > bv.z) bv.z = v.z, b2 = c;
      if (v.w > bv.w) bv.w = v.w, b3 = c;
    }
    const int best[4] = {b0, b1, b2, b3};
    if (argmax) *reinterpret_cast<uint32_t*>(argmax + i) = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
    if (label) {
      float lv[4];
      if (label_u8) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(label) + i));
#pragma unroll
        for (int k = 0; k < 4; ++k) lv[k] = static_cast<float>((w >> (8 * k)) & 0xFFu);
      } else {
        const float4 t = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(label) + i));
        lv[0] = t.x, lv[1] = t.y, lv[2] = t.z, lv[3] = t.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int li;
        const bool lvalid = label_class(lv[k], C, li);
        atomicAdd(&s_cnt[1][best[k]], 1u);
        if (lvalid) {
          atomicAdd(&s_cnt[2][li], 1u);
          if (li == best[k]) atomicAdd(&s_cnt[0][best[k]], 1u);
        }
      }
    }
  }
  __syncthreads();
  if (label && threadIdx.x < 96) {
    const int k = threadIdx.x / 32, c = threadIdx.x % 32;
    if (c < C && s_cnt[k][c]) atomicAdd(&counts[k * C + c], static_cast<unsigned long long>(s_cnt[k][c]));
  }
}

// dst += src over n floats: the partial accumulator slabs other ranks send to the slab owner (sharded sliding window)
__global__ void __launch_bounds__(256) accumulate_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4,
                                                             int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (int64_t q = i0; q < n4; q += stride) {
    float4 d = reinterpret_cast<float4*>(dst)[q];
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
    d.x += v.x, d.y += v.y, d.z += v.z, d.w += v.w;
    reinterpret_cast<float4*>(dst)[q] = d;
  }
  for (int64_t i = n4 * 4 + i0; i < n; i += stride) dst[i] += src[i];
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_sw_blend(void* acc, void* wsum, const float* tile_logits, const float* gauss, int c, int d, int h,
                             int w, int td, int th, int tw, int d0, int h0, int w0, int acc_bytes, int d_outer,
                             mmpl_stream_t stream) {
  MMPL_REQUIRE(d0 >= 0 && h0 >= 0 && w0 >= 0 && d0 + td <= d && h0 + th <= h && w0 + tw <= w, MMPL_E_SHAPE,
               "sw_blend: tile (%d,%d,%d)+(%d,%d,%d) outside volume (%d,%d,%d)", d0, h0, w0, td, th, tw, d, h, w);
  MMPL_REQUIRE(acc_bytes == 4 || acc_bytes == 8, MMPL_E_DTYPE, "sw_blend: acc_bytes=%d", acc_bytes);
  const int64_t tvox = static_cast<int64_t>(td) * th * tw;
  const int blocks = static_cast<int>(std::min<int64_t>((tvox + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (acc_bytes == 8)
    sw_blend_kernel<double><<<blocks, 256, 0, s>>>(static_cast<double*>(acc), static_cast<double*>(wsum), tile_logits,
                                                  gauss, c, d, h, w, td, th, tw, d0, h0, w0, d_outer);
  else
    sw_blend_kernel<float><<<blocks, 256, 0, s>>>(static_cast<float*>(acc), static_cast<float*>(wsum), tile_logits, gauss,
                                                 c, d, h, w, td, th, tw, d0, h0, w0, d_outer);
  MMPL_CHECK_LAUNCH("sw_blend");
  return MMPL_OK;
}

extern "C" int mmpl_sw_finalize(const void* acc, const void* wsum, const void* label, int label_is_u8, float* out_logits,
                                uint8_t* argmax, long long* counts, int c, int64_t voxels, int64_t plane, int acc_bytes,
                                mmpl_stream_t stream) {
  MMPL_REQUIRE(c >= 1 && c <= 32, MMPL_E_SHAPE, "sw_finalize: classes=%d", c);
  MMPL_REQUIRE(acc_bytes == 4 || acc_bytes == 8, MMPL_E_DTYPE, "sw_finalize: acc_bytes=%d", acc_bytes);
  MMPL_REQUIRE(plane >= 0 && (plane == 0 || voxels % plane == 0), MMPL_E_SHAPE, "sw_finalize: plane=%lld", (long long)plane);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (counts) MMPL_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * 3 * c, s));
  const void* lab = counts ? label : nullptr;
  const int d_outer = plane > 0;
  const int64_t pl = plane > 0 ? plane : voxels;      // class-major: the whole volume is one "plane" per class
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts);
  const bool vec = acc_bytes == 4 && out_logits == nullptr && pl % 4 == 0 && reinterpret_cast<uintptr_t>(acc) % 16 == 0 &&
                   (argmax == nullptr || reinterpret_cast<uintptr_t>(argmax) % 4 == 0) &&
                   (lab == nullptr || reinterpret_cast<uintptr_t>(lab) % (label_is_u8 ? 4 : 16) == 0);
  if (vec) {
    // the fast path skips the division by wsum (argmax-invariant); callers that pass wsum get the same argmax
    const int blocks = static_cast<int>(std::min<int64_t>((voxels / 4 + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
    if (d_outer)
      sw_finalize_f32x4_kernel<true><<<blocks, 256, 0, s>>>(static_cast<const float*>(acc), lab, label_is_u8, argmax, cnt, c, voxels, pl);
    else
      sw_finalize_f32x4_kernel<false><<<blocks, 256, 0, s>>>(static_cast<const float*>(acc), lab, label_is_u8, argmax, cnt, c, voxels, pl);
    MMPL_CHECK_LAUNCH("sw_finalize");
    return MMPL_OK;
  }
  const int blocks = static_cast<int>(std::min<int64_t>((voxels + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  if (acc_bytes == 8)
    sw_finalize_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(acc), static_cast<const double*>(wsum), lab,
                                                     label_is_u8, out_logits, argmax, cnt, c, voxels, pl, d_outer);
  else
    sw_finalize_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(acc), static_cast<const float*>(wsum), lab,
                                                    label_is_u8, out_logits, argmax, cnt, c, voxels, pl, d_outer);
  MMPL_CHECK_LAUNCH("sw_finalize");
  return MMPL_OK;
}

extern "C" int mmpl_accumulate_f32(float* dst, const float* src, int64_t n, mmpl_stream_t stream) {
  MMPL_REQUIRE(n >= 0, MMPL_E_SHAPE, "accumulate_f32: n=%lld", (long long)n);
  if (n == 0) return MMPL_OK;
  const bool vec = (reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) % 16 == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  const int blocks = static_cast<int>(std::min<int64_t>((std::max<int64_t>(n4, 1) + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  accumulate_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, n4, n);
  MMPL_CHECK_LAUNCH("accumulate_f32");
  return MMPL_OK;
}
