// GPU input pipeline for the train loop: the per-sample preprocessing of AMOSDataSet_newatlas.__getitem__
// (MOTSDataset.py:299-395) and the intensity augmentations of get_train_transform (:33-52) as device kernels, so that a
// 13 ms train step is not fed by a CPU DataLoader doing numpy padding / cropping / normalisation per sample.
//
// Reference order of operations per sample (volume layout [h][w][d], d fastest, as the reference holds it):
//   atlas  : nearest-neighbour resize of the [K][ha][wa][da] atlas to the image shape (:357, F.interpolate default mode)
//   pad    : zero-pad image / label / atlas at the END of each axis up to crop + 5 (:370-372, pad_image :269-297)
//   scale  : CT  -> clip to [-325, 325] HU, divide by 325;  MRI -> (x - mean) / std over the whole PADDED volume (:374, :171-186)
//   crop   : window [b, b+crop_h) x [c, c+crop_w) x [a, a+crop_d) (:377-383; the random origin stays on the host)
//   layout : transpose to [1][D][H][W] (:389-391), fp32
// One kernel does pad + scale + crop + transpose for image and label (shared-memory tiled transpose so both the strided
// source and the dense destination are accessed in full sectors), one does the same for the atlas including the resize.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace mmpl {
namespace {

template <typename S>
__device__ __forceinline__ float src_value(const S* p, int64_t i) {
  return static_cast<float>(p[i]);
}

// sum and sum of squares of the source volume in fp64 (MRI z-score, :184-185: np.mean / np.std over the padded volume --
// the zero padding contributes to the count only, which the caller passes)
template <typename S>
__global__ void __launch_bounds__(256)
volume_moments_kernel(const S* __restrict__ v, int64_t n, double* __restrict__ out) {
  double s = 0, q = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double x = static_cast<double>(v[i]);
    s += x;
    q += x * x;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  __shared__ double ss[8], sq[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) ss[warp] = s, sq[warp] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; ++w) a += ss[w], b += sq[w];
    atomicAdd(&out[0], a);
    atomicAdd(&out[1], b);
  }
}

struct PatchGeo {
  int h, w, d;             // source extents (before padding), layout [h][w][d]
  int b, c, a;             // crop origin along h, w, d
  int ch, cw, cd;          // crop extents = output [cd][ch][cw]
};

// grid: (ceil(cw/32), ceil(cd/32), ch); block 32 x 8.  Tile = 32 (x along w) x 32 (z along d) of one output row y.
// mode 0: CT clip/scale; 1: MRI z-score with moments (sum, sumsq) and the padded voxel count; 2: copy (labels)
template <typename S, typename O>
__global__ void __launch_bounds__(256)
prepare_patch_kernel(const S* __restrict__ src, O* __restrict__ dst, PatchGeo g, int mode, const double* __restrict__ moments,
                     double padded_count) {
  __shared__ float tile[32][33];
  const int y = blockIdx.z;
  const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
  // integer sources follow numpy's promotion: the reference's arithmetic runs in float64 and is rounded to fp32 once at
  // the end (int16 / 325.0, (int16 - mean) / std); fp32 sources stay in fp32 like numpy keeps them
  constexpr bool kWide = !std::is_same<S, float>::value;
  double mu = 0, sd = 1;
  if (mode == 1) {
    mu = moments[0] / padded_count;
    const double var = moments[1] / padded_count - mu * mu;
    sd = sqrt(var > 0 ? var : 0);
  }
  // load: threads run along z (the fastest source axis)
  const int sy = g.b + y;
#pragma unroll
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int x = x0 + i, z = z0 + threadIdx.x;
    const int sx = g.c + x, sz = g.a + z;
    float v = 0.f;                                            // zero padding (pad_image)
    if (x < g.cw && z < g.cd && sy < g.h && sx < g.w && sz < g.d)
      v = src_value(src, (static_cast<int64_t>(sy) * g.w + sx) * g.d + sz);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  // store: threads run along x (the fastest destination axis)
#pragma unroll
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int z = z0 + i, x = x0 + threadIdx.x;
    if (z >= g.cd || x >= g.cw) continue;
    float v = tile[threadIdx.x][i];
    if (mode == 0) {                                          // truncate(), :171-183 (subtract = 0, divide = 325)
      const float cl = fminf(fmaxf(v, -325.f), 325.f);
      v = kWide ? static_cast<float>(static_cast<double>(cl) / 325.0) : cl / 325.f;
    } else if (mode == 1) {                                   // :184-185
      v = kWide ? static_cast<float>((static_cast<double>(v) - mu) / sd)
                : (v - static_cast<float>(mu)) / static_cast<float>(sd);
    }
    dst[(static_cast<int64_t>(z) * g.ch + y) * g.cw + x] = static_cast<O>(v);
  }
}

// atlas [K][ha][wa][da] --nearest resize to (h, w, d)--> zero-pad --> crop --> [K][cd][ch][cw]
// nearest index as ATen computes it: src = min(int(floorf(dst * (float)in / out)), in - 1)
__global__ void __launch_bounds__(256)
atlas_patch_kernel(const float* __restrict__ atlas, float* __restrict__ dst, int K, int ha, int wa, int da, PatchGeo g) {
  const int64_t per = static_cast<int64_t>(g.cd) * g.ch * g.cw;
  const int64_t total = per * K;
  const float sh = static_cast<float>(ha) / g.h, sw = static_cast<float>(wa) / g.w, sd = static_cast<float>(da) / g.d;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i / per);
    int64_t r = i - k * per;
    const int x = static_cast<int>(r % g.cw);
    r /= g.cw;
    const int y = static_cast<int>(r % g.ch);
    const int z = static_cast<int>(r / g.ch);
    const int sy = g.b + y, sx = g.c + x, sz = g.a + z;
    float v = 0.f;
    if (sy < g.h && sx < g.w && sz < g.d) {
      const int ay = min(static_cast<int>(floorf(sy * sh)), ha - 1);
      const int ax = min(static_cast<int>(floorf(sx * sw)), wa - 1);
      const int az = min(static_cast<int>(floorf(sz * sd)), da - 1);
      v = atlas[((static_cast<int64_t>(k) * ha + ay) * wa + ax) * da + az];
    }
    dst[i] = v;
  }
}

// ---- intensity augmentations (batchgenerators, as configured at :33-52), parameters drawn by the caller ---------------
//   x <- x + N(0, noise_std)                          GaussianNoiseTransform (its "variance" is used as the std)
//   x <- x * mult                                     BrightnessMultiplicativeTransform
//   x <- x + add                                      BrightnessTransform (additive, one draw per channel)
//   x <- clip((x - mean) * contrast + mean, lo, hi)   ContrastAugmentationTransform (preserve_range: lo/hi = min/max before)
// A step is skipped by its neutral parameter (noise_std 0, mult 1, add 0, contrast 1).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float gauss_from_counter(uint64_t seed, uint64_t i) {
  const uint32_t a = mix32(static_cast<uint32_t>(i) ^ static_cast<uint32_t>(seed));
  const uint32_t b = mix32(static_cast<uint32_t>(i >> 32) ^ static_cast<uint32_t>(seed >> 32) ^ (a * 0x9e3779b9u));
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * (1.0f / 16777216.0f);      // (0, 1]
  const float u2 = static_cast<float>(b >> 8) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);                // Box-Muller
}

__global__ void __launch_bounds__(256)
augment_kernel(float* __restrict__ x, int64_t n, float noise_std, uint64_t seed, float mult, float add, float contrast,
               const double* __restrict__ stats /* sum, min, max (contrast only) */) {
  float mean = 0.f, lo = 0.f, hi = 0.f;
  const bool do_contrast = contrast != 1.0f && stats != nullptr;
  if (do_contrast) {
    mean = static_cast<float>(stats[0] / static_cast<double>(n));
    lo = static_cast<float>(stats[1]);
    hi = static_cast<float>(stats[2]);
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v = x[i];
    if (noise_std > 0.f) v += noise_std * gauss_from_counter(seed, static_cast<uint64_t>(i));
    v = v * mult + add;
    if (do_contrast) v = fminf(fmaxf((v - mean) * contrast + mean, lo), hi);
    x[i] = v;
  }
}

// sum, min, max of a patch (fp64 [3]; min / max via ordered-int atomics on the fp32 bit pattern would lose nothing, but a
// two-level reduction through fp64 atomics is simpler at 2.4 M elements)
__global__ void __launch_bounds__(256)
patch_stats_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ out, unsigned long long* __restrict__ mm) {
  double s = 0;
  float lo = INFINITY, hi = -INFINITY;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    s += v;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], s);
    // order-preserving map of fp32 to uint32 so that integer atomicMin/Max order floats
    auto key = [](float f) {
      const uint32_t u = __float_as_uint(f);
      return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    };
    atomicMin(reinterpret_cast<unsigned int*>(mm), key(lo));
    atomicMax(reinterpret_cast<unsigned int*>(mm) + 1, key(hi));
  }
}
__global__ void patch_stats_finish_kernel(double* out, const unsigned long long* mm) {
  auto unkey = [](uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); };
  const unsigned int* m = reinterpret_cast<const unsigned int*>(mm);
  out[1] = static_cast<double>(unkey(m[0]));
  out[2] = static_cast<double>(unkey(m[1]));
}

// 1-D Gaussian correlation along one axis of a [D][H][W] fp32 volume with scipy.ndimage 'reflect' boundary
// (d c b a | a b c d | d c b a): GaussianBlurTransform = gaussian_filter(img, sigma, order=0), truncate 4.0.
__global__ void __launch_bounds__(256)
blur_axis_kernel(const float* __restrict__ src, float* __restrict__ dst, int D, int H, int W, int axis,
                 const float* __restrict__ taps, int radius) {
  const int64_t total = static_cast<int64_t>(D) * H * W;
  const int n = axis == 0 ? D : axis == 1 ? H : W;
  const int64_t stride = axis == 0 ? static_cast<int64_t>(H) * W : axis == 1 ? W : 1;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>((i / stride) % n);
    const int64_t base = i - p * stride;
    float acc = 0.f;
    for (int t = -radius; t <= radius; ++t) {
      int q = p + t;
      // reflect (half-sample symmetric), valid for any offset
      const int period = 2 * n;
      q %= period;
      if (q < 0) q += period;
      if (q >= n) q = period - 1 - q;
      acc = fmaf(taps[t + radius], src[base + q * stride], acc);
    }
    dst[i] = acc;
  }
}

template <typename S, typename O>
int launch_prepare(const void* src, void* dst, const PatchGeo& g, int mode, const double* moments, double padded_count,
                   cudaStream_t s) {
  const dim3 grid((g.cw + 31) / 32, (g.cd + 31) / 32, g.ch);
  MMPL_REQUIRE(grid.y <= 65535 && grid.z <= 65535, MMPL_E_SHAPE, "prepare_patch: crop (%d,%d,%d) exceeds the grid", g.ch, g.cw, g.cd);
  prepare_patch_kernel<S, O><<<grid, dim3(32, 8), 0, s>>>(static_cast<const S*>(src), static_cast<O*>(dst), g, mode, moments,
                                                         padded_count);
  return MMPL_OK;
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

static int64_t padded_extent(int e, int crop) { return std::max<int64_t>(e, crop + 5); }

// src_dtype: 0 = fp32, 1 = int16, 2 = uint8
extern "C" int mmpl_volume_moments(const void* volume, int src_dtype, int64_t count, double* moments, mmpl_stream_t stream) {
  MMPL_REQUIRE(count > 0 && volume && moments, MMPL_E_SHAPE, "volume_moments: empty input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * 2, s));
  const int blocks = static_cast<int>(std::min<int64_t>((count + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  if (src_dtype == 0)
    volume_moments_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(volume), count, moments);
  else if (src_dtype == 1)
    volume_moments_kernel<int16_t><<<blocks, 256, 0, s>>>(static_cast<const int16_t*>(volume), count, moments);
  else
    MMPL_FAIL(MMPL_E_DTYPE, "volume_moments: src_dtype=%d (0 = fp32, 1 = int16)", src_dtype);
  MMPL_CHECK_LAUNCH("volume_moments");
  return MMPL_OK;
}

extern "C" int mmpl_prepare_patch(const void* volume, int src_dtype, void* out, int out_is_u8, int h, int w, int d, int b,
                                  int c, int a, int crop_h, int crop_w, int crop_d, int mode, const double* moments,
                                  mmpl_stream_t stream) {
  MMPL_REQUIRE(h > 0 && w > 0 && d > 0 && crop_h > 0 && crop_w > 0 && crop_d > 0, MMPL_E_SHAPE, "prepare_patch: empty input");
  MMPL_REQUIRE(b >= 0 && c >= 0 && a >= 0 && b + crop_h <= padded_extent(h, crop_h) && c + crop_w <= padded_extent(w, crop_w) &&
                   a + crop_d <= padded_extent(d, crop_d),
               MMPL_E_SHAPE, "prepare_patch: crop origin (%d,%d,%d) outside the padded volume", b, c, a);
  MMPL_REQUIRE(mode >= 0 && mode <= 2 && (mode != 1 || moments != nullptr), MMPL_E_SHAPE, "prepare_patch: mode=%d", mode);
  MMPL_REQUIRE(!out_is_u8 || mode == 2, MMPL_E_DTYPE, "prepare_patch: uint8 output is for labels (mode 2)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PatchGeo g{h, w, d, b, c, a, crop_h, crop_w, crop_d};
  const double padded = static_cast<double>(padded_extent(h, crop_h)) * padded_extent(w, crop_w) * padded_extent(d, crop_d);
  int rc;
  if (src_dtype == 0)
    rc = out_is_u8 ? launch_prepare<float, uint8_t>(volume, out, g, mode, moments, padded, s)
                   : launch_prepare<float, float>(volume, out, g, mode, moments, padded, s);
  else if (src_dtype == 1)
    rc = out_is_u8 ? launch_prepare<int16_t, uint8_t>(volume, out, g, mode, moments, padded, s)
                   : launch_prepare<int16_t, float>(volume, out, g, mode, moments, padded, s);
  else if (src_dtype == 2)
    rc = out_is_u8 ? launch_prepare<uint8_t, uint8_t>(volume, out, g, mode, moments, padded, s)
                   : launch_prepare<uint8_t, float>(volume, out, g, mode, moments, padded, s);
  else
    MMPL_FAIL(MMPL_E_DTYPE, "prepare_patch: src_dtype=%d (0 = fp32, 1 = int16, 2 = uint8)", src_dtype);
  if (rc) return rc;
  MMPL_CHECK_LAUNCH("prepare_patch");
  return MMPL_OK;
}

extern "C" int mmpl_atlas_patch(const float* atlas, float* out, int k, int ha, int wa, int da, int h, int w, int d, int b,
                                int c, int a, int crop_h, int crop_w, int crop_d, mmpl_stream_t stream) {
  MMPL_REQUIRE(k > 0 && ha > 0 && wa > 0 && da > 0 && h > 0 && w > 0 && d > 0, MMPL_E_SHAPE, "atlas_patch: empty input");
  PatchGeo g{h, w, d, b, c, a, crop_h, crop_w, crop_d};
  const int64_t total = static_cast<int64_t>(k) * crop_h * crop_w * crop_d;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  atlas_patch_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(atlas, out, k, ha, wa, da, g);
  MMPL_CHECK_LAUNCH("atlas_patch");
  return MMPL_OK;
}

extern "C" int mmpl_patch_stats(const float* x, int64_t n, double* stats /*[4]: sum, min, max, scratch*/,
                                mmpl_stream_t stream) {
  MMPL_REQUIRE(n > 0, MMPL_E_SHAPE, "patch_stats: empty input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_CUDA(cudaMemsetAsync(stats, 0, sizeof(double), s));
  unsigned long long* mm = reinterpret_cast<unsigned long long*>(stats + 3);
  const unsigned int init[2] = {0xFFFFFFFFu, 0u};
  MMPL_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, s));
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  patch_stats_kernel<<<blocks, 256, 0, s>>>(x, n, stats, mm);
  patch_stats_finish_kernel<<<1, 1, 0, s>>>(stats, mm);
  MMPL_CHECK_LAUNCH("patch_stats");
  return MMPL_OK;
}

extern "C" int mmpl_augment_patch(float* x, int64_t n, float noise_std, uint64_t seed, float mult, float add, float contrast,
                                  const double* stats, mmpl_stream_t stream) {
  MMPL_REQUIRE(n > 0, MMPL_E_SHAPE, "augment_patch: empty input");
  MMPL_REQUIRE(contrast == 1.0f || stats != nullptr, MMPL_E_SHAPE, "augment_patch: contrast needs mmpl_patch_stats");
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  augment_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, noise_std, seed, mult, add, contrast, stats);
  MMPL_CHECK_LAUNCH("augment_patch");
  return MMPL_OK;
}

extern "C" int mmpl_blur_axis(const float* src, float* dst, int d, int h, int w, int axis, const float* taps_dev, int radius,
                              mmpl_stream_t stream) {
  MMPL_REQUIRE(d > 0 && h > 0 && w > 0 && axis >= 0 && axis <= 2 && radius >= 0 && src != dst, MMPL_E_SHAPE,
               "blur_axis: bad arguments");
  const int64_t total = static_cast<int64_t>(d) * h * w;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  blur_axis_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, d, h, w, axis, taps_dev, radius);
  MMPL_CHECK_LAUNCH("blur_axis");
  return MMPL_OK;
}
