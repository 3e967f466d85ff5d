// Small kernels for the other two networks of the train loop: the refiner unet3D_g (unet3D.py:1507-1623)
// and the discriminator norm_style_discriminator_output (:1907-1947), driven at train_amos_atlas_final.py:277-368.
// Their convolutions run on the tcgen05 kernels of conv_tc.cu / wgrad_tc.cu:
//   * refiner widths 24/48/96/192 are zero-padded per GroupNorm group to 32/64/128/256 (mmpl_gn_relu_* real_cpg);
//   * the discriminator's 4x4x4 stride-2 padding-1 convolutions become 3x3x3 stride-1 padding-1 convolutions over a
//     space-to-depth(2) copy of their input: tap t of an axis reads position 2o - 1 + t, i.e. block o-1 parity 1 (t = 0),
//     block o parity 0 / 1 (t = 1 / 2), block o+1 parity 0 (t = 3), so the 8C-channel 3^3 filter holds the 4^3 taps with
//     (4/6)^3 = 30 % density -- 3.4x the MMA work of the 4^3 conv, but on tensor cores and with no new conv kernel.
// What is left for this file is data movement and element-wise work:
//   mmpl_space_to_depth2_{fwd,bwd}   [N,D,H,W,C] <-> [N,D/2,H/2,W/2,Cp] with channel (pd,ph,pw,c) (Cp >= 8C, zero padded)
//   mmpl_bias_lrelu_{fwd,bwd}        y = leaky_relu(x + bias[c], slope) on NDHWC rows; bwd: dx and dbias
//   mmpl_upsample2x_ncdhw_{fwd,bwd}  nn.Upsample(scale_factor=2, mode='trilinear') of fp32 NCDHW logits (unet3D.py:1621)
#include <algorithm>

#include "common.cuh"

namespace mmpl {
namespace {

// one thread per element of the s2d tensor (E = uint16_t for bf16, uint32_t for fp32: pure data movement)
template <typename E>
__global__ void __launch_bounds__(256)
s2d_kernel(const E* __restrict__ x, E* __restrict__ y, int N, int D, int H, int W, int C, int Cp, int inverse) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const int64_t total = static_cast<int64_t>(N) * Do * Ho * Wo * Cp;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cp = static_cast<int>(i % Cp);
    int64_t r = i / Cp;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    r /= Ho;
    const int d_o = static_cast<int>(r % Do);
    const int n = static_cast<int>(r / Do);
    const int par = cp / C, c = cp - par * C;
    if (par >= 8) {                       // padding channels
      if (!inverse) y[i] = 0;
      continue;
    }
    const int d = 2 * d_o + (par >> 2), h = 2 * ho + ((par >> 1) & 1), w = 2 * wo + (par & 1);
    const int64_t j = (((static_cast<int64_t>(n) * D + d) * H + h) * W + w) * C + c;
    if (!inverse)
      y[i] = x[j];
    else
      y[j] = x[i];                        // depth-to-space: every element of the full-resolution tensor is written once
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bias_lrelu_fwd_kernel(const T* __restrict__ x, const float* __restrict__ bias, T* __restrict__ y, int64_t rows, int C,
                      float slope) {
  constexpr int VN = Vec<T>::N;
  const int vpr = C / VN;
  const int64_t total = rows * vpr;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(i % vpr);
    Vec<T> v;
    v.load(x + i * VN);
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      const float t = v.v[k] + bias[cv * VN + k];
      v.v[k] = t > 0.f ? t : t * slope;
    }
    v.store(y + i * VN);
  }
}

// dx = dy * (y > 0 ? 1 : slope) (the sign of the output equals the sign of x + bias for slope > 0); dbias[c] = sum dx
template <typename T>
__global__ void __launch_bounds__(256)
bias_lrelu_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, float* __restrict__ dbias,
                      int64_t rows, int C, float slope) {
  constexpr int VN = Vec<T>::N;
  const int vpr = C / VN;                              // 256 % vpr == 0 (host check): a thread keeps its channel vector
  const int cv = threadIdx.x % vpr;
  const int rl = threadIdx.x / vpr, rstep = blockDim.x / vpr;
  float acc[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) acc[k] = 0.f;
  for (int64_t r = blockIdx.x * static_cast<int64_t>(rstep) + rl; r < rows; r += static_cast<int64_t>(gridDim.x) * rstep) {
    Vec<T> o, g;
    o.load(y + r * C + cv * VN);
    g.load(dy + r * C + cv * VN);
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      g.v[k] = o.v[k] > 0.f ? g.v[k] : g.v[k] * slope;
      acc[k] += g.v[k];
    }
    g.store(dx + r * C + cv * VN);
  }
#pragma unroll
  for (int k = 0; k < VN; ++k) atomicAdd(&dbias[cv * VN + k], acc[k]);
}

// align_corners = False, exact 2x: out[2i] = .25 in[i-1] + .75 in[i], out[2i+1] = .75 in[i] + .25 in[i+1], clamped
__device__ __forceinline__ void up_taps(int o, int n, int& i0, int& i1, float& w0, float& w1) {
  const int i = o >> 1;
  if (o & 1) {
    i0 = i, i1 = min(i + 1, n - 1), w0 = 0.75f, w1 = 0.25f;
  } else {
    i0 = max(i - 1, 0), i1 = i, w0 = 0.25f, w1 = 0.75f;
  }
}

__global__ void __launch_bounds__(256)
up2_ncdhw_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t planes, int D, int H, int W) {
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const int64_t total = planes * Do * Ho * Wo;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(i % Wo);
    int64_t r = i / Wo;
    const int ho = static_cast<int>(r % Ho);
    r /= Ho;
    const int d_o = static_cast<int>(r % Do);
    const int64_t p = r / Do;
    int d0, d1, h0, h1, w0, w1;
    float a0, a1, b0, b1, c0, c1;
    up_taps(d_o, D, d0, d1, a0, a1);
    up_taps(ho, H, h0, h1, b0, b1);
    up_taps(wo, W, w0, w1, c0, c1);
    const float* xp = x + p * D * H * W;
    auto at = [&](int d, int h, int w) { return xp[(static_cast<int64_t>(d) * H + h) * W + w]; };
    const float v0 = b0 * (c0 * at(d0, h0, w0) + c1 * at(d0, h0, w1)) + b1 * (c0 * at(d0, h1, w0) + c1 * at(d0, h1, w1));
    const float v1 = b0 * (c0 * at(d1, h0, w0) + c1 * at(d1, h0, w1)) + b1 * (c0 * at(d1, h1, w0) + c1 * at(d1, h1, w1));
    y[i] = a0 * v0 + a1 * v1;
  }
}

// transpose of the above in gather form: input voxel i receives from outputs 2i-1 .. 2i+2 per axis (clamping folds the
// out-of-range taps of the border outputs back onto the border input)
__device__ __forceinline__ float up_weight(int o, int i, int n) {
  // coefficient of in[i] in out[o]
  int i0, i1;
  float w0, w1;
  up_taps(o, n, i0, i1, w0, w1);
  return (i0 == i ? w0 : 0.f) + (i1 == i ? w1 : 0.f);
}

__global__ void __launch_bounds__(256)
up2_ncdhw_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t planes, int D, int H, int W) {
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const int64_t total = planes * D * H * W;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    int64_t r = i / W;
    const int h = static_cast<int>(r % H);
    r /= H;
    const int d = static_cast<int>(r % D);
    const int64_t p = r / D;
    const float* gp = dy + p * Do * Ho * Wo;
    float acc = 0.f;
    for (int od = max(2 * d - 1, 0); od <= min(2 * d + 2, Do - 1); ++od) {
      const float a = up_weight(od, d, D);
      if (a == 0.f) continue;
      for (int oh = max(2 * h - 1, 0); oh <= min(2 * h + 2, Ho - 1); ++oh) {
        const float b = up_weight(oh, h, H);
        if (b == 0.f) continue;
        for (int ow = max(2 * w - 1, 0); ow <= min(2 * w + 2, Wo - 1); ++ow) {
          const float c = up_weight(ow, w, W);
          if (c != 0.f) acc = fmaf(a * b * c, gp[(static_cast<int64_t>(od) * Ho + oh) * Wo + ow], acc);
        }
      }
    }
    dx[i] = acc;
  }
}

int blocks_for(int64_t n) { return static_cast<int>(std::min<int64_t>((n + 255) / 256, static_cast<int64_t>(num_sms()) * 8)); }

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_space_to_depth2(const void* x, void* y, int n, int d, int h, int w, int c, int cp, int inverse,
                                    int dtype, mmpl_stream_t stream) {
  MMPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && c > 0, MMPL_E_SHAPE, "space_to_depth2: empty tensor");
  MMPL_REQUIRE(d % 2 == 0 && h % 2 == 0 && w % 2 == 0, MMPL_E_SHAPE, "space_to_depth2: extents (%d,%d,%d) must be even", d, h, w);
  MMPL_REQUIRE(cp >= 8 * c, MMPL_E_SHAPE, "space_to_depth2: cp=%d < 8*c=%d", cp, 8 * c);
  const int64_t total = static_cast<int64_t>(n) * (d / 2) * (h / 2) * (w / 2) * cp;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == MMPL_BF16)
    s2d_kernel<uint16_t><<<blocks_for(total), 256, 0, s>>>(static_cast<const uint16_t*>(x), static_cast<uint16_t*>(y), n, d, h, w,
                                                          c, cp, inverse);
  else if (dtype == MMPL_F32)
    s2d_kernel<uint32_t><<<blocks_for(total), 256, 0, s>>>(static_cast<const uint32_t*>(x), static_cast<uint32_t*>(y), n, d, h, w,
                                                          c, cp, inverse);
  else
    MMPL_FAIL(MMPL_E_DTYPE, "space_to_depth2: dtype=%d", dtype);
  MMPL_CHECK_LAUNCH("space_to_depth2");
  return MMPL_OK;
}

extern "C" int mmpl_bias_lrelu_fwd(const void* x, const float* bias, void* y, int64_t rows, int c, float slope, int dtype,
                                   mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(rows > 0 && c % vn == 0, MMPL_E_SHAPE, "bias_lrelu: rows=%lld C=%d", (long long)rows, c);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (bias_lrelu_fwd_kernel<T><<<blocks_for(rows * (c / vn)), 256, 0, s>>>(
                                    static_cast<const T*>(x), bias, static_cast<T*>(y), rows, c, slope)));
  MMPL_CHECK_LAUNCH("bias_lrelu_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_bias_lrelu_bwd(const void* y, const void* dy, void* dx, float* dbias, int64_t rows, int c, float slope,
                                   int dtype, mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(rows > 0 && c % vn == 0 && c / vn <= 256 && 256 % (c / vn) == 0, MMPL_E_SHAPE,
               "bias_lrelu_bwd: rows=%lld C=%d", (long long)rows, c);
  MMPL_REQUIRE(slope > 0.f, MMPL_E_SHAPE, "bias_lrelu_bwd: the gate is read from the output, slope must be > 0");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * c, s));
  const int rstep = 256 / (c / vn);
  const int blocks = static_cast<int>(std::min<int64_t>((rows + rstep - 1) / rstep, static_cast<int64_t>(num_sms()) * 4));
  MMPL_DISPATCH_DTYPE(dtype, T, (bias_lrelu_bwd_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(y), static_cast<const T*>(dy),
                                                                               static_cast<T*>(dx), dbias, rows, c, slope)));
  MMPL_CHECK_LAUNCH("bias_lrelu_bwd");
  return MMPL_OK;
}

extern "C" int mmpl_upsample2x_ncdhw_fwd(const float* x, float* y, int64_t planes, int d, int h, int w,
                                         mmpl_stream_t stream) {
  MMPL_REQUIRE(planes > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x_ncdhw: empty tensor");
  up2_ncdhw_fwd_kernel<<<blocks_for(planes * 8 * d * h * w), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, planes, d, h, w);
  MMPL_CHECK_LAUNCH("upsample2x_ncdhw_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_upsample2x_ncdhw_bwd(const float* dy, float* dx, int64_t planes, int d, int h, int w,
                                         mmpl_stream_t stream) {
  MMPL_REQUIRE(planes > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x_ncdhw: empty tensor");
  up2_ncdhw_bwd_kernel<<<blocks_for(planes * d * h * w), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, dx, planes, d, h, w);
  MMPL_CHECK_LAUNCH("upsample2x_ncdhw_bwd");
  return MMPL_OK;
}
