// Trilinear x2 up-sampling (align_corners=False) fused with the additive skip connection, and its backward.
// Reference: nn.Upsample(scale_factor=2, mode='trilinear') + `x = x + skipN`, unet3D.py:608, :686-687.
// With an exact factor of 2 the source coordinate (dst+0.5)/2-0.5 gives fixed weights per axis:
//   out[2i]   = 1/4 in[i-1] + 3/4 in[i]      out[2i+1] = 3/4 in[i] + 1/4 in[i+1]     (indices clamped to [0, n-1])
// The backward is written as a gather (no atomics): in[i] collects 1/4,3/4,3/4,1/4 of dy[2i-1..2i+2], where a tap
// that falls outside is redirected to the edge output that clamped onto in[i].
#include "common.cuh"

namespace mmpl {
namespace {

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_add_fwd_kernel(const T* __restrict__ xlo, const T* __restrict__ skip, T* __restrict__ y, int N, int D, int H,
                          int W, int C) {
  constexpr int VN = Vec<T>::N;
  const int vpv = C / VN;
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const int64_t total = static_cast<int64_t>(N) * Do * Ho * Wo * vpv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % vpv);
    int64_t r = idx / vpv;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    r /= Ho;
    const int dd = static_cast<int>(r % Do);
    const int n = static_cast<int>(r / Do);
    int i0[3], i1[3];
    float w0[3], w1[3];
    const int o[3] = {dd, ho, wo}, lim[3] = {D, H, W};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int i = o[a] >> 1;
      if (o[a] & 1) {
        i0[a] = i, w0[a] = 0.75f, i1[a] = min(i + 1, lim[a] - 1), w1[a] = 0.25f;
      } else {
        i0[a] = max(i - 1, 0), w0[a] = 0.25f, i1[a] = i, w1[a] = 0.75f;
      }
    }
    Vec<T> acc;
    acc.load(skip + (idx - cv) * VN + cv * VN);
    const T* base = xlo + static_cast<int64_t>(n) * D * H * W * C + cv * VN;
    // PyTorch (upsample_trilinear3d) sums the eight taps as w_d*(w_h*(w_w a + w_w b) + ...); keep that nesting.
    float up[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) up[k] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int id = a ? i1[0] : i0[0];
      const float wd = a ? w1[0] : w0[0];
      float ph[VN];
#pragma unroll
      for (int k = 0; k < VN; ++k) ph[k] = 0.f;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ih = b ? i1[1] : i0[1];
        const float wh = b ? w1[1] : w0[1];
        Vec<T> p, q;
        const int64_t row = (static_cast<int64_t>(id) * H + ih) * W;
        p.load(base + (row + i0[2]) * C);
        q.load(base + (row + i1[2]) * C);
#pragma unroll
        for (int k = 0; k < VN; ++k) ph[k] += wh * (w0[2] * p.v[k] + w1[2] * q.v[k]);
      }
#pragma unroll
      for (int k = 0; k < VN; ++k) up[k] += wd * ph[k];
    }
#pragma unroll
    for (int k = 0; k < VN; ++k) acc.v[k] += up[k];
    acc.store(y + idx * VN);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dxlo, int N, int D, int H, int W, int C) {
  constexpr int VN = Vec<T>::N;
  const int vpv = C / VN;
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t total = static_cast<int64_t>(N) * D * H * W * vpv;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % vpv);
    int64_t r = idx / vpv;
    const int wi = static_cast<int>(r % W);
    r /= W;
    const int hi = static_cast<int>(r % H);
    r /= H;
    const int di = static_cast<int>(r % D);
    const int n = static_cast<int>(r / D);
    int oi[3][4];
    const int in[3] = {di, hi, wi}, lim[3] = {D, H, W};
    const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int i = in[a];
      oi[a][0] = i > 0 ? 2 * i - 1 : 0;
      oi[a][1] = 2 * i;
      oi[a][2] = 2 * i + 1;
      oi[a][3] = i < lim[a] - 1 ? 2 * i + 2 : 2 * lim[a] - 1;
    }
    float acc[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) acc[k] = 0.f;
    const T* base = dy + static_cast<int64_t>(n) * (2 * D) * Ho * Wo * C + cv * VN;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        const float wab = wt[a] * wt[b];
        const int64_t row = (static_cast<int64_t>(oi[0][a]) * Ho + oi[1][b]) * Wo;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          Vec<T> g;
          g.load(base + (row + oi[2][c]) * C);
          const float w3 = wab * wt[c];
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[k] = fmaf(w3, g.v[k], acc[k]);
        }
      }
    Vec<T> o;
#pragma unroll
    for (int k = 0; k < VN; ++k) o.v[k] = acc[k];
    o.store(dxlo + idx * VN);
  }
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_upsample2x_add_fwd(const void* x_lo, const void* skip, void* y, int n, int d, int h, int w, int c,
                                       int dtype, mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(c % vn == 0 && n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x: bad shape C=%d", c);
  const int64_t total = static_cast<int64_t>(n) * d * h * w * 8 * (c / vn);
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(num_sms()) * 16));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (upsample2x_add_fwd_kernel<T><<<blocks, 256, 0, s>>>(
                                    static_cast<const T*>(x_lo), static_cast<const T*>(skip), static_cast<T*>(y), n, d,
                                    h, w, c)));
  MMPL_CHECK_LAUNCH("upsample2x_add_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_upsample2x_bwd(const void* dy, void* dx_lo, int n, int d, int h, int w, int c, int dtype,
                                   mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(c % vn == 0 && n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x: bad shape C=%d", c);
  const int64_t total = static_cast<int64_t>(n) * d * h * w * (c / vn);
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(num_sms()) * 16));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (upsample2x_bwd_kernel<T><<<blocks, 256, 0, s>>>(
                                    static_cast<const T*>(dy), static_cast<T*>(dx_lo), n, d, h, w, c)));
  MMPL_CHECK_LAUNCH("upsample2x_bwd");
  return MMPL_OK;
}
