// Trilinear x2 up-sampling (align_corners=False) fused with the additive skip connection, and its backward.
// Reference: nn.Upsample(scale_factor=2, mode='trilinear') + `x = x + skipN`, unet3D.py:608, :686-687.
// With an exact factor of 2 the source coordinate (dst+0.5)/2-0.5 gives fixed weights per axis:
//   out[2i]   = 1/4 in[i-1] + 3/4 in[i]      out[2i+1] = 3/4 in[i] + 1/4 in[i+1]     (indices clamped to [0, n-1])
// The backward is written as a gather (no atomics): in[i] collects 1/4,3/4,3/4,1/4 of dy[2i-1..2i+2], where a tap
// that falls outside is redirected to the edge output that clamped onto in[i].
#include <algorithm>

#include "common.cuh"

namespace mmpl {
namespace {

// Both kernels exploit the separability of the interpolation along the depth axis: a thread owns one (h, w, channel
// vector) column and WALKS the depth axis, carrying the in-plane (H,W) interpolation of the neighbouring planes in
// registers.  Forward: 2 gathered loads + 6 flops per output element instead of 8 loads + 15 flops; backward: 32 loads
// per input element instead of 64.  All index arithmetic is 32-bit and per-thread constant.
//
// Forward grid: (ceil(2W * C/VN / 256), 2H, N).  A thread keeps the same channel vector for the whole kernel, so it can
// also carry the GroupNorm(16) partial sums of the OUTPUT (the next block's gn1 / downsample.0 statistics) in
// registers: one fp64 atomic per (group, moment) per block at the end.
template <typename T, bool STATS>
__global__ void __launch_bounds__(256)
upsample2x_add_fwd_kernel(const T* __restrict__ xlo, const T* __restrict__ skip, T* __restrict__ y,
                          double* __restrict__ stats, int D, int H, int W, int C, int groups) {
  constexpr int VN = Vec<T>::N;
  const int vpv = C / VN;
  const int Ho = 2 * H, Wo = 2 * W;
  const int n = blockIdx.z, ho = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;
  const bool active = e < Wo * vpv;
  const int wo = active ? e / vpv : 0, cv = active ? e - wo * vpv : 0;
  float ds[VN], dq[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) ds[i] = dq[i] = 0.f;
  if (active) {
    int ih0, ih1, iw0, iw1;
    float wh0, wh1, ww0, ww1;
    if (ho & 1) {
      ih0 = ho >> 1, wh0 = 0.75f, ih1 = min((ho >> 1) + 1, H - 1), wh1 = 0.25f;
    } else {
      ih0 = max((ho >> 1) - 1, 0), wh0 = 0.25f, ih1 = ho >> 1, wh1 = 0.75f;
    }
    if (wo & 1) {
      iw0 = wo >> 1, ww0 = 0.75f, iw1 = min((wo >> 1) + 1, W - 1), ww1 = 0.25f;
    } else {
      iw0 = max((wo >> 1) - 1, 0), ww0 = 0.25f, iw1 = wo >> 1, ww1 = 0.75f;
    }
    const int64_t plane = static_cast<int64_t>(H) * W * C;
    const T* xn = xlo + static_cast<int64_t>(n) * D * plane + cv * VN;
    const T* p00 = xn + (static_cast<int64_t>(ih0) * W + iw0) * C;
    const T* p01 = xn + (static_cast<int64_t>(ih0) * W + iw1) * C;
    const T* p10 = xn + (static_cast<int64_t>(ih1) * W + iw0) * C;
    const T* p11 = xn + (static_cast<int64_t>(ih1) * W + iw1) * C;
    const int64_t oplane = static_cast<int64_t>(Ho) * Wo * C;
    const int64_t o0 = static_cast<int64_t>(n) * (2 * D) * oplane + (static_cast<int64_t>(ho) * Wo + wo) * C + cv * VN;
    // in-plane interpolation of an input plane, nested as PyTorch's upsample_trilinear3d does: w_h*(w_w a + w_w b) + ...
    auto hw = [&](const Vec<T>& a, const Vec<T>& b, const Vec<T>& c, const Vec<T>& q, float (&up)[VN]) {
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        float ph = 0.f;
        ph += wh0 * (ww0 * a.v[k] + ww1 * b.v[k]);
        ph += wh1 * (ww0 * c.v[k] + ww1 * q.v[k]);
        up[k] = ph;
      }
    };
    auto emit = [&](int dd, Vec<T>& acc, const float (&u0)[VN], float w0, const float (&u1)[VN], float w1) {
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        float up = 0.f;
        up += w0 * u0[k];
        up += w1 * u1[k];
        acc.v[k] += up;
      }
      acc.store(y + o0 + static_cast<int64_t>(dd) * oplane);
      if (STATS) {
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          const float r = to_f32<T>(from_f32<T>(acc.v[k]));   // statistics of the values as stored
          ds[k] += r;
          dq[k] = fmaf(r, r, dq[k]);
        }
      }
    };
    float prev[VN], cur[VN], nxt[VN];
    {
      Vec<T> a, b, c, q;
      a.load(p00), b.load(p01), c.load(p10), q.load(p11);
      hw(a, b, c, q, cur);
    }
#pragma unroll
    for (int k = 0; k < VN; ++k) prev[k] = cur[k];
    for (int d = 0; d < D; ++d) {
      // all six loads of this step are issued before any arithmetic or store (loads in flight bound this kernel)
      Vec<T> a, b, c, q, s0, s1;
      const bool more = d + 1 < D;
      if (more) {
        const int64_t off = static_cast<int64_t>(d + 1) * plane;
        a.load(p00 + off), b.load(p01 + off), c.load(p10 + off), q.load(p11 + off);
      }
      s0.load(skip + o0 + static_cast<int64_t>(2 * d) * oplane);
      s1.load(skip + o0 + static_cast<int64_t>(2 * d + 1) * oplane);
      if (more) {
        hw(a, b, c, q, nxt);
      } else {
#pragma unroll
        for (int k = 0; k < VN; ++k) nxt[k] = cur[k];
      }
      emit(2 * d, s0, prev, 0.25f, cur, 0.75f);        // taps (max(d-1,0), d)
      emit(2 * d + 1, s1, cur, 0.75f, nxt, 0.25f);     // taps (d, min(d+1,D-1))
#pragma unroll
      for (int k = 0; k < VN; ++k) prev[k] = cur[k], cur[k] = nxt[k];
    }
  }
  if (STATS) {
    __shared__ double sg[32][2];
    if (threadIdx.x < 64) (&sg[0][0])[threadIdx.x] = 0.0;
    __syncthreads();
    const int cpg = C / groups;
    const int gpt = cpg >= VN ? 1 : VN / cpg;  // groups per thread
    const int cpp = VN / gpt;                  // channels per partial
    // lanes that share a channel vector are vpv apart (the host guarantees 256 % vpv == 0, so also 32 % vpv == 0 or
    // vpv % 32 == 0); inactive threads contribute zeros
    for (int j = 0; j < gpt; ++j) {
      double a = 0, b = 0;
      for (int i = 0; i < cpp; ++i) a += static_cast<double>(ds[j * cpp + i]), b += static_cast<double>(dq[j * cpp + i]);
      for (int o = 16; o >= vpv && o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      const bool leader = vpv >= 32 ? true : ((threadIdx.x & 31) < vpv);
      if (leader) {
        const int g = ((threadIdx.x % vpv) * VN + j * cpp) / cpg;
        atomicAdd(&sg[g][0], a);
        atomicAdd(&sg[g][1], b);
      }
    }
    __syncthreads();
    if (threadIdx.x < groups * 2) {
      const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
      atomicAdd(&stats[(static_cast<int64_t>(n) * groups + g) * 2 + k], sg[g][k]);
    }
  }
}

// Backward grid: (ceil(W * C/VN / 256), H, N); a thread owns input column (h, w, channel vector).  With
// t[o] = sum_{b,c} wt[b] wt[c] dy[o, oh[b], ow[c]] (16 taps of output plane o):
//   dx[d] = 1/4 t[2d-1] + 3/4 t[2d] + 3/4 t[2d+1] + 1/4 t[2d+2], out-of-range planes redirected to the edge plane that
// clamped onto d in the forward.  Two new t planes per step, two carried over.
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dxlo, int D, int H, int W, int C) {
  using V = VecH<T>;   // 8-byte vectors: half the per-thread state, twice the threads
  constexpr int VN = V::N;
  const int vpv = C / VN;
  const int Ho = 2 * H, Wo = 2 * W;
  const int n = blockIdx.z, hi = blockIdx.y;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= W * vpv) return;
  const int wi = e / vpv, cv = e - wi * vpv;
  const int oh[4] = {hi > 0 ? 2 * hi - 1 : 0, 2 * hi, 2 * hi + 1, hi < H - 1 ? 2 * hi + 2 : 2 * H - 1};
  const int ow[4] = {wi > 0 ? 2 * wi - 1 : 0, 2 * wi, 2 * wi + 1, wi < W - 1 ? 2 * wi + 2 : 2 * W - 1};
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  const int64_t oplane = static_cast<int64_t>(Ho) * Wo * C;
  const T* dyn = dy + static_cast<int64_t>(n) * (2 * D) * oplane + cv * VN;
  int64_t roff[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) roff[b] = static_cast<int64_t>(oh[b]) * Wo * C;
  int coff[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) coff[c] = ow[c] * C;
  auto tplane = [&](int o, float (&t)[VN]) {
    const T* pl = dyn + static_cast<int64_t>(o) * oplane;
#pragma unroll
    for (int k = 0; k < VN; ++k) t[k] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      V g[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) g[c].load(pl + roff[b] + coff[c]);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float r = (0.25f * g[0].v[k] + 0.75f * g[1].v[k]) + (0.75f * g[2].v[k] + 0.25f * g[3].v[k]);
        t[k] = fmaf(wt[b], r, t[k]);
      }
    }
  };
  const int64_t plane = static_cast<int64_t>(H) * W * C;
  T* out = dxlo + static_cast<int64_t>(n) * D * plane + (static_cast<int64_t>(hi) * W + wi) * C + cv * VN;
  float tm[VN], ta[VN], tb[VN], tp[VN];
  tplane(0, ta);
  tplane(1, tb);
#pragma unroll
  for (int k = 0; k < VN; ++k) tm[k] = ta[k];
  for (int d = 0; d < D; ++d) {
    if (d + 1 < D) {
      tplane(2 * d + 2, tp);
    } else {
#pragma unroll
      for (int k = 0; k < VN; ++k) tp[k] = tb[k];
    }
    V o;
#pragma unroll
    for (int k = 0; k < VN; ++k) o.v[k] = (0.25f * tm[k] + 0.75f * ta[k]) + (0.75f * tb[k] + 0.25f * tp[k]);
    o.store(out + static_cast<int64_t>(d) * plane);
#pragma unroll
    for (int k = 0; k < VN; ++k) tm[k] = tb[k], ta[k] = tp[k];
    if (d + 1 < D) tplane(2 * d + 3, tb);
  }
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_upsample2x_add_fwd(const void* x_lo, const void* skip, void* y, int n, int d, int h, int w, int c,
                                       int dtype, void* gn_stats, mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(c % vn == 0 && n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x: bad shape C=%d", c);
  MMPL_REQUIRE(gn_stats == nullptr || (256 % (c / vn) == 0 && c % 16 == 0 && c <= 512), MMPL_E_SHAPE,
               "upsample2x: fused GroupNorm statistics need C/%d to divide 256 and C %% 16 == 0 (C=%d)", vn, c);
  MMPL_REQUIRE(2 * h <= 65535 && n <= 65535, MMPL_E_SHAPE, "upsample2x: H=%d N=%d exceed the grid limits", h, n);
  dim3 grid(ceil_div(static_cast<int64_t>(2 * w) * (c / vn), 256), 2 * h, n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, {
    if (gn_stats)
      upsample2x_add_fwd_kernel<T, true><<<grid, 256, 0, s>>>(static_cast<const T*>(x_lo), static_cast<const T*>(skip),
                                                             static_cast<T*>(y), static_cast<double*>(gn_stats), d, h, w,
                                                             c, 16);
    else
      upsample2x_add_fwd_kernel<T, false><<<grid, 256, 0, s>>>(static_cast<const T*>(x_lo), static_cast<const T*>(skip),
                                                              static_cast<T*>(y), nullptr, d, h, w, c, 16);
  });
  MMPL_CHECK_LAUNCH("upsample2x_add_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_upsample2x_bwd(const void* dy, void* dx_lo, int n, int d, int h, int w, int c, int dtype,
                                   mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(c % vn == 0 && n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x: bad shape C=%d", c);
  MMPL_REQUIRE(h <= 65535 && n <= 65535, MMPL_E_SHAPE, "upsample2x: H=%d N=%d exceed the grid limits", h, n);
  dim3 grid(ceil_div(static_cast<int64_t>(w) * (c / (vn / 2)), 256), h, n);   // 8-byte vectors in the backward
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (upsample2x_bwd_kernel<T><<<grid, 256, 0, s>>>(
                                    static_cast<const T*>(dy), static_cast<T*>(dx_lo), d, h, w, c)));
  MMPL_CHECK_LAUNCH("upsample2x_bwd");
  return MMPL_OK;
}
