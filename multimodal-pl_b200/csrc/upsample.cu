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

// Forward: grid (row chunks, N); a block walks a contiguous range of output rows (do, ho) so all index arithmetic
// per element is 32-bit and per-row quantities are block-uniform.  With 256 % (C / VN) == 0 a thread keeps the same
// channel vector for the whole kernel, which lets it carry the GroupNorm(16) partial sums of the OUTPUT (the next
// block's gn1 / downsample.0 statistics) in registers: one fp64 atomic per (group, moment) per block at the end.
template <typename T, bool STATS>
__global__ void __launch_bounds__(256)
upsample2x_add_fwd_kernel(const T* __restrict__ xlo, const T* __restrict__ skip, T* __restrict__ y,
                          double* __restrict__ stats, int D, int H, int W, int C, int groups, int rows_per_block) {
  constexpr int VN = Vec<T>::N;
  const int vpv = C / VN;
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
  const int n = blockIdx.y;
  const int rows = Do * Ho, row_elems = Wo * vpv;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows);
  const T* xn = xlo + static_cast<int64_t>(n) * D * H * W * C;
  const int64_t out_n = static_cast<int64_t>(n) * rows * Wo * C;
  float ds[VN], dq[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) ds[i] = dq[i] = 0.f;
  for (int row = r0; row < r1; ++row) {
    const int dd = row / Ho, ho = row - dd * Ho;
    int id0, id1, ih0, ih1;
    float wd0, wd1, wh0, wh1;
    if (dd & 1) {
      id0 = dd >> 1, wd0 = 0.75f, id1 = min((dd >> 1) + 1, D - 1), wd1 = 0.25f;
    } else {
      id0 = max((dd >> 1) - 1, 0), wd0 = 0.25f, id1 = dd >> 1, wd1 = 0.75f;
    }
    if (ho & 1) {
      ih0 = ho >> 1, wh0 = 0.75f, ih1 = min((ho >> 1) + 1, H - 1), wh1 = 0.25f;
    } else {
      ih0 = max((ho >> 1) - 1, 0), wh0 = 0.25f, ih1 = ho >> 1, wh1 = 0.75f;
    }
    const T* r00 = xn + (static_cast<int64_t>(id0) * H + ih0) * W * C;
    const T* r01 = xn + (static_cast<int64_t>(id0) * H + ih1) * W * C;
    const T* r10 = xn + (static_cast<int64_t>(id1) * H + ih0) * W * C;
    const T* r11 = xn + (static_cast<int64_t>(id1) * H + ih1) * W * C;
    const int64_t orow = out_n + static_cast<int64_t>(row) * Wo * C;
    for (int e = threadIdx.x; e < row_elems; e += 256) {
      const int wo = e / vpv, cv = e - wo * vpv;
      int iw0, iw1;
      float ww0, ww1;
      if (wo & 1) {
        iw0 = wo >> 1, ww0 = 0.75f, iw1 = min((wo >> 1) + 1, W - 1), ww1 = 0.25f;
      } else {
        iw0 = max((wo >> 1) - 1, 0), ww0 = 0.25f, iw1 = wo >> 1, ww1 = 0.75f;
      }
      const int o0 = iw0 * C + cv * VN, o1 = iw1 * C + cv * VN;
      Vec<T> acc, p00, q00, p01, q01, p10, q10, p11, q11;
      acc.load(skip + orow + static_cast<int64_t>(e) * VN);
      p00.load(r00 + o0), q00.load(r00 + o1);
      p01.load(r01 + o0), q01.load(r01 + o1);
      p10.load(r10 + o0), q10.load(r10 + o1);
      p11.load(r11 + o0), q11.load(r11 + o1);
      // PyTorch (upsample_trilinear3d) sums the eight taps as w_d*(w_h*(w_w a + w_w b) + ...); keep that nesting.
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        float ph0 = 0.f, ph1 = 0.f, up = 0.f;
        ph0 += wh0 * (ww0 * p00.v[k] + ww1 * q00.v[k]);
        ph0 += wh1 * (ww0 * p01.v[k] + ww1 * q01.v[k]);
        ph1 += wh0 * (ww0 * p10.v[k] + ww1 * q10.v[k]);
        ph1 += wh1 * (ww0 * p11.v[k] + ww1 * q11.v[k]);
        up += wd0 * ph0;
        up += wd1 * ph1;
        acc.v[k] += up;
      }
      acc.store(y + orow + static_cast<int64_t>(e) * VN);
      if (STATS) {
#pragma unroll
        for (int k = 0; k < VN; ++k) {
          const float r = to_f32<T>(from_f32<T>(acc.v[k]));   // statistics of the values as stored
          ds[k] += r;
          dq[k] = fmaf(r, r, dq[k]);
        }
      }
    }
  }
  if (STATS) {
    __shared__ double sg[32][2];
    if (threadIdx.x < 64) (&sg[0][0])[threadIdx.x] = 0.0;
    __syncthreads();
    const int cv = threadIdx.x % vpv;          // constant per thread: the host guarantees 256 % vpv == 0
    const int cpg = C / groups;
    const int gpt = cpg >= VN ? 1 : VN / cpg;  // groups per thread
    const int cpp = VN / gpt;                  // channels per partial
    for (int j = 0; j < gpt; ++j) {
      double a = 0, b = 0;
      for (int i = 0; i < cpp; ++i) a += static_cast<double>(ds[j * cpp + i]), b += static_cast<double>(dq[j * cpp + i]);
      for (int o = 16; o >= vpv && o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      const bool leader = vpv >= 32 ? true : ((threadIdx.x & 31) < vpv);
      if (leader) {
        const int g = (cv * VN + j * cpp) / cpg;
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

// Backward: same decomposition over INPUT rows (d, h); each element gathers its 4 x 4 x 4 output taps (L1-resident).
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dxlo, int D, int H, int W, int C, int rows_per_block) {
  constexpr int VN = Vec<T>::N;
  const int vpv = C / VN;
  const int Ho = 2 * H, Wo = 2 * W;
  const int n = blockIdx.y;
  const int rows = D * H, row_elems = W * vpv;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows);
  const T* dyn = dy + static_cast<int64_t>(n) * (2 * D) * Ho * Wo * C;
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (int row = r0; row < r1; ++row) {
    const int di = row / H, hi = row - di * H;
    const int od[4] = {di > 0 ? 2 * di - 1 : 0, 2 * di, 2 * di + 1, di < D - 1 ? 2 * di + 2 : 2 * D - 1};
    const int oh[4] = {hi > 0 ? 2 * hi - 1 : 0, 2 * hi, 2 * hi + 1, hi < H - 1 ? 2 * hi + 2 : 2 * H - 1};
    const int64_t orow = (static_cast<int64_t>(n) * rows + row) * W * C;
    for (int e = threadIdx.x; e < row_elems; e += 256) {
      const int wi = e / vpv, cv = e - wi * vpv;
      const int ow[4] = {wi > 0 ? 2 * wi - 1 : 0, 2 * wi, 2 * wi + 1, wi < W - 1 ? 2 * wi + 2 : 2 * W - 1};
      float acc[VN];
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float wab = wt[a] * wt[b];
          const T* rowp = dyn + (static_cast<int64_t>(od[a]) * Ho + oh[b]) * Wo * C + cv * VN;
          Vec<T> g[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) g[c].load(rowp + ow[c] * C);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float w3 = wab * wt[c];
#pragma unroll
            for (int k = 0; k < VN; ++k) acc[k] = fmaf(w3, g[c].v[k], acc[k]);
          }
        }
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) o.v[k] = acc[k];
      o.store(dxlo + orow + static_cast<int64_t>(e) * VN);
    }
  }
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

static int rows_per_block_for(int rows, int n, int target_blocks) {
  const int per_sample = std::max(1, target_blocks / std::max(n, 1));
  return std::max(1, (rows + per_sample - 1) / per_sample);
}

extern "C" int mmpl_upsample2x_add_fwd(const void* x_lo, const void* skip, void* y, int n, int d, int h, int w, int c,
                                       int dtype, void* gn_stats, mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(c % vn == 0 && n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x: bad shape C=%d", c);
  MMPL_REQUIRE(gn_stats == nullptr || (256 % (c / vn) == 0 && c % 16 == 0 && c <= 512), MMPL_E_SHAPE,
               "upsample2x: fused GroupNorm statistics need C/%d to divide 256 and C %% 16 == 0 (C=%d)", vn, c);
  const int rows = 4 * d * h;
  const int rpb = rows_per_block_for(rows, n, num_sms() * 8);
  dim3 grid((rows + rpb - 1) / rpb, n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, {
    if (gn_stats)
      upsample2x_add_fwd_kernel<T, true><<<grid, 256, 0, s>>>(static_cast<const T*>(x_lo), static_cast<const T*>(skip),
                                                             static_cast<T*>(y), static_cast<double*>(gn_stats), d, h, w,
                                                             c, 16, rpb);
    else
      upsample2x_add_fwd_kernel<T, false><<<grid, 256, 0, s>>>(static_cast<const T*>(x_lo), static_cast<const T*>(skip),
                                                              static_cast<T*>(y), nullptr, d, h, w, c, 16, rpb);
  });
  MMPL_CHECK_LAUNCH("upsample2x_add_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_upsample2x_bwd(const void* dy, void* dx_lo, int n, int d, int h, int w, int c, int dtype,
                                   mmpl_stream_t stream) {
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(c % vn == 0 && n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "upsample2x: bad shape C=%d", c);
  const int rows = d * h;
  const int rpb = rows_per_block_for(rows, n, num_sms() * 8);
  dim3 grid((rows + rpb - 1) / rpb, n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (upsample2x_bwd_kernel<T><<<grid, 256, 0, s>>>(
                                    static_cast<const T*>(dy), static_cast<T*>(dx_lo), d, h, w, c, rpb)));
  MMPL_CHECK_LAUNCH("upsample2x_bwd");
  return MMPL_OK;
}
