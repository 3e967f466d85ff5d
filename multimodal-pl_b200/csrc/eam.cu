// Kernels for the class-token attention maps of unet3D_with_feam3 (reference unet3D.py:142-212, :1051-1068, :1127-1175).
//
// What the model consumes of EAM.forward is only ``attn.mean(1)`` -- the mean over heads of the UNSCALED q.k^T logits.
// The head mean of per-head dot products is one dot product over all channels:
//     amap[t][v] = (1/H) * sum_c q[t][c] * k[v][c],   k[v] = Wk * LayerNorm2(x[v]),   q[t] = Wq * LayerNorm3(token[t])
// so with M = q * Wk (a 15 x C matrix, a few kFLOP, host-side autograd) the map is a 15-class "classifier" over the
// LayerNorm-ed voxel rows:  amap = (M (.) gamma2) xhat / H + (M beta2) / H.   The device work is therefore
//   (1) mmpl_ln_rows_{fwd,bwd}: LayerNorm over the channel axis of an NDHWC tensor without affine (the affine is folded
//       into M), one pass each;
//   (2) the classifier kernels (mmpl_cls_fwd / mmpl_cls_bwd), here extended by a generic-width pair for the channel
//       counts the fast kernels do not cover (128 at 1/8 resolution: eam84 and the deepout1 head);
//   (3) mmpl_token_stats / mmpl_token_ema: renew_token -- per-class masked mean of a feature map (nearest-neighbour
//       down-sampled label volume) and the EMA update of the class tokens, without the reference's 3 x 15 host syncs.
#include <algorithm>

#include "common.cuh"

namespace mmpl {
namespace {

// ------------------------------------------------------------------------------------------------ LayerNorm rows
// rows x C (C / Vec<T>::N lanes per row, a power of two <= 32).  y = (x - mean) * rstd, biased variance, eps inside sqrt.
template <typename T>
__global__ void __launch_bounds__(256)
ln_rows_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, float* __restrict__ rstd_out, int64_t rows, int C, float eps) {
  constexpr int VN = Vec<T>::N;
  const int lpr = C / VN;                                   // lanes per row
  const int lane = threadIdx.x & 31;
  const int sub = lane / lpr, li = lane % lpr, rpw = 32 / lpr;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r0 = warp * rpw; r0 < rows; r0 += nwarps * rpw) {
    const int64_t r = r0 + sub;
    const bool ok = r < rows;
    Vec<T> v;
    if (ok) {
      v.load(x + r * C + li * VN);
    } else {
#pragma unroll
      for (int k = 0; k < VN; ++k) v.v[k] = 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VN; ++k) s += v.v[k];
    for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      const float d = v.v[k] - mean;
      q = fmaf(d, d, q);
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
    if (ok) {
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) o.v[k] = (v.v[k] - mean) * rstd;
      o.store(y + r * C + li * VN);
      if (li == 0) rstd_out[r] = rstd;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), xhat = the stored forward output
template <typename T>
__global__ void __launch_bounds__(256)
ln_rows_bwd_kernel(const T* __restrict__ xhat, const float* __restrict__ rstd, const T* __restrict__ g, T* __restrict__ dx,
                   int64_t rows, int C) {
  constexpr int VN = Vec<T>::N;
  const int lpr = C / VN;
  const int lane = threadIdx.x & 31;
  const int sub = lane / lpr, li = lane % lpr, rpw = 32 / lpr;
  const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r0 = warp * rpw; r0 < rows; r0 += nwarps * rpw) {
    const int64_t r = r0 + sub;
    const bool ok = r < rows;
    Vec<T> xv, gv;
    if (ok) {
      xv.load(xhat + r * C + li * VN);
      gv.load(g + r * C + li * VN);
    } else {
#pragma unroll
      for (int k = 0; k < VN; ++k) xv.v[k] = gv.v[k] = 0.f;
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      s1 += gv.v[k];
      s2 = fmaf(gv.v[k], xv.v[k], s2);
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (ok) {
      const float m1 = s1 / C, m2 = s2 / C, rs = rstd[r];
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) o.v[k] = rs * (gv.v[k] - m1 - xv.v[k] * m2);
      o.store(dx + r * C + li * VN);
    }
  }
}

// ------------------------------------------------------------------------------------------------ generic classifier
// Any cin that is a multiple of 8 (bf16) / 4 (fp32), classes <= 16.  Only small tensors come here (128 channels at 1/8
// resolution: 9 216 voxels per sample at cfg2), so clarity beats tuning: weights in shared memory, one thread per voxel.
template <typename T>
__global__ void __launch_bounds__(256)
cls_fwd_generic_kernel(const T* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ bias,
                       float* __restrict__ logits, int N, int64_t S, int cin, int classes) {
  extern __shared__ float sw[];      // [cin][16]
  for (int i = threadIdx.x; i < 16 * cin; i += blockDim.x) {
    const int c = i / cin, k = i % cin;
    sw[k * 16 + c] = c < classes ? wc[c * cin + k] : 0.f;
  }
  __syncthreads();
  constexpr int VN = Vec<T>::N;
  const int64_t total = static_cast<int64_t>(N) * S;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = c < classes ? bias[c] : 0.f;
    for (int k0 = 0; k0 < cin; k0 += VN) {
      Vec<T> x;
      x.load(a + v * cin + k0);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float* w = sw + (k0 + k) * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = fmaf(x.v[k], w[c], acc[c]);
      }
    }
    const int64_t n = v / S, s = v - n * S;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < classes) logits[(n * classes + c) * S + s] = acc[c];
  }
}

// da[v][k] = sum_c dl[c][v] W[c][k]
template <typename T>
__global__ void __launch_bounds__(256)
cls_bwd_da_generic_kernel(const float* __restrict__ wc, const float* __restrict__ dl, T* __restrict__ da, int N, int64_t S,
                          int cin, int classes) {
  extern __shared__ float sw[];      // [cin][16]
  for (int i = threadIdx.x; i < 16 * cin; i += blockDim.x) {
    const int c = i / cin, k = i % cin;
    sw[k * 16 + c] = c < classes ? wc[c * cin + k] : 0.f;
  }
  __syncthreads();
  constexpr int VN = Vec<T>::N;
  const int64_t total = static_cast<int64_t>(N) * S;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t n = v / S, s = v - n * S;
    float g[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) g[c] = c < classes ? dl[(n * classes + c) * S + s] : 0.f;
    for (int k0 = 0; k0 < cin; k0 += VN) {
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        const float* w = sw + (k0 + k) * 16;
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) t = fmaf(g[c], w[c], t);
        o.v[k] = t;
      }
      o.store(da + v * cin + k0);
    }
  }
}

// dW[c][k] += sum_v dl[c][v] a[v][k], db[c] += sum_v dl[c][v].  Thread = one input channel k of one voxel sub-stream.
template <typename T>
__global__ void __launch_bounds__(256)
cls_bwd_dw_generic_kernel(const T* __restrict__ a, const float* __restrict__ dl, float* __restrict__ dwc,
                          float* __restrict__ dbias, int N, int64_t S, int cin, int classes, int64_t vox_per_block) {
  const int streams = blockDim.x / cin;               // cin <= 256 and divides 256 (32, 64, 128, 256)
  const int k = threadIdx.x % cin, sub = threadIdx.x / cin;
  const int64_t total = static_cast<int64_t>(N) * S;
  const int64_t v0 = blockIdx.x * vox_per_block, v1 = min(v0 + vox_per_block, total);
  float acc[16], accb[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = accb[c] = 0.f;
  if (sub < streams) {
    for (int64_t v = v0 + sub; v < v1; v += streams) {
      const int64_t n = v / S, s = v - n * S;
      const float x = to_f32<T>(a[v * cin + k]);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float g = c < classes ? dl[(n * classes + c) * S + s] : 0.f;     // same address across the k threads: broadcast
        acc[c] = fmaf(g, x, acc[c]);
        accb[c] += g;
      }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < classes) {
        atomicAdd(&dwc[c * cin + k], acc[c]);
        if (k == 0) atomicAdd(&dbias[c], accb[c]);
      }
  }
}

// ------------------------------------------------------------------------------------------------ renew_token
// sums[l][c] += x[n][v][c] over the voxels whose (nearest-neighbour down-sampled) label is l + 1; cnt[l] += 1.
// mask: [N][Dm][Hm][Wm] class ids (fp32 or uint8), features NDHWC [N][D][H][W][C]; nearest: src = floor(dst * Dm / D).
template <typename T>
__global__ void __launch_bounds__(256)
token_stats_kernel(const T* __restrict__ x, const void* __restrict__ mask, int mask_u8, float* __restrict__ sums,
                   float* __restrict__ cnt, int N, int D, int H, int W, int C, int Dm, int Hm, int Wm, int ntok) {
  extern __shared__ float s_acc[];      // [ntok][C] + [ntok]
  float* s_cnt = s_acc + ntok * C;
  for (int i = threadIdx.x; i < ntok * C + ntok; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  constexpr int VN = Vec<T>::N;
  const int lpv = C / VN;                                  // lanes per voxel
  const int64_t vox = static_cast<int64_t>(N) * D * H * W;
  const int64_t units = vox * lpv;
  for (int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; u < units;
       u += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t v = u / lpv;
    const int li = static_cast<int>(u - v * lpv);
    int64_t r = v;
    const int w = static_cast<int>(r % W);
    r /= W;
    const int h = static_cast<int>(r % H);
    r /= H;
    const int d = static_cast<int>(r % D);
    const int n = static_cast<int>(r / D);
    // F.interpolate(mode="nearest"): src = floor(dst * in / out)
    const int md = min(static_cast<int>((static_cast<int64_t>(d) * Dm) / D), Dm - 1);
    const int mh = min(static_cast<int>((static_cast<int64_t>(h) * Hm) / H), Hm - 1);
    const int mw = min(static_cast<int>((static_cast<int64_t>(w) * Wm) / W), Wm - 1);
    const int64_t mi = ((static_cast<int64_t>(n) * Dm + md) * Hm + mh) * Wm + mw;
    const float lv = mask_u8 ? static_cast<float>(static_cast<const uint8_t*>(mask)[mi]) : static_cast<const float*>(mask)[mi];
    const int l = static_cast<int>(lv);
    if (static_cast<float>(l) != lv || l < 1 || l > ntok) continue;
    Vec<T> xv;
    xv.load(x + v * C + li * VN);
#pragma unroll
    for (int k = 0; k < VN; ++k) atomicAdd(&s_acc[(l - 1) * C + li * VN + k], xv.v[k]);
    if (li == 0) atomicAdd(&s_cnt[l - 1], 1.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntok * C; i += blockDim.x)
    if (s_acc[i] != 0.f) atomicAdd(&sums[i], s_acc[i]);
  for (int i = threadIdx.x; i < ntok; i += blockDim.x)
    if (s_cnt[i] != 0.f) atomicAdd(&cnt[i], s_cnt[i]);
}

// token[l] = token[l] * (1 - alpha) + mean_l * alpha where class l + 1 occurs (cnt > 0); untouched otherwise
__global__ void token_ema_kernel(float* __restrict__ token, const float* __restrict__ sums, const float* __restrict__ cnt,
                                 int ntok, int C, float alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntok * C) return;
  const float n = cnt[i / C];
  if (n > 0.f) token[i] = token[i] * (1.f - alpha) + (sums[i] / n) * alpha;
}

}  // namespace

int cls_fwd_generic(const void* a, const float* wc, const float* bias, float* logits, int n, int64_t spatial, int cin,
                    int classes, int dtype, cudaStream_t s) {
  const int64_t total = static_cast<int64_t>(n) * spatial;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  const size_t smem = sizeof(float) * 16 * cin;
  MMPL_REQUIRE(smem <= 48 * 1024, MMPL_E_UNSUPPORTED, "cls_fwd: cin=%d too wide for the generic kernel", cin);
  MMPL_DISPATCH_DTYPE(dtype, T, (cls_fwd_generic_kernel<T><<<blocks, 256, smem, s>>>(static_cast<const T*>(a), wc, bias, logits,
                                                                                   n, spatial, cin, classes)));
  return MMPL_OK;
}

int cls_bwd_generic(const void* a, const float* wc, const float* dl, void* da, float* dwc, float* dbias, int n,
                    int64_t spatial, int cin, int classes, int dtype, cudaStream_t s) {
  const int64_t total = static_cast<int64_t>(n) * spatial;
  const size_t smem = sizeof(float) * 16 * cin;
  MMPL_REQUIRE(smem <= 48 * 1024 && cin <= 256 && 256 % cin == 0, MMPL_E_UNSUPPORTED,
               "cls_bwd: cin=%d is not covered by the generic kernel (a divisor of 256)", cin);
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(num_sms()) * 8));
  const int wblocks = static_cast<int>(std::min<int64_t>((total + 63) / 64, static_cast<int64_t>(num_sms()) * 4));
  const int64_t vpb = (total + wblocks - 1) / wblocks;
  MMPL_DISPATCH_DTYPE(dtype, T, {
    cls_bwd_da_generic_kernel<T><<<blocks, 256, smem, s>>>(wc, dl, static_cast<T*>(da), n, spatial, cin, classes);
    cls_bwd_dw_generic_kernel<T><<<wblocks, 256, 0, s>>>(static_cast<const T*>(a), dl, dwc, dbias, n, spatial, cin, classes, vpb);
  });
  return MMPL_OK;
}

}  // namespace mmpl

using namespace mmpl;

static int ln_check(int64_t rows, int c, int dtype) {
  MMPL_REQUIRE(rows > 0, MMPL_E_SHAPE, "ln_rows: empty input");
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  const int lpr = c / vn;
  MMPL_REQUIRE(c % vn == 0 && lpr >= 1 && lpr <= 32 && (lpr & (lpr - 1)) == 0, MMPL_E_SHAPE,
               "ln_rows: C=%d (C / %d must be a power of two <= 32)", c, vn);
  return MMPL_OK;
}

extern "C" int mmpl_ln_rows_fwd(const void* x, void* y, float* rstd, int64_t rows, int c, float eps, int dtype,
                                mmpl_stream_t stream) {
  if (int e = ln_check(rows, c, dtype)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  const int64_t rpw = 32 / (c / vn);
  const int blocks = static_cast<int>(std::min<int64_t>((rows + rpw * 8 - 1) / (rpw * 8), static_cast<int64_t>(num_sms()) * 8));
  MMPL_DISPATCH_DTYPE(dtype, T, (ln_rows_fwd_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(x), static_cast<T*>(y), rstd,
                                                                            rows, c, eps)));
  MMPL_CHECK_LAUNCH("ln_rows_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_ln_rows_bwd(const void* xhat, const float* rstd, const void* g, void* dx, int64_t rows, int c, int dtype,
                                mmpl_stream_t stream) {
  if (int e = ln_check(rows, c, dtype)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  const int64_t rpw = 32 / (c / vn);
  const int blocks = static_cast<int>(std::min<int64_t>((rows + rpw * 8 - 1) / (rpw * 8), static_cast<int64_t>(num_sms()) * 8));
  MMPL_DISPATCH_DTYPE(dtype, T, (ln_rows_bwd_kernel<T><<<blocks, 256, 0, s>>>(static_cast<const T*>(xhat), rstd,
                                                                            static_cast<const T*>(g), static_cast<T*>(dx), rows, c)));
  MMPL_CHECK_LAUNCH("ln_rows_bwd");
  return MMPL_OK;
}

extern "C" int mmpl_token_stats(const void* x, const void* mask, int mask_is_u8, float* sums, float* counts, int n, int d,
                                int h, int w, int c, int dm, int hm, int wm, int ntok, int dtype, mmpl_stream_t stream) {
  MMPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && dm > 0 && hm > 0 && wm > 0, MMPL_E_SHAPE, "token_stats: empty input");
  MMPL_REQUIRE(ntok >= 1 && ntok <= 31, MMPL_E_SHAPE, "token_stats: ntok=%d", ntok);
  const int vn = dtype == MMPL_BF16 ? 8 : 4;
  MMPL_REQUIRE(c % vn == 0, MMPL_E_SHAPE, "token_stats: C=%d", c);
  const size_t smem = sizeof(float) * (static_cast<size_t>(ntok) * c + ntok);
  MMPL_REQUIRE(smem <= 48 * 1024, MMPL_E_UNSUPPORTED, "token_stats: ntok*C=%d too large", ntok * c);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * ntok * c, s));
  MMPL_CUDA(cudaMemsetAsync(counts, 0, sizeof(float) * ntok, s));
  const int64_t units = static_cast<int64_t>(n) * d * h * w * (c / vn);
  const int blocks = static_cast<int>(std::min<int64_t>((units + 255) / 256, static_cast<int64_t>(num_sms()) * 4));
  MMPL_DISPATCH_DTYPE(dtype, T, (token_stats_kernel<T><<<blocks, 256, smem, s>>>(static_cast<const T*>(x), mask, mask_is_u8, sums,
                                                                               counts, n, d, h, w, c, dm, hm, wm, ntok)));
  MMPL_CHECK_LAUNCH("token_stats");
  return MMPL_OK;
}

extern "C" int mmpl_token_ema(float* token, const float* sums, const float* counts, int ntok, int c, float alpha,
                              mmpl_stream_t stream) {
  MMPL_REQUIRE(ntok >= 1 && c >= 1, MMPL_E_SHAPE, "token_ema: ntok=%d c=%d", ntok, c);
  token_ema_kernel<<<(ntok * c + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(token, sums, counts, ntok, c, alpha);
  MMPL_CHECK_LAUNCH("token_ema");
  return MMPL_OK;
}
