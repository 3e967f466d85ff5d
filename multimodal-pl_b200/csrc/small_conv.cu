// The two convolutions at the ends of the network that are bandwidth-bound, not tensor-core work:
//  * stem   : conv3x3x3(1 -> base) on the fp32 image (unet3D.py:594, :666); K = 27, output-write bound.
//  * cls    : nn.Conv3d(base, classes, 1) with bias (unet3D.py:632, :713); reads NDHWC activations, writes the
//             NCDHW fp32 logits the reference returns, so no layout transpose is ever materialised.
#include "common.cuh"

namespace mmpl {
namespace {

// ------------------------------------------------------------------------------------------------ stem forward
template <typename T, int COUT>
__global__ void __launch_bounds__(128)
stem_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w_hat, T* __restrict__ y, int N, int D, int H,
                int W) {
  __shared__ __align__(16) float sw[27][COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i % 27][i / 27] = w_hat[i];  // w_hat is [COUT][27]
  __syncthreads();
  const int64_t total = static_cast<int64_t>(N) * D * H * W;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < total;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(v % W);
    int64_t r = v / W;
    const int yy = static_cast<int>(r % H);
    r /= H;
    const int z = static_cast<int>(r % D);
    const int n = static_cast<int>(r / D);
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
    const float* base = img + static_cast<int64_t>(n) * D * H * W;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int zz = z + kd - 1;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int y2 = yy + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int x2 = x + kw - 1;
          float xv = 0.f;
          if (zz >= 0 && zz < D && y2 >= 0 && y2 < H && x2 >= 0 && x2 < W)
            xv = base[(static_cast<int64_t>(zz) * H + y2) * W + x2];
          const int t = (kd * 3 + kh) * 3 + kw;
#pragma unroll
          for (int c = 0; c < COUT; c += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(&sw[t][c]);   // warp-broadcast LDS.128
            acc[c] = fmaf(xv, w4.x, acc[c]);
            acc[c + 1] = fmaf(xv, w4.y, acc[c + 1]);
            acc[c + 2] = fmaf(xv, w4.z, acc[c + 2]);
            acc[c + 3] = fmaf(xv, w4.w, acc[c + 3]);
          }
        }
      }
    }
    constexpr int VN = Vec<T>::N;
    T* out = y + v * COUT;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += VN) {
      Vec<T> o;
#pragma unroll
      for (int k = 0; k < VN; ++k) o.v[k] = acc[c0 + k];
      o.store(out + c0);
    }
  }
}

// ------------------------------------------------------------------------------------------------ stem wgrad
// dW[t][co] = sum_v x[v + t - 1] * dy[v][co].  The image is first copied into a zero-padded buffer so the 27 window
// loads need no bounds checks; block = 16 voxel lanes x (COUT/2) channel pairs, each thread keeps 27 x 2 accumulators.
__global__ void __launch_bounds__(256)
pad_image_kernel(const float* __restrict__ img, float* __restrict__ pad, int N, int D, int H, int W) {
  const int64_t total = static_cast<int64_t>(N) * (D + 2) * (H + 2) * (W + 2);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % (W + 2)) - 1;
    int64_t r = i / (W + 2);
    const int y = static_cast<int>(r % (H + 2)) - 1;
    r /= (H + 2);
    const int z = static_cast<int>(r % (D + 2)) - 1;
    const int n = static_cast<int>(r / (D + 2));
    float v = 0.f;
    if (x >= 0 && x < W && y >= 0 && y < H && z >= 0 && z < D)
      v = img[((static_cast<int64_t>(n) * D + z) * H + y) * W + x];
    pad[i] = v;
  }
}

template <typename T>
__device__ __forceinline__ void load2(const T* p, float& a, float& b);
template <>
__device__ __forceinline__ void load2<float>(const float* p, float& a, float& b) {
  const float2 t = *reinterpret_cast<const float2*>(p);
  a = t.x, b = t.y;
}
template <>
__device__ __forceinline__ void load2<__nv_bfloat16>(const __nv_bfloat16* p, float& a, float& b) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  a = __uint_as_float(u << 16), b = __uint_as_float(u & 0xFFFF0000u);
}

template <typename T, int COUT>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ pad, const T* __restrict__ dy, float* __restrict__ dw_tapmajor, int N, int D,
                  int H, int W) {
  constexpr int CP = COUT / 2;       // channel pairs
  constexpr int LANES = 256 / CP;    // voxel lanes per block
  const int cp = threadIdx.x % CP, vl = threadIdx.x / CP;
  float acc[27][2];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t][0] = acc[t][1] = 0.f;
  const int64_t total = static_cast<int64_t>(N) * D * H * W;
  const int64_t prow = W + 2, pplane = static_cast<int64_t>(H + 2) * (W + 2);
  for (int64_t v = blockIdx.x * static_cast<int64_t>(LANES) + vl; v < total;
       v += static_cast<int64_t>(gridDim.x) * LANES) {
    const int x = static_cast<int>(v % W);
    int64_t r = v / W;
    const int y = static_cast<int>(r % H);
    r /= H;
    const int z = static_cast<int>(r % D);
    const int n = static_cast<int>(r / D);
    float g0, g1;
    load2<T>(dy + v * COUT + 2 * cp, g0, g1);
    const float* win = pad + (static_cast<int64_t>(n) * (D + 2) + z) * pplane + y * prow + x;  // window origin
#pragma unroll
    for (int kd = 0; kd < 3; ++kd)
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* row = win + kd * pplane + kh * prow;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float xv = row[kw];
          acc[(kd * 3 + kh) * 3 + kw][0] = fmaf(xv, g0, acc[(kd * 3 + kh) * 3 + kw][0]);
          acc[(kd * 3 + kh) * 3 + kw][1] = fmaf(xv, g1, acc[(kd * 3 + kh) * 3 + kw][1]);
        }
      }
  }
  // reduce the voxel lanes through shared memory, nine taps at a time (keeps static shared memory under 48 KB)
  __shared__ float red[LANES][9][COUT + 1];
#pragma unroll
  for (int t0 = 0; t0 < 27; t0 += 9) {
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 9; ++t) red[vl][t][2 * cp] = acc[t0 + t][0], red[vl][t][2 * cp + 1] = acc[t0 + t][1];
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * COUT; i += 256) {
      const int t = i / COUT, c = i % COUT;
      float s = 0.f;
      for (int l = 0; l < LANES; ++l) s += red[l][t][c];
      atomicAdd(&dw_tapmajor[(t0 + t) * COUT + c], s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ stem im2col
// X27[v][t] = image[v + tap(t) - 1] (zero padded), t = 0..26, stored bf16 NDHWC.  With it the Cin = 1 stem becomes a
// 1x1x1 convolution that runs on the tcgen05 kernels (forward AND weight gradient) at HBM speed instead of 864
// CUDA-core FMAs per voxel.  SPLIT = 0: 32 channels (27 taps + 5 zeros).  SPLIT = 1: 64 channels, the fp32 value is
// carried as hi + lo bf16 parts (channels t and 32 + t; 16 mantissa bits), so only the weights are rounded to bf16 as
// in every other layer.  One thread per voxel.
// Grid (ceil(W/256), H, N*D): rows are block-uniform, all index arithmetic is 32-bit.
template <int SPLIT>
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int N, int D, int H, int W) {
  constexpr int CH = SPLIT ? 64 : 32;
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y, z = blockIdx.z % D, n = blockIdx.z / D;
  const float* base = img + static_cast<int64_t>(n) * D * H * W;
  float t[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) t[i] = 0.f;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int zz = z + kd - 1;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = y + kh - 1;
      const bool row_ok = zz >= 0 && zz < D && yy >= 0 && yy < H;
      const float* row = base + (static_cast<int64_t>(row_ok ? zz : 0) * H + (row_ok ? yy : 0)) * W;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int xx = x + kw - 1;
        t[(kd * 3 + kh) * 3 + kw] = (row_ok && xx >= 0 && xx < W) ? row[xx] : 0.f;
      }
    }
  }
  __nv_bfloat16* dst = out + (((static_cast<int64_t>(n) * D + z) * H + y) * W + x) * CH;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) {
    Vec<__nv_bfloat16> o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = t[c0 + k];
    o.store(dst + c0);
    if (SPLIT) {
      Vec<__nv_bfloat16> l;
#pragma unroll
      for (int k = 0; k < 8; ++k) l.v[k] = t[c0 + k] - __bfloat162float(__float2bfloat16(t[c0 + k]));
      l.store(dst + 32 + c0);
    }
  }
}

// ------------------------------------------------------------------------------------------------ classifier fwd
template <typename T, int CIN>
__global__ void __launch_bounds__(256)
cls_fwd_kernel(const T* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ bias,
               float* __restrict__ logits, int N, int64_t S, int classes) {
  __shared__ __align__(16) float sw[CIN][16];   // transposed: [input channel][class]
  __shared__ float sb[16];
  for (int i = threadIdx.x; i < 16 * CIN; i += blockDim.x) sw[i % CIN][i / CIN] = (i / CIN) < classes ? wc[i] : 0.f;
  if (threadIdx.x < 16) sb[threadIdx.x] = threadIdx.x < classes ? bias[threadIdx.x] : 0.f;
  __syncthreads();
  constexpr int VN = Vec<T>::N;
  constexpr int VT = 2;   // voxels per thread: every weight LDS.128 feeds 8 FMAs (no grid-stride loop: keeps registers low)
  const int64_t total = static_cast<int64_t>(N) * S;
  const int64_t vb = blockIdx.x * static_cast<int64_t>(blockDim.x) * VT + threadIdx.x;
  float acc[VT][16];
#pragma unroll
  for (int u = 0; u < VT; ++u)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[u][c] = sb[c];
#pragma unroll
  for (int k0 = 0; k0 < CIN; k0 += VN) {
    Vec<T> x[VT];
#pragma unroll
    for (int u = 0; u < VT; ++u) {
      const int64_t v = vb + u * static_cast<int64_t>(blockDim.x);
      if (v < total) {
        x[u].load(a + v * CIN + k0);
      } else {
#pragma unroll
        for (int k = 0; k < VN; ++k) x[u].v[k] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < VN; ++k)
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&sw[k0 + k][c]);
#pragma unroll
        for (int u = 0; u < VT; ++u) {
          acc[u][c] = fmaf(x[u].v[k], w4.x, acc[u][c]);
          acc[u][c + 1] = fmaf(x[u].v[k], w4.y, acc[u][c + 1]);
          acc[u][c + 2] = fmaf(x[u].v[k], w4.z, acc[u][c + 2]);
          acc[u][c + 3] = fmaf(x[u].v[k], w4.w, acc[u][c + 3]);
        }
      }
  }
#pragma unroll
  for (int u = 0; u < VT; ++u) {
    const int64_t v = vb + u * static_cast<int64_t>(blockDim.x);
    if (v >= total) continue;
    const int64_t n = v / S, s = v - n * S;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < classes) logits[(n * classes + c) * S + s] = acc[u][c];
  }
}

// ------------------------------------------------------------------------------------------------ classifier bwd
// One pass over dlogits and a: da[v][k] = sum_c dl[c][v] W[c][k];  dW[c][k] = sum_v dl[c][v] a[v][k];  db[c] = sum_v dl.
// 1024 FMAs per voxel (CIN = 32) next to 192 bytes of HBM traffic, so the kernel is laid out to be FMA-issue bound, not
// shared-memory bound: tiles of 128 voxels are staged in shared memory (dl transposed to [voxel][class] with a 16-byte
// swizzle, a as fp32); phase 1 keeps W in registers (thread = 4 input channels x 16 classes, 1 LDS.128 per 16 FMAs),
// phase 2 register-tiles dW 4(c) x 4(k) per thread (1 LDS.128 per 8 FMAs).  The next tile's global loads are issued
// before the current tile's arithmetic (register prefetch).
template <typename T>
struct Raw16 {   // 16 raw bytes of T (8 bf16 or 4 fp32), converted only when stored to shared memory
  uint4 u;
  __device__ __forceinline__ void load(const T* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { u = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void to_smem(float* dst) const;
};
template <>
__device__ __forceinline__ void Raw16<float>::to_smem(float* dst) const {
  *reinterpret_cast<uint4*>(dst) = u;
}
template <>
__device__ __forceinline__ void Raw16<__nv_bfloat16>::to_smem(float* dst) const {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  float f[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) f[2 * i] = __uint_as_float(w[i] << 16), f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

template <typename T, int CIN, bool GN>
__global__ void __launch_bounds__(256, 2)
cls_bwd_kernel(const T* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ dl, T* __restrict__ da,
               float* __restrict__ dwc, float* __restrict__ dbias, const float* __restrict__ gn_beta,
               double* __restrict__ gn_ws, int N, int64_t S, int classes) {
  constexpr int TV = 128;               // voxels per tile
  constexpr int KB = CIN / 4;           // blocks of 4 input channels
  constexpr int VPP = 256 / KB;         // phase 1: voxels per pass over the block
  constexpr int SLOTS = 4 * KB;         // phase 2: (class block, channel block) pairs
  constexpr int VG = 256 / SLOTS;       // phase 2: voxel groups
  constexpr int VN = Vec<T>::N;
  constexpr int AV = TV * CIN / VN / 256;  // 16-byte loads of `a` per thread per tile
  __shared__ __align__(16) float s_g[TV][16];
  __shared__ __align__(16) float s_a[TV][CIN];
  const int tid = threadIdx.x;
  const int kb1 = tid % KB, vs1 = tid / KB;
  const int slot = tid % SLOTS, vg = tid / SLOTS, cb2 = slot / KB, kb2 = slot % KB;
  const int sj = tid % TV, sc = (tid / TV) * 8;
  float wreg[16][4];
#pragma unroll
  for (int c = 0; c < 16; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) wreg[c][i] = c < classes ? wc[c * CIN + kb1 * 4 + i] : 0.f;
  float accw[4][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) accw[i][j] = 0.f;
  // tiles never straddle samples (tile -> (n, t)): the fused GroupNorm-backward sums below are per sample
  const int tps = static_cast<int>((S + TV - 1) / TV);   // tiles per sample (32-bit tile arithmetic: no 64-bit divisions)
  const int ntiles = N * tps;
  // Fused first pass of the backward of precls_conv.0/1 = GroupNorm+ReLU (unet3D.py:629-631): `a` is its output, so
  // S1_c = sum_v da*[a > 0] and Q_c = sum_v da*a - beta_c*S1_c (see mmpl_gn_bwd_fuse) accumulate per thread in phase 1.
  constexpr bool gn_on = GN;
  float gs1[4] = {0.f, 0.f, 0.f, 0.f}, gs2[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t gn_n = -1;
  auto gn_flush = [&]() {
    if (!gn_on || gn_n < 0) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s1 = gs1[i], s2 = gs2[i];
      for (int o = 16; o >= KB; o >>= 1) {          // lanes KB apart own the same 4 channels
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if ((tid & 31) < KB) {
        const int c = kb1 * 4 + i;
        double* w = gn_ws + (gn_n * CIN + c) * 6;
        atomicAdd(w, static_cast<double>(s1));
        atomicAdd(w + 1, static_cast<double>(s2) - static_cast<double>(gn_beta[c]) * static_cast<double>(s1));
      }
      gs1[i] = gs2[i] = 0.f;
    }
  };

  float pg[8];
  Raw16<T> pa[AV];
  auto prefetch = [&](int tile) {
    const int tn_ = tile / tps;
    const int64_t n = tn_, sp0 = static_cast<int64_t>(tile - tn_ * tps) * TV;
    const int64_t sp = sp0 + sj;
    const bool ok = sp < S;
#pragma unroll
    for (int i = 0; i < 8; ++i) pg[i] = (ok && sc + i < classes) ? dl[(n * classes + sc + i) * S + sp] : 0.f;
#pragma unroll
    for (int r = 0; r < AV; ++r) {
      const int idx = tid + 256 * r;                 // 16-byte unit inside the tile
      const int j = idx / (CIN / VN);
      if (sp0 + j < S)
        pa[r].load(a + (n * S + sp0) * CIN + static_cast<int64_t>(idx) * VN);
      else
        pa[r].zero();
    }
  };

  int tile = blockIdx.x;
  if (tile < ntiles) prefetch(tile);
  for (; tile < ntiles; tile += gridDim.x) {
    const int tni = tile / tps;
    const int64_t tn = tni, sp0 = static_cast<int64_t>(tile - tni * tps) * TV;
    const int64_t v0 = tn * S + sp0;
    const int64_t total = (tn + 1) * S;              // end of this sample
    if (gn_on && tn != gn_n) {
      gn_flush();
      gn_n = tn;
    }
    __syncthreads();
    {
      const int sw = (sj >> 1) & 3;
      *reinterpret_cast<float4*>(&s_g[sj][((sc / 4) ^ sw) * 4]) = make_float4(pg[0], pg[1], pg[2], pg[3]);
      *reinterpret_cast<float4*>(&s_g[sj][((sc / 4 + 1) ^ sw) * 4]) = make_float4(pg[4], pg[5], pg[6], pg[7]);
#pragma unroll
      for (int r = 0; r < AV; ++r) pa[r].to_smem(&s_a[0][0] + static_cast<int64_t>(tid + 256 * r) * VN);
    }
    __syncthreads();
    if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x);
    // phase 1: da
#pragma unroll 2
    for (int p = 0; p < TV / VPP; ++p) {
      const int j = p * VPP + vs1;
      const int sw = (j >> 1) & 3;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const float4 g4 = *reinterpret_cast<const float4*>(&s_g[j][(cb ^ sw) * 4]);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = fmaf(g[c], wreg[cb * 4 + c][i], o[i]);
      }
      if (v0 + j < total) {
        store4(da + (v0 + j) * CIN + kb1 * 4, o);
        if (gn_on) {
          const float4 a4 = *reinterpret_cast<const float4*>(&s_a[j][kb1 * 4]);
          const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float e = to_f32<T>(from_f32<T>(o[i]));     // da as stored
            gs2[i] = fmaf(e, av[i], gs2[i]);
            gs1[i] += av[i] > 0.f ? e : 0.f;
          }
        }
      }
    }
    // phase 2: dW / dbias
#pragma unroll 4
    for (int j = vg; j < TV; j += VG) {
      const float4 x4 = *reinterpret_cast<const float4*>(&s_a[j][kb2 * 4]);
      const float4 g4 = *reinterpret_cast<const float4*>(&s_g[j][(cb2 ^ ((j >> 1) & 3)) * 4]);
      const float g[4] = {g4.x, g4.y, g4.z, g4.w};
      const float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) accw[i][k] = fmaf(g[i], x[k], accw[i][k]);
      if (kb2 == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) accb[i] += g[i];
      }
    }
  }
  gn_flush();
  // cross-group reduction through shared memory (reuse the staging tiles), then one atomic per (c,k) per block
  __syncthreads();
  float* red = &s_a[0][0];    // VG x 16 classes x CIN floats <= TV*CIN
  float* redb = &s_g[0][0];   // VG x 16 classes
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) red[(vg * 16 + cb2 * 4 + i) * CIN + kb2 * 4 + k] = accw[i][k];
  if (kb2 == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) redb[vg * 16 + cb2 * 4 + i] = accb[i];
  }
  __syncthreads();
  for (int i = tid; i < 16 * CIN; i += 256) {
    const int c = i / CIN, k = i % CIN;
    if (c >= classes) continue;
    float sum = 0.f;
    for (int w = 0; w < VG; ++w) sum += red[(w * 16 + c) * CIN + k];
    atomicAdd(&dwc[c * CIN + k], sum);
  }
  if (tid < classes) {
    float sum = 0.f;
    for (int w = 0; w < VG; ++w) sum += redb[w * 16 + tid];
    atomicAdd(&dbias[tid], sum);
  }
}

template <typename T>
void launch_cls_bwd(int cin, int blocks, cudaStream_t s, const T* a, const float* wc, const float* dl, T* da, float* dwc,
                    float* dbias, const float* gn_beta, double* gn_ws, int n, int64_t spatial, int classes) {
  if (cin == 32) {
    if (gn_ws)
      cls_bwd_kernel<T, 32, true><<<blocks, 256, 0, s>>>(a, wc, dl, da, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
    else
      cls_bwd_kernel<T, 32, false><<<blocks, 256, 0, s>>>(a, wc, dl, da, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
  } else {
    if (gn_ws)
      cls_bwd_kernel<T, 64, true><<<blocks, 256, 0, s>>>(a, wc, dl, da, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
    else
      cls_bwd_kernel<T, 64, false><<<blocks, 256, 0, s>>>(a, wc, dl, da, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
  }
}

}  // namespace
}  // namespace mmpl

namespace mmpl {
int cls_fwd_mma(const void*, const float*, const float*, float*, int, int64_t, int, int, cudaStream_t);
int cls_fwd_generic(const void*, const float*, const float*, float*, int, int64_t, int, int, int, cudaStream_t);
int cls_bwd_generic(const void*, const float*, const float*, void*, float*, float*, int, int64_t, int, int, int, cudaStream_t);
int cls_bwd_mma(const void*, const float*, const float*, void*, float*, float*, const float*, double*, int, int64_t, int,
                int, cudaStream_t);
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_stem_conv_fwd(const float* image, const float* w_hat, void* y, int n, int d, int h, int w, int cout,
                                  int dtype, mmpl_stream_t stream) {
  MMPL_REQUIRE(cout == 32 || cout == 64, MMPL_E_SHAPE, "stem: cout=%d (32 or 64)", cout);
  const int64_t total = static_cast<int64_t>(n) * d * h * w;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 127) / 128, static_cast<int64_t>(num_sms()) * 16));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, {
    if (cout == 32)
      stem_fwd_kernel<T, 32><<<blocks, 128, 0, s>>>(image, w_hat, static_cast<T*>(y), n, d, h, w);
    else
      stem_fwd_kernel<T, 64><<<blocks, 128, 0, s>>>(image, w_hat, static_cast<T*>(y), n, d, h, w);
  });
  MMPL_CHECK_LAUNCH("stem_conv_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_stem_im2col(const float* image, void* x27, int n, int d, int h, int w, int channels,
                                mmpl_stream_t stream) {
  MMPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "stem_im2col: empty image");
  MMPL_REQUIRE(channels == 32 || channels == 64, MMPL_E_SHAPE, "stem_im2col: channels=%d (32, or 64 = hi/lo split)",
               channels);
  MMPL_REQUIRE(h <= 65535 && static_cast<int64_t>(n) * d <= 65535, MMPL_E_SHAPE, "stem_im2col: H=%d, N*D=%lld exceed the grid",
               h, static_cast<long long>(n) * d);
  const dim3 blocks((w + 255) / 256, h, n * d);
  if (channels == 64)
    stem_im2col_kernel<1><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(image, static_cast<__nv_bfloat16*>(x27),
                                                                               n, d, h, w);
  else
    stem_im2col_kernel<0><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(image, static_cast<__nv_bfloat16*>(x27),
                                                                               n, d, h, w);
  MMPL_CHECK_LAUNCH("stem_im2col");
  return MMPL_OK;
}

extern "C" size_t mmpl_stem_conv_wgrad_workspace(int n, int d, int h, int w) {
  return sizeof(float) * static_cast<size_t>(n) * (d + 2) * (h + 2) * (w + 2);
}

extern "C" int mmpl_stem_conv_wgrad(const float* image, const void* dy, float* dw_tapmajor, int n, int d, int h, int w,
                                    int cout, int dtype, void* workspace, size_t workspace_bytes,
                                    mmpl_stream_t stream) {
  MMPL_REQUIRE(cout == 32 || cout == 64, MMPL_E_SHAPE, "stem: cout=%d (32 or 64)", cout);
  MMPL_REQUIRE(workspace != nullptr && workspace_bytes >= mmpl_stem_conv_wgrad_workspace(n, d, h, w), MMPL_E_SHAPE,
               "stem_conv_wgrad: workspace of %zu bytes required", mmpl_stem_conv_wgrad_workspace(n, d, h, w));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* pad = static_cast<float*>(workspace);
  const int64_t ptotal = static_cast<int64_t>(n) * (d + 2) * (h + 2) * (w + 2);
  pad_image_kernel<<<static_cast<int>(std::min<int64_t>((ptotal + 255) / 256, static_cast<int64_t>(num_sms()) * 8)), 256, 0, s>>>(
      image, pad, n, d, h, w);
  MMPL_CHECK_LAUNCH("pad_image");
  MMPL_CUDA(cudaMemsetAsync(dw_tapmajor, 0, sizeof(float) * 27 * cout, s));
  const int blocks = num_sms() * 2;
  MMPL_DISPATCH_DTYPE(dtype, T, {
    if (cout == 32)
      stem_wgrad_kernel<T, 32><<<blocks, 256, 0, s>>>(pad, static_cast<const T*>(dy), dw_tapmajor, n, d, h, w);
    else
      stem_wgrad_kernel<T, 64><<<blocks, 256, 0, s>>>(pad, static_cast<const T*>(dy), dw_tapmajor, n, d, h, w);
  });
  MMPL_CHECK_LAUNCH("stem_conv_wgrad");
  return MMPL_OK;
}

extern "C" int mmpl_cls_fwd(const void* a, const float* wc, const float* bias, float* logits, int n, int64_t spatial,
                            int cin, int classes, int dtype, mmpl_stream_t stream) {
  MMPL_REQUIRE(cin >= 8 && cin % 8 == 0 && classes >= 1 && classes <= 16, MMPL_E_SHAPE,
               "cls: cin=%d classes=%d (cin a multiple of 8, classes<=16)", cin, classes);
  const int64_t total = static_cast<int64_t>(n) * spatial;
  const int blocks = static_cast<int>((total + 511) / 512);   // 256 threads x 2 voxels
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cin != 32 && cin != 64) {   // widths outside the tuned kernels (128 at 1/8 resolution): generic kernel, eam.cu
    if (int e = cls_fwd_generic(a, wc, bias, logits, n, spatial, cin, classes, dtype, s)) return e;
    MMPL_CHECK_LAUNCH("cls_fwd");
    return MMPL_OK;
  }
  if (dtype == MMPL_BF16) {   // warp-MMA kernel (cls_mma.cu); the CUDA-core kernel below is the fp32 exact path
    if (int e = cls_fwd_mma(a, wc, bias, logits, n, spatial, cin, classes, s)) return e;
    MMPL_CHECK_LAUNCH("cls_fwd");
    return MMPL_OK;
  }
  MMPL_DISPATCH_DTYPE(dtype, T, {
    if (cin == 32)
      cls_fwd_kernel<T, 32><<<blocks, 256, 0, s>>>(static_cast<const T*>(a), wc, bias, logits, n, spatial, classes);
    else
      cls_fwd_kernel<T, 64><<<blocks, 256, 0, s>>>(static_cast<const T*>(a), wc, bias, logits, n, spatial, classes);
  });
  MMPL_CHECK_LAUNCH("cls_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_cls_bwd(const void* a, const float* wc, const float* dlogits, void* da, float* dwc, float* dbias,
                            const float* gn_beta, double* gn_ws, int n, int64_t spatial, int cin, int classes, int dtype,
                            mmpl_stream_t stream) {
  MMPL_REQUIRE((gn_beta == nullptr) == (gn_ws == nullptr), MMPL_E_SHAPE, "cls_bwd: gn_beta and gn_ws go together");
  MMPL_REQUIRE(cin >= 8 && cin % 8 == 0 && classes >= 1 && classes <= 16, MMPL_E_SHAPE,
               "cls: cin=%d classes=%d (cin a multiple of 8, classes<=16)", cin, classes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_CUDA(cudaMemsetAsync(dwc, 0, sizeof(float) * classes * cin, s));
  MMPL_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * classes, s));
  if (cin != 32 && cin != 64) {
    MMPL_REQUIRE(gn_ws == nullptr, MMPL_E_UNSUPPORTED, "cls_bwd: the fused GroupNorm-backward reduction needs cin 32 or 64");
    if (int e = cls_bwd_generic(a, wc, dlogits, da, dwc, dbias, n, spatial, cin, classes, dtype, s)) return e;
    MMPL_CHECK_LAUNCH("cls_bwd");
    return MMPL_OK;
  }
  if (dtype == MMPL_BF16) {   // warp-MMA kernel (cls_mma.cu)
    if (int e = cls_bwd_mma(a, wc, dlogits, da, dwc, dbias, gn_beta, gn_ws, n, spatial, cin, classes, s)) return e;
    MMPL_CHECK_LAUNCH("cls_bwd");
    return MMPL_OK;
  }
  const int64_t ntiles = static_cast<int64_t>(n) * ((spatial + 127) / 128);
  MMPL_REQUIRE(ntiles < (1ll << 31), MMPL_E_SHAPE, "cls_bwd: too many voxels");
  const int blocks = static_cast<int>(std::min<int64_t>(ntiles, static_cast<int64_t>(num_sms()) * 2));
  MMPL_DISPATCH_DTYPE(dtype, T, (launch_cls_bwd<T>(cin, blocks, s, static_cast<const T*>(a), wc, dlogits,
                                                   static_cast<T*>(da), dwc, dbias, gn_beta, gn_ws, n, spatial, classes)));
  MMPL_CHECK_LAUNCH("cls_bwd");
  return MMPL_OK;
}
