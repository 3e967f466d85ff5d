// Training head in two launches: classifier (precls_conv.2 = nn.Conv3d(base, classes, 1) with bias, unet3D.py:632, :713)
// fused with the partial-label loss (EDiceLoss_partial.forward, loss_functions/loss_partial.py:71-99) it feeds.
//
// The unfused chain writes fp32 logits (64 B/voxel), reads them twice (loss forward, loss backward), writes dlogits
// (64 B/voxel) and reads them back in the classifier backward: 4 launches, ~390 B/voxel.  Here the logits of a voxel only
// ever exist in the accumulator registers of a warp MMA:
//   forward : a (bf16, CIN channels) + label  ->  per-class sums (I, Z, Y, E)  ->  loss          2 CIN + L  B/voxel
//   backward: a + label -> logits (recomputed) -> softmax -> dz  ->  dA = dz W (bf16), dW += dz^T a, db += dz,
//             + the first pass of the GroupNorm+ReLU backward that produced `a`               4 CIN + L  B/voxel
// Work split: an m16n8k16 accumulator holds classes (2t, 2t+1, 8+2t, 8+2t+1) of voxel rows g and g+8 in thread
// (g, t) -- the 16 classes of a voxel live on the 4 threads of a quad, so the softmax costs 4 quad shuffles and every
// other per-(voxel, class) operation runs on all 32 lanes.  The backward re-uses the accumulator registers as the
// A operand of the dA product (row = voxel, k = class: the same fragment layout) and transposes them across the warp with
// movmatrix for the dW product (row = class, k = voxel).
// Precision: identical to the unfused kernels -- fp32 operands (W, dz) enter the MMAs as bf16 hi + lo pairs.
#include <algorithm>

#include "common.cuh"
#include "loss_math.cuh"
#include "mma.cuh"

namespace mmpl {
namespace {

constexpr int CL_THREADS = 256;
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr float kLog2e = 1.4426950408889634f;

// class of accumulator slot k (0..3) of a thread with quad index t
__device__ __forceinline__ int slot_class(int k, int t) { return (k >> 1) * 8 + 2 * t + (k & 1); }

// softmax over the 16 classes of one voxel row spread over a quad (4 slots per thread); invalid classes hold -inf
__device__ __forceinline__ void quad_softmax(float (&p)[4]) {
  float mx = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3]));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  const float mxs = mx * kLog2e;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    p[k] = ex2_approx(fmaf(p[k], kLog2e, -mxs));
    sum += p[k];
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k] *= inv;
}

template <bool TU8>
__device__ __forceinline__ float load_label(const void* target, int64_t idx) {
  if (TU8) return static_cast<float>(__ldg(static_cast<const uint8_t*>(target) + idx));
  return __ldg(static_cast<const float*>(target) + idx);
}

// B fragments of the logits product: W^T as (K = channel) x (N = class), hi / lo
template <int KS>
__device__ __forceinline__ void load_w_frags(const float* wc, int cin, int classes, int g, int t, uint32_t (&bh)[KS][2][2],
                                             uint32_t (&bl)[KS][2][2]) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = nt * 8 + g, k = ks * 16 + h * 8 + 2 * t;
        const float w0 = c < classes ? wc[c * cin + k] : 0.f, w1 = c < classes ? wc[c * cin + k + 1] : 0.f;
        split2(w0, w1, bh[ks][nt][h], bl[ks][nt][h]);
      }
}

// ------------------------------------------------------------------------------------------------ forward
// grid (blocks per sample, N).  A warp owns 32-voxel chunks of its sample; A fragments straight from global memory.
template <int CIN, bool TU8>
__global__ void __launch_bounds__(CL_THREADS, 2)
cls_loss_fwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ bias,
                    const void* __restrict__ target, const float* __restrict__ cw, const float* __restrict__ lut,
                    double* __restrict__ sums, float* __restrict__ loss, unsigned int* __restrict__ ticket, int N, int64_t S,
                    int C, int uce, int per_sample) {
  constexpr int KS = CIN / 16;
  __shared__ float s_lut[16];
  __shared__ float s_part[CL_WARPS][64];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t n = blockIdx.y;
  const int grp = per_sample ? static_cast<int>(n) : 0;
  const float* cw_g = cw + static_cast<int64_t>(grp) * C;
  const float* lut_g = lut ? lut + static_cast<int64_t>(grp) * C : nullptr;
  if (threadIdx.x < 16) s_lut[threadIdx.x] = (lut_g && threadIdx.x < C) ? lut_g[threadIdx.x] : static_cast<float>(threadIdx.x);
  uint32_t bh[KS][2][2], bl[KS][2][2];
  load_w_frags<KS>(wc, CIN, C, g, t, bh, bl);
  float bia[4];
  int cls[4];
  unsigned int ce_bits = 0;       // slots whose BCE term is needed (weight != 0)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    cls[k] = slot_class(k, t);
    bia[k] = cls[k] < C ? bias[cls[k]] : -INFINITY;       // classes beyond C never win the max and add exp(-inf) = 0
    if (uce && cls[k] < C && cw_g[cls[k]] != 0.f) ce_bits |= 1u << k;
  }
  __syncthreads();
  float aI[4], aZ[4], aY[4], aE[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) aI[k] = aZ[k] = aY[k] = aE[k] = 0.f;
  const __nv_bfloat16* an = a + (n * S) * CIN;
  const int64_t cps = (S + 31) / 32;          // 32-voxel chunks of this sample
  for (int64_t chunk = static_cast<int64_t>(blockIdx.x) * CL_WARPS + warp; chunk < cps;
       chunk += static_cast<int64_t>(gridDim.x) * CL_WARPS) {
    const int64_t s0 = chunk * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int64_t r0 = s0 + mt * 16 + g, r1 = r0 + 8;
      const bool ok0 = r0 < S, ok1 = r1 < S;
      uint32_t af[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int k = ks * 16 + 2 * t;
        af[ks][0] = ok0 ? *reinterpret_cast<const uint32_t*>(an + r0 * CIN + k) : 0u;
        af[ks][1] = ok1 ? *reinterpret_cast<const uint32_t*>(an + r1 * CIN + k) : 0u;
        af[ks][2] = ok0 ? *reinterpret_cast<const uint32_t*>(an + r0 * CIN + k + 8) : 0u;
        af[ks][3] = ok1 ? *reinterpret_cast<const uint32_t*>(an + r1 * CIN + k + 8) : 0u;
      }
      const float tv0 = ok0 ? load_label<TU8>(target, n * S + r0) : -1.f;
      const float tv1 = ok1 ? load_label<TU8>(target, n * S + r1) : -1.f;
      float acc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        acc[nt][0] = acc[nt][2] = bia[nt * 2], acc[nt][1] = acc[nt][3] = bia[nt * 2 + 1];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma_bf16(acc[nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bh[ks][nt][0], bh[ks][nt][1]);
          mma_bf16(acc[nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bl[ks][nt][0], bl[ks][nt][1]);
        }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = acc[k >> 1][half * 2 + (k & 1)];
        quad_softmax(p);
        const bool ok = half ? ok1 : ok0;
        const int tc = class_of(half ? tv1 : tv0, lut_g ? s_lut : nullptr, C);
        if (ok) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool hit = cls[k] == tc;
            aZ[k] = fmaf(p[k], p[k], aZ[k]);
            aI[k] += hit ? p[k] : 0.f;
            aY[k] += hit ? 1.f : 0.f;
            if (ce_bits & (1u << k)) {
              // nn.BCELoss semantics: log() of the fp32 probability, clamped at -100
              const float l = hit ? logf(p[k]) : logf(1.0f - p[k]);
              aE[k] -= fmaxf(l, -100.f);
            }
          }
        }
      }
    }
  }
  // rows live on the lanes that share t: reduce over g, then over the block's warps, then one fp64 atomic per sum
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      aI[k] += __shfl_xor_sync(0xffffffffu, aI[k], o);
      aZ[k] += __shfl_xor_sync(0xffffffffu, aZ[k], o);
      aY[k] += __shfl_xor_sync(0xffffffffu, aY[k], o);
      aE[k] += __shfl_xor_sync(0xffffffffu, aE[k], o);
    }
    if (g == 0) {
      s_part[warp][cls[k]] = aI[k];
      s_part[warp][16 + cls[k]] = aZ[k];
      s_part[warp][32 + cls[k]] = aY[k];
      s_part[warp][48 + cls[k]] = aE[k];
    }
  }
  __syncthreads();
  double* sums_g = sums + static_cast<int64_t>(grp) * 4 * C;
  if (threadIdx.x < 64) {
    const int q = threadIdx.x >> 4, c = threadIdx.x & 15;
    if (c < C) {
      double v = 0;
      for (int w = 0; w < CL_WARPS; ++w) v += static_cast<double>(s_part[w][threadIdx.x]);
      atomicAdd(&sums_g[q * C + c], v);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    partial_loss_finalize(sums, cw, N, S, C, uce, per_sample, loss);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// grid (blocks per sample, N).  A warp owns 16-voxel tiles; the 16 x CIN activation tile goes through a per-warp
// shared-memory tile (coalesced 16-byte loads) and comes back as A fragments of the logits product (ldmatrix), as the
// transposed B fragments of the dW product (ldmatrix.trans) and -- the same registers -- as the activation values of the
// GroupNorm-backward sums.
template <int CIN, bool GN, bool TU8>
__global__ void __launch_bounds__(CL_THREADS, CIN == 32 ? 2 : 1)
cls_loss_bwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ bias,
                    const void* __restrict__ target, const float* __restrict__ cw, const float* __restrict__ lut,
                    const double* __restrict__ sums, const float* __restrict__ grad_out, __nv_bfloat16* __restrict__ da,
                    float* __restrict__ dwc, float* __restrict__ dbias, const float* __restrict__ gn_beta,
                    double* __restrict__ gn_ws, int N, int64_t S, int C, int uce, int per_sample) {
  constexpr int NT = CIN / 8, KS = CIN / 16;
  constexpr int PITCH = CIN * 2 + 16;                           // bytes per voxel row of the per-warp activation tile
  __shared__ __align__(16) uint8_t s_tile[CL_WARPS][16 * PITCH];
  __shared__ float s_lut[16], s_ca[16], s_cb[16], s_ce[16];
  __shared__ float red[16 * CIN];
  __shared__ float redb[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t n = blockIdx.y;
  const int grp = per_sample ? static_cast<int>(n) : 0;
  const float* lut_g = lut ? lut + static_cast<int64_t>(grp) * C : nullptr;
  if (threadIdx.x < 16) {
    const int c = threadIdx.x;
    s_lut[c] = (lut_g && c < C) ? lut_g[c] : static_cast<float>(c);
    float ca = 0.f, cb = 0.f, ce = 0.f;
    if (c < C)
      partial_loss_coeffs(sums + static_cast<int64_t>(grp) * 4 * C, cw[grp * C + c], *grad_out, N, S, C, c, uce, per_sample,
                          ca, cb, ce);
    s_ca[c] = ca, s_cb[c] = cb, s_ce[c] = ce;
  }
  for (int i = threadIdx.x; i < 16 * CIN; i += CL_THREADS) red[i] = 0.f;
  if (threadIdx.x < 16) redb[threadIdx.x] = 0.f;
  __syncthreads();
  uint32_t bh[KS][2][2], bl[KS][2][2];
  load_w_frags<KS>(wc, CIN, C, g, t, bh, bl);
  // B operand of the dA product: W[c][k] as (K = class) x (N = channel), hi/lo
  uint32_t wh[NT][2], wl[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = h * 8 + 2 * t, k = nt * 8 + g;
      const float w0 = c < C ? wc[c * CIN + k] : 0.f, w1 = (c + 1) < C ? wc[(c + 1) * CIN + k] : 0.f;
      split2(w0, w1, wh[nt][h], wl[nt][h]);
    }
  float bia[4], ca[4], cb[4], ce[4];
  int cls[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    cls[k] = slot_class(k, t);
    bia[k] = cls[k] < C ? bias[cls[k]] : -INFINITY;
    ca[k] = s_ca[cls[k]], cb[k] = s_cb[cls[k]], ce[k] = s_ce[cls[k]];
  }
  uint8_t* stage = s_tile[warp];
  const uint32_t stage_u32 = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
  float dw[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) dw[nt][j] = 0.f;
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
  float gs1[NT][2], gs2[NT][2];        // GN: partial S1 / sum dA*a of channels nt*8 + 2t + j over this thread's rows
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) gs1[nt][0] = gs1[nt][1] = gs2[nt][0] = gs2[nt][1] = 0.f;
  const __nv_bfloat16* an = a + (n * S) * CIN;
  __nv_bfloat16* dan = da + (n * S) * CIN;
  const int64_t tps = (S + 15) / 16;            // 16-voxel tiles of this sample
  for (int64_t tile = static_cast<int64_t>(blockIdx.x) * CL_WARPS + warp; tile < tps;
       tile += static_cast<int64_t>(gridDim.x) * CL_WARPS) {
    const int64_t s0 = tile * 16;
    const int64_t r0 = s0 + g, r1 = r0 + 8;
    const bool ok0 = r0 < S, ok1 = r1 < S;
    const float tv0 = ok0 ? load_label<TU8>(target, n * S + r0) : -1.f;
    const float tv1 = ok1 ? load_label<TU8>(target, n * S + r1) : -1.f;
    // ---- activation tile -> shared memory (rows beyond the sample are zero)
    __syncwarp();
#pragma unroll
    for (int i = 0; i < (16 * CIN * 2) / (32 * 16); ++i) {
      const int u = lane + 32 * i, row = u / (CIN / 8), c16 = u % (CIN / 8);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (s0 + row < S) v = *reinterpret_cast<const uint4*>(an + (s0 + row) * CIN + c16 * 8);
      *reinterpret_cast<uint4*>(stage + row * PITCH + c16 * 16) = v;
    }
    __syncwarp();
    const int mi = lane >> 3, mr = lane & 7;
    // ---- logits: A fragments (a0: rows 0-7 / k 0-7, a1: rows 8-15 / k 0-7, a2: rows 0-7 / k 8-15, a3: rows 8-15 / k 8-15)
    uint32_t af[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldmatrix_x4(af[ks], stage_u32 + ((mi & 1) * 8 + mr) * PITCH + (ks * 16 + (mi >> 1) * 8) * 2);
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[nt][0] = acc[nt][2] = bia[nt * 2], acc[nt][1] = acc[nt][3] = bia[nt * 2 + 1];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        mma_bf16(acc[nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bh[ks][nt][0], bh[ks][nt][1]);
        mma_bf16(acc[nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bl[ks][nt][0], bl[ks][nt][1]);
      }
    // ---- dz = p (g - sum_k g_k p_k) for rows r0 (half 0) and r1 (half 1)
    float dz[2][4];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float p[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) p[k] = acc[k >> 1][half * 2 + (k & 1)];
      quad_softmax(p);
      const int tc = class_of(half ? tv1 : tv0, lut_g ? s_lut : nullptr, C);
      float gk[4], dot = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float tt = (cls[k] == tc) ? 1.f : 0.f;
        gk[k] = tt * ca[k] + p[k] * cb[k];
        if (ce[k] != 0.f) gk[k] += ce[k] * (p[k] - tt) / fmaxf(p[k] * (1.0f - p[k]), 1e-12f);
        dot = fmaf(gk[k], p[k], dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      const bool ok = half ? ok1 : ok0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dz[half][k] = ok ? p[k] * (gk[k] - dot) : 0.f;
        dbacc[k] += dz[half][k];
      }
    }
    // the accumulator layout IS the A-fragment layout of a (voxel x class) operand:
    //   e0 = (r0, classes 2t,2t+1)  e1 = (r1, same)  e2 = (r0, classes 8+2t,+1)  e3 = (r1, same)
    uint32_t ah[4], al[4];
    split2(dz[0][0], dz[0][1], ah[0], al[0]);
    split2(dz[1][0], dz[1][1], ah[1], al[1]);
    split2(dz[0][2], dz[0][3], ah[2], al[2]);
    split2(dz[1][2], dz[1][3], ah[3], al[3]);
    // ---- dA[v][k] = sum_c dz[v][c] W[c][k]
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      mma_bf16(o, ah[0], ah[1], ah[2], ah[3], wh[nt][0], wh[nt][1]);
      mma_bf16(o, al[0], al[1], al[2], al[3], wh[nt][0], wh[nt][1]);
      mma_bf16(o, ah[0], ah[1], ah[2], ah[3], wl[nt][0], wl[nt][1]);
      if (ok0) *reinterpret_cast<uint32_t*>(dan + r0 * CIN + nt * 8 + 2 * t) = pack_hi(o[0], o[1]);
      if (ok1) *reinterpret_cast<uint32_t*>(dan + r1 * CIN + nt * 8 + 2 * t) = pack_hi(o[2], o[3]);
      if (GN) {
        // (row g, channels nt*8 + 2t, +1) and (row g+8, same) are A-fragment registers of the logits product
        const uint32_t u0 = af[nt >> 1][(nt & 1) * 2], u1 = af[nt >> 1][(nt & 1) * 2 + 1];
        const float a00 = __uint_as_float(u0 << 16), a01 = __uint_as_float(u0 & 0xFFFF0000u);
        const float a10 = __uint_as_float(u1 << 16), a11 = __uint_as_float(u1 & 0xFFFF0000u);
        gs2[nt][0] = fmaf(o[0], a00, fmaf(o[2], a10, gs2[nt][0]));
        gs2[nt][1] = fmaf(o[1], a01, fmaf(o[3], a11, gs2[nt][1]));
        if (a00 > 0.f) gs1[nt][0] += o[0];
        if (a10 > 0.f) gs1[nt][0] += o[2];
        if (a01 > 0.f) gs1[nt][1] += o[1];
        if (a11 > 0.f) gs1[nt][1] += o[3];
      }
    }
    // ---- dW[c][k] += sum_v dz[v][c] a[v][k]: A = dz^T (row = class, k = voxel) by 8x8 transposes of the fragments
    //   (voxels 0-7 x classes 0-7)^T -> a0, (voxels 0-7 x classes 8-15)^T -> a1, (8-15 x 0-7)^T -> a2, (8-15 x 8-15)^T -> a3
    uint32_t th[4], tl[4];
    th[0] = movmatrix_trans(ah[0]), th[1] = movmatrix_trans(ah[2]), th[2] = movmatrix_trans(ah[1]), th[3] = movmatrix_trans(ah[3]);
    tl[0] = movmatrix_trans(al[0]), tl[1] = movmatrix_trans(al[2]), tl[2] = movmatrix_trans(al[1]), tl[3] = movmatrix_trans(al[3]);
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
      // matrices (v 0..7 | 8..15) x (n-tile nt | nt+1), transposed on the way out: b0, b1 of two n-tiles
      uint32_t b[4];
      ldmatrix_x4_trans(b, stage_u32 + ((mi & 1) * 8 + mr) * PITCH + (nt + (mi >> 1)) * 16);
      mma_bf16(dw[nt], th[0], th[1], th[2], th[3], b[0], b[1]);
      mma_bf16(dw[nt], tl[0], tl[1], tl[2], tl[3], b[0], b[1]);
      mma_bf16(dw[nt + 1], th[0], th[1], th[2], th[3], b[2], b[3]);
      mma_bf16(dw[nt + 1], tl[0], tl[1], tl[2], tl[3], b[2], b[3]);
    }
  }
  if (GN) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float s1 = gs1[nt][j], s2 = gs2[nt][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {      // rows live on the lanes that share t
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (g == 0) {
          const int c = nt * 8 + 2 * t + j;
          double* w = gn_ws + (n * CIN + c) * 6;
          atomicAdd(w, static_cast<double>(s1));
          atomicAdd(w + 1, static_cast<double>(s2) - static_cast<double>(gn_beta[c]) * static_cast<double>(s1));
        }
      }
  }
  // ---- block reduction of dW (fragment: rows c = g, g+8; cols k = nt*8 + 2t, +1) and db (slots of t, rows over g)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&red[(g + (j >> 1) * 8) * CIN + nt * 8 + 2 * t + (j & 1)], dw[nt][j]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v = dbacc[k];
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (g == 0) atomicAdd(&redb[cls[k]], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * CIN; i += CL_THREADS)
    if (i / CIN < C) atomicAdd(&dwc[i], red[i]);
  if (threadIdx.x < C) atomicAdd(&dbias[threadIdx.x], redb[threadIdx.x]);
}

int check_common(const void* a, const void* target, int n, int64_t spatial, int cin, int classes, const char* what) {
  MMPL_REQUIRE(cin == 32 || cin == 64, MMPL_E_UNSUPPORTED, "%s: %d input channels (32 or 64)", what, cin);
  MMPL_REQUIRE(classes >= 1 && classes <= 16, MMPL_E_UNSUPPORTED, "%s: %d classes (1..16)", what, classes);
  MMPL_REQUIRE(n > 0 && spatial > 0, MMPL_E_SHAPE, "%s: empty input", what);
  MMPL_REQUIRE(n <= 65535, MMPL_E_SHAPE, "%s: batch %d exceeds the grid limit", what, n);
  MMPL_REQUIRE(reinterpret_cast<uintptr_t>(a) % 16 == 0, MMPL_E_ALIGN, "%s: activations must be 16-byte aligned", what);
  MMPL_REQUIRE(a != nullptr && target != nullptr, MMPL_E_SHAPE, "%s: null input", what);
  return MMPL_OK;
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

// a [n][spatial][cin] bf16 (NDHWC); target [n][spatial] fp32 or uint8 class ids; class_weight / lut [G][classes] with
// G = n (per_sample) or 1; sums = G*4*classes + 1 doubles (written here, read by the backward); loss = one float.
extern "C" int mmpl_cls_loss_fwd(const void* a, const float* wc, const float* bias, const void* target, int target_is_u8,
                                 const float* class_weight, const float* lut, int per_sample, double* sums, float* loss,
                                 int n, int64_t spatial, int cin, int classes, int uce, mmpl_stream_t stream) {
  if (int rc = check_common(a, target, n, spatial, cin, classes, "cls_loss_fwd")) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = per_sample ? n : 1;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + static_cast<int64_t>(groups) * 4 * classes);
  MMPL_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (static_cast<int64_t>(groups) * 4 * classes + 1), s));
  const int64_t chunks = (spatial + 31) / 32;
  const dim3 grid(static_cast<unsigned>(std::min<int64_t>((chunks + CL_WARPS - 1) / CL_WARPS, std::max(1, num_sms() * 2 / n))), n);
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
#define MMPL_CL_FWD(CIN, TU8) \
  cls_loss_fwd_kernel<CIN, TU8><<<grid, CL_THREADS, 0, s>>>(ap, wc, bias, target, class_weight, lut, sums, loss, ticket, n, \
                                                            spatial, classes, uce, per_sample)
  if (cin == 32) {
    if (target_is_u8) MMPL_CL_FWD(32, true); else MMPL_CL_FWD(32, false);
  } else {
    if (target_is_u8) MMPL_CL_FWD(64, true); else MMPL_CL_FWD(64, false);
  }
#undef MMPL_CL_FWD
  MMPL_CHECK_LAUNCH("cls_loss_fwd");
  return MMPL_OK;
}

// da [n][spatial][cin] bf16, dwc [classes][cin], dbias [classes] fp32: all written (dwc / dbias are zeroed here first);
// gn_beta [cin] / gn_ws [n][cin][6] (both or neither): the fused first pass of the GroupNorm+ReLU backward of `a`.
extern "C" int mmpl_cls_loss_bwd(const void* a, const float* wc, const float* bias, const void* target, int target_is_u8,
                                 const float* class_weight, const float* lut, int per_sample, const double* sums,
                                 const float* grad_out, void* da, float* dwc, float* dbias, const float* gn_beta,
                                 double* gn_ws, int n, int64_t spatial, int cin, int classes, int uce,
                                 mmpl_stream_t stream) {
  if (int rc = check_common(a, target, n, spatial, cin, classes, "cls_loss_bwd")) return rc;
  MMPL_REQUIRE((gn_beta == nullptr) == (gn_ws == nullptr), MMPL_E_SHAPE, "cls_loss_bwd: gn_beta and gn_ws go together");
  MMPL_REQUIRE(reinterpret_cast<uintptr_t>(da) % 16 == 0, MMPL_E_ALIGN, "cls_loss_bwd: da must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_CUDA(cudaMemsetAsync(dwc, 0, sizeof(float) * classes * cin, s));
  MMPL_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * classes, s));
  const int64_t tiles = (spatial + 15) / 16;
  const int per_sm = cin == 32 ? 2 : 1;
  const dim3 grid(static_cast<unsigned>(std::min<int64_t>((tiles + CL_WARPS - 1) / CL_WARPS, std::max(1, num_sms() * per_sm / n))), n);
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
  __nv_bfloat16* dap = static_cast<__nv_bfloat16*>(da);
#define MMPL_CL_BWD(CIN, GN, TU8) \
  cls_loss_bwd_kernel<CIN, GN, TU8><<<grid, CL_THREADS, 0, s>>>(ap, wc, bias, target, class_weight, lut, sums, grad_out, dap, dwc, \
                                                                dbias, gn_beta, gn_ws, n, spatial, classes, uce, per_sample)
#define MMPL_CL_BWD2(CIN, GN) do { if (target_is_u8) MMPL_CL_BWD(CIN, GN, true); else MMPL_CL_BWD(CIN, GN, false); } while (0)
  const bool gn = gn_ws != nullptr;
  if (cin == 32) {
    if (gn) MMPL_CL_BWD2(32, true); else MMPL_CL_BWD2(32, false);
  } else {
    if (gn) MMPL_CL_BWD2(64, true); else MMPL_CL_BWD2(64, false);
  }
#undef MMPL_CL_BWD2
#undef MMPL_CL_BWD
  MMPL_CHECK_LAUNCH("cls_loss_bwd");
  return MMPL_OK;
}
