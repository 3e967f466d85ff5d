// bf16 classifier kernels on warp-level tensor-core MMAs (mma.sync m16n8k16, fp32 accumulate).
// Reference op: precls_conv.2 = nn.Conv3d(base, classes, 1) with bias and its autograd (unet3D.py:632, :713).
//
// Why not tcgen05: the contraction lengths are 32/64 (forward, dA) or 16 (classes), the GEMMs are skinny and the
// kernels are bound by HBM traffic and instruction issue, not by tensor throughput -- the CUDA-core versions in
// small_conv.cu spend ~1000 FMAs per voxel and are issue-bound at 2-3x the HBM time.  Warp MMAs cut the instruction
// count ~15x with no shared-memory staging: every fragment is loaded straight from global memory in the layout the
// instruction wants.
//
// Precision: activations are bf16 already; fp32 operands (W, dlogits) are split into hi + lo bf16 parts and
// multiplied in 2-3 MMAs, so the products carry 16 mantissa bits (logits within 1e-5 of the fp32 computation).
#include <algorithm>

#include "common.cuh"
#include "mma.cuh"

namespace mmpl {
namespace {

// ------------------------------------------------------------------------------------------------ forward
// logits[n][c][s] = bias[c] + sum_k a[n][s][k] W[c][k].  A warp owns chunks of 32 voxels of one sample (2 m16 tiles);
// A fragments come from `a` (row = voxel), B fragments (W^T, hi/lo) live in registers for the whole kernel.
template <int CIN>
__global__ void __launch_bounds__(256)
cls_fwd_mma_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ bias,
                   float* __restrict__ logits, int N, int64_t S, int classes) {
  constexpr int KS = CIN / 16;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint32_t bh[KS][2][2], bl[KS][2][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = nt * 8 + g, k = ks * 16 + h * 8 + 2 * t;
        const float w0 = c < classes ? wc[c * CIN + k] : 0.f, w1 = c < classes ? wc[c * CIN + k + 1] : 0.f;
        split2(w0, w1, bh[ks][nt][h], bl[ks][nt][h]);
      }
  float bia[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) bia[nt][j] = (nt * 8 + 2 * t + j) < classes ? bias[nt * 8 + 2 * t + j] : 0.f;
  const int cps = static_cast<int>((S + 31) / 32);          // 32-voxel chunks per sample
  const int nchunks = N * cps;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warps) {
    const int n = chunk / cps;
    const int64_t s0 = static_cast<int64_t>(chunk - n * cps) * 32;
    const __nv_bfloat16* an = a + (static_cast<int64_t>(n) * S) * CIN;
    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        acc[mt][nt][0] = acc[mt][nt][2] = bia[nt][0], acc[mt][nt][1] = acc[mt][nt][3] = bia[nt][1];
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
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma_bf16(acc[mt][nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bh[ks][nt][0], bh[ks][nt][1]);
          mma_bf16(acc[mt][nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bl[ks][nt][0], bl[ks][nt][1]);
        }
    }
    float* ln = logits + static_cast<int64_t>(n) * classes * S;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t r = s0 + mt * 16 + g + (j >> 1) * 8;
          const int c = nt * 8 + 2 * t + (j & 1);
          if (r < S && c < classes) ln[static_cast<int64_t>(c) * S + r] = acc[mt][nt][j];
        }
  }
}

// ------------------------------------------------------------------------------------------------ forward + blend
// Sliding-window inference (predict_sliding, evaluate_amos.py:244-276): the classifier of a tile is immediately
// weighted with the Gaussian importance map and accumulated into the volume accumulator -- acc[c][voxel] += g * logit --
// instead of writing a 151 MB fp32 logits tile that a blend kernel reads back.  Same MMA mainloop as cls_fwd; the
// epilogue is a read-modify-write of the fp32 accumulator (each element is touched by exactly one thread per launch and
// launches of one volume are stream-ordered, so no atomics).  The tile origin lives in DEVICE memory so that one captured
// CUDA graph serves every tile of the volume.
struct BlendGeo {
  float* acc;            // [C][D][H][W], or [D][C][H][W] when d_outer
  float* wsum;           // [D][H][W] or NULL (argmax / Dice do not need the normaliser)
  const float* gauss;    // [td][th][tw]
  const int* origin;     // device: d0, h0, w0
  int D, H, W, td, th, tw, d_outer;
};

template <int CIN>
__global__ void __launch_bounds__(256)
cls_blend_mma_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ bias,
                     BlendGeo q, int classes) {
  constexpr int KS = CIN / 16;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint32_t bh[KS][2][2], bl[KS][2][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = nt * 8 + g, k = ks * 16 + h * 8 + 2 * t;
        const float w0 = c < classes ? wc[c * CIN + k] : 0.f, w1 = c < classes ? wc[c * CIN + k + 1] : 0.f;
        split2(w0, w1, bh[ks][nt][h], bl[ks][nt][h]);
      }
  float bia[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) bia[nt][j] = (nt * 8 + 2 * t + j) < classes ? bias[nt * 8 + 2 * t + j] : 0.f;
  const int d0 = __ldg(q.origin), h0 = __ldg(q.origin + 1), w0 = __ldg(q.origin + 2);
  const int S = q.td * q.th * q.tw;
  const int64_t plane = static_cast<int64_t>(q.H) * q.W;
  __shared__ float tile[8][16][36];     // per warp: [class][voxel of the chunk], pitch 36 keeps the fragment stores conflict-free
  const int nchunks = (S + 31) / 32;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int chunk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warps) {
    const int s0 = chunk * 32;
    float acc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        acc[mt][nt][0] = acc[mt][nt][2] = bia[nt][0], acc[mt][nt][1] = acc[mt][nt][3] = bia[nt][1];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r0 = s0 + mt * 16 + g, r1 = r0 + 8;
      const bool ok0 = r0 < S, ok1 = r1 < S;
      uint32_t af[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int k = ks * 16 + 2 * t;
        af[ks][0] = ok0 ? *reinterpret_cast<const uint32_t*>(a + static_cast<int64_t>(r0) * CIN + k) : 0u;
        af[ks][1] = ok1 ? *reinterpret_cast<const uint32_t*>(a + static_cast<int64_t>(r1) * CIN + k) : 0u;
        af[ks][2] = ok0 ? *reinterpret_cast<const uint32_t*>(a + static_cast<int64_t>(r0) * CIN + k + 8) : 0u;
        af[ks][3] = ok1 ? *reinterpret_cast<const uint32_t*>(a + static_cast<int64_t>(r1) * CIN + k + 8) : 0u;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          mma_bf16(acc[mt][nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bh[ks][nt][0], bh[ks][nt][1]);
          mma_bf16(acc[mt][nt], af[ks][0], af[ks][1], af[ks][2], af[ks][3], bl[ks][nt][0], bl[ks][nt][1]);
        }
    }
    // Epilogue: transpose the 32 x 16 logits of the chunk through shared memory so that lane L owns voxel s0 + L and
    // every read-modify-write instruction of the warp covers 32 consecutive voxels of ONE class plane (full 128-byte
    // lines; the accumulator-fragment layout would touch four class planes with 32 bytes each).
    float(*tw_)[36] = tile[threadIdx.x >> 5];
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          tw_[nt * 8 + 2 * t + (i & 1)][mt * 16 + g + (i >> 1) * 8] = acc[mt][nt][i];
    __syncwarp();
    const int r = s0 + lane;
    if (r < S) {
      const int x = r % q.tw, yz = r / q.tw, y = yz % q.th, z = yz / q.th;
      const int vd = d0 + z, vh = h0 + y, vw = w0 + x;
      if (!(vd >= q.D || vh >= q.H || vw >= q.W || vd < 0 || vh < 0 || vw < 0)) {      // memory safety only
        const float gv = __ldg(q.gauss + r);
        const int64_t o = static_cast<int64_t>(vh) * q.W + vw;
        if (q.wsum != nullptr) q.wsum[vd * plane + o] += gv;
        float* dst = q.acc + (q.d_outer ? static_cast<int64_t>(vd) * classes * plane + o : vd * plane + o);
        const int64_t cstride = q.d_outer ? plane : static_cast<int64_t>(q.D) * plane;
        // all loads of the voxel's class column first, then the stores: 16 independent 128-byte lines in flight per warp
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < classes) v[c] = dst[c * cstride];
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < classes) dst[c * cstride] = v[c] + gv * tw_[c][lane];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward
// One pass over dlogits and a per 16-voxel tile (one tile per warp iteration):
//   dA[v][k]  = sum_c dl[c][v] W[c][k]       M = voxels, N = CIN, K = 16 classes   (dl hi/lo x W hi/lo, 3 MMAs)
//   dW[c][k] += sum_v dl[c][v] a[v][k]       M = classes, N = CIN, K = 16 voxels   (dl hi/lo, 2 MMAs)
//   db[c]    += sum_v dl[c][v]
// dW / db stay in registers for the whole kernel, are summed over the block's warps in shared memory and leave with
// one fp32 atomic per element per block.
// GN = true additionally folds in the first pass of the backward of precls_conv.0/1 (GroupNorm+ReLU, unet3D.py:629-631),
// whose output is `a`: S1_c = sum_v dA*[a > 0], Q_c = sum_v dA*a - beta_c*S1_c (see mmpl_gn_bwd_fuse) -> gn_ws[N][CIN][6].
template <int CIN, bool GN>
__global__ void __launch_bounds__(256)
cls_bwd_mma_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ wc, const float* __restrict__ dl,
                   __nv_bfloat16* __restrict__ da, float* __restrict__ dwc, float* __restrict__ dbias,
                   const float* __restrict__ gn_beta, double* __restrict__ gn_ws, int N, int64_t S, int classes) {
  constexpr int NT = CIN / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // B operand of the dA product: W[c][k] as (K = class) x (N = channel), hi/lo
  uint32_t wh[NT][2], wl[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = h * 8 + 2 * t, k = nt * 8 + g;
      const float w0 = c < classes ? wc[c * CIN + k] : 0.f, w1 = (c + 1) < classes ? wc[(c + 1) * CIN + k] : 0.f;
      split2(w0, w1, wh[nt][h], wl[nt][h]);
    }
  constexpr int PITCH = CIN * 2 + 16;                           // bytes per voxel row of the per-warp activation tile
  __shared__ __align__(16) uint8_t s_tile[8][16 * PITCH];
  uint8_t* stage = s_tile[warp];
  const uint32_t stage_u32 = static_cast<uint32_t>(__cvta_generic_to_shared(stage));
  float dw[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) dw[nt][j] = 0.f;
  float db0 = 0.f, db1 = 0.f;          // classes g and g + 8
  float gs1[NT][2], gs2[NT][2];        // GN: partial S1 / sum dA*a of channels nt*8 + 2t + j over this thread's rows
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) gs1[nt][0] = gs1[nt][1] = gs2[nt][0] = gs2[nt][1] = 0.f;
  int gn_n = -1;
  auto gn_flush = [&]() {
    if (!GN || gn_n < 0) return;
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
          double* w = gn_ws + (static_cast<int64_t>(gn_n) * CIN + c) * 6;
          atomicAdd(w, static_cast<double>(s1));
          atomicAdd(w + 1, static_cast<double>(s2) - static_cast<double>(gn_beta[c]) * static_cast<double>(s1));
        }
        gs1[nt][j] = gs2[nt][j] = 0.f;
      }
  };
  const int tps = static_cast<int>((S + 15) / 16);            // 16-voxel tiles per sample
  const int ntiles = N * tps;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += warps) {
    const int n = tile / tps;
    if (GN && n != gn_n) {
      gn_flush();
      gn_n = n;
    }
    const int64_t s0 = static_cast<int64_t>(tile - n * tps) * 16;
    const float* dln = dl + static_cast<int64_t>(n) * classes * S;
    const __nv_bfloat16* an = a + (static_cast<int64_t>(n) * S) * CIN;
    __nv_bfloat16* dan = da + (static_cast<int64_t>(n) * S) * CIN;
    // ---- dA: A[row v][col c] = dl[c][v]
    {
      const int64_t r0 = s0 + g, r1 = r0 + 8;
      const bool ok0 = r0 < S, ok1 = r1 < S;
      float e[4][2];     // [fragment register][class 2t+j (+8)]
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c0 = 2 * t + j, c1 = c0 + 8;
        e[0][j] = (ok0 && c0 < classes) ? dln[static_cast<int64_t>(c0) * S + r0] : 0.f;
        e[1][j] = (ok1 && c0 < classes) ? dln[static_cast<int64_t>(c0) * S + r1] : 0.f;
        e[2][j] = (ok0 && c1 < classes) ? dln[static_cast<int64_t>(c1) * S + r0] : 0.f;
        e[3][j] = (ok1 && c1 < classes) ? dln[static_cast<int64_t>(c1) * S + r1] : 0.f;
      }
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split2(e[i][0], e[i][1], ah[i], al[i]);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16(o, ah[0], ah[1], ah[2], ah[3], wh[nt][0], wh[nt][1]);
        mma_bf16(o, al[0], al[1], al[2], al[3], wh[nt][0], wh[nt][1]);
        mma_bf16(o, ah[0], ah[1], ah[2], ah[3], wl[nt][0], wl[nt][1]);
        if (ok0) *reinterpret_cast<uint32_t*>(dan + r0 * CIN + nt * 8 + 2 * t) = pack_hi(o[0], o[1]);
        if (ok1) *reinterpret_cast<uint32_t*>(dan + r1 * CIN + nt * 8 + 2 * t) = pack_hi(o[2], o[3]);
        if (GN) {
          const uint32_t u0 = ok0 ? *reinterpret_cast<const uint32_t*>(an + r0 * CIN + nt * 8 + 2 * t) : 0u;
          const uint32_t u1 = ok1 ? *reinterpret_cast<const uint32_t*>(an + r1 * CIN + nt * 8 + 2 * t) : 0u;
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
    }
    // ---- dW / db: A[row c][col v] = dl[c][v], B[row v][col k] = a[v][k]
    {
      const int64_t v0 = s0 + 2 * t;            // fragment columns: voxels v0, v0+1, v0+8, v0+9
      const bool okc0 = g < classes, okc1 = (g + 8) < classes;
      float e[4][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t va = v0 + j, vb = v0 + 8 + j;
        e[0][j] = (okc0 && va < S) ? dln[static_cast<int64_t>(g) * S + va] : 0.f;
        e[1][j] = (okc1 && va < S) ? dln[static_cast<int64_t>(g + 8) * S + va] : 0.f;
        e[2][j] = (okc0 && vb < S) ? dln[static_cast<int64_t>(g) * S + vb] : 0.f;
        e[3][j] = (okc1 && vb < S) ? dln[static_cast<int64_t>(g + 8) * S + vb] : 0.f;
      }
      db0 += (e[0][0] + e[0][1]) + (e[2][0] + e[2][1]);
      db1 += (e[1][0] + e[1][1]) + (e[3][0] + e[3][1]);
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split2(e[i][0], e[i][1], ah[i], al[i]);
      // B fragments: the 16 x CIN activation tile goes through a per-warp shared-memory tile (two coalesced 16-byte
      // loads per lane, row pitch CIN*2+16 bytes -> conflict-free) and comes back transposed with ldmatrix.x4.trans:
      // matrices (v 0..7 | 8..15) x (n-tile nt | nt+1) are exactly b0, b1 of two n-tiles.
      __syncwarp();
#pragma unroll
      for (int i = 0; i < (16 * CIN * 2) / (32 * 16); ++i) {
        const int u = lane + 32 * i, row = u / (CIN / 8), c16 = u % (CIN / 8);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (s0 + row < S) v = *reinterpret_cast<const uint4*>(an + (s0 + row) * CIN + c16 * 8);
        *reinterpret_cast<uint4*>(stage + row * PITCH + c16 * 16) = v;
      }
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < NT; nt += 2) {
        const int mi = lane >> 3, r = lane & 7;
        const uint32_t addr = stage_u32 + ((mi & 1) * 8 + r) * PITCH + (nt + (mi >> 1)) * 16;
        uint32_t b[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
                     : "r"(addr));
        mma_bf16(dw[nt], ah[0], ah[1], ah[2], ah[3], b[0], b[1]);
        mma_bf16(dw[nt], al[0], al[1], al[2], al[3], b[0], b[1]);
        mma_bf16(dw[nt + 1], ah[0], ah[1], ah[2], ah[3], b[2], b[3]);
        mma_bf16(dw[nt + 1], al[0], al[1], al[2], al[3], b[2], b[3]);
      }
    }
  }
  gn_flush();
  // ---- block reduction of dW (fragment: rows c = g, g+8; cols k = nt*8 + 2t, +1) and db
  __shared__ float red[16 * CIN];
  __shared__ float redb[16];
  for (int i = threadIdx.x; i < 16 * CIN; i += 256) red[i] = 0.f;
  if (threadIdx.x < 16) redb[threadIdx.x] = 0.f;
  __syncthreads();
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&red[(g + (j >> 1) * 8) * CIN + nt * 8 + 2 * t + (j & 1)], dw[nt][j]);
  db0 += __shfl_xor_sync(0xffffffffu, db0, 1);
  db0 += __shfl_xor_sync(0xffffffffu, db0, 2);
  db1 += __shfl_xor_sync(0xffffffffu, db1, 1);
  db1 += __shfl_xor_sync(0xffffffffu, db1, 2);
  if (t == 0) {
    atomicAdd(&redb[g], db0);
    atomicAdd(&redb[g + 8], db1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * CIN; i += 256)
    if (i / CIN < classes) atomicAdd(&dwc[i], red[i]);
  if (threadIdx.x < classes) atomicAdd(&dbias[threadIdx.x], redb[threadIdx.x]);
}

}  // namespace

int cls_fwd_mma(const void* a, const float* wc, const float* bias, float* logits, int n, int64_t spatial, int cin,
                int classes, cudaStream_t s) {
  const int64_t chunks = static_cast<int64_t>(n) * ((spatial + 31) / 32);
  MMPL_REQUIRE(chunks < (1ll << 31), MMPL_E_SHAPE, "cls_fwd: too many voxels");
  const int blocks = static_cast<int>(std::min<int64_t>((chunks + 7) / 8, static_cast<int64_t>(num_sms()) * 8));
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
  if (cin == 32)
    cls_fwd_mma_kernel<32><<<blocks, 256, 0, s>>>(ap, wc, bias, logits, n, spatial, classes);
  else
    cls_fwd_mma_kernel<64><<<blocks, 256, 0, s>>>(ap, wc, bias, logits, n, spatial, classes);
  return MMPL_OK;
}

}  // namespace mmpl

extern "C" int mmpl_cls_blend(const void* a, const float* wc, const float* bias, const float* gauss, float* acc,
                              float* wsum, const int* origin_dev, int classes, int d, int h, int w, int td, int th,
                              int tw, int cin, int d_outer, mmpl_stream_t stream) {
  using namespace mmpl;
  MMPL_REQUIRE(cin == 32 || cin == 64, MMPL_E_UNSUPPORTED, "cls_blend: cin=%d (32 or 64)", cin);
  MMPL_REQUIRE(classes >= 1 && classes <= 16, MMPL_E_SHAPE, "cls_blend: classes=%d (1..16)", classes);
  MMPL_REQUIRE(td > 0 && th > 0 && tw > 0 && td <= d && th <= h && tw <= w, MMPL_E_SHAPE,
               "cls_blend: tile (%d,%d,%d) vs volume (%d,%d,%d)", td, th, tw, d, h, w);
  MMPL_REQUIRE(a && wc && bias && gauss && acc && origin_dev, MMPL_E_SHAPE, "cls_blend: null argument");
  const int64_t S = static_cast<int64_t>(td) * th * tw;
  MMPL_REQUIRE(S < (1ll << 31), MMPL_E_SHAPE, "cls_blend: tile too large");
  BlendGeo q{acc, wsum, gauss, origin_dev, d, h, w, td, th, tw, d_outer};
  const int64_t chunks = (S + 31) / 32;
  const int blocks = static_cast<int>(std::min<int64_t>((chunks + 7) / 8, static_cast<int64_t>(num_sms()) * 8));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
  if (cin == 32)
    cls_blend_mma_kernel<32><<<blocks, 256, 0, s>>>(ap, wc, bias, q, classes);
  else
    cls_blend_mma_kernel<64><<<blocks, 256, 0, s>>>(ap, wc, bias, q, classes);
  MMPL_CHECK_LAUNCH("cls_blend");
  return MMPL_OK;
}

namespace mmpl {

// dwc / dbias must be zero on entry
int cls_bwd_mma(const void* a, const float* wc, const float* dlogits, void* da, float* dwc, float* dbias,
                const float* gn_beta, double* gn_ws, int n, int64_t spatial, int cin, int classes, cudaStream_t s) {
  const int64_t tiles = static_cast<int64_t>(n) * ((spatial + 15) / 16);
  MMPL_REQUIRE(tiles < (1ll << 31), MMPL_E_SHAPE, "cls_bwd: too many voxels");
  const int blocks = static_cast<int>(std::min<int64_t>((tiles + 7) / 8, static_cast<int64_t>(num_sms()) * 4));
  const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
  __nv_bfloat16* dap = static_cast<__nv_bfloat16*>(da);
  if (cin == 32) {
    if (gn_ws)
      cls_bwd_mma_kernel<32, true><<<blocks, 256, 0, s>>>(ap, wc, dlogits, dap, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
    else
      cls_bwd_mma_kernel<32, false><<<blocks, 256, 0, s>>>(ap, wc, dlogits, dap, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
  } else {
    if (gn_ws)
      cls_bwd_mma_kernel<64, true><<<blocks, 256, 0, s>>>(ap, wc, dlogits, dap, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
    else
      cls_bwd_mma_kernel<64, false><<<blocks, 256, 0, s>>>(ap, wc, dlogits, dap, dwc, dbias, gn_beta, gn_ws, n, spatial, classes);
  }
  return MMPL_OK;
}

}  // namespace mmpl
