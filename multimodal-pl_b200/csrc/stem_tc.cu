// Stem convolution conv3x3x3(1 -> base) on tcgen05 WITHOUT materialising the 27-tap expansion of the image.
// Reference op: self.conv1 of unet3D_baseline (unet3D.py:594, :666) -- a weight-standardised Conv3d with Cin = 1 -- and its
// weight gradient.
//
// Round 1 expanded the fp32 image once into a bf16 [N,D,H,W,64] tensor (27 shifted copies, hi + lo bf16 parts) and ran
// forward and weight gradient as 64 -> base 1x1x1 tensor-core convolutions: 1.5 GB of HBM traffic for an op whose
// algorithmic traffic is 4 + 64 bytes per voxel (321 MB at cfg2).  Here the K = 64 operand tile is BUILT IN SHARED MEMORY:
//   * a (TD+2) x 18 x 10 fp32 halo of the image is staged in shared memory (zero outside the volume = the padding);
//   * each thread expands one voxel row: 27 neighbours -> hi = bf16(x), lo = bf16(x - hi) -> one 128-byte K-major row
//     [27 hi | 5 zero | 27 lo | 5 zero] written with the 128-byte swizzle the UMMA descriptor expects
//     (16-byte chunk c of row r lands at chunk c ^ (r & 7));
//   * forward:  D[128 voxels x base] = A[128 x 64] * W^T[64 x base], four K = 16 tcgen05.mma per 128-voxel plane, epilogue =
//     tcgen05.ld -> bf16 -> 64-byte row stores + the GroupNorm(16) statistics of the stored output (layer0.0.gn1);
//   * weight gradient:  G[64 x base] += A^T[64 x 128] * dY[128 x base] with both operands MN-major (the same descriptor recipe
//     as wgrad_tc.cu: M-chunks one voxel row apart -- the second chunk is discarded), dY tiles by TMA, split-K partials
//     resident in TMEM and added to dW (tap-major [27][base]) once per CTA; hi and lo rows add to the same tap.
// HBM traffic: forward 4 B in + 2*base B out per voxel, weight gradient 4 + 2*base B in.
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace mmpl {
namespace {

using namespace ptx;

constexpr int ST_TH = 16, ST_TW = 8;             // one 128-row M tile = one d-plane of 16 x 8 voxels
constexpr int ST_PH = ST_TH + 2, ST_PW = ST_TW + 2;
constexpr int ST_ROW_BYTES = 128;                // K = 64 bf16

// ---- expand one voxel row from the staged halo: 27 taps -> hi/lo bf16 -> 8 swizzled 16-byte chunks ------------------
// halo: [planes][ST_PH][ST_PW] fp32, origin one voxel before the tile in every axis.
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void build_row(const float* __restrict__ halo, int plane, int r, uint8_t* __restrict__ tile) {
  const int h = r >> 3, w = r & 7;
  float t[28];
#pragma unroll
  for (int kd = 0; kd < 3; ++kd)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        t[(kd * 3 + kh) * 3 + kw] = halo[((plane + kd) * ST_PH + h + kh) * ST_PW + w + kw];
  t[27] = 0.f;
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 14; ++i) {
    hi[i] = pack_bf16(t[2 * i], t[2 * i + 1]);
    const float h0 = __uint_as_float(hi[i] << 16), h1 = __uint_as_float(hi[i] & 0xFFFF0000u);
    lo[i] = pack_bf16(t[2 * i] - h0, t[2 * i + 1] - h1);
  }
  hi[14] = hi[15] = lo[14] = lo[15] = 0u;
  uint8_t* row = tile + r * ST_ROW_BYTES;
  const int sw = r & 7;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
    *reinterpret_cast<uint4*>(row + (((c + 4) ^ sw) << 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
  }
}

// one halo element (index into [planes][ST_PH][ST_PW]) of the tile at (n, d0, h0, w0); zero outside the volume
__device__ __forceinline__ float halo_fetch(const float* __restrict__ img, int idx, int n, int d0, int h0, int w0, int D,
                                            int H, int W) {
  const int pw = idx % ST_PW, ph = (idx / ST_PW) % ST_PH, pd = idx / (ST_PW * ST_PH);
  const int d = d0 - 1 + pd, h = h0 - 1 + ph, w = w0 - 1 + pw;
  if (d < 0 || d >= D || h < 0 || h >= H || w < 0 || w >= W) return 0.f;
  return __ldg(img + ((static_cast<int64_t>(n) * D + d) * H + h) * W + w);
}

struct StemItem {
  int n, d0, h0, w0;
};
__device__ __forceinline__ StemItem stem_item(int item, int DT, int HT, int WT, int TD) {
  StemItem t;
  t.w0 = (item % WT) * ST_TW;
  item /= WT;
  t.h0 = (item % HT) * ST_TH;
  item /= HT;
  t.d0 = (item % DT) * TD;
  t.n = item / DT;
  return t;
}

// ================================================================================================ forward
constexpr int SF_TD = 4;                                   // planes (= M tiles) per work item
constexpr int SF_THREADS = 256;
constexpr int SF_HALO = (SF_TD + 2) * ST_PH * ST_PW;       // 1080 floats
constexpr int SF_HALO_PER_THREAD = (SF_HALO + SF_THREADS - 1) / SF_THREADS;

template <int NT>
struct StemFwdCfg {
  static constexpr int A_BYTES = SF_TD * 128 * ST_ROW_BYTES;          // 64 KB
  static constexpr int B_BYTES = NT * ST_ROW_BYTES;                   // 4 / 8 KB
  static constexpr int HALO_BYTES = (SF_HALO * 4 + 127) / 128 * 128;
  static constexpr int SMEM_BYTES = A_BYTES + B_BYTES + HALO_BYTES + 1024 /*align*/ + 64 /*barrier, tmem slot*/;
  static constexpr int TMEM_COLS = SF_TD * NT <= 128 ? 128 : 256;
};

// Two CTAs per SM.  The kernel is instruction-issue bound (ncu, profiles/r02_ncu_stem.md: 72 M warp instructions, issue
// slots 31 % busy at 16 resident warps); a third CTA per SM (80 registers, 3 x 74 KB of shared memory) measured SLOWER
// (324 us vs 268 us): the register cap spills the 27-tap expansion.
template <int NT>
__global__ void __launch_bounds__(SF_THREADS, 2)
stem_tc_fwd_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ wpk, __nv_bfloat16* __restrict__ y,
                   double* __restrict__ stats, int N, int D, int H, int W, int DT, int HT, int WT, int total_items) {
  using Cfg = StemFwdCfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic on the shared array (an integer round trip would hide the address space from
  // the compiler and turn every halo / tile access below into a generic LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_tiles = smem;
  uint8_t* b_tile = smem + Cfg::A_BYTES;
  float* halo = reinterpret_cast<float*>(b_tile + Cfg::B_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(halo) + Cfg::HALO_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  // weights [NT][64] bf16 (K-major rows of 128 bytes) -> swizzled B tile, once per CTA
  for (int u = tid; u < NT * 8; u += SF_THREADS) {
    const int row = u >> 3, c = u & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(wpk) + row * ST_ROW_BYTES + c * 16);
    *reinterpret_cast<uint4*>(b_tile + row * ST_ROW_BYTES + ((c ^ (row & 7)) << 4)) = v;
  }
  // halo of the first item
  StemItem cur = stem_item(blockIdx.x < total_items ? blockIdx.x : 0, DT, HT, WT, SF_TD);
  if (blockIdx.x < total_items)
    for (int i = tid; i < SF_HALO; i += SF_THREADS) halo[i] = halo_fetch(img, i, cur.n, cur.d0, cur.h0, cur.w0, D, H, W);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
  const uint64_t desc_fix = make_smem_desc(0, 16, 8 * ST_ROW_BYTES, SWZ_128B, 0);
  const uint32_t a_base = smem_u32(a_tiles), b_base = smem_u32(b_tile);

  // GroupNorm(16) statistics of the stored output: per-thread fp32 partials over this CTA's items, flushed per sample
  constexpr int CPG = NT / 16;
  float gsum[16], gsq[16];
#pragma unroll
  for (int g = 0; g < 16; ++g) gsum[g] = gsq[g] = 0.f;
  int stat_n = -1;
  auto flush_stats = [&]() {
    if (stats == nullptr || stat_n < 0) return;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      const float a = warp_sum(gsum[g]), b = warp_sum(gsq[g]);
      if (lane == 0) {
        atomicAdd(&stats[(static_cast<int64_t>(stat_n) * 16 + g) * 2 + 0], static_cast<double>(a));
        atomicAdd(&stats[(static_cast<int64_t>(stat_n) * 16 + g) * 2 + 1], static_cast<double>(b));
      }
      gsum[g] = gsq[g] = 0.f;
    }
  };

  uint32_t phase = 0;
  for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
    // ---- prefetch the next item's halo into registers: the global-load latency hides under build + MMA + epilogue
    const int nitem = item + gridDim.x;
    const bool has_next = nitem < total_items;
    const StemItem nxt = stem_item(has_next ? nitem : 0, DT, HT, WT, SF_TD);
    float pre[SF_HALO_PER_THREAD];
#pragma unroll
    for (int i = 0; i < SF_HALO_PER_THREAD; ++i) {
      const int idx = tid + i * SF_THREADS;
      pre[i] = (has_next && idx < SF_HALO) ? halo_fetch(img, idx, nxt.n, nxt.d0, nxt.h0, nxt.w0, D, H, W) : 0.f;
    }
    // ---- build the SF_TD operand tiles (2 rows per thread)
#pragma unroll
    for (int i = 0; i < SF_TD * 128 / SF_THREADS; ++i) {
      const int rowid = tid + i * SF_THREADS;
      const int plane = rowid >> 7, r = rowid & 127;
      build_row(halo, plane, r, a_tiles + plane * 128 * ST_ROW_BYTES);
    }
    fence_proxy_async();             // generic-proxy writes -> visible to the tensor core's async-proxy reads
    __syncthreads();
    // ---- MMAs: one thread, SF_TD planes x four K = 16 steps
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int pl = 0; pl < SF_TD; ++pl)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = desc_fix | static_cast<uint64_t>(((a_base + pl * 128 * ST_ROW_BYTES + ks * 32) >> 4) & 0x3FFF);
          const uint64_t bd = desc_fix | static_cast<uint64_t>(((b_base + ks * 32) >> 4) & 0x3FFF);
          umma_f16(tmem_base + pl * NT, ad, bd, idesc, ks != 0 ? 1u : 0u);
        }
      umma_commit(bar);
    }
    // the halo buffer is free (all rows built): install the prefetched one for the next iteration
    if (has_next) {
#pragma unroll
      for (int i = 0; i < SF_HALO_PER_THREAD; ++i) {
        const int idx = tid + i * SF_THREADS;
        if (idx < SF_HALO) halo[idx] = pre[i];
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: warp w reads TMEM lane quarter w & 3 of planes (w >> 2) * 2 + {0, 1}
    if (stats != nullptr && cur.n != stat_n) {
      flush_stats();
      stat_n = cur.n;
    }
    const int q = warp & 3, row = q * 32 + lane, rh = row >> 3, rw = row & 7;
    const int hh = cur.h0 + rh, ww = cur.w0 + rw;
#pragma unroll
    for (int pi = 0; pi < SF_TD / 2; ++pi) {
      const int pl = (warp >> 2) * (SF_TD / 2) + pi;
      const int dd = cur.d0 + pl;
      const bool valid = dd < D && hh < H && ww < W;
      __nv_bfloat16* dst = y + ((((static_cast<int64_t>(cur.n) * D + dd) * H + hh) * W + ww) * NT);
#pragma unroll
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + pl * NT + c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              o[k] = pack_bf16(__uint_as_float(r[v * 8 + 2 * k]), __uint_as_float(r[v * 8 + 2 * k + 1]));
              // statistics of the value as stored (bf16-rounded)
              const float x0 = __uint_as_float(o[k] << 16), x1 = __uint_as_float(o[k] & 0xFFFF0000u);
              const int g0 = (c0 + v * 8 + 2 * k) / CPG, g1 = (c0 + v * 8 + 2 * k + 1) / CPG;
              if (g0 == g1) {                 // compile-time: both channels of the pair belong to one group
                gsum[g0] += x0 + x1;
                gsq[g0] = fmaf(x0, x0, fmaf(x1, x1, gsq[g0]));
              } else {
                gsum[g0] += x0;
                gsq[g0] = fmaf(x0, x0, gsq[g0]);
                gsum[g1] += x1;
                gsq[g1] = fmaf(x1, x1, gsq[g1]);
              }
            }
            *reinterpret_cast<uint4*>(dst + c0 + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();        // TMEM, operand tiles and the halo buffer are reused by the next item
    cur = nxt;
  }
  flush_stats();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ================================================================================================ weight gradient
constexpr int SW_TD = 2;                                   // planes per voxel block
constexpr int SW_THREADS = 192;                            // warp 0: dY TMA, warp 1: MMA, warps 2..5: builders + epilogue
constexpr int SW_HALO = (SW_TD + 2) * ST_PH * ST_PW;       // 720 floats
constexpr int SW_HALO_PER_THREAD = (SW_HALO + 127) / 128;
constexpr int SW_NS = 2;

template <int NCO>
struct StemWgCfg {
  static constexpr int RBY = NCO * 2;
  static constexpr uint32_t SWY = RBY == 128 ? SWZ_128B : SWZ_64B;
  static constexpr int X_BYTES = SW_TD * 128 * ST_ROW_BYTES;            // 32 KB
  static constexpr int X_STAGE = X_BYTES + 1024;                        // slack: the discarded M-chunk over-reads one row
  static constexpr int Y_BYTES = SW_TD * 128 * RBY;
  static constexpr int Y_STAGE = (Y_BYTES + 1023) / 1024 * 1024;
  static constexpr int HALO_BYTES = (SW_HALO * 4 + 127) / 128 * 128;
  static constexpr int SMEM_BYTES = SW_NS * (X_STAGE + Y_STAGE) + HALO_BYTES + 1024 + 256;
};

struct StemWgParams {
  const float* img;
  float* dw;                // tap-major [27][cout], zeroed by the host
  int N, D, H, W, cout;
  int DT, HT, WT, total_blocks, ksplit, n_co;
};

template <int NCO>
__global__ void __launch_bounds__(SW_THREADS, 2)
stem_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ StemWgParams p) {
  using Cfg = StemWgCfg<NCO>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic on the shared array (an integer round trip would hide the address space from
  // the compiler and turn every halo / tile access below into a generic LD / ST instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* x_stage = smem;
  uint8_t* y_stage = smem + SW_NS * Cfg::X_STAGE;
  float* halo = reinterpret_cast<float*>(y_stage + SW_NS * Cfg::Y_STAGE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(halo) + Cfg::HALO_BYTES);
  uint64_t* full = bars;                 // count 1 (TMA expect_tx arrive) + 4 (one elected arrive per builder warp)
  uint64_t* empty = full + SW_NS;
  uint64_t* done = empty + SW_NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SW_NS; ++i) mbar_init(&full[i], 5), mbar_init(&empty[i], 1);
    mbar_init(done, 1);
    fence_barrier_init();
    prefetch_tmap(&tmY);
  }
  // the slack rows behind each X stage are read by the discarded M-chunk: keep them finite (zero)
  for (int i = threadIdx.x; i < SW_NS * 1024 / 16; i += SW_THREADS) {
    const int s = i / 64, o = i % 64;
    *reinterpret_cast<uint4*>(x_stage + s * Cfg::X_STAGE + Cfg::X_BYTES + o * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 1) tmem_alloc<128>(tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int co_i = blockIdx.x % p.n_co, split = blockIdx.x / p.n_co;
  const int co0 = co_i * NCO;

  if (warp == 0) {
    // ---- dY producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int b = split; b < p.total_blocks; b += p.ksplit, ++it) {
        const StemItem t = stem_item(b, p.DT, p.HT, p.WT, SW_TD);
        const uint32_t s = it % SW_NS, ph = (it / SW_NS) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], Cfg::Y_BYTES);
        tma_load_5d(y_stage + s * Cfg::Y_STAGE, &tmY, &full[s], co0, t.w0, t.h0, t.d0, t.n);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: G[64(+64 discarded) x NCO] += A^T * dY, both operands MN-major, K = 16 voxels per MMA
    const uint32_t idesc = make_idesc_bf16(128, NCO, 1, 1);
    const uint32_t xb0 = smem_u32(x_stage), yb0 = smem_u32(y_stage);
    // A: M-chunks (64 k-values = one 128-byte row) one voxel row apart (LBO), 8-voxel w-lines 1024 bytes apart (SBO)
    const uint64_t a_fix = make_smem_desc(0, ST_ROW_BYTES, ST_TW * ST_ROW_BYTES, SWZ_128B, 0);
    const uint64_t b_fix = make_smem_desc(0, 64 * Cfg::RBY, ST_TW * Cfg::RBY, Cfg::SWY, 0);
    uint32_t it = 0;
    for (int b = split; b < p.total_blocks; b += p.ksplit, ++it) {
      const uint32_t s = it % SW_NS, ph = (it / SW_NS) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xs = xb0 + s * Cfg::X_STAGE, ys = yb0 + s * Cfg::Y_STAGE;
#pragma unroll
        for (int pl = 0; pl < SW_TD; ++pl)
#pragma unroll
          for (int hl = 0; hl < ST_TH; hl += 2) {
            const uint32_t xo = xs + ((pl * ST_TH + hl) * ST_TW) * ST_ROW_BYTES;
            const uint32_t yo = ys + ((pl * ST_TH + hl) * ST_TW) * Cfg::RBY;
            umma_f16(tmem_base, a_fix | static_cast<uint64_t>((xo >> 4) & 0x3FFF), b_fix | static_cast<uint64_t>((yo >> 4) & 0x3FFF),
                     idesc, (it | pl | hl) != 0 ? 1u : 0u);
          }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    // ---- builders (warps 2..5, 128 threads): one voxel row per thread and plane
    const int bt = threadIdx.x - 64;                 // 0..127
    auto bar_builders = [] { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    uint32_t it = 0;
    float pre[SW_HALO_PER_THREAD];
    {
      const bool any = split < p.total_blocks;
      const StemItem t0 = stem_item(any ? split : 0, p.DT, p.HT, p.WT, SW_TD);
#pragma unroll
      for (int i = 0; i < SW_HALO_PER_THREAD; ++i) {
        const int idx = bt + i * 128;
        pre[i] = (any && idx < SW_HALO) ? halo_fetch(p.img, idx, t0.n, t0.d0, t0.h0, t0.w0, p.D, p.H, p.W) : 0.f;
      }
    }
    for (int b = split; b < p.total_blocks; b += p.ksplit, ++it) {
      const uint32_t s = it % SW_NS, ph = (it / SW_NS) & 1;
      // install this block's halo (prefetched), then start fetching the next one
#pragma unroll
      for (int i = 0; i < SW_HALO_PER_THREAD; ++i) {
        const int idx = bt + i * 128;
        if (idx < SW_HALO) halo[idx] = pre[i];
      }
      const int nb = b + p.ksplit;
      const bool has_next = nb < p.total_blocks;
      const StemItem nxt = stem_item(has_next ? nb : 0, p.DT, p.HT, p.WT, SW_TD);
#pragma unroll
      for (int i = 0; i < SW_HALO_PER_THREAD; ++i) {
        const int idx = bt + i * 128;
        pre[i] = (has_next && idx < SW_HALO) ? halo_fetch(p.img, idx, nxt.n, nxt.d0, nxt.h0, nxt.w0, p.D, p.H, p.W) : 0.f;
      }
      bar_builders();                                 // halo complete
      mbar_wait(&empty[s], ph ^ 1);                   // the MMAs that read this stage two blocks ago are done
      uint8_t* xs = x_stage + s * Cfg::X_STAGE;
#pragma unroll
      for (int pl = 0; pl < SW_TD; ++pl) build_row(halo, pl, bt, xs + pl * 128 * ST_ROW_BYTES);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      bar_builders();                                 // everyone has read the halo before it is overwritten
    }
    // ---- epilogue: lanes 0..63 of the accumulator are k = 0..63 -> tap k (hi rows) / k - 32 (lo rows)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(done, 0);
    tc_fence_after();
    const bool has_work = split < p.total_blocks;
    const int tap = row < 27 ? row : (row >= 32 && row < 59 ? row - 32 : -1);
#pragma unroll
    for (int c0 = 0; c0 < NCO; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, r);
      tmem_ld_wait();
      if (has_work && tap >= 0) {
        float* dst = p.dw + static_cast<int64_t>(tap) * p.cout + co0 + c0;
#pragma unroll
        for (int c = 0; c < 32; ++c) atomicAdd(dst + c, __uint_as_float(r[c]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<128>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn stem_encode() {
  bind_primary_context();
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &st) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}

template <int NT>
int launch_stem_fwd(const float* img, const void* wpk, void* y, double* stats, int n, int d, int h, int w, cudaStream_t s) {
  using Cfg = StemFwdCfg<NT>;
  const int DT = ceil_div(d, SF_TD), HT = ceil_div(h, ST_TH), WT = ceil_div(w, ST_TW);
  const int64_t items = static_cast<int64_t>(n) * DT * HT * WT;
  MMPL_REQUIRE(items < (1ll << 31), MMPL_E_SHAPE, "stem_tc_fwd: too many work items");
  MMPL_CUDA(cudaFuncSetAttribute(stem_tc_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = static_cast<int>(std::min<int64_t>(items, static_cast<int64_t>(num_sms()) * 2));
  stem_tc_fwd_kernel<NT><<<grid, SF_THREADS, Cfg::SMEM_BYTES, s>>>(img, static_cast<const __nv_bfloat16*>(wpk),
                                                                 static_cast<__nv_bfloat16*>(y), stats, n, d, h, w, DT, HT, WT,
                                                                 static_cast<int>(items));
  return MMPL_OK;
}

template <int NCO>
int launch_stem_wgrad(const float* img, const void* dy, float* dw, int n, int d, int h, int w, int cout, cudaStream_t s) {
  using Cfg = StemWgCfg<NCO>;
  EncodeTiledFn enc = stem_encode();
  MMPL_REQUIRE(enc != nullptr, MMPL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  CUtensorMap tmY;
  cuuint64_t gd[5] = {(cuuint64_t)cout, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)d, (cuuint64_t)n};
  cuuint64_t gs[4] = {(cuuint64_t)cout * 2, (cuuint64_t)w * cout * 2, (cuuint64_t)h * w * cout * 2, (cuuint64_t)d * h * w * cout * 2};
  cuuint32_t bx[5] = {(cuuint32_t)NCO, ST_TW, ST_TH, SW_TD, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, NCO * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMPL_REQUIRE(r == CUDA_SUCCESS, MMPL_E_CUDA, "cuTensorMapEncodeTiled(stem dY) failed: %d", (int)r);
  StemWgParams p;
  p.img = img, p.dw = dw, p.N = n, p.D = d, p.H = h, p.W = w, p.cout = cout;
  p.DT = ceil_div(d, SW_TD), p.HT = ceil_div(h, ST_TH), p.WT = ceil_div(w, ST_TW);
  const int64_t blocks = static_cast<int64_t>(n) * p.DT * p.HT * p.WT;
  MMPL_REQUIRE(blocks < (1ll << 31), MMPL_E_SHAPE, "stem_tc_wgrad: too many voxel blocks");
  p.total_blocks = static_cast<int>(blocks);
  p.n_co = cout / NCO;
  int ks = 2 * num_sms() / p.n_co;          // two CTAs per SM: twice the builder warps per SM
  if (ks < 1) ks = 1;
  if (ks > p.total_blocks) ks = p.total_blocks;
  p.ksplit = ks;
  MMPL_CUDA(cudaFuncSetAttribute(stem_tc_wgrad_kernel<NCO>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  MMPL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 27 * cout, s));
  stem_tc_wgrad_kernel<NCO><<<p.n_co * ks, SW_THREADS, Cfg::SMEM_BYTES, s>>>(tmY, p);
  return MMPL_OK;
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_stem_tc_fwd(const float* image, const void* w_packed, void* y, double* gn_stats_out, int n, int d,
                                int h, int w, int cout, mmpl_stream_t stream) {
  MMPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "stem_tc_fwd: empty image");
  MMPL_REQUIRE(cout == 32 || cout == 64, MMPL_E_UNSUPPORTED, "stem_tc_fwd: cout=%d (32 or 64)", cout);
  MMPL_REQUIRE((reinterpret_cast<uintptr_t>(w_packed) | reinterpret_cast<uintptr_t>(y)) % 16 == 0, MMPL_E_ALIGN,
               "stem_tc_fwd: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = cout == 32 ? launch_stem_fwd<32>(image, w_packed, y, gn_stats_out, n, d, h, w, s)
                      : launch_stem_fwd<64>(image, w_packed, y, gn_stats_out, n, d, h, w, s);
  if (rc) return rc;
  MMPL_CHECK_LAUNCH("stem_tc_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_stem_tc_wgrad(const float* image, const void* dy, float* dw_tapmajor, int n, int d, int h, int w,
                                  int cout, mmpl_stream_t stream) {
  MMPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, MMPL_E_SHAPE, "stem_tc_wgrad: empty image");
  MMPL_REQUIRE(cout == 32 || cout == 64, MMPL_E_UNSUPPORTED, "stem_tc_wgrad: cout=%d (32 or 64)", cout);
  MMPL_REQUIRE((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw_tapmajor)) % 16 == 0, MMPL_E_ALIGN,
               "stem_tc_wgrad: pointers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = cout == 32 ? launch_stem_wgrad<32>(image, dy, dw_tapmajor, n, d, h, w, cout, s)
                      : launch_stem_wgrad<64>(image, dy, dw_tapmajor, n, d, h, w, cout, s);
  if (rc) return rc;
  MMPL_CHECK_LAUNCH("stem_tc_wgrad");
  return MMPL_OK;
}
