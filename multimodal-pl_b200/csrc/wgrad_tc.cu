// tcgen05 weight-gradient kernels (bf16 in, fp32 accumulate in TMEM) for every convolution of the backbone:
// 3x3x3 / 1x1x1, stride 1 / 2.  Reference op: autograd of F.conv3d in Conv3d.forward (unet3D.py:27) w.r.t. the weight.
//
//   dW[(kd,kh,kw)][co][ci] = sum_o dY[o][co] * X[stride*o + (kd,kh,kw) - pad][ci]
//
// GEMM view: K = voxels (millions), M/N = channels (tiny).  Both operands are channel-contiguous (NDHWC), i.e.
// MN-major for the tensor core, which tcgen05 supports directly for bf16 -- no transposes.
//  * Work block = TD x 16 x 8 output voxels (same TMA boxes as the forward kernel: halo block(s) of X, dense block of
//    dY; TMA zero-fill makes both the padding and the partial edge tiles contribute exact zeros).
//  * kw-packing: the A operand is the X block viewed MN-major with the M dimension made of 64-/128-byte channel chunks
//    that are ONE VOXEL ROW apart (LBO = row pitch).  Chunk j is therefore the block shifted by j voxels along w, so a
//    single M=128 MMA produces the gradients of several kw taps at once; unused chunks are dropped in the epilogue.
//    Verified on B200: tools/probe_umma.cu T2.
//  * K = 16 per MMA = two 8-voxel w-lines (h, h+1); SBO is the line pitch (PW rows for X, 8 rows for dY).
//  * Stride 2 reads the parity-split copy P of X (mmpl_parity_split): tap k of an axis lives in parity (k != 1) at
//    shift (k != 0) relative to a block that starts one voxel before the tile, so every tap is again a plain shifted
//    view; a CTA owns one (pd, ph) parity pair and loads the two pw blocks per voxel block.
//  * An "accumulator group" table (<= 12 entries, built on the host) tells the MMA issuer which X block / row shift
//    feeds which TMEM accumulator and tells the epilogue which filter tap each M-chunk of it is.
//  * Split-K lives in TMEM: a CTA keeps its accumulators resident across all the voxel blocks assigned to it and only at
//    the very end adds them to dW with coalesced fp32 red.global (<= 148 partials per element).
#include "common.cuh"
#include "ptx.cuh"

namespace mmpl {
namespace {

using namespace ptx;

constexpr int WG_TH = 16, WG_TW = 8;
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_GROUPS = 12;

// HALO: extra voxels per axis of the X block (2: 3x3x3 s1, 1: 3x3x3 s2 from P, 0: 1x1x1)
// XB: X blocks per stage; PDE: extra planes of the X block (TD + PDE planes)
template <int KC, int NCO, int TD, int HALO, int XB, int PDE>
struct WgCfg {
  static constexpr int RBX = KC * 2, RBY = NCO * 2;
  static constexpr uint32_t SWX = RBX == 128 ? SWZ_128B : SWZ_64B;
  static constexpr uint32_t SWY = RBY == 128 ? SWZ_128B : SWZ_64B;
  static constexpr int PH = WG_TH + HALO, PW = WG_TW + HALO, PDX = TD + PDE;
  static constexpr int XBLK_BYTES = PDX * PH * PW * RBX;
  static constexpr int XBLK_STRIDE = (XBLK_BYTES + 1023) / 1024 * 1024;
  static constexpr int Y_BYTES = TD * WG_TH * WG_TW * RBY;
  static constexpr int X_STAGE = XB * XBLK_STRIDE + 1024;   // slack: unused M-chunks over-read a few rows
  static constexpr int Y_STAGE = (Y_BYTES + 1023) / 1024 * 1024;
  static constexpr int NS = 2;
  static constexpr int SMEM_BYTES = NS * (X_STAGE + Y_STAGE) + 1024 + 256;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct WgGroup {
  int xblk;        // which X block of the stage
  int row_shift;   // (sd*PH + sh)*PW + sw, in voxel rows
  int tap[4];      // filter tap produced by M-chunk j (-1: discard)
};

struct WgParams {
  float* dw;
  int N, D, H, W;           // extents of the dY tensor (tile-grid domain)
  int cin, cout;
  int DT, HT, WT;
  int n_ci, n_co, n_var;    // CTA "combo" = (ci chunk, co chunk, variant); variant selects the group table
  int combos, ksplit;
  int total_blocks;
  int xscale;               // 2: X coordinates are 2*o (element-strided TMA, 1x1x1 stride 2), else 1
  int xlo;                  // X block origin relative to the tile origin (-1 with halo, 0 without)
  int ngroups[4];                       // per variant
  int xd_off[4];                        // per variant: extra d offset of the X block (kd for the KDS=1 split)
  int xn_off[4][2];                     // per variant, per X block: batch-plane offset (parity plane * N)
  WgGroup groups[4][WG_MAX_GROUPS];     // per variant
};

// PM (plane-major): only for the all-27-taps variant (HALO 2, PDE 2, one X block).  For X plane s, line pair hl and kh the
// SAME A view pairs with the dY planes p = s - kd, so one MMA whose B operand spans those planes (N-chunks one dY
// plane apart, LBO = plane pitch) accumulates into the adjacent accumulators (kh, 2-kd): half the MMAs and A reads.
template <int KC, int NCO, int TD, int HALO, int XB, int PDE, bool PM>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ WgParams p) {
  using Cfg = WgCfg<KC, NCO, TD, HALO, XB, PDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* x_stage = smem;
  uint8_t* y_stage = smem + Cfg::NS * Cfg::X_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(y_stage + Cfg::NS * Cfg::Y_STAGE);
  uint64_t* full = bars;
  uint64_t* empty = full + Cfg::NS;
  uint64_t* done = empty + Cfg::NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::NS; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    mbar_init(done, 1);
    fence_barrier_init();
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmY);
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int combo = blockIdx.x % p.combos, split = blockIdx.x / p.combos;
  const int ci_i = combo % p.n_ci;
  const int co_i = (combo / p.n_ci) % p.n_co;
  const int var = combo / (p.n_ci * p.n_co);
  const int ci0 = ci_i * KC, co0 = co_i * NCO;
  const int ng = p.ngroups[var];

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int b = split; b < p.total_blocks; b += p.ksplit, ++it) {
        int r = b;
        const int w0 = (r % p.WT) * WG_TW;
        r /= p.WT;
        const int h0 = (r % p.HT) * WG_TH;
        r /= p.HT;
        const int d0 = (r % p.DT) * TD;
        const int n = r / p.DT;
        const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], XB * Cfg::XBLK_BYTES + Cfg::Y_BYTES);
#pragma unroll
        for (int xb = 0; xb < XB; ++xb)
          tma_load_5d(x_stage + s * Cfg::X_STAGE + xb * Cfg::XBLK_STRIDE, &tmX, &full[s], ci0, p.xscale * w0 + p.xlo,
                      p.xscale * h0 + p.xlo, p.xscale * d0 + p.xlo + p.xd_off[var], p.xn_off[var][xb] + n);
        tma_load_5d(y_stage + s * Cfg::Y_STAGE, &tmY, &full[s], co0, w0, h0, d0, n);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, one elected lane issues
    const uint32_t idesc = make_idesc_bf16(128, NCO, 1, 1);
    const uint32_t xb0 = smem_u32(x_stage), yb0 = smem_u32(y_stage);
    const uint64_t a_fix = make_smem_desc(0, Cfg::RBX, Cfg::PW * Cfg::RBX, Cfg::SWX, 0);
    const uint64_t b_fix = make_smem_desc(0, 64 * Cfg::RBY, WG_TW * Cfg::RBY, Cfg::SWY, 0);
    const uint32_t a_hi = static_cast<uint32_t>(a_fix >> 32), a_lo_fix = static_cast<uint32_t>(a_fix);
    const uint32_t b_hi = static_cast<uint32_t>(b_fix >> 32), b_lo_fix = static_cast<uint32_t>(b_fix);
    uint32_t it = 0;
    for (int b = split; b < p.total_blocks; b += p.ksplit, ++it) {
      const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t xs = xb0 + s * Cfg::X_STAGE, ys = yb0 + s * Cfg::Y_STAGE;
      if (PM) {
        if (elect_one()) {
          constexpr uint32_t YPLANE = WG_TH * WG_TW * Cfg::RBY;           // bytes between dY planes
          // B descriptor for the multi-plane operand: N-chunks (NCO channels each) are one dY plane apart
          const uint64_t bp_fix = make_smem_desc(0, YPLANE, WG_TW * Cfg::RBY, Cfg::SWY, 0);
          const uint32_t bp_hi = static_cast<uint32_t>(bp_fix >> 32), bp_lo_fix = static_cast<uint32_t>(bp_fix);
#pragma unroll
          for (int sp = 0; sp < TD + 2; ++sp) {
            const int kd_hi = sp < 2 ? sp : 2;
            const int kd_lo = sp - (TD - 1) > 0 ? sp - (TD - 1) : 0;
            const int nblk = kd_hi - kd_lo + 1;
            const int p_lo = sp - kd_hi;
            const bool fresh = sp <= 2;          // accumulators (kd = sp, kh) see their first MMA at (sp, hl = 0)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const uint32_t acc = tmem_base + (kh * 3 + (2 - kd_hi)) * NCO;
#pragma unroll
              for (int hl = 0; hl < WG_TH; hl += 2) {
                const uint32_t a_lo = a_lo_fix | ((xs + (((sp * Cfg::PH) + hl + kh) * Cfg::PW) * Cfg::RBX) >> 4);
                const uint32_t b_lo = bp_lo_fix | ((ys + ((p_lo * WG_TH + hl) * WG_TW) * Cfg::RBY) >> 4);
                if (fresh && hl == 0) {
                  umma_f16_lohi(acc, a_lo, a_hi, b_lo, bp_hi, make_idesc_bf16(128, NCO, 1, 1), it != 0 ? 1u : 0u);
                  if (nblk > 1)
                    umma_f16_lohi(acc + NCO, a_lo, a_hi, b_lo + (YPLANE >> 4), bp_hi,
                                  make_idesc_bf16(128, NCO * (nblk - 1), 1, 1), 1u);
                } else {
                  umma_f16_lohi(acc, a_lo, a_hi, b_lo, bp_hi, make_idesc_bf16(128, NCO * nblk, 1, 1), 1u);
                }
              }
            }
          }
          umma_commit(&empty[s]);
        }
        __syncwarp();
        continue;
      }
      if (elect_one()) {
        for (int g = 0; g < ng; ++g) {
          const uint32_t acc = tmem_base + g * NCO;
          const uint32_t xg = xs + p.groups[var][g].xblk * Cfg::XBLK_STRIDE + p.groups[var][g].row_shift * Cfg::RBX;
          const uint32_t a_lo = a_lo_fix | (xg >> 4), b_lo = b_lo_fix | (ys >> 4);
#pragma unroll
          for (int pl = 0; pl < TD; ++pl) {
#pragma unroll
            for (int hl = 0; hl < WG_TH; hl += 2) {
              umma_f16_lohi(acc, a_lo + ((((pl * Cfg::PH + hl) * Cfg::PW) * Cfg::RBX) >> 4), a_hi,
                            b_lo + ((((pl * WG_TH + hl) * WG_TW) * Cfg::RBY) >> 4), b_hi, idesc,
                            (it | pl | hl) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
  } else {
    // epilogue warps 2..5: once, after the last block
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int j = row / KC, ci_l = row % KC;
    mbar_wait(done, 0);
    tc_fence_after();
    const bool has_work = split < p.total_blocks;
#pragma unroll 1
    for (int g = 0; g < (PM ? 9 : ng); ++g) {
      int tap = p.groups[var][PM ? 0 : g].tap[j & 3];
      if (PM) tap = j < 3 ? ((2 - g % 3) * 3 + g / 3) * 3 + j : -1;
      const bool valid = has_work && tap >= 0;
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * NCO + c0, r);
        tmem_ld_wait();
        if (valid) {
          float* dst = p.dw + (static_cast<int64_t>(tap) * p.cout + co0 + c0) * p.cin + ci0 + ci_l;
#pragma unroll
          for (int c = 0; c < 32; ++c) atomicAdd(dst + static_cast<int64_t>(c) * p.cin, __uint_as_float(r[c]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg_encode() {
  bind_primary_context();
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int wg_act_map(CUtensorMap* m, const void* ptr, int64_t N, int D, int H, int W, int C, int bc, int bd, int bh, int bw,
               int estride) {
  EncodeTiledFn enc = wg_encode();
  MMPL_REQUIRE(enc != nullptr, MMPL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  const cuuint32_t e = static_cast<cuuint32_t>(estride);
  cuuint64_t gd[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gs[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t bx[5] = {(cuuint32_t)bc, (cuuint32_t)bw * e, (cuuint32_t)bh * e, (cuuint32_t)bd * e, 1};
  cuuint32_t es[5] = {1, e, e, e, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMPL_REQUIRE(r == CUDA_SUCCESS, MMPL_E_CUDA, "cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r);
  return MMPL_OK;
}

// X source description: tensor extents and how block coordinates derive from the tile origin.
struct WgSource {
  const void* x;
  int64_t xN;
  int xD, xH, xW;
  int estride;
};

template <int KC, int NCO, int TD, int HALO, int XB, int PDE, bool PM = false>
int launch_wg(const WgSource& src, const void* dy, WgParams p, int taps, cudaStream_t s) {
  using Cfg = WgCfg<KC, NCO, TD, HALO, XB, PDE>;
  CUtensorMap tmX, tmY;
  if (int e = wg_act_map(&tmX, src.x, src.xN, src.xD, src.xH, src.xW, p.cin, KC, Cfg::PDX, Cfg::PH, Cfg::PW, src.estride)) return e;
  if (int e = wg_act_map(&tmY, dy, p.N, p.D, p.H, p.W, p.cout, NCO, TD, WG_TH, WG_TW, 1)) return e;
  p.DT = ceil_div(p.D, TD), p.HT = ceil_div(p.H, WG_TH), p.WT = ceil_div(p.W, WG_TW);
  p.n_ci = p.cin / KC, p.n_co = p.cout / NCO;
  p.combos = p.n_ci * p.n_co * p.n_var;
  for (int v = 0; v < p.n_var; ++v)
    MMPL_REQUIRE(p.ngroups[v] * NCO <= 512 && p.ngroups[v] <= WG_MAX_GROUPS, MMPL_E_UNSUPPORTED, "wgrad_tc: accumulators exceed TMEM");
  const int64_t blocks = static_cast<int64_t>(p.N) * p.DT * p.HT * p.WT;
  MMPL_REQUIRE(blocks < (1ll << 31), MMPL_E_SHAPE, "wgrad_tc: too many voxel blocks");
  p.total_blocks = static_cast<int>(blocks);
  int ks = num_sms() / p.combos;
  if (ks < 1) ks = 1;
  if (ks > p.total_blocks) ks = p.total_blocks;
  p.ksplit = ks;
  // the attribute is per device: one flag per device ordinal for every instantiation
  static bool attr_set_dev[64] = {};
  int dev_ord = 0;
  cudaGetDevice(&dev_ord);
  bool& attr_set = attr_set_dev[dev_ord & 63];
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<KC, NCO, TD, HALO, XB, PDE, PM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::SMEM_BYTES));
    attr_set = true;
  }
  MMPL_CUDA(cudaMemsetAsync(p.dw, 0, sizeof(float) * taps * p.cout * p.cin, s));
  wgrad_tc_kernel<KC, NCO, TD, HALO, XB, PDE, PM><<<p.combos * ks, WG_THREADS, Cfg::SMEM_BYTES, s>>>(tmX, tmY, p);
  MMPL_CHECK_LAUNCH("wgrad_tc");
  return MMPL_OK;
}

void clear_params(WgParams& p) {
  memset(&p, 0, sizeof(p));
  for (int v = 0; v < 4; ++v)
    for (int g = 0; g < WG_MAX_GROUPS; ++g)
      for (int j = 0; j < 4; ++j) p.groups[v][g].tap[j] = -1;
}

}  // namespace

size_t conv_tc_wgrad_workspace(int, int, int, int, int, int) { return 0; }

bool conv_tc_wgrad_supported(int cin, int cout) {
  if (cin == 32) return cout == 32 || cout % 64 == 0;
  if (cin % 64 != 0) return false;
  return cout == 32 || cout % 64 == 0;
}

// x: NDHWC input (stride 1, and 1x1x1 stride 2) or the parity-split tensor P (3x3x3 stride 2).
// (d, h, w) are the extents of the conv INPUT; dy has the output extents.
int conv_tc_wgrad(const void* x, const void* dy, float* dw, int N, int d, int h, int w, int cin, int cout, int ksize,
                  int stride, cudaStream_t s) {
  MMPL_REQUIRE(conv_tc_wgrad_supported(cin, cout), MMPL_E_UNSUPPORTED, "wgrad_tc: cin=%d cout=%d", cin, cout);
  MMPL_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw)) % 16 == 0,
               MMPL_E_ALIGN, "wgrad_tc: pointers must be 16-byte aligned");
  const int Do = stride == 1 ? d : (d + 1) / 2, Ho = stride == 1 ? h : (h + 1) / 2, Wo = stride == 1 ? w : (w + 1) / 2;
  const int kc = cin == 32 ? 32 : 64;
  const int nchunk = 128 / kc;            // M-chunks per MMA (4 for 32 channels, 2 for 64)
  WgParams p;
  clear_params(p);
  p.dw = dw;
  p.N = N, p.D = Do, p.H = Ho, p.W = Wo, p.cin = cin, p.cout = cout;
  p.xscale = 1;
  const int taps = ksize * ksize * ksize;

  if (ksize == 1) {
    // one group, chunk 0 is the only tap
    p.n_var = 1, p.ngroups[0] = 1, p.xlo = 0;
    p.groups[0][0].xblk = 0, p.groups[0][0].row_shift = 0, p.groups[0][0].tap[0] = 0;
    WgSource src{x, N, d, h, w, stride};
    p.xscale = stride;
    if (kc == 32) {
      if (cout == 32) return launch_wg<32, 32, 4, 0, 1, 0>(src, dy, p, taps, s);
      return launch_wg<32, 64, 4, 0, 1, 0>(src, dy, p, taps, s);
    }
    if (cout == 32) return launch_wg<64, 32, 2, 0, 1, 0>(src, dy, p, taps, s);
    return launch_wg<64, 64, 2, 0, 1, 0>(src, dy, p, taps, s);
  }

  if (stride == 1) {
    WgSource src{x, N, d, h, w, 1};
    p.xlo = -1;
    if (kc == 32 && cout == 32) {
      // all 27 taps in one CTA: 9 groups (kd,kh), chunk j = kw
      constexpr int PH = 18, PW = 10;
      p.n_var = 1, p.ngroups[0] = 9;
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh) {
          WgGroup& g = p.groups[0][kd * 3 + kh];
          g.xblk = 0, g.row_shift = (kd * PH + kh) * PW;
          for (int j = 0; j < 3; ++j) g.tap[j] = (kd * 3 + kh) * 3 + j;
        }
      return launch_wg<32, 32, 4, 2, 1, 2, true>(src, dy, p, taps, s);
    }
    // one kd per CTA variant (X block of TD planes at d0-1+kd): groups (kh, km)
    constexpr int PW = 10;
    p.n_var = 3;
    for (int kd = 0; kd < 3; ++kd) {
      p.xd_off[kd] = kd;
      int ng = 0;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw0 = 0; kw0 < 3; kw0 += nchunk) {
          WgGroup& g = p.groups[kd][ng++];
          g.xblk = 0, g.row_shift = kh * PW + kw0;
          for (int j = 0; j < nchunk && kw0 + j < 3; ++j) g.tap[j] = (kd * 3 + kh) * 3 + kw0 + j;
        }
      p.ngroups[kd] = ng;
    }
    if (kc == 32) return launch_wg<32, 64, 2, 2, 1, 0>(src, dy, p, taps, s);
    if (cout == 32) return launch_wg<64, 32, 2, 2, 1, 0>(src, dy, p, taps, s);
    return launch_wg<64, 64, 2, 2, 1, 0>(src, dy, p, taps, s);
  }

  // 3x3x3 stride 2 from the parity-split tensor: variant = (pd, ph); X blocks = pw 0/1.
  // Axis tap k lives in parity (k != 1) at shift (k != 0) of a block starting at o0 - 1.
  {
    constexpr int PH = 17, PW = 9;
    WgSource src{x, static_cast<int64_t>(8) * N, Do, Ho, Wo, 1};
    p.xlo = -1;
    p.n_var = 4;
    for (int pd = 0; pd < 2; ++pd)
      for (int ph = 0; ph < 2; ++ph) {
        const int v = pd * 2 + ph;
        p.xn_off[v][0] = (pd * 4 + ph * 2 + 0) * N;
        p.xn_off[v][1] = (pd * 4 + ph * 2 + 1) * N;
        int ng = 0;
        for (int id = 0; id < (pd ? 2 : 1); ++id)
          for (int ih = 0; ih < (ph ? 2 : 1); ++ih) {
            const int kd = pd ? 2 * id : 1, kh = ph ? 2 * ih : 1;
            const int sd = kd != 0, sh = kh != 0;
            // pw = 1 block: kw = 0 (shift 0) and kw = 2 (shift 1) are adjacent chunks
            WgGroup& g1 = p.groups[v][ng++];
            g1.xblk = 1, g1.row_shift = (sd * PH + sh) * PW + 0;
            g1.tap[0] = (kd * 3 + kh) * 3 + 0, g1.tap[1] = (kd * 3 + kh) * 3 + 2;
            // pw = 0 block: kw = 1 at shift 1
            WgGroup& g0 = p.groups[v][ng++];
            g0.xblk = 0, g0.row_shift = (sd * PH + sh) * PW + 1;
            g0.tap[0] = (kd * 3 + kh) * 3 + 1;
          }
        p.ngroups[v] = ng;
      }
    if (kc == 32) return launch_wg<32, 64, 2, 1, 2, 1>(src, dy, p, taps, s);
    return launch_wg<64, 64, 1, 1, 2, 1>(src, dy, p, taps, s);
  }
}

}  // namespace mmpl
