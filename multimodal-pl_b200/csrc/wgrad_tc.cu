// tcgen05 weight-gradient kernel for the 3x3x3 stride-1 convolutions (bf16 in, fp32 accumulate in TMEM).
// Reference op: autograd of F.conv3d in Conv3d.forward (unet3D.py:27) w.r.t. the (standardised) weight.
//
//   dW[(kd,kh,kw)][co][ci] = sum_v dY[v][co] * X[v + (kd,kh,kw) - 1][ci]
//
// GEMM view: K = voxels (millions), M/N = channels (tiny).  Both operands are channel-contiguous (NDHWC), i.e.
// MN-major for the tensor core, which tcgen05 supports directly for bf16 -- no transposes.
//  * Work block = TD x 16 x 8 voxels (same TMA boxes as the forward kernel: halo block of X, dense block of dY;
//    TMA zero-fill makes both the padding and the partial edge tiles contribute exact zeros).
//  * kw-packing: the A operand is the X halo block viewed MN-major with the M dimension made of 64-byte/128-byte
//    channel chunks that are ONE VOXEL ROW apart (LBO = row pitch).  Chunk j is therefore the block shifted by j
//    voxels along w, so a single M=128 MMA produces the gradients of kw = 0..3 (KC=32) or kw = kwbase, kwbase+1
//    (KC=64) at once; the unused chunk (kw=3) is discarded in the epilogue (75 % useful rows instead of the 25 % a
//    per-tap M=Cout formulation would give for 32-channel layers).  Verified on B200: tools/probe_umma.cu T2.
//  * K = 16 per MMA = two 8-voxel w-lines (h, h+1); SBO is the line pitch (10 rows for X, 8 rows for dY).
//  * Split-K lives in TMEM: a CTA keeps its 6..9 accumulators resident across all the voxel blocks assigned to it and
//    only at the very end adds them to dW with coalesced fp32 red.global (<= 148 partials per element).
#include "common.cuh"
#include "ptx.cuh"

namespace mmpl {
namespace {

using namespace ptx;

constexpr int WG_TH = 16, WG_TW = 8, WG_PH = 18, WG_PW = 10;
constexpr int WG_THREADS = 192;

template <int KC, int NCO, int TD, int KDS>
struct WgCfg {
  static constexpr int RBX = KC * 2, RBY = NCO * 2;
  static constexpr uint32_t SWX = RBX == 128 ? SWZ_128B : SWZ_64B;
  static constexpr uint32_t SWY = RBY == 128 ? SWZ_128B : SWZ_64B;
  static constexpr int PDX = TD + KDS - 1;
  static constexpr int X_BYTES = PDX * WG_PH * WG_PW * RBX;
  static constexpr int Y_BYTES = TD * WG_TH * WG_TW * RBY;
  static constexpr int X_STAGE = (X_BYTES + 16 * RBX + 1023) / 1024 * 1024;  // slack: the kw=3 chunk over-reads a few rows
  static constexpr int Y_STAGE = (Y_BYTES + 1023) / 1024 * 1024;
  static constexpr int NS = 2;
  static constexpr int KWM = KC == 32 ? 1 : 2;     // MMAs per (kd,kh) needed to cover kw = 0..2
  static constexpr int NACC = KDS * 3 * KWM;
  static constexpr int ACC_COLS = NACC * NCO;
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = NS * (X_STAGE + Y_STAGE) + 1024 + 256;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct WgParams {
  float* dw;
  int N, D, H, W, cin, cout;
  int DT, HT, WT;
  int n_ci, n_co, n_kdg, combos, ksplit;
  int total_blocks;
};

template <int KC, int NCO, int TD, int KDS>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const WgParams p) {
  using Cfg = WgCfg<KC, NCO, TD, KDS>;
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
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int combo = blockIdx.x % p.combos, split = blockIdx.x / p.combos;
  const int ci_i = combo % p.n_ci;
  const int co_i = (combo / p.n_ci) % p.n_co;
  const int kdg = combo / (p.n_ci * p.n_co);
  const int ci0 = ci_i * KC, co0 = co_i * NCO, kd0 = kdg * KDS;

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
        mbar_expect_tx(&full[s], Cfg::X_BYTES + Cfg::Y_BYTES);
        tma_load_5d(x_stage + s * Cfg::X_STAGE, &tmX, &full[s], ci0, w0 - 1, h0 - 1, d0 - 1 + kd0, n);
        tma_load_5d(y_stage + s * Cfg::Y_STAGE, &tmY, &full[s], co0, w0, h0, d0, n);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, NCO, 1, 1);
      const uint32_t xb = smem_u32(x_stage), yb = smem_u32(y_stage);
      uint32_t it = 0;
      for (int b = split; b < p.total_blocks; b += p.ksplit, ++it) {
        const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t xs = xb + s * Cfg::X_STAGE, ys = yb + s * Cfg::Y_STAGE;
#pragma unroll 1
        for (int kdl = 0; kdl < KDS; ++kdl) {
#pragma unroll 1
          for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int km = 0; km < Cfg::KWM; ++km) {
              const uint32_t acc = tmem_base + ((kdl * 3 + kh) * Cfg::KWM + km) * NCO;
#pragma unroll 1
              for (int pl = 0; pl < TD; ++pl) {
#pragma unroll
                for (int hl = 0; hl < WG_TH; hl += 2) {
                  const uint32_t xa = xs + ((((pl + kdl) * WG_PH + hl + kh) * WG_PW) + 2 * km) * Cfg::RBX;
                  const uint32_t ya = ys + ((pl * WG_TH + hl) * WG_TW) * Cfg::RBY;
                  const uint64_t ad = make_smem_desc(xa, Cfg::RBX, WG_PW * Cfg::RBX, Cfg::SWX, 0);
                  const uint64_t bd = make_smem_desc(ya, 64 * Cfg::RBY, WG_TW * Cfg::RBY, Cfg::SWY, 0);
                  umma_f16(acc, ad, bd, idesc, (it | pl | hl) != 0 ? 1u : 0u);
                }
              }
            }
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
  } else {
    // epilogue warps 2..5: once, after the last block
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int j = row / KC, ci_l = row % KC;
    mbar_wait(done, 0);
    tc_fence_after();
    const bool has_work = split < p.total_blocks;
#pragma unroll 1
    for (int a = 0; a < Cfg::NACC; ++a) {
      const int km = a % Cfg::KWM, kh = (a / Cfg::KWM) % 3, kdl = a / (Cfg::KWM * 3);
      const int kw = 2 * km + j;
      const int kd = kd0 + kdl;
      const bool valid = has_work && kw <= 2 && kd <= 2;
      const int tap = (kd * 3 + kh) * 3 + kw;
#pragma unroll
      for (int c0 = 0; c0 < NCO; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * NCO + c0, r);
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
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int wg_act_map(CUtensorMap* m, const void* ptr, int N, int D, int H, int W, int C, int bc, int bd, int bh, int bw) {
  EncodeTiledFn enc = wg_encode();
  MMPL_REQUIRE(enc != nullptr, MMPL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gd[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gs[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t bx[5] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMPL_REQUIRE(r == CUDA_SUCCESS, MMPL_E_CUDA, "cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r);
  return MMPL_OK;
}

template <int KC, int NCO, int TD, int KDS>
int launch_wg(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int cin, int cout, cudaStream_t s) {
  using Cfg = WgCfg<KC, NCO, TD, KDS>;
  CUtensorMap tmX, tmY;
  if (int e = wg_act_map(&tmX, x, N, D, H, W, cin, KC, Cfg::PDX, WG_PH, WG_PW)) return e;
  if (int e = wg_act_map(&tmY, dy, N, D, H, W, cout, NCO, TD, WG_TH, WG_TW)) return e;
  WgParams p;
  p.dw = dw;
  p.N = N, p.D = D, p.H = H, p.W = W, p.cin = cin, p.cout = cout;
  p.DT = ceil_div(D, TD), p.HT = ceil_div(H, WG_TH), p.WT = ceil_div(W, WG_TW);
  p.n_ci = cin / KC, p.n_co = cout / NCO, p.n_kdg = 3 / KDS;
  p.combos = p.n_ci * p.n_co * p.n_kdg;
  const int64_t blocks = static_cast<int64_t>(N) * p.DT * p.HT * p.WT;
  MMPL_REQUIRE(blocks < (1ll << 31), MMPL_E_SHAPE, "wgrad_tc: too many voxel blocks");
  p.total_blocks = static_cast<int>(blocks);
  int ks = num_sms() / p.combos;
  if (ks < 1) ks = 1;
  if (ks > p.total_blocks) ks = p.total_blocks;
  p.ksplit = ks;
  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<KC, NCO, TD, KDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  MMPL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 27 * cout * cin, s));
  wgrad_tc_kernel<KC, NCO, TD, KDS><<<p.combos * ks, WG_THREADS, Cfg::SMEM_BYTES, s>>>(tmX, tmY, p);
  MMPL_CHECK_LAUNCH("wgrad_tc");
  return MMPL_OK;
}

}  // namespace

size_t conv_tc_wgrad_workspace(int, int, int, int, int, int) { return 0; }

bool conv_tc_wgrad_supported(int cin, int cout) {
  if (cin == 32) return cout == 32;
  if (cin % 64 != 0) return false;
  return cout == 32 || cout % 64 == 0;
}

int conv_tc_wgrad_3x3x3_s1(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int cin, int cout,
                           void*, size_t, cudaStream_t s) {
  MMPL_REQUIRE(conv_tc_wgrad_supported(cin, cout), MMPL_E_UNSUPPORTED, "wgrad_tc: cin=%d cout=%d", cin, cout);
  MMPL_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw)) % 16 == 0,
               MMPL_E_ALIGN, "wgrad_tc: pointers must be 16-byte aligned");
  if (cin == 32) return launch_wg<32, 32, 4, 3>(x, dy, dw, N, D, H, W, cin, cout, s);
  if (cout == 32) return launch_wg<64, 32, 2, 1>(x, dy, dw, N, D, H, W, cin, cout, s);
  return launch_wg<64, 64, 2, 1>(x, dy, dw, N, D, H, W, cin, cout, s);
}

}  // namespace mmpl
