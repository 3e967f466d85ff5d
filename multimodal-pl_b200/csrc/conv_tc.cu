// tcgen05 / TMEM / TMA implicit-GEMM 3x3x3 stride-1 convolution for sm_100a (bf16 in, fp32 accumulate in TMEM).
// Reference op: F.conv3d in Conv3d.forward (unet3D.py:27) for the 23 stride-1 3x3x3 sites of the backbone, and -- with
// the flipped/transposed dgrad packing -- their autograd data gradient.
//
// Design (B200-first, not an im2col translation):
//  * The activation tensor is NDHWC bf16.  One work item is a TD x 16 x 8 block of output voxels (x one tile of NT
//    output channels).  For each 64-(or 32-)channel chunk of Cin, ONE 5-D TMA box load brings the halo block
//    (TD+2) x 18 x 10 voxels x KC channels into shared memory (hardware zero-fill implements the padding).
//  * The 27 filter taps are NOT materialised: each tap is the same shared-memory block viewed through a UMMA
//    K-major descriptor whose start address is shifted by ((kd*18 + kh)*10 + kw) rows and whose 8-row-group stride
//    (SBO) is the 10-voxel line pitch.  Hardware swizzling is a pure function of the shared-memory address
//    (verified on B200 by tools/probe_umma.cu, profiles/r01_probe_umma.log), so any row shift is legal.
//    L2->SMEM traffic is therefore ~2.1x the activation instead of the 27x of a tap-by-tap im2col.
//  * M = 128 rows = 16 h-lines x 8 w-voxels of one d-plane; TD planes share every weight tile (TD accumulators in
//    TMEM), N = NT output channels, K = 16 per tcgen05.mma.  Accumulators are double-buffered in TMEM so the
//    epilogue of item i overlaps the MMAs of item i+1.
//  * Warp roles: 0 = activation TMA producer, 1 = weight TMA producer, 2 = MMA issuer (one elected thread) and TMEM
//    allocator, 3..6 = epilogue (tcgen05.ld -> +residual -> bf16 -> 16-byte global stores).
#include "common.cuh"
#include "ptx.cuh"

namespace mmpl {
namespace {

using namespace ptx;

constexpr int TC_TH = 16, TC_TW = 8, TC_PH = TC_TH + 2, TC_PW = TC_TW + 2;
constexpr int TC_THREADS = 224;

template <int KC, int NT, int TD>
struct TcCfg {
  static constexpr int RB = KC * 2;
  static constexpr uint32_t SWZ = RB == 128 ? SWZ_128B : SWZ_64B;
  static constexpr int PD = TD + 2;
  static constexpr int A_BYTES = PD * TC_PH * TC_PW * RB;
  static constexpr int A_STAGE = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int NA = 2;
  static constexpr int B_BYTES = NT * RB;
  static constexpr int SMEM_LIMIT = 227 * 1024 - 2048;
  static constexpr int NB_FIT = (SMEM_LIMIT - NA * A_STAGE) / B_BYTES;
  static constexpr int NB = NB_FIT > 8 ? 8 : NB_FIT;
  static constexpr int ACC_COLS = TD * NT;
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = NA * A_STAGE + NB * B_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(B_BYTES % 1024 == 0, "weight stage must keep 1024-byte alignment");
  static_assert(NB >= 2, "need at least two weight stages");
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
};

struct TcParams {
  __nv_bfloat16* y;
  const __nv_bfloat16* residual;
  int N, D, H, W;
  int nch;         // Cin / KC
  int cout_total;  // full output-channel count (row stride of y)
  int DT, HT, WT, NTILES;
  int total_items;
};

template <int KC, int NT, int TD>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using Cfg = TcCfg<KC, NT, TD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_stage = smem;
  uint8_t* b_stage = smem + Cfg::NA * Cfg::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_stage + Cfg::NB * Cfg::B_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + Cfg::NA;
  uint64_t* b_full = a_empty + Cfg::NA;
  uint64_t* b_empty = b_full + Cfg::NB;
  uint64_t* acc_full = b_empty + Cfg::NB;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::NA; ++i) mbar_init(&a_full[i], 1), mbar_init(&a_empty[i], 1);
    for (int i = 0; i < Cfg::NB; ++i) mbar_init(&b_full[i], 1), mbar_init(&b_empty[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&acc_full[i], 1), mbar_init(&acc_empty[i], 4);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_coords = [&](int item, int& nt, int& n, int& d0, int& h0, int& w0) {
    w0 = (item % p.WT) * TC_TW;
    item /= p.WT;
    h0 = (item % p.HT) * TC_TH;
    item /= p.HT;
    d0 = (item % p.DT) * TD;
    item /= p.DT;
    n = item % p.N;
    nt = item / p.N;
  };

  if (warp == 0) {
    // ===================================================== activation producer: one halo block per (item, chunk)
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int nt, n, d0, h0, w0;
        item_coords(item, nt, n, d0, h0, w0);
        for (int ch = 0; ch < p.nch; ++ch, ++it) {
          const uint32_t s = it % Cfg::NA, ph = (it / Cfg::NA) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          mbar_expect_tx(&a_full[s], Cfg::A_BYTES);
          tma_load_5d(a_stage + s * Cfg::A_STAGE, &tmA, &a_full[s], ch * KC, w0 - 1, h0 - 1, d0 - 1, n);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== weight producer: one [NT x KC] tile per (item, chunk, tap)
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int nt, n, d0, h0, w0;
        item_coords(item, nt, n, d0, h0, w0);
        for (int ch = 0; ch < p.nch; ++ch) {
          for (int tap = 0; tap < 27; ++tap, ++it) {
            const uint32_t s = it % Cfg::NB, ph = (it / Cfg::NB) & 1;
            mbar_wait(&b_empty[s], ph ^ 1);
            mbar_expect_tx(&b_full[s], Cfg::B_BYTES);
            tma_load_3d(b_stage + s * Cfg::B_BYTES, &tmB, &b_full[s], ch * KC, nt * NT, tap);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
      const uint32_t a_base = smem_u32(a_stage), b_base = smem_u32(b_stage);
      uint32_t ita = 0, itb = 0, iti = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iti) {
        const uint32_t buf = iti & 1, bph = (iti >> 1) & 1;
        mbar_wait(&acc_empty[buf], bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_COLS;
        for (int ch = 0; ch < p.nch; ++ch, ++ita) {
          const uint32_t sa = ita % Cfg::NA, pha = (ita / Cfg::NA) & 1;
          mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          const uint32_t a_addr = a_base + sa * Cfg::A_STAGE;
          for (int tap = 0; tap < 27; ++tap, ++itb) {
            const uint32_t sb = itb % Cfg::NB, phb = (itb / Cfg::NB) & 1;
            mbar_wait(&b_full[sb], phb);
            tc_fence_after();
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            const uint32_t b_addr = b_base + sb * Cfg::B_BYTES;
#pragma unroll
            for (int pl = 0; pl < TD; ++pl) {
              const uint32_t a_row = a_addr + (((pl + kd) * TC_PH + kh) * TC_PW + kw) * Cfg::RB;
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t ad = make_smem_desc(a_row + ks * 32, 16, TC_PW * Cfg::RB, Cfg::SWZ, 0);
                const uint64_t bd = make_smem_desc(b_addr + ks * 32, 16, 8 * Cfg::RB, Cfg::SWZ, 0);
                umma_f16(d_tmem + pl * NT, ad, bd, idesc, (ch | tap | ks) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&b_empty[sb]);
          }
          umma_commit(&a_empty[sa]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================================================== epilogue warps 3..6 (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rh = row >> 3, rw = row & 7;
    uint32_t iti = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iti) {
      int nt, n, d0, h0, w0;
      item_coords(item, nt, n, d0, h0, w0);
      const uint32_t buf = iti & 1, bph = (iti >> 1) & 1;
      mbar_wait(&acc_full[buf], bph);
      tc_fence_after();
      const int hh = h0 + rh, ww = w0 + rw;
      const bool in_hw = hh < p.H && ww < p.W;
#pragma unroll
      for (int pl = 0; pl < TD; ++pl) {
        const int dd = d0 + pl;
        const bool valid = in_hw && dd < p.D;
        const int64_t off = ((((static_cast<int64_t>(n) * p.D + dd) * p.H + hh) * p.W + ww) * p.cout_total) + nt * NT;
#pragma unroll
        for (int c0 = 0; c0 < NT; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg::ACC_COLS + pl * NT + c0, r);
          tmem_ld_wait();
          if (valid) {
            if (p.residual) {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                const uint4 rv = *reinterpret_cast<const uint4*>(p.residual + off + c0 + v * 8);
                const uint32_t u[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  r[v * 8 + 2 * k] = __float_as_uint(__uint_as_float(r[v * 8 + 2 * k]) + __uint_as_float(u[k] << 16));
                  r[v * 8 + 2 * k + 1] =
                      __float_as_uint(__uint_as_float(r[v * 8 + 2 * k + 1]) + __uint_as_float(u[k] & 0xFFFF0000u));
                }
              }
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint32_t o[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 h2 =
                    __floats2bfloat162_rn(__uint_as_float(r[v * 8 + 2 * k]), __uint_as_float(r[v * 8 + 2 * k + 1]));
                o[k] = *reinterpret_cast<uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(p.y + off + c0 + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_act_map(CUtensorMap* m, const void* ptr, int N, int D, int H, int W, int C, int kc, int pd, int ph, int pw,
                 int rb) {
  EncodeTiledFn enc = get_encode();
  MMPL_REQUIRE(enc != nullptr, MMPL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gd[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gs[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t bx[5] = {(cuuint32_t)kc, (cuuint32_t)pw, (cuuint32_t)ph, (cuuint32_t)pd, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMPL_REQUIRE(r == CUDA_SUCCESS, MMPL_E_CUDA, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  return MMPL_OK;
}

int make_weight_map(CUtensorMap* m, const void* ptr, int taps, int cout, int cin, int kc, int nt, int rb) {
  EncodeTiledFn enc = get_encode();
  MMPL_REQUIRE(enc != nullptr, MMPL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gd[3] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)taps};
  cuuint64_t gs[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cout * cin * 2};
  cuuint32_t bx[3] = {(cuuint32_t)kc, (cuuint32_t)nt, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMPL_REQUIRE(r == CUDA_SUCCESS, MMPL_E_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return MMPL_OK;
}

template <int KC, int NT, int TD>
int launch_tc(const void* x, const void* wp, const void* residual, void* y, int N, int D, int H, int W, int cin,
              int cout, cudaStream_t s) {
  using Cfg = TcCfg<KC, NT, TD>;
  CUtensorMap tmA, tmB;
  if (int e = make_act_map(&tmA, x, N, D, H, W, cin, KC, Cfg::PD, TC_PH, TC_PW, Cfg::RB)) return e;
  if (int e = make_weight_map(&tmB, wp, 27, cout, cin, KC, NT, Cfg::RB)) return e;
  TcParams p;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.N = N, p.D = D, p.H = H, p.W = W;
  p.nch = cin / KC;
  p.cout_total = cout;
  p.DT = ceil_div(D, TD), p.HT = ceil_div(H, TC_TH), p.WT = ceil_div(W, TC_TW), p.NTILES = cout / NT;
  const int64_t items = static_cast<int64_t>(p.NTILES) * N * p.DT * p.HT * p.WT;
  MMPL_REQUIRE(items < (1ll << 31), MMPL_E_SHAPE, "conv_tc: too many work items");
  p.total_items = static_cast<int>(items);
  static bool attr_set = false;
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(conv3_tc_kernel<KC, NT, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = static_cast<int>(std::min<int64_t>(items, num_sms()));
  conv3_tc_kernel<KC, NT, TD><<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(tmA, tmB, p);
  MMPL_CHECK_LAUNCH("conv3_tc");
  return MMPL_OK;
}

}  // namespace

// x [N,D,H,W,cin] bf16, wp [27][cout][cin] bf16 (either packing), y [N,D,H,W,cout] bf16 (+ residual).
int conv_tc_3x3x3_s1(const void* x, const void* wp, const void* residual, void* y, int N, int D, int H, int W, int cin,
                     int cout, cudaStream_t s) {
  MMPL_REQUIRE((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wp) | reinterpret_cast<uintptr_t>(y) |
                reinterpret_cast<uintptr_t>(residual)) % 16 == 0,
               MMPL_E_ALIGN, "conv_tc: pointers must be 16-byte aligned");
  const int kc = cin < 64 ? cin : 64;
  MMPL_REQUIRE(cin == 32 || cin % 64 == 0, MMPL_E_UNSUPPORTED, "conv_tc: cin=%d (32 or a multiple of 64)", cin);
  int nt = cout;
  if (cout > 256) {
    MMPL_REQUIRE(cout % 256 == 0, MMPL_E_UNSUPPORTED, "conv_tc: cout=%d", cout);
    nt = 256;
  }
  MMPL_REQUIRE(nt == 32 || nt == 64 || nt == 128 || nt == 256, MMPL_E_UNSUPPORTED, "conv_tc: cout=%d", cout);
  if (kc == 32) {
    if (nt == 32) return launch_tc<32, 32, 4>(x, wp, residual, y, N, D, H, W, cin, cout, s);
    if (nt == 64) return launch_tc<32, 64, 4>(x, wp, residual, y, N, D, H, W, cin, cout, s);
    MMPL_FAIL(MMPL_E_UNSUPPORTED, "conv_tc: cin=32 with cout=%d", cout);
  }
  if (nt == 32) return launch_tc<64, 32, 2>(x, wp, residual, y, N, D, H, W, cin, cout, s);
  if (nt == 64) return launch_tc<64, 64, 2>(x, wp, residual, y, N, D, H, W, cin, cout, s);
  if (nt == 128) return launch_tc<64, 128, 2>(x, wp, residual, y, N, D, H, W, cin, cout, s);
  return launch_tc<64, 256, 1>(x, wp, residual, y, N, D, H, W, cin, cout, s);
}

}  // namespace mmpl
