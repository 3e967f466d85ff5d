// tcgen05 / TMEM / TMA implicit-GEMM convolutions for sm_100a (bf16 in, fp32 accumulate in TMEM).
// Reference op: F.conv3d in Conv3d.forward (unet3D.py:27) -- all 35 weight-standardised convolutions of the backbone
// (3x3x3 and 1x1x1, stride 1 and 2) and their autograd data gradients.
//
// Design (B200-first, not an im2col translation):
//  * Activations are NDHWC bf16.  One work item is a TD x 16 x 8 block of output voxels x one tile of NT output
//    channels.  For each 64-(or 32-)channel chunk of the reduction channels ONE 5-D TMA box load brings the halo block
//    PD x PH x PW voxels x KC channels into shared memory (hardware zero-fill implements the padding).
//  * Filter taps are NOT materialised: each tap is the same shared-memory block viewed through a UMMA K-major
//    descriptor whose start address is shifted by ((sd*PH + sh)*PW + sw) rows and whose 8-row-group stride (SBO) is the
//    PW-voxel line pitch.  Hardware swizzling is a pure function of the shared-memory address (verified on B200 by
//    tools/probe_umma.cu, profiles/r01_probe_umma.log), so any row shift is legal.  L2->SMEM traffic is ~2x the
//    activation instead of the 27x of a tap-by-tap im2col.
//  * M = 128 rows = 16 h-lines x 8 w-voxels of one d-plane; TD planes share every weight tile (TD accumulators in
//    TMEM), N = NT, K = 16 per tcgen05.mma.  Accumulators are double-buffered in TMEM so the epilogue of item i
//    overlaps the MMAs of item i+1.
//  * Stride 2 never uses strided gathers in the MMA path:
//      fprop  reads a parity-split copy P[8*N][D/2][H/2][W/2][C] of the input (P_p[i] = X[2i+p], written by
//             mmpl_parity_split); tap k of parity p is P_p shifted by (k != 0) -> 8 chunks x (1..8 taps) per item;
//      dgrad  runs per parity class of dX: a 1..8-tap stride-1 correlation over dY whose epilogue stores to 2i+p.
//  * Warp roles: 0 = activation TMA producer, 1 = weight TMA producer, 2 = MMA issuer (warp-uniform loop, one elected
//    lane issues) and TMEM allocator, 3..6 = epilogue (tcgen05.ld -> +residual -> bf16 -> 16-byte global stores, plus
//    the fused GroupNorm statistics / GroupNorm-backward reduction).  A 12-warp variant with two epilogue warpgroups on
//    alternate items and a setmaxnreg register re-partition is kept behind MMPL_TC_TWO_GROUPS (see below).
//  * WRES: for 32->32 layers all 27 weight tiles (55 KB) stay resident in shared memory for the life of the CTA.
//  * EPI: the epilogue is compiled per variant (conv_tc_problem.cuh) -- plain, forward (statistics / residual), fused
//    GroupNorm backward -- so that a launch only carries the code it executes: with a single warp per scheduler the
//    epilogue is latency bound per instruction, and one kernel holding all variants spent most of its stall samples on
//    instruction fetches.
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "conv_tc_problem.cuh"
#include "ptx.cuh"

namespace mmpl {
namespace {

using namespace ptx;

constexpr int TC_TH = 16, TC_TW = 8;
// weight handling: one tap tile per ring stage / all 27 tiles resident / the three kd tiles of a (kh, kw) per ring stage
enum : int { WM_STREAM = 0, WM_RESIDENT = 1, WM_PLANE_MAJOR = 2 };
// Optional 12-warp layout (MMPL_TC_TWO_GROUPS=1): 0 = activation TMA, 1 = weight TMA, 2 = MMA issuer, 3 = idle, 4..7 and
// 8..11 = two epilogue warpgroups that take alternate work items (one per TMEM accumulator buffer), so an epilogue has two
// MMA periods to finish.  The register file is re-partitioned with setmaxnreg: 64 for warps 0..3, 216 for the epilogue
// warpgroups (launch: 384 x 168; the decrease must free more than the increase takes, or the second group waits forever).
// MMPL_TC_TWO_GROUPS = 0 (default) keeps ONE epilogue warpgroup (warps 3..6, 7 warps, no register re-partition): measured
// on B200 the register-lean epilogue keeps up with the MMAs on its own, and the 12-warp layout costs the plain
// launches ~18 % (profiles/r01_ncu_summary.md).
#ifndef MMPL_TC_TWO_GROUPS
#define MMPL_TC_TWO_GROUPS 0
#endif
constexpr bool TC_TWO_GROUPS = MMPL_TC_TWO_GROUPS != 0;
constexpr int TC_THREADS = TC_TWO_GROUPS ? 384 : 224;
constexpr int TC_EPI0 = TC_TWO_GROUPS ? 4 : 3;      // first epilogue warp
constexpr int TC_REGS_SPECIAL = 64, TC_REGS_EPILOGUE = 216;


template <int MODE>
struct Geo {
  static constexpr int KS = (MODE == MODE_S1K3 || MODE == MODE_S2F || MODE == MODE_S2D) ? 3 : 1;
  static constexpr int EXTRA = MODE == MODE_S1K3 ? 2 : (MODE == MODE_S2F || MODE == MODE_S2D) ? 1 : 0;  // halo voxels
  static constexpr int LO = (MODE == MODE_S1K3 || MODE == MODE_S2F) ? -1 : 0;   // block origin relative to the tile
  static constexpr int PH = TC_TH + EXTRA, PW = TC_TW + EXTRA;
  static constexpr bool STRIDED_OUT = (MODE == MODE_S2D || MODE == MODE_S2K1D);
  static constexpr bool PARITY_CHUNKS = (MODE == MODE_S2F);
};

template <int KC, int NT, int TD, int MODE, int WM, int NA_>
struct TcCfg {
  static constexpr bool WRES = WM == WM_RESIDENT, WPM = WM == WM_PLANE_MAJOR;
  using G = Geo<MODE>;
  static constexpr int RB = KC * 2;
  static constexpr uint32_t SWZ = RB == 128 ? SWZ_128B : SWZ_64B;
  static constexpr int PD = TD + G::EXTRA;
  static constexpr int A_BYTES = PD * G::PH * G::PW * RB;
  static constexpr int A_STAGE = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int NA = NA_;
  static constexpr int B_BYTES = NT * RB;                        // one tap tile
  static constexpr int B_STAGE = (WPM ? 3 : 1) * B_BYTES;        // plane-major: the three kd tiles of one (kh, kw)
  static constexpr int SMEM_LIMIT = 227 * 1024 - 2048;
  static constexpr int NB_FIT = (SMEM_LIMIT - NA * A_STAGE) / B_STAGE;
  static constexpr int NB_CAP = WPM ? 4 : 8;
  static constexpr int NB = WRES ? 27 : (NB_FIT > NB_CAP ? NB_CAP : NB_FIT);
  static constexpr int ACC_COLS = TD * NT;
  static constexpr int TMEM_COLS = 2 * ACC_COLS <= 32 ? 32 : 2 * ACC_COLS <= 64 ? 64 : 2 * ACC_COLS <= 128 ? 128 : 2 * ACC_COLS <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = NA * A_STAGE + NB * B_STAGE + 1024 /*align slack*/ + 512 /*barriers*/;
  static_assert(B_BYTES % 1024 == 0, "weight stage must keep 1024-byte alignment");
  static_assert(NB >= 2, "need at least two weight stages");
  static_assert(2 * ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(WM == WM_STREAM || (MODE == MODE_S1K3), "resident / plane-major weights only for the 3x3x3 stride-1 mode");
  static_assert(!WPM || (TD >= 2 && NT * (TD >= 3 ? 3 : 2) <= 256), "plane-major issue: N = NT * planes sharing one A view <= 256");
};

struct TcParams {
  __nv_bfloat16* y;
  const __nv_bfloat16* residual;
  int N;            // batch
  int D, H, W;      // extents of the OUTPUT tensor y (full dims, also for the strided modes)
  int Ds, Hs, Ws;   // extents of the tile grid domain (== D,H,W except strided-output modes: the parity sub-grid)
  int nch;          // reduction channels / KC
  int cout_total;   // row stride of y in elements
  int DT, HT, WT, NTILES, NPAR;
  int total_items;
  double* stats;    // optional GroupNorm(16) raw sums of the OUTPUT [N][16][2] (requires cout_total == NT), else NULL
  // Optional fused first pass of the GroupNorm+ReLU backward (dgrad launches): the output of this launch is dA, the
  // gradient w.r.t. a = relu(gn(x)) = this convolution's forward input, still available as gn_a (NDHWC like y, or the
  // parity-split copy P when gn_psplit).  Because a = gamma*xhat + beta where the ReLU passes and 0 elsewhere,
  //   S1_c = sum_v g = sum_v dA*[a > 0]          and          gamma_c * sum_v g*xhat = sum_v dA*a - beta_c * S1_c,
  // so the epilogue needs one extra row read and 4 instructions per element, no per-channel constants.  It accumulates
  // S1 and Q = gamma*sum(g*xhat) into gn_ws[n][c][gn_head*2 + {0,1}] (the workspace of mmpl_gn_relu_bwd).
  const __nv_bfloat16* gn_a;
  const float* gn_beta;
  double* gn_ws;
  int gn_ws_stride;   // doubles per (n, c) entry of gn_ws
  int gn_head;
  int gn_psplit;
  int epi_groups;     // 1 or 2 epilogue warpgroups in use
};

// Tap enumeration shared by the weight producer and the MMA issuer.
// For the parity modes an axis of parity 1 sees taps {0,2}, parity 0 sees tap {1}.
__device__ __forceinline__ int axis_ntaps(int par) { return par ? 2 : 1; }
__device__ __forceinline__ int axis_tap(int par, int i) { return par ? 2 * i : 1; }

template <int KC, int NT, int TD, int MODE, int WM, int NA_, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using Cfg = TcCfg<KC, NT, TD, MODE, WM, NA_>;
  using G = Geo<MODE>;
  constexpr bool WRES = Cfg::WRES, WPM = Cfg::WPM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_stage = smem;
  uint8_t* b_stage = smem + Cfg::NA * Cfg::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_stage + Cfg::NB * Cfg::B_STAGE);
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

  // item -> (nt, parity class, n, tile origin in the tile-grid domain)
  // The parity class is the FASTEST index: the 8 classes of a stride-2 dgrad tile run on neighbouring CTAs at the same
  // time, so their interleaved 64-byte stores merge into full lines in L2 and the dY tile is fetched from HBM once.
  auto item_coords = [&](int item, int& nt, int& pc, int& n, int& d0, int& h0, int& w0) {
    pc = item % p.NPAR;
    item /= p.NPAR;
    w0 = (item % p.WT) * TC_TW;
    item /= p.WT;
    h0 = (item % p.HT) * TC_TH;
    item /= p.HT;
    d0 = (item % p.DT) * TD;
    item /= p.DT;
    n = item % p.N;
    nt = item / p.N;
  };
  // number of (A-chunk) loads per item and taps per chunk
  const int chunks_per_item = G::PARITY_CHUNKS ? p.nch * 8 : p.nch;

  if (warp < TC_EPI0) {
  if (TC_TWO_GROUPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_SPECIAL));
  if (warp == 0) {
    // ===================================================== activation producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int nt, pc, n, d0, h0, w0;
        item_coords(item, nt, pc, n, d0, h0, w0);
        for (int c = 0; c < chunks_per_item; ++c, ++it) {
          const int ch = G::PARITY_CHUNKS ? (c >> 3) : c;
          const uint32_t s = it % Cfg::NA, ph = (it / Cfg::NA) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          mbar_expect_tx(&a_full[s], Cfg::A_BYTES);
          int cw = w0 + G::LO, chh = h0 + G::LO, cd = d0 + G::LO, cn = n;
          if (MODE == MODE_S2F) cn = (c & 7) * p.N + n;          // parity plane of the split tensor
          if (MODE == MODE_S2K1F) cw = 2 * w0, chh = 2 * h0, cd = 2 * d0;
          tma_load_5d(a_stage + s * Cfg::A_STAGE, &tmA, &a_full[s], ch * KC, cw, chh, cd, cn);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== weight producer
    if (lane == 0) {
      if (WRES) {
        // all 27 tiles once; every stage has its own barrier, completed exactly once
        // slot order (kh, kw, 2-kd): for a fixed (kh,kw) the three kd tiles are consecutive rows of ONE [3*NT x KC]
        // B matrix, so a single MMA with N = 3*NT feeds the accumulators of three output planes (see the MMA issuer)
        for (int tap = 0; tap < 27; ++tap) {
          const int kd = tap / 9, khw = tap % 9, slot = khw * 3 + (2 - kd);
          mbar_expect_tx(&b_full[slot], Cfg::B_BYTES);
          tma_load_3d(b_stage + slot * Cfg::B_BYTES, &tmB, &b_full[slot], 0, 0, tap);
        }
      } else if (WPM) {
        // per (chunk, kh, kw): the three kd tiles land in ONE stage, slot order (2 - kd) like the resident layout
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
          int nt, pc, n, d0, h0, w0;
          item_coords(item, nt, pc, n, d0, h0, w0);
          for (int c = 0; c < p.nch; ++c)
            for (int khw = 0; khw < 9; ++khw, ++it) {
              const uint32_t s = it % Cfg::NB, ph = (it / Cfg::NB) & 1;
              mbar_wait(&b_empty[s], ph ^ 1);
              mbar_expect_tx(&b_full[s], Cfg::B_STAGE);
#pragma unroll
              for (int kd = 0; kd < 3; ++kd)
                tma_load_3d(b_stage + s * Cfg::B_STAGE + (2 - kd) * Cfg::B_BYTES, &tmB, &b_full[s], c * KC, nt * NT, kd * 9 + khw);
            }
        }
      } else {
        uint32_t it = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
          int nt, pc, n, d0, h0, w0;
          item_coords(item, nt, pc, n, d0, h0, w0);
          for (int c = 0; c < chunks_per_item; ++c) {
            const int ch = G::PARITY_CHUNKS ? (c >> 3) : c;
            const int par = G::PARITY_CHUNKS ? (c & 7) : pc;
            const int nd = G::KS == 1 ? 1 : (MODE == MODE_S1K3 ? 3 : axis_ntaps(par >> 2));
            const int nh = G::KS == 1 ? 1 : (MODE == MODE_S1K3 ? 3 : axis_ntaps((par >> 1) & 1));
            const int nw = G::KS == 1 ? 1 : (MODE == MODE_S1K3 ? 3 : axis_ntaps(par & 1));
            for (int id = 0; id < nd; ++id)
              for (int ih = 0; ih < nh; ++ih)
                for (int iw = 0; iw < nw; ++iw, ++it) {
                  int tap = 0;
                  if (MODE == MODE_S1K3) tap = (id * 3 + ih) * 3 + iw;
                  if (MODE == MODE_S2F || MODE == MODE_S2D)
                    tap = (axis_tap(par >> 2, id) * 3 + axis_tap((par >> 1) & 1, ih)) * 3 + axis_tap(par & 1, iw);
                  const uint32_t s = it % Cfg::NB, ph = (it / Cfg::NB) & 1;
                  mbar_wait(&b_empty[s], ph ^ 1);
                  mbar_expect_tx(&b_full[s], Cfg::B_BYTES);
                  tma_load_3d(b_stage + s * Cfg::B_BYTES, &tmB, &b_full[s], ch * KC, nt * NT, tap);
                }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================== MMA issuer: warp-uniform control flow, one lane issues
    const uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
    const uint32_t a_base = smem_u32(a_stage), b_base = smem_u32(b_stage);
    const uint64_t a_fix = make_smem_desc(0, 16, G::PW * Cfg::RB, Cfg::SWZ, 0);
    const uint64_t b_fix = make_smem_desc(0, 16, 8 * Cfg::RB, Cfg::SWZ, 0);
    const uint32_t a_hi = static_cast<uint32_t>(a_fix >> 32), a_lo_fix = static_cast<uint32_t>(a_fix);
    const uint32_t b_hi = static_cast<uint32_t>(b_fix >> 32), b_lo_fix = static_cast<uint32_t>(b_fix);
    uint32_t ita = 0, itb = 0, iti = 0;
    if (WRES) {
      for (int tap = 0; tap < 27; ++tap) mbar_wait(&b_full[tap], 0);
    }
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++iti) {
      int nt, pc, n, d0, h0, w0;
      item_coords(item, nt, pc, n, d0, h0, w0);
      const uint32_t buf = iti & 1, bph = (iti >> 1) & 1;
      mbar_wait(&acc_empty[buf], bph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * Cfg::ACC_COLS;
      if (WRES) {
        // ---- plane-major issue (resident weights, one reduction chunk): for input plane s and tap (kh,kw) the SAME A view
        // multiplies W[kd] for every output plane p = s - kd, so one MMA with N = NT * (#valid kd) accumulates into
        // the adjacent TMEM accumulators of planes p_lo..p_hi.  Halves the MMA count and the A-operand reads.
        const uint32_t sa = ita % Cfg::NA, pha = (ita / Cfg::NA) & 1;
        mbar_wait(&a_full[sa], pha);
        tc_fence_after();
        const uint32_t a_addr = a_base + sa * Cfg::A_STAGE;
        if (elect_one()) {
#pragma unroll
          for (int sp = 0; sp < TD + 2; ++sp) {
            constexpr int dummy = 0;
            (void)dummy;
            const int kd_hi = sp < 2 ? sp : 2;                       // p_lo = sp - kd_hi
            const int kd_lo = sp - (TD - 1) > 0 ? sp - (TD - 1) : 0;  // p_hi = sp - kd_lo
            const int nblk = kd_hi - kd_lo + 1;
            const bool fresh = sp < TD;                              // plane p_hi == sp gets its first contribution here
            const uint32_t d_lo = d_tmem + (sp - kd_hi) * NT;
#pragma unroll
            for (int khw = 0; khw < 9; ++khw) {
              const int kh = khw / 3, kw = khw % 3;
              const uint32_t a_lo = a_lo_fix | ((a_addr + (((sp * G::PH) + kh) * G::PW + kw) * Cfg::RB) >> 4);
              const uint32_t b_lo = b_lo_fix | ((b_base + (khw * 3 + (2 - kd_hi)) * Cfg::B_BYTES) >> 4);
#pragma unroll
              for (int ks = 0; ks < KC / 16; ++ks) {
                if (fresh && khw == 0 && ks == 0) {
                  // first touch of plane sp: older planes accumulate, the fresh one is overwritten
                  if (nblk > 1)
                    umma_f16_lohi(d_lo, a_lo, a_hi, b_lo, b_hi, make_idesc_bf16(128, NT * (nblk - 1), 0, 0), 1u);
                  umma_f16_lohi(d_lo + (nblk - 1) * NT, a_lo, a_hi, b_lo + (((nblk - 1) * Cfg::B_BYTES) >> 4), b_hi,
                                make_idesc_bf16(128, NT, 0, 0), 0u);
                } else {
                  umma_f16_lohi(d_lo, a_lo + ((ks * 32) >> 4), a_hi, b_lo + ((ks * 32) >> 4), b_hi,
                                make_idesc_bf16(128, NT * nblk, 0, 0), 1u);
                }
              }
            }
          }
          umma_commit(&a_empty[sa]);
        }
        __syncwarp();
        ++ita;
      } else if (WPM) {
        // ---- plane-major issue with streamed weights: as above, but the weight ring delivers the three kd tiles of one
        // (kh, kw) per stage, so the (kh, kw) loop is the outer one and runs per reduction chunk
        for (int c = 0; c < p.nch; ++c, ++ita) {
          const uint32_t sa = ita % Cfg::NA, pha = (ita / Cfg::NA) & 1;
          mbar_wait(&a_full[sa], pha);
          tc_fence_after();
          const uint32_t a_addr = a_base + sa * Cfg::A_STAGE;
          for (int khw = 0; khw < 9; ++khw, ++itb) {
            const uint32_t sb = itb % Cfg::NB;
            mbar_wait(&b_full[sb], (itb / Cfg::NB) & 1);
            tc_fence_after();
            const int kh = khw / 3, kw = khw - 3 * kh;
            const uint32_t a_tap = a_addr + (kh * G::PW + kw) * Cfg::RB;
            const uint32_t b_tap = b_base + sb * Cfg::B_STAGE;
            const bool first_tap = c == 0 && khw == 0;
            if (elect_one()) {
#pragma unroll
              for (int sp = 0; sp < TD + 2; ++sp) {
                const int kd_hi = sp < 2 ? sp : 2;                       // p_lo = sp - kd_hi
                const int kd_lo = sp - (TD - 1) > 0 ? sp - (TD - 1) : 0;  // p_hi = sp - kd_lo
                const int nblk = kd_hi - kd_lo + 1;
                const uint32_t d_lo = d_tmem + (sp - kd_hi) * NT;
                const uint32_t a_lo = a_lo_fix | ((a_tap + sp * (G::PH * G::PW * Cfg::RB)) >> 4);
                const uint32_t b_lo = b_lo_fix | ((b_tap + (2 - kd_hi) * Cfg::B_BYTES) >> 4);
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                  if (sp < TD && ks == 0 && first_tap) {
                    // first touch of plane sp: older planes accumulate, the fresh one is overwritten
                    if (nblk > 1)
                      umma_f16_lohi(d_lo, a_lo, a_hi, b_lo, b_hi, make_idesc_bf16(128, NT * (nblk - 1), 0, 0), 1u);
                    umma_f16_lohi(d_lo + (nblk - 1) * NT, a_lo, a_hi, b_lo + (((nblk - 1) * Cfg::B_BYTES) >> 4), b_hi,
                                  make_idesc_bf16(128, NT, 0, 0), 0u);
                  } else {
                    umma_f16_lohi(d_lo, a_lo + ((ks * 32) >> 4), a_hi, b_lo + ((ks * 32) >> 4), b_hi,
                                  make_idesc_bf16(128, NT * nblk, 0, 0), 1u);
                  }
                }
              }
              umma_commit(&b_empty[sb]);
            }
            __syncwarp();
          }
          if (elect_one()) umma_commit(&a_empty[sa]);
          __syncwarp();
        }
      } else {
      uint32_t first = 1;
      for (int c = 0; c < chunks_per_item; ++c, ++ita) {
        const int par = G::PARITY_CHUNKS ? (c & 7) : pc;
        const uint32_t sa = ita % Cfg::NA, pha = (ita / Cfg::NA) & 1;
        mbar_wait(&a_full[sa], pha);
        tc_fence_after();
        const uint32_t a_addr = a_base + sa * Cfg::A_STAGE;
        const int nd = G::KS == 1 ? 1 : (MODE == MODE_S1K3 ? 3 : axis_ntaps(par >> 2));
        const int nh = G::KS == 1 ? 1 : (MODE == MODE_S1K3 ? 3 : axis_ntaps((par >> 1) & 1));
        const int nw = G::KS == 1 ? 1 : (MODE == MODE_S1K3 ? 3 : axis_ntaps(par & 1));
        for (int id = 0; id < nd; ++id)
          for (int ih = 0; ih < nh; ++ih)
            for (int iw = 0; iw < nw; ++iw, ++itb) {
              // row shift of this tap inside the halo block
              int sd = 0, sh = 0, sw = 0, tap = 0;
              if (MODE == MODE_S1K3) sd = id, sh = ih, sw = iw, tap = (id * 3 + ih) * 3 + iw;
              if (MODE == MODE_S2F) {   // tap k of parity p: P_p[o + (k != 0) - 1], block origin is o0 - 1
                sd = axis_tap(par >> 2, id) != 0, sh = axis_tap((par >> 1) & 1, ih) != 0, sw = axis_tap(par & 1, iw) != 0;
              }
              if (MODE == MODE_S2D) {   // tap t' of parity p: dY[i' + (t' == 2)], block origin is i0'
                sd = axis_tap(par >> 2, id) == 2, sh = axis_tap((par >> 1) & 1, ih) == 2, sw = axis_tap(par & 1, iw) == 2;
              }
              uint32_t sb;
              if (WRES) {
                sb = tap;
              } else {
                sb = itb % Cfg::NB;
                mbar_wait(&b_full[sb], (itb / Cfg::NB) & 1);
                tc_fence_after();
              }
              // start-address fields (>>4) of this tap; every MMA below adds a compile-time constant
              const uint32_t b_lo = b_lo_fix | ((b_base + sb * Cfg::B_BYTES) >> 4);
              const uint32_t a_lo = a_lo_fix | ((a_addr + ((sd * G::PH + sh) * G::PW + sw) * Cfg::RB) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int pl = 0; pl < TD; ++pl) {
#pragma unroll
                  for (int ks = 0; ks < KC / 16; ++ks) {
                    umma_f16_lohi(d_tmem + pl * NT, a_lo + ((pl * (G::PH * G::PW * Cfg::RB) + ks * 32) >> 4), a_hi,
                                  b_lo + ((ks * 32) >> 4), b_hi, idesc, (first && ks == 0) ? 0u : 1u);
                  }
                }
                if (!WRES) umma_commit(&b_empty[sb]);
              }
              __syncwarp();
              first = 0;
            }
        if (elect_one()) umma_commit(&a_empty[sa]);
        __syncwarp();
      }
      }
      if (elect_one()) umma_commit(&acc_full[buf]);
      __syncwarp();
    }
  }
  } else {
    // ===================================================== epilogue warpgroups (TMEM lane quarter = warp % 4)
    if (TC_TWO_GROUPS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_EPILOGUE));
    const int q = warp & 3;
    // epilogue group: with p.epi_groups == 2 group g owns accumulator buffer g and the items of parity g; with 1 the
    // first group takes every item (plain launches: the second group would only add scheduler pressure)
    const int eg = (warp - TC_EPI0) >> 2;
    const int epi_groups = TC_TWO_GROUPS ? p.epi_groups : 1;
    if (eg < epi_groups) {
    const int row = q * 32 + lane;
    const int rh = row >> 3, rw = row & 7;
    // fused GroupNorm statistics of the stored output: 16 groups of NT/16 channels; per-thread fp32 partials over this
    // CTA's items, reduced across the warp and added to the fp64 buffer only when the sample changes / at the end
    constexpr int CPG = NT / 16;
    // p1/p2: per-thread partials of the fused GroupNorm-backward reduction (dgrad launches).  The forward statistics
    // (fprop launches) never coexist with them and live in the first half of p1: gsum = p1[0..15], gsq = p1[16..31].
    constexpr bool E_FWD = EPI == EPI_FWD, E_GN = EPI == EPI_GN;
    float p1[(E_FWD || E_GN) ? 32 : 1], p2[E_GN ? 32 : 1];
#pragma unroll
    for (int j = 0; j < ((E_FWD || E_GN) ? 32 : 1); ++j) p1[j] = 0.f;
#pragma unroll
    for (int j = 0; j < (E_GN ? 32 : 1); ++j) p2[j] = 0.f;
    float (&gstat)[(E_FWD || E_GN) ? 32 : 1] = p1;
    const bool st_on = E_FWD && p.stats != nullptr;
    int stat_n = -1;
    auto flush_stats = [&]() {
      if (!E_FWD) return;
      if (!st_on || stat_n < 0) return;
#pragma unroll
      for (int g = 0; g < 16; ++g) {
        const float a = warp_sum(gstat[g]), b = warp_sum(gstat[16 + g]);
        if (lane == 0) {
          atomicAdd(&p.stats[(static_cast<int64_t>(stat_n) * 16 + g) * 2 + 0], static_cast<double>(a));
          atomicAdd(&p.stats[(static_cast<int64_t>(stat_n) * 16 + g) * 2 + 1], static_cast<double>(b));
        }
        gstat[g] = gstat[16 + g] = 0.f;
      }
    };
    // ---- fused GroupNorm+ReLU backward reduction (see TcParams::gn_*)
    // p1[j] / p2[j]: this thread's (= output row's) partial S1 and sum dA*a for column j of the current chunk.  Summing over
    // the warp's 32 rows is a 31-shuffle transpose-reduce; with a single column chunk (NT == 32) the partials simply
    // keep accumulating over all items of the CTA and are reduced once per sample, otherwise once per item and chunk.
    constexpr bool gn_on = E_GN;            // the GN variant is only launched with a workspace (launch_tc)
    constexpr bool GN_PERSIST = NT == 32;
    float gacc1[NT / 32], gacc2[NT / 32];
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) gacc1[i] = gacc2[i] = 0.f;
    int gn_n = -1, gn_nt = -1;
    // sum over the warp's 32 rows of each of 32 per-lane values; afterwards v[0] of lane l is the total of index l
    auto transpose_reduce = [&](auto& v) {
      if (!E_GN) return;
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) {
        const bool up = (lane & sft) != 0;
#pragma unroll
        for (int i = 0; i < sft; ++i) {
          const float send = up ? v[i] : v[i + sft];
          const float keep = up ? v[i + sft] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
        }
      }
    };
    auto gn_flush = [&]() {
      if (!E_GN) return;
      if (gn_n < 0) return;
      if (GN_PERSIST) {
        transpose_reduce(p1);
        transpose_reduce(p2);
        gacc1[0] = p1[0], gacc2[0] = p2[0];
#pragma unroll
        for (int j = 0; j < 32; ++j) p1[j] = p2[j] = 0.f;
      }
#pragma unroll
      for (int ci = 0; ci < NT / 32; ++ci) {
        const int c = gn_nt * NT + ci * 32 + lane;
        double* w = p.gn_ws + (static_cast<int64_t>(gn_n) * p.cout_total + c) * p.gn_ws_stride + p.gn_head * 2;
        const double s1 = static_cast<double>(gacc1[ci]);
        atomicAdd(w, s1);
        atomicAdd(w + 1, static_cast<double>(gacc2[ci]) - static_cast<double>(p.gn_beta[c]) * s1);
        gacc1[ci] = gacc2[ci] = 0.f;
      }
    };
    // The epilogue reads one auxiliary row per output row (the residual of conv2, or the activation for the fused
    // GroupNorm-backward reduction).  With a single epilogue warp per scheduler a ~1 us load latency per (chunk, plane)
    // step would be fully exposed, so the rows run through a 2-step register ring that is kept two steps AHEAD of the
    // arithmetic across item boundaries (the coordinates of the next item are known in advance).
    constexpr int TOT = TD * (NT / 32);          // (column chunk, plane) steps per item, chunk-major
    static_assert(TOT % 2 == 0, "the 2-deep ring needs an even number of steps per item");
    const __nv_bfloat16* aux = E_GN ? p.gn_a : (E_FWD ? p.residual : nullptr);
    struct ItemPos {
      int nt, pc, n, d0, h0, w0;
      bool ok;
    };
    auto locate = [&](int item) {
      ItemPos t;
      t.ok = item < p.total_items;
      if (t.ok) item_coords(item, t.nt, t.pc, t.n, t.d0, t.h0, t.w0);
      return t;
    };
    // element offset of step i of an item in y (and in a same-layout aux tensor); valid = row inside the tensor
    auto step_off = [&](const ItemPos& t, int i, bool& valid) -> int64_t {
      const int c0 = (i / TD) * 32, pl = i % TD;
      int dd = t.d0 + pl, hh = t.h0 + rh, ww = t.w0 + rw;
      if (G::STRIDED_OUT) dd = 2 * dd + (t.pc >> 2), hh = 2 * hh + ((t.pc >> 1) & 1), ww = 2 * ww + (t.pc & 1);
      valid = t.ok && hh < p.H && ww < p.W && dd < p.D;
      return ((((static_cast<int64_t>(t.n) * p.D + dd) * p.H + hh) * p.W + ww) * p.cout_total) + t.nt * NT + c0;
    };
    uint4 ring[2][4];
    auto prefetch = [&](const ItemPos& t, int i, uint4 (&dst)[4]) {
      if (EPI == EPI_PLAIN) return;
      bool valid;
      int64_t off = step_off(t, i, valid);
      if (G::STRIDED_OUT && gn_on && p.gn_psplit)   // a lives in the parity-split copy: plane (pc, n), sub-grid coordinates
        off = ((((static_cast<int64_t>(t.pc) * p.N + t.n) * p.Ds + t.d0 + (i % TD)) * p.Hs + t.h0 + rh) * p.Ws + t.w0 + rw) *
                  p.cout_total + t.nt * NT + (i / TD) * 32;
#pragma unroll
      for (int v = 0; v < 4; ++v)
        dst[v] = (aux != nullptr && valid) ? *reinterpret_cast<const uint4*>(aux + off + v * 8) : make_uint4(0u, 0u, 0u, 0u);
    };
    const int item_step = epi_groups * gridDim.x;
    ItemPos cur = locate(blockIdx.x + eg * gridDim.x);
    prefetch(cur, 0, ring[0]);
    prefetch(cur, 1, ring[1]);
    uint32_t iti = eg;
    for (int item = blockIdx.x + eg * gridDim.x; item < p.total_items; item += item_step, iti += epi_groups) {
      const ItemPos nxt = locate(item + item_step);
      const int nt = cur.nt, n = cur.n;
      if (st_on && n != stat_n) {
        flush_stats();
        stat_n = n;
      }
      if (gn_on && (n != gn_n || nt != gn_nt)) {
        gn_flush();
        gn_n = n, gn_nt = nt;
      }
      const uint32_t buf = iti & 1, bph = (iti >> 1) & 1;
      mbar_wait(&acc_full[buf], bph);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < TOT; ++i) {
        const int c0 = (i / TD) * 32, pl = i % TD;
        if (gn_on && !GN_PERSIST && pl == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) p1[j] = p2[j] = 0.f;
        }
        bool valid;
        const int64_t off = step_off(cur, i, valid);
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg::ACC_COLS + pl * NT + c0, r);
        tmem_ld_wait();
        uint4 (&row)[4] = ring[i % 2];
        if (valid) {
          if (E_FWD && p.residual) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const uint4 rv = row[v];
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
            const uint32_t xu[4] = {row[v].x, row[v].y, row[v].z, row[v].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float e0 = __uint_as_float(r[v * 8 + 2 * k]), e1 = __uint_as_float(r[v * 8 + 2 * k + 1]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
              o[k] = *reinterpret_cast<uint32_t*>(&h2);
              if (st_on) {   // statistics of the value as stored (bf16-rounded)
                const float x0 = __uint_as_float(o[k] << 16), x1 = __uint_as_float(o[k] & 0xFFFF0000u);
                const int g0 = (c0 + v * 8 + 2 * k) / CPG, g1 = (c0 + v * 8 + 2 * k + 1) / CPG;
                gstat[g0] += x0;
                gstat[16 + g0] = fmaf(x0, x0, gstat[16 + g0]);
                gstat[g1] += x1;
                gstat[16 + g1] = fmaf(x1, x1, gstat[16 + g1]);
              }
              if (gn_on) {   // dA (fp32, before the bf16 rounding of the store) against a = relu(gn(x)): gate = [a > 0]
                const int j = v * 8 + 2 * k;
                const float a0 = __uint_as_float(xu[k] << 16), a1 = __uint_as_float(xu[k] & 0xFFFF0000u);
                p2[j] = fmaf(e0, a0, p2[j]);
                p2[j + 1] = fmaf(e1, a1, p2[j + 1]);
                if (a0 > 0.f) p1[j] += e0;
                if (a1 > 0.f) p1[j + 1] += e1;
              }
            }
            *reinterpret_cast<uint4*>(p.y + off + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        if (i + 2 < TOT)          // the slot is consumed: refill it two steps ahead
          prefetch(cur, i + 2, ring[i % 2]);
        else
          prefetch(nxt, i + 2 - TOT, ring[i % 2]);
        if (gn_on && !GN_PERSIST && pl == TD - 1) {
          transpose_reduce(p1);
          transpose_reduce(p2);
          gacc1[c0 / 32] += p1[0];
          gacc2[c0 / 32] += p2[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      cur = nxt;
    }
    gn_flush();
    flush_stats();
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
  bind_primary_context();
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 5-D map over an NDHWC bf16 tensor [N][D][H][W][C]; box counts are OUTPUT elements, estride the traversal stride.
int make_act_map(CUtensorMap* m, const void* ptr, int64_t N, int D, int H, int W, int C, int kc, int pd, int ph, int pw,
                 int estride) {
  EncodeTiledFn enc = get_encode();
  MMPL_REQUIRE(enc != nullptr, MMPL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  const cuuint32_t e = static_cast<cuuint32_t>(estride);
  cuuint64_t gd[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gs[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)D * H * W * C * 2};
  cuuint32_t bx[5] = {(cuuint32_t)kc, (cuuint32_t)pw * e, (cuuint32_t)ph * e, (cuuint32_t)pd * e, 1};
  cuuint32_t es[5] = {1, e, e, e, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMPL_REQUIRE(r == CUDA_SUCCESS, MMPL_E_CUDA, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  return MMPL_OK;
}

int make_weight_map(CUtensorMap* m, const void* ptr, int taps, int cout, int cin, int kc, int nt) {
  EncodeTiledFn enc = get_encode();
  MMPL_REQUIRE(enc != nullptr, MMPL_E_CUDA, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gd[3] = {(cuuint64_t)cin, (cuuint64_t)cout, (cuuint64_t)taps};
  cuuint64_t gs[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cout * cin * 2};
  cuuint32_t bx[3] = {(cuuint32_t)kc, (cuuint32_t)nt, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MMPL_REQUIRE(r == CUDA_SUCCESS, MMPL_E_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return MMPL_OK;
}

template <int KC, int NT, int TD, int MODE, int WM, int NA_, int EPI>
int launch_tc(const TcProblem& q, cudaStream_t s) {
  static_assert(EPI == EPI_PLAIN || EPI == EPI_FWD || EPI == EPI_GN, "epilogue variant");
  MMPL_REQUIRE((EPI == EPI_GN) == (q.gn != nullptr), MMPL_E_UNSUPPORTED, "conv_tc: epilogue variant %d vs fused GroupNorm backward", EPI);
  MMPL_REQUIRE(EPI == EPI_FWD || q.residual == nullptr, MMPL_E_UNSUPPORTED, "conv_tc: a residual needs the forward epilogue");
  using Cfg = TcCfg<KC, NT, TD, MODE, WM, NA_>;
  using G = Geo<MODE>;
  CUtensorMap tmA, tmB;
  if (int e = make_act_map(&tmA, q.a, q.aN, q.aD, q.aH, q.aW, q.kred, KC, Cfg::PD, G::PH, G::PW, MODE == MODE_S2K1F ? 2 : 1)) return e;
  if (int e = make_weight_map(&tmB, q.wp, G::KS * G::KS * G::KS, q.nout, q.kred, KC, NT)) return e;
  TcParams p;
  p.y = static_cast<__nv_bfloat16*>(q.y);
  p.residual = static_cast<const __nv_bfloat16*>(q.residual);
  p.N = q.N, p.D = q.D, p.H = q.H, p.W = q.W;
  p.Ds = G::STRIDED_OUT ? (q.D + 1) / 2 : q.D;
  p.Hs = G::STRIDED_OUT ? (q.H + 1) / 2 : q.H;
  p.Ws = G::STRIDED_OUT ? (q.W + 1) / 2 : q.W;
  p.nch = q.kred / KC;
  p.cout_total = q.nout;
  p.DT = ceil_div(p.Ds, TD), p.HT = ceil_div(p.Hs, TC_TH), p.WT = ceil_div(p.Ws, TC_TW), p.NTILES = q.nout / NT;
  p.NPAR = MODE == MODE_S2D ? 8 : 1;
  p.stats = (EPI == EPI_FWD && q.stats != nullptr && q.nout == NT && !G::STRIDED_OUT) ? q.stats : nullptr;
  if (q.stats_fused) *q.stats_fused = p.stats != nullptr;
  p.gn_a = nullptr, p.gn_beta = nullptr, p.gn_ws = nullptr, p.gn_ws_stride = 6, p.gn_head = 0, p.gn_psplit = 0;
  // two epilogue groups pay off when the epilogue has extra work per row (residual read, fused GroupNorm reduction)
  static const int force_groups = [] { const char* e = getenv("MMPL_TC_EPI_GROUPS"); return e ? atoi(e) : 0; }();
  p.epi_groups = force_groups == 1 || force_groups == 2 ? force_groups : ((q.gn != nullptr || q.residual != nullptr) ? 2 : 1);
  if (q.gn != nullptr) {
    MMPL_REQUIRE(!q.gn->a_is_parity_split || MODE == MODE_S2D, MMPL_E_UNSUPPORTED,
                 "conv_tc: a parity-split activation only pairs with the stride-2 3x3x3 dgrad");
    p.gn_a = static_cast<const __nv_bfloat16*>(q.gn->a);
    p.gn_beta = q.gn->beta, p.gn_ws = q.gn->ws, p.gn_head = q.gn->head, p.gn_psplit = q.gn->a_is_parity_split;
  }
  const int64_t items = static_cast<int64_t>(p.NTILES) * p.NPAR * q.N * p.DT * p.HT * p.WT;
  MMPL_REQUIRE(items < (1ll << 31), MMPL_E_SHAPE, "conv_tc: too many work items");
  p.total_items = static_cast<int>(items);
  // the attribute is per device: one flag per device ordinal for every instantiation
  static bool attr_set_dev[64] = {};
  int dev_ord = 0;
  cudaGetDevice(&dev_ord);
  bool& attr_set = attr_set_dev[dev_ord & 63];
  if (!attr_set) {
    MMPL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KC, NT, TD, MODE, WM, NA_, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = static_cast<int>(std::min<int64_t>(items, num_sms()));
  conv_tc_kernel<KC, NT, TD, MODE, WM, NA_, EPI><<<grid, TC_THREADS, Cfg::SMEM_BYTES, s>>>(tmA, tmB, p);
  MMPL_CHECK_LAUNCH("conv_tc");
  return MMPL_OK;
}

}  // namespace

namespace {
inline bool plane_major() {
  static const bool on = [] { const char* e = getenv("MMPL_TC_PLANE_MAJOR"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace

template <int MODE, int EPI>
int dispatch_tc(const TcProblem& q, cudaStream_t s) {
  const int kred = q.kred, nout = q.nout;
  MMPL_REQUIRE(kred == 32 || kred % 64 == 0, MMPL_E_UNSUPPORTED, "conv_tc: reduction channels %d (32 or a multiple of 64)", kred);
  int nt = nout;
  if (nout > 256) {
    MMPL_REQUIRE(nout % 256 == 0, MMPL_E_UNSUPPORTED, "conv_tc: output channels %d", nout);
    nt = 256;
  }
  MMPL_REQUIRE(nt == 32 || nt == 64 || nt == 128 || nt == 256, MMPL_E_UNSUPPORTED, "conv_tc: output channels %d", nout);
  if (kred == 32) {
    if (nt == 32) return launch_tc<32, 32, 4, MODE, MODE == MODE_S1K3 ? WM_RESIDENT : WM_STREAM, 2, EPI>(q, s);
    if constexpr (MODE == MODE_S1K3) {
      if (nt == 64 && plane_major()) return launch_tc<32, 64, 4, MODE, WM_PLANE_MAJOR, 2, EPI>(q, s);
    }
    if (nt == 64) return launch_tc<32, 64, 4, MODE, WM_STREAM, 2, EPI>(q, s);
    MMPL_FAIL(MMPL_E_UNSUPPORTED, "conv_tc: 32 reduction channels with %d output channels", nout);
  }
  if constexpr (MODE == MODE_S1K3) {
    if (nt >= 128) {
      // Lowest-resolution layers (cfg2: 2 x 4 x 12 x 12 voxels, 256 channels): a 128-row x 256-column tiling yields 16
      // work items for 148 SMs.  Narrow column tiles and single planes give 4-8x the items; the extra A-operand traffic
      // is irrelevant at this size.  Two activation stages: with one plane per item the MMAs of a 64-channel chunk take
      // about as long as the TMA round trip of the next chunk.
      const int64_t sp = static_cast<int64_t>(q.N) * ceil_div(q.H, TC_TH) * ceil_div(q.W, TC_TW);
      const int64_t items_default = sp * ceil_div(q.D, nt == 128 ? 2 : 1) * (nout / nt);
      if (items_default * 2 <= num_sms()) return launch_tc<64, 64, 1, MODE, WM_STREAM, 2, EPI>(q, s);
    }
    // Streamed-weight 64-channel configs: ONE activation stage, more planes per item and a deeper weight ring (the weight
    // tiles, one TMA round trip per tap, are the latency-critical stream).
    // Plane-major issue (like the resident-weight kernel): one A view per input plane and (kh, kw) feeds the accumulators of
    // up to three output planes (N = 2-3 x NT), halving the A-operand reads -- the tcgen05 operand fetch from shared memory
    // (measured ~85 B/clk/SM here) is what bounds these kernels, not the MMA rate.
    if (plane_major()) {
      if (nt == 32) return launch_tc<64, 32, 4, MODE, WM_PLANE_MAJOR, 1, EPI>(q, s);
      if (nt == 64) return launch_tc<64, 64, 4, MODE, WM_PLANE_MAJOR, 1, EPI>(q, s);
      if (nt == 128) return launch_tc<64, 128, 2, MODE, WM_PLANE_MAJOR, 1, EPI>(q, s);
    }
    if (nt == 32) return launch_tc<64, 32, 4, MODE, WM_STREAM, 1, EPI>(q, s);
    if (nt == 64) return launch_tc<64, 64, 4, MODE, WM_STREAM, 1, EPI>(q, s);
    if (nt == 128) return launch_tc<64, 128, 2, MODE, WM_STREAM, 1, EPI>(q, s);
    return launch_tc<64, 256, 1, MODE, WM_STREAM, 1, EPI>(q, s);
  }
  if constexpr (MODE == MODE_S2D) {
    // stride-2 dgrad: every parity class re-reads the dY halo block, so deeper tiles (4 planes, one activation stage)
    // cut the L2->SMEM traffic per output voxel
    if (nt == 32) return launch_tc<64, 32, 4, MODE, WM_STREAM, 1, EPI>(q, s);
    if (nt == 64) return launch_tc<64, 64, 4, MODE, WM_STREAM, 1, EPI>(q, s);
  }
  if (nt == 32) return launch_tc<64, 32, 2, MODE, WM_STREAM, 2, EPI>(q, s);
  if (nt == 64) return launch_tc<64, 64, 2, MODE, WM_STREAM, 2, EPI>(q, s);
  if (nt == 128) return launch_tc<64, 128, 2, MODE, WM_STREAM, 2, EPI>(q, s);
  return launch_tc<64, 256, 1, MODE, WM_STREAM, 2, EPI>(q, s);
}

}  // namespace mmpl
