// Fused partial-label loss (Dice + class-gated BCE on softmax probabilities), forward and backward.
// Reference: EDiceLoss_partial.forward (loss_functions/loss_partial.py:71-99) -> DiceLoss.forward (:38-57) and
// _dice_loss (:24-36), BCELoss per class (:90-92); class weights are mask[0] (:87,:92).
//
// Forward = ONE pass over logits/target producing four per-class sums (fp64 across threads):
//   I_c = sum p_c t_c,  Z_c = sum p_c^2,  Y_c = sum t_c,  E_c = sum -[t_c max(log p_c,-100) + (1-t_c) max(log(1-p_c),-100)]
//   L   = (1/C) sum_c w_c (1 - (2 I_c + s)/(Z_c + Y_c + s)) + sum_c w_c E_c / N_v,   s = 1e-5
// Backward = one more pass (closed form):
//   g_c  = (w_c/C)(-2 t_c/D_c + 2 p_c (2 I_c + s)/D_c^2) + (w_c/N_v)(p_c - t_c)/max(p_c(1-p_c), 1e-12),  D_c = Z_c+Y_c+s
//   dz_c = p_c (g_c - sum_k g_k p_k) * grad_out
// No one-hot tensor, no host synchronisation (the reference does 16 .item() syncs, loss_partial.py:55).
// Algorithmic HBM bytes per voxel (fp32 logits, fp32 labels): fwd 4C+4, bwd 8C+4.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "loss_math.cuh"
#include "ptx.cuh"

namespace mmpl {
namespace {

constexpr int kThreads = 256;

// softmax over the class axis, in registers.  exp(z - mx) as one FMA + one MUFU: 2^(z*log2(e) - mx*log2(e)),
// ex2.approx (max rel. error 2^-22; the arguments are <= 0 so flush-to-zero only affects probabilities below 1e-38).
template <int MAXC>
__device__ __forceinline__ void softmax_inplace(float (&p)[MAXC], int C) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) mx = fmaxf(mx, c < C ? p[c] : -INFINITY);
  constexpr float kLog2e = 1.4426950408889634f;
  const float mxs = mx * kLog2e;
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    p[c] = c < C ? ex2_approx(fmaf(p[c], kLog2e, -mxs)) : 0.f;
    sum += p[c];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) p[c] *= inv;
}

// VEC consecutive voxels of one class plane / of the label volume.  VEC = 4 needs S % 4 == 0 and 16-byte aligned
// planes (checked by the host); the loads are then 16 bytes per thread and class plane -- four times the bytes in flight
// of the scalar form, which is what an HBM-bound kernel with a 16-deep dependent softmax needs to cover the latency.
template <int VEC>
struct VoxVec {
  float v[VEC];
  __device__ __forceinline__ void load(const float* p) {
    if (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = t.x, v[1 % VEC] = t.y, v[2 % VEC] = t.z, v[3 % VEC] = t.w;
    } else {
      v[0] = __ldg(p);
    }
  }
  __device__ __forceinline__ void store(float* p) const {
    if (VEC == 4)
      *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    else
      *p = v[0];
  }
};

// labels: fp32 class ids like the reference's label tensors (train_amos_atlas_final.py:214) or uint8 class ids
template <int VEC, bool TU8>
__device__ __forceinline__ void load_labels(const void* target, int64_t idx, float (&t)[VEC]) {
  if (TU8) {
    const uint8_t* p = static_cast<const uint8_t*>(target) + idx;
    if (VEC == 4) {
      const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
#pragma unroll
      for (int i = 0; i < VEC; ++i) t[i] = static_cast<float>((w >> (8 * i)) & 0xFFu);
    } else {
      t[0] = static_cast<float>(__ldg(p));
    }
  } else {
    VoxVec<VEC> tv;
    tv.load(static_cast<const float*>(target) + idx);
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = tv.v[i];
  }
}

// Sums layout: double [G][4][C] (G = N in per-sample mode, else 1) followed by the ticket slot.
template <int MAXC, int NTHR, int VEC, bool TU8>
__global__ void __launch_bounds__(NTHR, (MAXC <= 16 && VEC == 1) ? 2 : 1)
partial_loss_fwd_kernel(const float* __restrict__ logits, const void* __restrict__ target,
                        const float* __restrict__ cw, const float* __restrict__ lut, double* __restrict__ sums,
                        float* __restrict__ loss, unsigned int* __restrict__ ticket, int N, int64_t S, int C, int uce,
                        int per_sample) {
  __shared__ float s_lut[MAXC];
  __shared__ float s_part[NTHR / 32][4 * MAXC];
  __shared__ bool s_last;
  __shared__ unsigned int s_ce_mask;   // classes whose BCE term is needed (weight != 0)
  // I_c and Y_c only ever receive the voxel's own class: per-thread accumulators in shared memory ([class][thread],
  // conflict-free) take a dynamic class index, which registers cannot; Z_c and E_c stay in registers.  One slot more
  // than MAXC collects voxels whose label is not a class id.
  __shared__ float s_I[MAXC + 1][NTHR], s_Y[MAXC + 1][NTHR];
  const int64_t n = blockIdx.y;                  // grid = (blocks per sample, N): no 64-bit division per voxel
  const int grp = per_sample ? static_cast<int>(n) : 0;
  const float* cw_g = cw + static_cast<int64_t>(grp) * C;
  const float* lut_g = lut ? lut + static_cast<int64_t>(grp) * C : nullptr;
  if (threadIdx.x < MAXC) s_lut[threadIdx.x] = (lut_g && threadIdx.x < C) ? lut_g[threadIdx.x] : static_cast<float>(threadIdx.x);
  if (threadIdx.x == 0) {
    unsigned int m = 0;
    for (int c = 0; c < C; ++c)
      if (uce && cw_g[c] != 0.f) m |= 1u << c;
    s_ce_mask = m;
  }
#pragma unroll
  for (int c = 0; c <= MAXC; ++c) s_I[c][threadIdx.x] = s_Y[c][threadIdx.x] = 0.f;
  __syncthreads();
  const unsigned int ce_mask = s_ce_mask;
  float aZ[MAXC], aE[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) aZ[c] = aE[c] = 0.f;
  const float* zb = logits + n * C * S;
  const int64_t SV = S / VEC;
  for (int64_t sv = blockIdx.x * static_cast<int64_t>(NTHR) + threadIdx.x; sv < SV;
       sv += static_cast<int64_t>(gridDim.x) * NTHR) {
    const int64_t s = sv * VEC;
    VoxVec<VEC> z[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) z[c].load(zb + c * S + s);
    float tv[VEC];
    load_labels<VEC, TU8>(target, n * S + s, tv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float p[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) p[c] = c < C ? z[c].v[i] : 0.f;
      softmax_inplace<MAXC>(p, C);
      const int tc = class_of(tv[i], lut_g ? s_lut : nullptr, C);
      float ptc = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const bool t = (c == tc);
          ptc = t ? p[c] : ptc;
          aZ[c] = fmaf(p[c], p[c], aZ[c]);
          if (ce_mask & (1u << c)) {
            // nn.BCELoss semantics: log() of the fp32 probability, clamped at -100
            const float l = t ? logf(p[c]) : logf(1.0f - p[c]);
            aE[c] -= fmaxf(l, -100.f);
          }
        }
      }
      const int slot = tc >= 0 ? tc : MAXC;
      s_I[slot][threadIdx.x] += ptc;
      s_Y[slot][threadIdx.x] += 1.f;
    }
  }
  float aI[MAXC], aY[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) aI[c] = s_I[c][threadIdx.x], aY[c] = s_Y[c][threadIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const float a = warp_sum(aI[c]), b = warp_sum(aZ[c]), d = warp_sum(aY[c]), e = warp_sum(aE[c]);
    if (lane == 0) {
      s_part[warp][c] = a;
      s_part[warp][MAXC + c] = b;
      s_part[warp][2 * MAXC + c] = d;
      s_part[warp][3 * MAXC + c] = e;
    }
  }
  __syncthreads();
  double* sums_g = sums + static_cast<int64_t>(grp) * 4 * C;
  if (threadIdx.x < 4 * MAXC) {
    const int k = threadIdx.x / MAXC, c = threadIdx.x % MAXC;
    if (c < C) {
      double t = 0;
      for (int w = 0; w < NTHR / 32; ++w) t += static_cast<double>(s_part[w][threadIdx.x]);
      atomicAdd(&sums_g[k * C + c], t);
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

template <int MAXC, int VEC, bool TU8>
__global__ void __launch_bounds__(kThreads, (MAXC <= 16 && VEC == 1) ? 2 : 1)
partial_loss_bwd_kernel(const float* __restrict__ logits, const void* __restrict__ target,
                        const float* __restrict__ cw, const float* __restrict__ lut, const double* __restrict__ sums,
                        const float* __restrict__ grad_out, float* __restrict__ dlogits, int N, int64_t S, int C,
                        int uce, int per_sample) {
  __shared__ float s_lut[MAXC], s_a[MAXC], s_b[MAXC], s_e[MAXC];
  const int64_t n = blockIdx.y;
  const int grp = per_sample ? static_cast<int>(n) : 0;
  const float* lut_g = lut ? lut + static_cast<int64_t>(grp) * C : nullptr;
  if (threadIdx.x < MAXC) {
    const int c = threadIdx.x;
    s_lut[c] = (lut_g && c < C) ? lut_g[c] : static_cast<float>(c);
    if (c < C) {
      const double* sg = sums + static_cast<int64_t>(grp) * 4 * C;
      const double sm = 1e-5, I = sg[c], Z = sg[C + c], Y = sg[2 * C + c], w = cw[grp * C + c];
      const double go = static_cast<double>(*grad_out) / (per_sample ? N : 1);
      const double nv = static_cast<double>(per_sample ? S : static_cast<int64_t>(N) * S);   // voxels in the BCE mean
      const double Dc = Z + Y + sm;
      s_a[c] = static_cast<float>(go * (w / C) * (-2.0 / Dc));
      s_b[c] = static_cast<float>(go * (w / C) * 2.0 * (2.0 * I + sm) / (Dc * Dc));
      s_e[c] = uce ? static_cast<float>(go * w / nv) : 0.f;
    } else {
      s_a[c] = s_b[c] = s_e[c] = 0.f;
    }
  }
  __syncthreads();
  const float* zb = logits + n * C * S;
  float* db = dlogits + n * C * S;
  const int64_t SV = S / VEC;
  for (int64_t sv = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; sv < SV;
       sv += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int64_t s = sv * VEC;
    VoxVec<VEC> z[MAXC];
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) z[c].load(zb + c * S + s);
    float tv[VEC];
    load_labels<VEC, TU8>(target, n * S + s, tv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float p[MAXC];
#pragma unroll
      for (int c = 0; c < MAXC; ++c) p[c] = c < C ? z[c].v[i] : 0.f;
      softmax_inplace<MAXC>(p, C);
      const int tc = class_of(tv[i], lut_g ? s_lut : nullptr, C);
      // g_c is cheap: evaluate it twice (once for the dot product, once for the output) instead of keeping 16 more
      // registers live
      auto gfun = [&](int c) {
        const float t = (c == tc) ? 1.f : 0.f;
        float gc = t * s_a[c] + p[c] * s_b[c];
        if (s_e[c] != 0.f) gc += s_e[c] * (p[c] - t) / fmaxf(p[c] * (1.0f - p[c]), 1e-12f);   // warp-uniform branch
        return gc;
      };
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) dot = fmaf(gfun(c), p[c], dot);
#pragma unroll
      for (int c = 0; c < MAXC; ++c) z[c].v[i] = c < C ? p[c] * (gfun(c) - dot) : 0.f;
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c)
      if (c < C) z[c].store(db + c * S + s);
  }
}

// ------------------------------------------------------------------------------------------------ staged kernels
// A producer warp streams [C planes x V voxels] + labels into shared memory with 1-D bulk-async copies (TMA engine,
// mbarrier completion) three stages ahead, and eight consumer warps read their voxels from shared memory: the loads no
// longer occupy registers, issue slots or the load/store unit of the warps that do the arithmetic.  Requirements:
// S % 16 == 0 and 16-byte aligned planes (else the register kernels run).  See the dispatch for which direction uses it.
constexpr int LS_V = 512;                 // voxels per stage
constexpr int LS_NS = 3;                  // stages
constexpr int LS_CONSUMERS = 256;         // 8 consumer warps; warp 8 is the producer
constexpr int LS_THREADS = LS_CONSUMERS + 32;

template <int MAXC>
struct LossStage {
  float z[MAXC][LS_V];
  float t[LS_V];                          // fp32 labels, or LS_V bytes of uint8 labels in the first quarter
};

template <int MAXC>
__device__ __forceinline__ void loss_produce(LossStage<MAXC>* stages, uint64_t* full, uint64_t* empty, const float* zb,
                                             const void* tb, bool tu8, int64_t S, int C, int chunk0, int chunk_step,
                                             int nchunks) {
  uint32_t it = 0;
  for (int ch = chunk0; ch < nchunks; ch += chunk_step, ++it) {
    const uint32_t s = it % LS_NS, ph = (it / LS_NS) & 1;
    ptx::mbar_wait(&empty[s], ph ^ 1);
    const int64_t s0 = static_cast<int64_t>(ch) * LS_V;
    const uint32_t nv = static_cast<uint32_t>(min(static_cast<int64_t>(LS_V), S - s0));
    const uint32_t lab_bytes = tu8 ? nv : nv * 4u;
    ptx::mbar_expect_tx(&full[s], nv * 4u * C + lab_bytes);
    for (int c = 0; c < C; ++c) ptx::bulk_load_1d(stages[s].z[c], zb + c * S + s0, nv * 4u, &full[s]);
    if (tu8)
      ptx::bulk_load_1d(stages[s].t, static_cast<const uint8_t*>(tb) + s0, lab_bytes, &full[s]);
    else
      ptx::bulk_load_1d(stages[s].t, static_cast<const float*>(tb) + s0, lab_bytes, &full[s]);
  }
}

template <int MAXC, bool TU8>
__global__ void __launch_bounds__(LS_THREADS, 2)
partial_loss_fwd_staged_kernel(const float* __restrict__ logits, const void* __restrict__ target,
                               const float* __restrict__ cw, const float* __restrict__ lut, double* __restrict__ sums,
                               float* __restrict__ loss, unsigned int* __restrict__ ticket, int N, int64_t S, int C, int uce,
                               int per_sample) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer arithmetic on the shared array, so that the accesses below stay LDS (see stem_tc.cu)
  LossStage<MAXC>* stages = reinterpret_cast<LossStage<MAXC>*>(smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u));
  __shared__ uint64_t full[LS_NS], empty[LS_NS];
  __shared__ float s_lut[MAXC];
  __shared__ float s_part[LS_CONSUMERS / 32][4 * MAXC];
  __shared__ bool s_last;
  __shared__ unsigned int s_ce_mask;
  const int64_t n = blockIdx.y;
  const int grp = per_sample ? static_cast<int>(n) : 0;
  const float* cw_g = cw + static_cast<int64_t>(grp) * C;
  const float* lut_g = lut ? lut + static_cast<int64_t>(grp) * C : nullptr;
  if (threadIdx.x == 0) {
    for (int i = 0; i < LS_NS; ++i) ptx::mbar_init(&full[i], 1), ptx::mbar_init(&empty[i], LS_CONSUMERS / 32);
    ptx::fence_barrier_init();
    unsigned int m = 0;
    for (int c = 0; c < C; ++c)
      if (uce && cw_g[c] != 0.f) m |= 1u << c;
    s_ce_mask = m;
  }
  if (threadIdx.x < MAXC) s_lut[threadIdx.x] = (lut_g && threadIdx.x < C) ? lut_g[threadIdx.x] : static_cast<float>(threadIdx.x);
  __syncthreads();
  const int nchunks = static_cast<int>((S + LS_V - 1) / LS_V);
  const float* zb = logits + n * C * S;
  const uint8_t* tb = static_cast<const uint8_t*>(target) + n * S * (TU8 ? 1 : 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float aI[MAXC], aZ[MAXC], aY[MAXC], aE[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) aI[c] = aZ[c] = aY[c] = aE[c] = 0.f;
  if (warp == LS_CONSUMERS / 32) {
    if (lane == 0) loss_produce<MAXC>(stages, full, empty, zb, tb, TU8, S, C, blockIdx.x, gridDim.x, nchunks);
  } else {
    const unsigned int ce_mask = s_ce_mask;
    uint32_t it = 0;
    for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
      const uint32_t s = it % LS_NS, ph = (it / LS_NS) & 1;
      ptx::mbar_wait(&full[s], ph);
      const int nv = static_cast<int>(min(static_cast<int64_t>(LS_V), S - static_cast<int64_t>(ch) * LS_V));
#pragma unroll
      for (int j = 0; j < LS_V / LS_CONSUMERS; ++j) {
        const int v = threadIdx.x + j * LS_CONSUMERS;
        if (v < nv) {
          float p[MAXC];
#pragma unroll
          for (int c = 0; c < MAXC; ++c) p[c] = c < C ? stages[s].z[c][v] : 0.f;
          softmax_inplace<MAXC>(p, C);
          const float tv = TU8 ? static_cast<float>(reinterpret_cast<const uint8_t*>(stages[s].t)[v]) : stages[s].t[v];
          const int tc = class_of(tv, lut_g ? s_lut : nullptr, C);
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            if (c < C) {
              const bool t = (c == tc);
              aZ[c] = fmaf(p[c], p[c], aZ[c]);
              aI[c] += t ? p[c] : 0.f;
              aY[c] += t ? 1.f : 0.f;
              if (ce_mask & (1u << c)) {
                const float l = t ? logf(p[c]) : logf(1.0f - p[c]);     // nn.BCELoss: log of the fp32 probability, clamp -100
                aE[c] -= fmaxf(l, -100.f);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      const float a = warp_sum(aI[c]), b = warp_sum(aZ[c]), d = warp_sum(aY[c]), e = warp_sum(aE[c]);
      if (lane == 0) {
        s_part[warp][c] = a;
        s_part[warp][MAXC + c] = b;
        s_part[warp][2 * MAXC + c] = d;
        s_part[warp][3 * MAXC + c] = e;
      }
    }
  }
  __syncthreads();
  double* sums_g = sums + static_cast<int64_t>(grp) * 4 * C;
  if (threadIdx.x < 4 * MAXC) {
    const int k = threadIdx.x / MAXC, c = threadIdx.x % MAXC;
    if (c < C) {
      double t = 0;
      for (int w = 0; w < LS_CONSUMERS / 32; ++w) t += static_cast<double>(s_part[w][threadIdx.x]);
      atomicAdd(&sums_g[k * C + c], t);
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

template <int MAXC, bool TU8>
__global__ void __launch_bounds__(LS_THREADS, 2)
partial_loss_bwd_staged_kernel(const float* __restrict__ logits, const void* __restrict__ target,
                               const float* __restrict__ cw, const float* __restrict__ lut, const double* __restrict__ sums,
                               const float* __restrict__ grad_out, float* __restrict__ dlogits, int N, int64_t S, int C,
                               int uce, int per_sample) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer arithmetic on the shared array, so that the accesses below stay LDS (see stem_tc.cu)
  LossStage<MAXC>* stages = reinterpret_cast<LossStage<MAXC>*>(smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u));
  __shared__ uint64_t full[LS_NS], empty[LS_NS];
  __shared__ float s_lut[MAXC], s_a[MAXC], s_b[MAXC], s_e[MAXC];
  const int64_t n = blockIdx.y;
  const int grp = per_sample ? static_cast<int>(n) : 0;
  const float* lut_g = lut ? lut + static_cast<int64_t>(grp) * C : nullptr;
  if (threadIdx.x == 0) {
    for (int i = 0; i < LS_NS; ++i) ptx::mbar_init(&full[i], 1), ptx::mbar_init(&empty[i], LS_CONSUMERS / 32);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < MAXC) {
    const int c = threadIdx.x;
    s_lut[c] = (lut_g && c < C) ? lut_g[c] : static_cast<float>(c);
    if (c < C) {
      const double* sg = sums + static_cast<int64_t>(grp) * 4 * C;
      const double sm = 1e-5, I = sg[c], Z = sg[C + c], Y = sg[2 * C + c], w = cw[grp * C + c];
      const double go = static_cast<double>(*grad_out) / (per_sample ? N : 1);
      const double nv = static_cast<double>(per_sample ? S : static_cast<int64_t>(N) * S);
      const double Dc = Z + Y + sm;
      s_a[c] = static_cast<float>(go * (w / C) * (-2.0 / Dc));
      s_b[c] = static_cast<float>(go * (w / C) * 2.0 * (2.0 * I + sm) / (Dc * Dc));
      s_e[c] = uce ? static_cast<float>(go * w / nv) : 0.f;
    } else {
      s_a[c] = s_b[c] = s_e[c] = 0.f;
    }
  }
  __syncthreads();
  const int nchunks = static_cast<int>((S + LS_V - 1) / LS_V);
  const float* zb = logits + n * C * S;
  float* db = dlogits + n * C * S;
  const uint8_t* tb = static_cast<const uint8_t*>(target) + n * S * (TU8 ? 1 : 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == LS_CONSUMERS / 32) {
    if (lane == 0) loss_produce<MAXC>(stages, full, empty, zb, tb, TU8, S, C, blockIdx.x, gridDim.x, nchunks);
    return;
  }
  uint32_t it = 0;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x, ++it) {
    const uint32_t s = it % LS_NS, ph = (it / LS_NS) & 1;
    ptx::mbar_wait(&full[s], ph);
    const int64_t s0 = static_cast<int64_t>(ch) * LS_V;
    const int nv = static_cast<int>(min(static_cast<int64_t>(LS_V), S - s0));
#pragma unroll
    for (int j = 0; j < LS_V / LS_CONSUMERS; ++j) {
      const int v = threadIdx.x + j * LS_CONSUMERS;
      if (v < nv) {
        float p[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) p[c] = c < C ? stages[s].z[c][v] : 0.f;
        softmax_inplace<MAXC>(p, C);
        const float tv = TU8 ? static_cast<float>(reinterpret_cast<const uint8_t*>(stages[s].t)[v]) : stages[s].t[v];
        const int tc = class_of(tv, lut_g ? s_lut : nullptr, C);
        auto gfun = [&](int c) {
          const float t = (c == tc) ? 1.f : 0.f;
          float gc = t * s_a[c] + p[c] * s_b[c];
          if (s_e[c] != 0.f) gc += s_e[c] * (p[c] - t) / fmaxf(p[c] * (1.0f - p[c]), 1e-12f);   // warp-uniform branch
          return gc;
        };
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) dot = fmaf(gfun(c), p[c], dot);
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) db[c * S + s0 + v] = p[c] * (gfun(c) - dot);
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[s]);
  }
}

template <int MAXC, int NTHR, int VEC, bool TU8>
void launch_fwd(const float* logits, const void* target, const float* cw, const float* lut, double* sums, float* loss,
                unsigned int* ticket, int n, int64_t S, int C, int uce, int per_sample, int blocks_per_sm, cudaStream_t s) {
  const int bx = static_cast<int>(std::min<int64_t>((S / VEC + NTHR - 1) / NTHR, std::max(1, num_sms() * blocks_per_sm / n)));
  partial_loss_fwd_kernel<MAXC, NTHR, VEC, TU8><<<dim3(bx, n), NTHR, 0, s>>>(logits, target, cw, lut, sums, loss, ticket, n,
                                                                            S, C, uce, per_sample);
}

template <int MAXC, int VEC, bool TU8>
void launch_bwd(const float* logits, const void* target, const float* cw, const float* lut, const double* sums,
                const float* grad_out, float* dlogits, int n, int64_t S, int C, int uce, int per_sample, cudaStream_t s) {
  const int bx = static_cast<int>(std::min<int64_t>((S / VEC + kThreads - 1) / kThreads, std::max(1, num_sms() * (VEC == 4 ? 1 : 8) / n)));
  partial_loss_bwd_kernel<MAXC, VEC, TU8><<<dim3(bx, n), kThreads, 0, s>>>(logits, target, cw, lut, sums, grad_out, dlogits,
                                                                          n, S, C, uce, per_sample);
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

// Measured on B200 (cfg2, profiles/r02_bench_kernels.txt): the 4-voxel form needs ~220 registers, i.e. 8 resident warps
// per SM, and its long compute phases between load batches then expose the HBM latency (fwd 211 us vs 140 us scalar at 24
// warps per SM).  The scalar form stays the default; MMPL_LOSS_VEC4=1 selects the vector form for experiments.
static bool vec4_ok(const void* logits, const void* target, const void* dlogits, int64_t S, int u8) {
  static const bool enabled = [] { const char* e = getenv("MMPL_LOSS_VEC4"); return e && e[0] == '1'; }();
  const uintptr_t a = reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits);
  return enabled && S % 4 == 0 && a % 16 == 0 && reinterpret_cast<uintptr_t>(target) % (u8 ? 4 : 16) == 0;
}

// staged (bulk-async) kernels: <= 16 classes, plane size a multiple of 16 voxels (every chunk then starts 16-byte aligned
// for fp32 planes and uint8 labels alike and has a size that is a multiple of 16 bytes), 16-byte aligned bases
static bool staged_ok(const void* logits, const void* target, const void* dlogits, int64_t S, int classes) {
  static const bool enabled = [] { const char* e = getenv("MMPL_LOSS_STAGED"); return !(e && e[0] == '0'); }();
  const uintptr_t a = reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits) | reinterpret_cast<uintptr_t>(target);
  return enabled && classes <= 16 && S % 16 == 0 && a % 16 == 0;
}

extern "C" int mmpl_partial_loss_fwd(const float* logits, const void* target, int target_is_u8,
                                     const float* class_weight, const float* lut, int per_sample, double* sums,
                                     float* loss, int n, int64_t spatial, int classes, int uce, mmpl_stream_t stream) {
  MMPL_REQUIRE(classes >= 1 && classes <= 32, MMPL_E_SHAPE, "partial_loss: classes=%d (1..32 supported)", classes);
  MMPL_REQUIRE(n > 0 && spatial > 0, MMPL_E_SHAPE, "partial_loss: empty input");
  MMPL_REQUIRE(n <= 65535, MMPL_E_SHAPE, "partial_loss: batch %d exceeds the grid limit", n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the last-block ticket is the trailing slot of the caller's workspace: nothing is shared between calls in flight on
  // different streams or devices, and nothing is allocated here (a first call may sit inside a CUDA-graph capture)
  const int groups = per_sample ? n : 1;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + static_cast<int64_t>(groups) * 4 * classes);
  MMPL_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * (static_cast<int64_t>(groups) * 4 * classes + 1), s));
  // forward: ncu shows both forms issue ~78 M warp instructions (about 540 per voxel: the kernel is instruction-issue bound,
  // not bandwidth bound), and the register form keeps 24 warps per SM busy against 16 for the staged one -- 167 us vs
  // 272 us at cfg2.  The staged forward is kept behind MMPL_LOSS_STAGED_FWD=1; the staged BACKWARD is the default (230 us
  // vs 255 us: its stores no longer compete with the loads for the load/store unit).
  static const bool staged_fwd = [] { const char* e = getenv("MMPL_LOSS_STAGED_FWD"); return e && e[0] == '1'; }();
  if (staged_fwd && staged_ok(logits, target, nullptr, spatial, classes)) {
    const size_t smem = sizeof(LossStage<16>) * LS_NS + 128;
    const int nchunks = static_cast<int>((spatial + LS_V - 1) / LS_V);
    const dim3 grid(std::min(nchunks, std::max(1, num_sms() * 2 / n)), n);
    if (target_is_u8) {
      MMPL_CUDA(cudaFuncSetAttribute(partial_loss_fwd_staged_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      partial_loss_fwd_staged_kernel<16, true><<<grid, LS_THREADS, smem, s>>>(logits, target, class_weight, lut, sums, loss, ticket,
                                                                           n, spatial, classes, uce, per_sample);
    } else {
      MMPL_CUDA(cudaFuncSetAttribute(partial_loss_fwd_staged_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      partial_loss_fwd_staged_kernel<16, false><<<grid, LS_THREADS, smem, s>>>(logits, target, class_weight, lut, sums, loss, ticket,
                                                                            n, spatial, classes, uce, per_sample);
    }
    MMPL_CHECK_LAUNCH("partial_loss_fwd");
    return MMPL_OK;
  }
  const bool v4 = vec4_ok(logits, target, nullptr, spatial, target_is_u8);
#define MMPL_LOSS_FWD(MAXC, NTHR, VEC, TU8, BPS) \
  launch_fwd<MAXC, NTHR, VEC, TU8>(logits, target, class_weight, lut, sums, loss, ticket, n, spatial, classes, uce, per_sample, BPS, s)
  if (classes <= 16) {
    if (v4 && target_is_u8) MMPL_LOSS_FWD(16, 256, 4, true, 1);
    else if (v4) MMPL_LOSS_FWD(16, 256, 4, false, 1);
    else if (target_is_u8) MMPL_LOSS_FWD(16, 256, 1, true, 6);
    else MMPL_LOSS_FWD(16, 256, 1, false, 6);
  } else {
    if (target_is_u8) MMPL_LOSS_FWD(32, 128, 1, true, 8);
    else MMPL_LOSS_FWD(32, 128, 1, false, 8);
  }
#undef MMPL_LOSS_FWD
  MMPL_CHECK_LAUNCH("partial_loss_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_partial_loss_bwd(const float* logits, const void* target, int target_is_u8,
                                     const float* class_weight, const float* lut, int per_sample, const double* sums,
                                     const float* grad_out, float* dlogits, int n, int64_t spatial, int classes, int uce,
                                     mmpl_stream_t stream) {
  MMPL_REQUIRE(classes >= 1 && classes <= 32, MMPL_E_SHAPE, "partial_loss: classes=%d (1..32 supported)", classes);
  MMPL_REQUIRE(n > 0 && spatial > 0, MMPL_E_SHAPE, "partial_loss: empty input");
  MMPL_REQUIRE(n <= 65535, MMPL_E_SHAPE, "partial_loss: batch %d exceeds the grid limit", n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (staged_ok(logits, target, dlogits, spatial, classes)) {
    const size_t smem = sizeof(LossStage<16>) * LS_NS + 128;
    const int nchunks = static_cast<int>((spatial + LS_V - 1) / LS_V);
    const dim3 grid(std::min(nchunks, std::max(1, num_sms() * 2 / n)), n);
    if (target_is_u8) {
      MMPL_CUDA(cudaFuncSetAttribute(partial_loss_bwd_staged_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      partial_loss_bwd_staged_kernel<16, true><<<grid, LS_THREADS, smem, s>>>(logits, target, class_weight, lut, sums, grad_out,
                                                                           dlogits, n, spatial, classes, uce, per_sample);
    } else {
      MMPL_CUDA(cudaFuncSetAttribute(partial_loss_bwd_staged_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      partial_loss_bwd_staged_kernel<16, false><<<grid, LS_THREADS, smem, s>>>(logits, target, class_weight, lut, sums, grad_out,
                                                                            dlogits, n, spatial, classes, uce, per_sample);
    }
    MMPL_CHECK_LAUNCH("partial_loss_bwd");
    return MMPL_OK;
  }
  const bool v4 = vec4_ok(logits, target, dlogits, spatial, target_is_u8);
#define MMPL_LOSS_BWD(MAXC, VEC, TU8) \
  launch_bwd<MAXC, VEC, TU8>(logits, target, class_weight, lut, sums, grad_out, dlogits, n, spatial, classes, uce, per_sample, s)
  if (classes <= 16) {
    if (v4 && target_is_u8) MMPL_LOSS_BWD(16, 4, true);
    else if (v4) MMPL_LOSS_BWD(16, 4, false);
    else if (target_is_u8) MMPL_LOSS_BWD(16, 1, true);
    else MMPL_LOSS_BWD(16, 1, false);
  } else {
    if (target_is_u8) MMPL_LOSS_BWD(32, 1, true);
    else MMPL_LOSS_BWD(32, 1, false);
  }
#undef MMPL_LOSS_BWD
  MMPL_CHECK_LAUNCH("partial_loss_bwd");
  return MMPL_OK;
}
