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
#include <algorithm>

#include "common.cuh"

namespace mmpl {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int class_of(float tv, const float* lut, int C) {
  int ti = static_cast<int>(tv);
  if (static_cast<float>(ti) != tv || ti < 0 || ti >= C) return -1;
  if (lut) {
    tv = lut[ti];
    ti = static_cast<int>(tv);
    if (static_cast<float>(ti) != tv || ti < 0 || ti >= C) return -1;
  }
  return ti;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MAXC>
__device__ __forceinline__ void softmax_regs(const float* __restrict__ logits, int64_t base, int64_t S, int C,
                                             float (&p)[MAXC]) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    p[c] = c < C ? logits[base + c * S] : -INFINITY;
    mx = fmaxf(mx, p[c]);
  }
  // exp(z - mx) as one FMA + one MUFU: 2^(z*log2(e) - mx*log2(e)), ex2.approx (max rel. error 2^-22; the arguments are
  // <= 0 so flush-to-zero only affects probabilities below 1e-38).  The kernels are instruction-issue bound.
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

template <int MAXC, int NTHR>
__global__ void __launch_bounds__(NTHR, MAXC <= 16 ? 3 : 1)
partial_loss_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                        const float* __restrict__ cw, const float* __restrict__ lut, double* __restrict__ sums,
                        float* __restrict__ loss, unsigned int* __restrict__ ticket, int N, int64_t S, int C, int uce) {
  __shared__ float s_lut[MAXC];
  __shared__ float s_part[NTHR / 32][4 * MAXC];
  __shared__ bool s_last;
  __shared__ unsigned int s_ce_mask;   // classes whose BCE term is needed (weight != 0)
  // I_c and Y_c only ever receive the voxel's own class: per-thread accumulators in shared memory ([class][thread],
  // conflict-free) take a dynamic class index, which registers cannot; Z_c and E_c stay in registers.  One slot more
  // than MAXC collects voxels whose label is not a class id.
  __shared__ float s_I[MAXC + 1][NTHR], s_Y[MAXC + 1][NTHR];
  if (threadIdx.x < MAXC) s_lut[threadIdx.x] = (lut && threadIdx.x < C) ? lut[threadIdx.x] : static_cast<float>(threadIdx.x);
  if (threadIdx.x == 0) {
    unsigned int m = 0;
    for (int c = 0; c < C; ++c)
      if (uce && cw[c] != 0.f) m |= 1u << c;
    s_ce_mask = m;
  }
#pragma unroll
  for (int c = 0; c <= MAXC; ++c) s_I[c][threadIdx.x] = s_Y[c][threadIdx.x] = 0.f;
  __syncthreads();
  const unsigned int ce_mask = s_ce_mask;
  float aZ[MAXC], aE[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) aZ[c] = aE[c] = 0.f;
  const int64_t total = static_cast<int64_t>(N) * S;
  // grid = (blocks per sample, N): no 64-bit division per voxel
  const int64_t n = blockIdx.y;
  for (int64_t s = blockIdx.x * static_cast<int64_t>(NTHR) + threadIdx.x; s < S;
       s += static_cast<int64_t>(gridDim.x) * NTHR) {
    float p[MAXC];
    softmax_regs<MAXC>(logits, n * C * S + s, S, C, p);
    const int tc = class_of(target[n * S + s], lut ? s_lut : nullptr, C);
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
  if (threadIdx.x < 4 * MAXC) {
    const int k = threadIdx.x / MAXC, c = threadIdx.x % MAXC;
    if (c < C) {
      double t = 0;
      for (int w = 0; w < NTHR / 32; ++w) t += static_cast<double>(s_part[w][threadIdx.x]);
      atomicAdd(&sums[k * C + c], t);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    const double sm = 1e-5, nv = static_cast<double>(total);
    double dice = 0, ce = 0;
    for (int c = 0; c < C; ++c) {
      const volatile double* vs = sums;
      const double I = vs[c], Z = vs[C + c], Y = vs[2 * C + c], E = vs[3 * C + c], w = cw[c];
      dice += w * (1.0 - (2.0 * I + sm) / (Z + Y + sm));
      ce += w * (E / nv);
    }
    *loss = static_cast<float>(dice / C + (uce ? ce : 0.0));
    *ticket = 0;
  }
}

template <int MAXC>
__global__ void __launch_bounds__(kThreads, MAXC <= 16 ? 2 : 1)
partial_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                        const float* __restrict__ cw, const float* __restrict__ lut, const double* __restrict__ sums,
                        const float* __restrict__ grad_out, float* __restrict__ dlogits, int N, int64_t S, int C,
                        int uce) {
  __shared__ float s_lut[MAXC], s_a[MAXC], s_b[MAXC], s_e[MAXC];
  const int64_t total = static_cast<int64_t>(N) * S;   // voxels in the BCE mean
  if (threadIdx.x < MAXC) {
    const int c = threadIdx.x;
    s_lut[c] = (lut && c < C) ? lut[c] : static_cast<float>(c);
    if (c < C) {
      const double sm = 1e-5, I = sums[c], Z = sums[C + c], Y = sums[2 * C + c], w = cw[c], go = *grad_out;
      const double Dc = Z + Y + sm;
      s_a[c] = static_cast<float>(go * (w / C) * (-2.0 / Dc));
      s_b[c] = static_cast<float>(go * (w / C) * 2.0 * (2.0 * I + sm) / (Dc * Dc));
      s_e[c] = uce ? static_cast<float>(go * w / static_cast<double>(total)) : 0.f;
    } else {
      s_a[c] = s_b[c] = s_e[c] = 0.f;
    }
  }
  __syncthreads();
  const int64_t n = blockIdx.y;
  for (int64_t s = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; s < S;
       s += static_cast<int64_t>(gridDim.x) * kThreads) {
    float p[MAXC];
    softmax_regs<MAXC>(logits, n * C * S + s, S, C, p);
    const int tc = class_of(target[n * S + s], lut ? s_lut : nullptr, C);
    // g_c is cheap: evaluate it twice (once for the dot product, once for the output) instead of keeping 16 more
    // registers live -- the kernel is bound by loads in flight, i.e. by occupancy
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
      if (c < C) dlogits[n * C * S + c * S + s] = p[c] * (gfun(c) - dot);
  }
}

unsigned int* ticket_buffer() {
  static unsigned int* t = nullptr;
  if (!t) {
    if (cudaMalloc(&t, sizeof(unsigned int)) != cudaSuccess) return nullptr;
    cudaMemset(t, 0, sizeof(unsigned int));
  }
  return t;
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_partial_loss_fwd(const float* logits, const float* target, const float* class_weight,
                                     const float* lut, double* sums, float* loss, int n, int64_t spatial, int classes,
                                     int uce, mmpl_stream_t stream) {
  MMPL_REQUIRE(classes >= 1 && classes <= 32, MMPL_E_SHAPE, "partial_loss: classes=%d (1..32 supported)", classes);
  MMPL_REQUIRE(n > 0 && spatial > 0, MMPL_E_SHAPE, "partial_loss: empty input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned int* ticket = ticket_buffer();
  MMPL_REQUIRE(ticket != nullptr, MMPL_E_CUDA, "partial_loss: ticket allocation failed");
  MMPL_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * classes, s));
  const int64_t total = static_cast<int64_t>(n) * spatial;
  MMPL_REQUIRE(n <= 65535, MMPL_E_SHAPE, "partial_loss: batch %d exceeds the grid limit", n);
  if (classes <= 16) {
    const int bx = static_cast<int>(std::min<int64_t>((spatial + 255) / 256, std::max(1, num_sms() * 6 / n)));
    const dim3 blocks(bx, n);
    partial_loss_fwd_kernel<16, 256><<<blocks, 256, 0, s>>>(logits, target, class_weight, lut, sums, loss, ticket, n,
                                                           spatial, classes, uce);
  } else {
    const int bx = static_cast<int>(std::min<int64_t>((spatial + 127) / 128, std::max(1, num_sms() * 8 / n)));
    const dim3 blocks(bx, n);
    partial_loss_fwd_kernel<32, 128><<<blocks, 128, 0, s>>>(logits, target, class_weight, lut, sums, loss, ticket, n,
                                                           spatial, classes, uce);
  }
  MMPL_CHECK_LAUNCH("partial_loss_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_partial_loss_bwd(const float* logits, const float* target, const float* class_weight,
                                     const float* lut, const double* sums, const float* grad_out, float* dlogits, int n,
                                     int64_t spatial, int classes, int uce, mmpl_stream_t stream) {
  MMPL_REQUIRE(classes >= 1 && classes <= 32, MMPL_E_SHAPE, "partial_loss: classes=%d (1..32 supported)", classes);
  MMPL_REQUIRE(n > 0 && spatial > 0, MMPL_E_SHAPE, "partial_loss: empty input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_REQUIRE(n <= 65535, MMPL_E_SHAPE, "partial_loss: batch %d exceeds the grid limit", n);
  const int bx = static_cast<int>(std::min<int64_t>((spatial + kThreads - 1) / kThreads, std::max(1, num_sms() * 8 / n)));
  const dim3 blocks(bx, n);
  if (classes <= 16)
    partial_loss_bwd_kernel<16><<<blocks, kThreads, 0, s>>>(logits, target, class_weight, lut, sums, grad_out, dlogits, n,
                                                           spatial, classes, uce);
  else
    partial_loss_bwd_kernel<32><<<blocks, kThreads, 0, s>>>(logits, target, class_weight, lut, sums, grad_out, dlogits, n,
                                                           spatial, classes, uce);
  MMPL_CHECK_LAUNCH("partial_loss_bwd");
  return MMPL_OK;
}
