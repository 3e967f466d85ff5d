// Fused binary Dice over a voxel gate (+ optional BCE-with-logits), forward and backward.
// Reference: DiceLoss._dice_loss (loss_functions/loss_partial.py:24-36) as called by EDiceLoss_full2.forward
// (:150-170) from the pseudo-label terms of get_loss (losses.py:165-176): score = sigmoid(x) (or x) [V], soft target
// t [V], gate m [V] (confidence mask):
//   I = sum_m p t,  Y = sum_m t^2,  Z = sum_m p^2,   dice = 1 - (2I + s)/(Z + Y + s),  s = 1e-5
//   bce  = mean over ALL voxels of  max(x,0) - x t + log(1 + exp(-|x|))          (nn.BCEWithLogitsLoss, :168)
// The reference materialises boolean-index gathers (nonzero + index) and ~10 reductions per call and calls this
// (organs x 4 scales) times per step; here it is one pass forward, one pass backward, no host synchronisation.
//   d dice / d p_i = m_i (-2 t_i / D + 2 p_i (2I + s)/D^2),  D = Z + Y + s;   d p/d x = p (1 - p) with the sigmoid
//   d dice / d t_i = m_i (-2 p_i / D + 2 t_i (2I + s)/D^2)
//   d bce / d x_i  = (sigmoid(x_i) - t_i)/V,   d bce / d t_i = -x_i / V
// Algorithmic HBM bytes per voxel (fp32): fwd 12, bwd 12 + 4 (+4 with the target gradient).
#include <algorithm>

#include "common.cuh"

namespace mmpl {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(kThreads)
masked_dice_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ m,
                       double* __restrict__ sums, float* __restrict__ loss, unsigned int* __restrict__ ticket, int64_t V,
                       int sigmoid, int uce) {
  float aI = 0.f, aY = 0.f, aZ = 0.f, aE = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < V;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const float xv = x[i], tv = t[i];
    const bool on = m == nullptr || m[i] != 0.f;
    const float p = sigmoid ? sigmoidf(xv) : xv;
    if (on) {
      aI = fmaf(p, tv, aI);
      aY = fmaf(tv, tv, aY);
      aZ = fmaf(p, p, aZ);
    }
    if (uce) aE += fmaxf(xv, 0.f) - xv * tv + log1pf(__expf(-fabsf(xv)));
  }
  __shared__ float s_part[kThreads / 32][4];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float a = warp_sum(aI), b = warp_sum(aY), c = warp_sum(aZ), d = warp_sum(aE);
  if (lane == 0) s_part[warp][0] = a, s_part[warp][1] = b, s_part[warp][2] = c, s_part[warp][3] = d;
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0;
    for (int w = 0; w < kThreads / 32; ++w) s += static_cast<double>(s_part[w][threadIdx.x]);
    atomicAdd(&sums[threadIdx.x], s);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    const volatile double* vs = sums;
    const double sm = 1e-5, I = vs[0], Y = vs[1], Z = vs[2], E = vs[3];
    *loss = static_cast<float>(1.0 - (2.0 * I + sm) / (Z + Y + sm) + (uce ? E / static_cast<double>(V) : 0.0));
  }
}

__global__ void __launch_bounds__(kThreads)
masked_dice_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ m,
                       const double* __restrict__ sums, const float* __restrict__ grad_out, float* __restrict__ dx,
                       float* __restrict__ dt, int64_t V, int sigmoid, int uce) {
  const double sm = 1e-5, I = sums[0], Y = sums[1], Z = sums[2], D = Z + Y + sm, go = *grad_out;
  const float ca = static_cast<float>(go * (-2.0 / D));
  const float cb = static_cast<float>(go * 2.0 * (2.0 * I + sm) / (D * D));
  const float ce = uce ? static_cast<float>(go / static_cast<double>(V)) : 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; i < V;
       i += static_cast<int64_t>(gridDim.x) * kThreads) {
    const float xv = x[i], tv = t[i];
    const bool on = m == nullptr || m[i] != 0.f;
    const float sg = (sigmoid || uce) ? sigmoidf(xv) : 0.f;
    const float p = sigmoid ? sg : xv;
    float gp = on ? fmaf(ca, tv, cb * p) : 0.f;        // d dice / d p
    if (sigmoid) gp *= p * (1.0f - p);
    dx[i] = gp + ce * (sg - tv);
    if (dt) dt[i] = (on ? fmaf(ca, p, cb * tv) : 0.f) - ce * xv;
  }
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_masked_dice_fwd(const float* x, const float* target, const float* gate, double* sums, float* loss,
                                    int64_t voxels, int sigmoid, int uce, mmpl_stream_t stream) {
  MMPL_REQUIRE(voxels > 0 && x && target && sums && loss, MMPL_E_SHAPE, "masked_dice: empty input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + 4);   // trailing slot of the caller's workspace
  MMPL_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 5, s));
  const int blocks = static_cast<int>(std::min<int64_t>((voxels + kThreads - 1) / kThreads, static_cast<int64_t>(num_sms()) * 8));
  masked_dice_fwd_kernel<<<blocks, kThreads, 0, s>>>(x, target, gate, sums, loss, ticket, voxels, sigmoid, uce);
  MMPL_CHECK_LAUNCH("masked_dice_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_masked_dice_bwd(const float* x, const float* target, const float* gate, const double* sums,
                                    const float* grad_out, float* dx, float* dtarget, int64_t voxels, int sigmoid, int uce,
                                    mmpl_stream_t stream) {
  MMPL_REQUIRE(voxels > 0 && x && target && sums && grad_out && dx, MMPL_E_SHAPE, "masked_dice: empty input");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int blocks = static_cast<int>(std::min<int64_t>((voxels + kThreads - 1) / kThreads, static_cast<int64_t>(num_sms()) * 8));
  masked_dice_bwd_kernel<<<blocks, kThreads, 0, s>>>(x, target, gate, sums, grad_out, dx, dtarget, voxels, sigmoid, uce);
  MMPL_CHECK_LAUNCH("masked_dice_bwd");
  return MMPL_OK;
}
