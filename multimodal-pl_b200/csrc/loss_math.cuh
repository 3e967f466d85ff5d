// Per-voxel pieces of the partial-label loss shared by the stand-alone loss kernels (loss.cu) and the fused
// classifier + loss kernels (cls_loss.cu).  Reference: loss_functions/loss_partial.py:24-57, :71-99.
#pragma once
#include "common.cuh"

namespace mmpl {

// a label value -> class id in [0, C) or -1 (not a class id); `lut` is the per-sample cmask remap
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

// The loss from the per-class sums [G][4][C] = (I, Z, Y, E), run by one thread of the last block to finish.
// pooled (reference, loss_partial.py:87,92): one group over batch and voxels, class weights = mask[0].
// per-sample: the reference formula evaluated per sample with that sample's weights, averaged over the batch.
__device__ __forceinline__ void partial_loss_finalize(const double* sums, const float* cw, int N, int64_t S, int C, int uce,
                                                      int per_sample, float* loss) {
  const int G = per_sample ? N : 1;
  const double sm = 1e-5, nv = static_cast<double>(per_sample ? S : static_cast<int64_t>(N) * S);
  double total = 0;
  for (int g = 0; g < G; ++g) {
    const volatile double* vs = sums + static_cast<int64_t>(g) * 4 * C;
    double dice = 0, ce = 0;
    for (int c = 0; c < C; ++c) {
      const double I = vs[c], Z = vs[C + c], Y = vs[2 * C + c], E = vs[3 * C + c], w = cw[g * C + c];
      dice += w * (1.0 - (2.0 * I + sm) / (Z + Y + sm));
      ce += w * (E / nv);
    }
    total += dice / C + (uce ? ce : 0.0);
  }
  *loss = static_cast<float>(total / G);
}

// Backward coefficients of class c (closed form), grad_out folded in:
//   g_c = t_c * ca + p_c * cb + ce * (p_c - t_c) / max(p_c (1 - p_c), 1e-12),   dz_c = p_c (g_c - sum_k g_k p_k)
__device__ __forceinline__ void partial_loss_coeffs(const double* sums_g, float w, float grad_out, int N, int64_t S, int C,
                                                    int c, int uce, int per_sample, float& ca, float& cb, float& ce) {
  const double sm = 1e-5, I = sums_g[c], Z = sums_g[C + c], Y = sums_g[2 * C + c];
  const double go = static_cast<double>(grad_out) / (per_sample ? N : 1);
  const double nv = static_cast<double>(per_sample ? S : static_cast<int64_t>(N) * S);   // voxels in the BCE mean
  const double Dc = Z + Y + sm;
  ca = static_cast<float>(go * (w / C) * (-2.0 / Dc));
  cb = static_cast<float>(go * (w / C) * 2.0 * (2.0 * I + sm) / (Dc * Dc));
  ce = uce ? static_cast<float>(go * w / nv) : 0.f;
}

}  // namespace mmpl
