// Fused SGD-with-momentum step over a flat fp32 parameter buffer.
// Reference: torch.optim.SGD(lr, momentum=0.9, weight_decay=1e-4) at train_amos_atlas_final.py:132-135, .step() :378
// (the reference runs it as several foreach kernels per step); poly LR from utils.py:53-60 arrives as a device scalar.
#include "common.cuh"

namespace mmpl {
namespace {
__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, int64_t count,
           const float* __restrict__ lr_dev, float momentum, float wd, float gscale, int first, int head) {
  const float lr = *lr_dev;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // `head` scalar elements bring the three (equally misaligned) pointers to a 16-byte boundary: a range of the flat
  // buffers may start at any parameter
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < head; i += stride) {
    const float d = fmaf(wd, p[i], g[i] * gscale);
    const float b = first ? d : fmaf(momentum, buf[i], d);
    buf[i] = b;
    p[i] = fmaf(-lr, b, p[i]);
  }
  p += head, g += head, buf += head, count -= head;
  const int64_t n4 = count / 4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
    float* pp = &pv.x;
    const float* gg = &gv.x;
    float* bb = &bv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = fmaf(wd, pp[k], gg[k] * gscale);
      bb[k] = first ? d : fmaf(momentum, bb[k], d);
      pp[k] = fmaf(-lr, bb[k], pp[k]);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(buf)[i] = bv;
  }
  for (int64_t i = n4 * 4 + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count; i += stride) {
    const float d = fmaf(wd, p[i], g[i] * gscale);
    const float b = first ? d : fmaf(momentum, buf[i], d);
    buf[i] = b;
    p[i] = fmaf(-lr, b, p[i]);
  }
}
}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_sgd_step(float* p, const float* grad, float* buf, int64_t count, const float* lr_dev,
                             float momentum, float weight_decay, float grad_scale, int first_step,
                             mmpl_stream_t stream) {
  MMPL_REQUIRE(count >= 0, MMPL_E_SHAPE, "sgd: count=%lld", static_cast<long long>(count));
  if (count == 0) return MMPL_OK;
  const uintptr_t mis = reinterpret_cast<uintptr_t>(p) % 16;
  MMPL_REQUIRE(mis % 4 == 0 && reinterpret_cast<uintptr_t>(grad) % 16 == mis && reinterpret_cast<uintptr_t>(buf) % 16 == mis,
               MMPL_E_ALIGN, "sgd: the three buffers must be 4-byte aligned and share their offset from a 16-byte boundary");
  const int head = static_cast<int>(std::min<int64_t>(count, ((16 - mis) % 16) / 4));
  const int blocks = static_cast<int>(std::min<int64_t>((count / 4 + 255) / 256 + 1, static_cast<int64_t>(num_sms()) * 8));
  sgd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, grad, buf, count, lr_dev, momentum, weight_decay,
                                                                   grad_scale, first_step, head);
  MMPL_CHECK_LAUNCH("sgd_step");
  return MMPL_OK;
}
