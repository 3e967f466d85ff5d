// Shared host/device helpers for libmmpl_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mmpl_b200.h"

namespace mmpl {

// ---------------------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define MMPL_FAIL(code, ...)     \
  do {                           \
    ::mmpl::set_error(__VA_ARGS__); \
    return (code);               \
  } while (0)

#define MMPL_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) MMPL_FAIL(code, __VA_ARGS__); \
  } while (0)

#define MMPL_CHECK_LAUNCH(name)                                                         \
  do {                                                                                  \
    ::mmpl::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) MMPL_FAIL(MMPL_E_CUDA, "%s: %s", name, cudaGetErrorString(e_)); \
  } while (0)

#define MMPL_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) MMPL_FAIL(MMPL_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

int num_sms();

// The tensor-map encoders are DRIVER entry points: they need the primary context current on the calling thread, which the
// runtime only guarantees after the thread's first runtime call.  A backward pass that starts with a convolution runs on
// an autograd worker thread that may not have made one yet (cuTensorMapEncodeTiled then fails with INVALID_CONTEXT).
inline void bind_primary_context() {
  static thread_local bool bound = false;
  if (!bound) {
    cudaFree(nullptr);
    bound = true;
  }
}

static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------- vector access
// VecT<T> moves 16 bytes: 8 bf16 or 4 fp32 channels.
template <typename T>
struct Vec;
template <>
struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(u[i] << 16);
      v[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      u[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
  }
};

// VecH<T> moves 8 bytes: 4 bf16 or 2 fp32 channels (half the per-thread state of Vec<T>: used where per-channel
// constants would otherwise push the register count past the occupancy a streaming kernel needs).
template <typename T>
struct VecH;
template <>
struct VecH<float> {
  static constexpr int N = 2;
  float v[2];
  __device__ __forceinline__ void load(const float* p) {
    float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x, v[1] = t.y;
  }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
template <>
struct VecH<__nv_bfloat16> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xFFFF0000u);
    v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xFFFF0000u);
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
};

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) {
  return x;
}
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) {
  return __bfloat162float(x);
}
template <typename T>
__device__ __forceinline__ T from_f32(float x);
template <>
__device__ __forceinline__ float from_f32<float>(float x) {
  return x;
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) {
  return __float2bfloat16(x);
}

// ---------------------------------------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Dispatch on the activation dtype.
#define MMPL_DISPATCH_DTYPE(dtype, T, ...)                            \
  do {                                                                \
    if ((dtype) == MMPL_F32) {                                        \
      using T = float;                                                \
      __VA_ARGS__;                                                    \
    } else if ((dtype) == MMPL_BF16) {                                \
      using T = __nv_bfloat16;                                        \
      __VA_ARGS__;                                                    \
    } else {                                                          \
      MMPL_FAIL(MMPL_E_DTYPE, "unsupported dtype %d", (int)(dtype)); \
    }                                                                 \
  } while (0)

}  // namespace mmpl
