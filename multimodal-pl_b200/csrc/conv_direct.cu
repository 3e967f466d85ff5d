// CUDA-core (FFMA) implicit-GEMM convolution: fprop, dgrad and wgrad for k in {1,3}, stride in {1,2}.
// Reference op: F.conv3d inside Conv3d.forward (unet3D.py:27) and its autograd.  This is the exact-arithmetic path
// (fp32 activations: products and sums in fp32, like the reference) and the general-shape path for the layers the
// tcgen05 kernels in conv_tc.cu do not cover.  It is a GPU kernel, not a fallback to the CPU.
//
// Gather rule shared by fprop and dgrad: out voxel o, tap t reads input coordinate q = o*SO + t - pad; the read is
// valid when q % SI == 0 and 0 <= q/SI < Din.   fprop: SO = stride, SI = 1.   dgrad: SO = 1, SI = stride with the
// flipped/transposed packing written by mmpl_ws_weight_fwd (so both are plain correlations).
//
// Accumulation on the fp32 (exact) path is two-level: 32 products are summed in fp32, the 32-term partial sums in
// fp64.  A plain fp32 chain over K = 27*Cin <= 6912 terms carries a relative error of ~sqrt(K)*2^-24 = 5e-6, enough to
// resolve ~1e-5 of the following ReLU gates differently from the reference and to move whole-network gradients by 2e-3
// (tests/test_gpu_parity_strict.py::test_unet_fp32_all_gradients_vs_fp64_oracle); the blocked sum is as accurate as the
// reference's oneDNN kernels.  The bf16 instantiation keeps one fp32 accumulator (its inputs carry 2^-9 already).
#include <type_traits>

#include "common.cuh"

namespace mmpl {
namespace {

constexpr int TM = 64;   // output voxels per block
constexpr int TK = 32;   // channels per k-chunk

struct ConvDims {
  int N, Di, Hi, Wi, Do, Ho, Wo, Cin, Cout, k, pad, SO, SI;
  // Output sub-grid (stride-2 dgrad is run as 8 parity classes, each with only the taps that can hit it):
  // the kernel enumerates voxels (osd + 2*d', osh + 2*h', osw + 2*w') when ostep == 2; sub-grid extents Ds,Hs,Ws.
  int ostep, osd, osh, osw, Ds, Hs, Ws;
  int ntaps;
  int taplist[27];
};

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  Vec<__nv_bfloat16> t;
  t.load(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = t.v[i];
}

template <typename T, int TN>
__global__ void __launch_bounds__(256)
conv_direct_kernel(const T* __restrict__ x, const T* __restrict__ wp, const T* __restrict__ addend, T* __restrict__ y,
                   ConvDims dm) {
  constexpr int CN = TN / 16;  // output channels per thread
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ int s_n[TM], s_d[TM], s_h[TM], s_w[TM];
  __shared__ int64_t s_off[TM];
  const int tid = threadIdx.x;
  const int64_t M = static_cast<int64_t>(dm.N) * dm.Ds * dm.Hs * dm.Ws;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * TM;
  const int co0 = blockIdx.y * TN;
  if (tid < TM) {
    int64_t m = m0 + tid;
    if (m < M) {
      s_w[tid] = dm.osw + dm.ostep * static_cast<int>(m % dm.Ws);
      m /= dm.Ws;
      s_h[tid] = dm.osh + dm.ostep * static_cast<int>(m % dm.Hs);
      m /= dm.Hs;
      s_d[tid] = dm.osd + dm.ostep * static_cast<int>(m % dm.Ds);
      s_n[tid] = static_cast<int>(m / dm.Ds);
    } else {
      s_n[tid] = -1;
    }
  }
  __syncthreads();  // coordinates are read by every thread in the epilogue (and the tap loop may be empty)
  const int tm = tid / 16, tn = tid % 16;
  constexpr bool kTwoLevel = std::is_same<T, float>::value;
  using AccT = typename std::conditional<kTwoLevel, double, float>::type;
  AccT acc[4][CN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) acc[i][j] = 0;
  const int lv = tid / 4, lc = (tid % 4) * 8;  // loader mapping: voxel / weight row, 8 channels
  for (int ti = 0; ti < dm.ntaps; ++ti) {
    const int t = dm.taplist[ti];
    const int kd = t / (dm.k * dm.k), kh = (t / dm.k) % dm.k, kw = t % dm.k;
    __syncthreads();
    if (tid < TM) {
      int64_t off = -1;
      if (s_n[tid] >= 0) {
        const int qd = s_d[tid] * dm.SO + kd - dm.pad, qh = s_h[tid] * dm.SO + kh - dm.pad,
                  qw = s_w[tid] * dm.SO + kw - dm.pad;
        if (qd >= 0 && qh >= 0 && qw >= 0 && qd % dm.SI == 0 && qh % dm.SI == 0 && qw % dm.SI == 0) {
          const int id = qd / dm.SI, ih = qh / dm.SI, iw = qw / dm.SI;
          if (id < dm.Di && ih < dm.Hi && iw < dm.Wi)
            off = (((static_cast<int64_t>(s_n[tid]) * dm.Di + id) * dm.Hi + ih) * dm.Wi + iw) * dm.Cin;
        }
      }
      s_off[tid] = off;
    }
    __syncthreads();
    for (int c0 = 0; c0 < dm.Cin; c0 += TK) {
      float av[8];
      const int64_t off = s_off[lv];
      if (off >= 0) {
        load8<T>(x + off + c0 + lc, av);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) As[lc + i][lv] = av[i];
      if (lv < TN) {
        float bv[8];
        load8<T>(wp + (static_cast<int64_t>(t) * dm.Cout + co0 + lv) * dm.Cin + c0 + lc, bv);
#pragma unroll
        for (int i = 0; i < 8; ++i) Bs[lc + i][lv] = bv[i];
      }
      __syncthreads();
      float part[4][CN];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) part[i][j] = kTwoLevel ? 0.f : static_cast<float>(acc[i][j]);
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][tm * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float b[CN];
#pragma unroll
        for (int j = 0; j < CN; ++j) b[j] = Bs[kk][tn * CN + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < CN; ++j) part[i][j] = fmaf(a[i], b[j], part[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = kTwoLevel ? acc[i][j] + static_cast<AccT>(part[i][j]) : static_cast<AccT>(part[i][j]);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tm * 4 + i;
    if (m0 + r >= M) continue;
    const int64_t o = ((((static_cast<int64_t>(s_n[r]) * dm.Do + s_d[r]) * dm.Ho + s_h[r]) * dm.Wo + s_w[r])) * dm.Cout +
                      co0 + tn * CN;
#pragma unroll
    for (int j = 0; j < CN; ++j) {
      float v = static_cast<float>(acc[i][j]);
      if (addend) v += to_f32<T>(addend[o + j]);
      y[o + j] = from_f32<T>(v);
    }
  }
}

// wgrad: dw[t][co][ci] += sum over a slice of output voxels of dy[o][co] * x[gather(o,t)][ci]
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_direct_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, ConvDims dm,
                         int64_t vox_per_block) {
  constexpr int TV = 64;
  __shared__ float Ys[TV][32 + 2];
  __shared__ float Xs[TV][32 + 2];
  const int tid = threadIdx.x;
  const int t = blockIdx.y;
  const int ci_tiles = dm.Cin / 32;
  const int co0 = (blockIdx.z / ci_tiles) * 32, ci0 = (blockIdx.z % ci_tiles) * 32;
  const int kd = t / (dm.k * dm.k), kh = (t / dm.k) % dm.k, kw = t % dm.k;
  const int64_t M = static_cast<int64_t>(dm.N) * dm.Do * dm.Ho * dm.Wo;
  const int64_t mbeg = static_cast<int64_t>(blockIdx.x) * vox_per_block;
  const int64_t mend = min(mbeg + vox_per_block, M);
  const int tco = tid / 16, tci = tid % 16;
  constexpr bool kTwoLevel = std::is_same<T, float>::value;
  using AccT = typename std::conditional<kTwoLevel, double, float>::type;
  AccT acc[2][2] = {{0, 0}, {0, 0}};
  const int lv = tid / 4, lc = (tid % 4) * 8;
  for (int64_t m0 = mbeg; m0 < mend; m0 += TV) {
    const int64_t m = m0 + lv;
    float yv[8], xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) yv[i] = xv[i] = 0.f;
    if (m < mend) {
      load8<T>(dy + m * dm.Cout + co0 + lc, yv);
      int64_t r = m;
      const int ow = static_cast<int>(r % dm.Wo);
      r /= dm.Wo;
      const int oh = static_cast<int>(r % dm.Ho);
      r /= dm.Ho;
      const int od = static_cast<int>(r % dm.Do);
      const int n = static_cast<int>(r / dm.Do);
      const int id = od * dm.SO + kd - dm.pad, ih = oh * dm.SO + kh - dm.pad, iw = ow * dm.SO + kw - dm.pad;
      if (id >= 0 && id < dm.Di && ih >= 0 && ih < dm.Hi && iw >= 0 && iw < dm.Wi)
        load8<T>(x + (((static_cast<int64_t>(n) * dm.Di + id) * dm.Hi + ih) * dm.Wi + iw) * dm.Cin + ci0 + lc, xv);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) Ys[lv][lc + i] = yv[i], Xs[lv][lc + i] = xv[i];
    __syncthreads();
    float part[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) part[i][j] = kTwoLevel ? 0.f : static_cast<float>(acc[i][j]);
#pragma unroll 16
    for (int v = 0; v < TV; ++v) {
      const float2 a = *reinterpret_cast<const float2*>(&Ys[v][tco * 2]);
      const float2 b = *reinterpret_cast<const float2*>(&Xs[v][tci * 2]);
      part[0][0] = fmaf(a.x, b.x, part[0][0]);
      part[0][1] = fmaf(a.x, b.y, part[0][1]);
      part[1][0] = fmaf(a.y, b.x, part[1][0]);
      part[1][1] = fmaf(a.y, b.y, part[1][1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc[i][j] = kTwoLevel ? acc[i][j] + static_cast<AccT>(part[i][j]) : static_cast<AccT>(part[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      atomicAdd(&dw[(static_cast<int64_t>(t) * dm.Cout + co0 + tco * 2 + i) * dm.Cin + ci0 + tci * 2 + j],
                static_cast<float>(acc[i][j]));
}

template <typename T>
int launch_direct(const void* x, const void* wp, const void* addend, void* y, const ConvDims& dm, cudaStream_t s) {
  const int64_t M = static_cast<int64_t>(dm.N) * dm.Ds * dm.Hs * dm.Ws;
  if (M == 0) return MMPL_OK;
  const int mt = ceil_div(M, TM);
  if (dm.Cout % 64 == 0)
    conv_direct_kernel<T, 64><<<dim3(mt, dm.Cout / 64), 256, 0, s>>>(static_cast<const T*>(x), static_cast<const T*>(wp),
                                                                    static_cast<const T*>(addend), static_cast<T*>(y), dm);
  else
    conv_direct_kernel<T, 32><<<dim3(mt, dm.Cout / 32), 256, 0, s>>>(static_cast<const T*>(x), static_cast<const T*>(wp),
                                                                    static_cast<const T*>(addend), static_cast<T*>(y), dm);
  return MMPL_OK;
}

}  // namespace

int conv_out_dim(int in, int k, int stride) { return (in + 2 * (k / 2) - k) / stride + 1; }

static void full_grid(ConvDims& dm) {
  dm.ostep = 1, dm.osd = dm.osh = dm.osw = 0;
  dm.Ds = dm.Do, dm.Hs = dm.Ho, dm.Ws = dm.Wo;
  dm.ntaps = dm.k * dm.k * dm.k;
  for (int t = 0; t < dm.ntaps; ++t) dm.taplist[t] = t;
}

int conv_direct_fprop(const void* x, const void* w, const void* residual, void* y, int n, int d, int h, int wd, int cin,
                      int cout, int k, int stride, int dtype, cudaStream_t s) {
  ConvDims dm{n, d, h, wd, conv_out_dim(d, k, stride), conv_out_dim(h, k, stride), conv_out_dim(wd, k, stride),
              cin, cout, k, k / 2, stride, 1};
  full_grid(dm);
  MMPL_DISPATCH_DTYPE(dtype, T, launch_direct<T>(x, w, residual, y, dm, s));
  MMPL_CHECK_LAUNCH("conv_direct_fprop");
  return MMPL_OK;
}

// dgrad: "input" is dy [n, Do,Ho,Wo, cout], "output" is dx [n, d,h,w, cin]; weights are the dgrad packing.
int conv_direct_dgrad(const void* dy, const void* w, const void* addend, void* dx, int n, int d, int h, int wd, int cin,
                      int cout, int k, int stride, int dtype, cudaStream_t s) {
  ConvDims dm{n, conv_out_dim(d, k, stride), conv_out_dim(h, k, stride), conv_out_dim(wd, k, stride), d, h, wd,
              cout, cin, k, k / 2, 1, stride};
  if (stride == 1) {
    full_grid(dm);
    MMPL_DISPATCH_DTYPE(dtype, T, launch_direct<T>(dy, w, addend, dx, dm, s));
    MMPL_CHECK_LAUNCH("conv_direct_dgrad");
    return MMPL_OK;
  }
  // stride 2: one launch per parity class of dx; a tap t' reaches parity p only if (p + t' - pad) is even per axis
  const int pad = k / 2;
  for (int pc = 0; pc < 8; ++pc) {
    const int pd = pc >> 2, ph = (pc >> 1) & 1, pw = pc & 1;
    dm.ostep = 2, dm.osd = pd, dm.osh = ph, dm.osw = pw;
    dm.Ds = (d - pd + 1) / 2, dm.Hs = (h - ph + 1) / 2, dm.Ws = (wd - pw + 1) / 2;
    dm.ntaps = 0;
    for (int t = 0; t < k * k * k; ++t) {
      const int kd = t / (k * k), kh = (t / k) % k, kw = t % k;
      if (((pd + kd - pad) & 1) == 0 && ((ph + kh - pad) & 1) == 0 && ((pw + kw - pad) & 1) == 0) dm.taplist[dm.ntaps++] = t;
    }
    // parity classes no tap can reach (k = 1: everything but the even class) are plain zeros (+ addend)
    MMPL_DISPATCH_DTYPE(dtype, T, launch_direct<T>(dy, w, addend, dx, dm, s));
    MMPL_CHECK_LAUNCH("conv_direct_dgrad");
  }
  return MMPL_OK;
}

int conv_direct_wgrad(const void* x, const void* dy, float* dw, int n, int d, int h, int wd, int cin, int cout, int k,
                      int stride, int dtype, cudaStream_t s) {
  ConvDims dm{n, d, h, wd, conv_out_dim(d, k, stride), conv_out_dim(h, k, stride), conv_out_dim(wd, k, stride),
              cin, cout, k, k / 2, stride, 1};
  const int taps = k * k * k;
  MMPL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * taps * cout * cin, s));
  const int64_t M = static_cast<int64_t>(n) * dm.Do * dm.Ho * dm.Wo;
  const int tiles = (cout / 32) * (cin / 32);
  int64_t want = static_cast<int64_t>(num_sms()) * 8 / (static_cast<int64_t>(taps) * tiles);
  if (want < 1) want = 1;
  int64_t vpb = (M + want - 1) / want;
  vpb = (vpb + 63) / 64 * 64;
  const int ksplit = ceil_div(M, vpb);
  MMPL_DISPATCH_DTYPE(dtype, T, (conv_wgrad_direct_kernel<T><<<dim3(ksplit, taps, tiles), 256, 0, s>>>(
                                    static_cast<const T*>(x), static_cast<const T*>(dy), dw, dm, vpb)));
  MMPL_CHECK_LAUNCH("conv_direct_wgrad");
  return MMPL_OK;
}

}  // namespace mmpl
