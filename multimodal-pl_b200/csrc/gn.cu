// GroupNorm(16)+ReLU forward/backward on NDHWC activations (reference: NoBottleneck.forward, unet3D.py:59-60,64-65;
// downsample.0/1, unet3D.py:645-646).  All kernels are HBM-bound streaming kernels: 16-byte vector loads along the
// channel dimension, fp32 math, fp64 cross-thread accumulation of the statistics.
//
// Thread mapping shared by all kernels: a block of 256 threads covers 256/VPV voxels per iteration, where
// VPV = C / (channels per 16-byte vector); thread t owns vector column t % VPV for the whole kernel, so per-channel
// parameters and partial sums live in registers.
#include <type_traits>

#include "common.cuh"

namespace mmpl {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxC = 512;

template <typename T>
__global__ void __launch_bounds__(kThreads) gn_stats_kernel(const T* __restrict__ x, double* __restrict__ stats,
                                                            int64_t spatial, int C, int groups, int64_t vox_per_block) {
  constexpr int VN = Vec<T>::N;
  const int vpv = C / VN;
  const int cv = threadIdx.x % vpv, vl = threadIdx.x / vpv, vstep = kThreads / vpv;
  const int n = blockIdx.y;
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * vox_per_block;
  const int64_t v1 = min(v0 + vox_per_block, spatial);
  const T* base = x + (static_cast<int64_t>(n) * spatial) * C + cv * VN;
  __shared__ double sg[32][2];
  if (threadIdx.x < 64) (&sg[0][0])[threadIdx.x] = 0.0;
  __syncthreads();
  // each thread sums <= a few hundred values per channel in fp32 (short runs), fp64 is used only across threads
  float ds[VN], dq[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) ds[i] = dq[i] = 0.f;
#pragma unroll 2
  for (int64_t v = v0 + vl; v < v1; v += vstep) {
    Vec<T> a;
    a.load(base + v * C);
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      ds[i] += a.v[i];
      dq[i] = fmaf(a.v[i], a.v[i], dq[i]);
    }
  }
  const int cpg = C / groups;
  // reduce the VN channels of this thread into per-group partials, then across lanes that share the column
  const int gpt = cpg >= VN ? 1 : VN / cpg;  // groups per thread
  const int cpp = VN / gpt;                  // channels per partial
  for (int j = 0; j < gpt; ++j) {
    double a = 0, b = 0;
    for (int i = 0; i < cpp; ++i) a += static_cast<double>(ds[j * cpp + i]), b += static_cast<double>(dq[j * cpp + i]);
    for (int o = 16; o >= vpv && o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const bool leader = vpv >= 32 ? true : ((threadIdx.x & 31) < vpv);
    if (leader) {
      const int g = (cv * VN + j * cpp) / cpg;
      atomicAdd(&sg[g][0], a);
      atomicAdd(&sg[g][1], b);
    }
  }
  __syncthreads();
  if (threadIdx.x < groups * 2) {
    const int g = threadIdx.x >> 1, k = threadIdx.x & 1;
    atomicAdd(&stats[(static_cast<int64_t>(n) * groups + g) * 2 + k], sg[g][k]);
  }
}

__device__ __forceinline__ void mean_rstd(const double* st, double m, float eps, float& mean, float& rstd) {
  const double mu = st[0] / m;
  double var = st[1] / m - mu * mu;
  if (var < 0) var = 0;
  mean = static_cast<float>(mu);
  rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// The ReLU gate of the backward must reproduce the forward's rounding exactly (same fused multiply-add on the same
// folded scale/shift), otherwise elements whose pre-activation is within rounding of zero get a gate that disagrees
// with the stored output.
__device__ __forceinline__ bool relu_gate(float x, float mean, float rstd, float gamma, float beta) {
  const float sc = rstd * gamma;
  const float sh = beta - mean * sc;
  return fmaf(x, sc, sh) > 0.f;
}

// Second head on the even voxels only (the 1x1x1 stride-2 downsample of NoBottleneck, unet3D.py:645-651, reads nothing
// else): its tensor is stored compact, [N][ceil(D/2)][ceil(H/2)][ceil(W/2)][C].  Walks (d, h, w) of a strided voxel run.
struct EvenWalk {
  int d, h, w, H, W, Hc, Wc;
  int64_t sp2;     // voxels of the compact tensor per sample
  __device__ __forceinline__ void init(int64_t v, int D_, int H_, int W_) {
    H = H_, W = W_, Hc = (H_ + 1) >> 1, Wc = (W_ + 1) >> 1;
    sp2 = static_cast<int64_t>((D_ + 1) >> 1) * Hc * Wc;
    const int64_t hw = static_cast<int64_t>(H_) * W_;
    d = static_cast<int>(v / hw);
    const int r = static_cast<int>(v - d * hw);
    h = r / W_, w = r - h * W_;
  }
  __device__ __forceinline__ void step(int n) {
    w += n;
    while (w >= W) {
      w -= W;
      if (++h == H) h = 0, ++d;
    }
  }
  __device__ __forceinline__ bool even() const { return ((d | h | w) & 1) == 0; }
  __device__ __forceinline__ int64_t index() const { return (static_cast<int64_t>(d >> 1) * Hc + (h >> 1)) * Wc + (w >> 1); }
};

template <typename T, bool DUAL>
__global__ void __launch_bounds__(kThreads)
gn_relu_fwd_kernel(const T* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                   const float* __restrict__ beta, T* __restrict__ y, const float* __restrict__ gamma2,
                   const float* __restrict__ beta2, T* __restrict__ y2, int64_t spatial, int C, int groups, int cnt_cpg,
                   float eps, int64_t vox_per_block, int cd, int ch, int cw, int layout) {
  constexpr int VN = Vec<T>::N;
  const int vpv = C / VN;
  const int cv = threadIdx.x % vpv, vl = threadIdx.x / vpv, vstep = kThreads / vpv;
  const int n = blockIdx.y;
  const int cpg = C / groups;
  const double m = static_cast<double>(cnt_cpg) * static_cast<double>(spatial);   // elements that count per group
  float sc[VN], sh[VN], sc2[VN], sh2[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    const int c = cv * VN + i;
    float mean, rstd;
    mean_rstd(stats + (static_cast<int64_t>(n) * groups + c / cpg) * 2, m, eps, mean, rstd);
    sc[i] = rstd * gamma[c];
    sh[i] = beta[c] - mean * sc[i];
    if (DUAL) {
      sc2[i] = rstd * gamma2[c];
      sh2[i] = beta2[c] - mean * sc2[i];
    }
  }
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * vox_per_block;
  const int64_t v1 = min(v0 + vox_per_block, spatial);
  const int64_t off = (static_cast<int64_t>(n) * spatial) * C + cv * VN;
  const bool compact = DUAL && (layout & 2) != 0;   // second head on the even voxels only
  const bool psplit = (layout & 1) != 0;            // first head in the parity-split layout P[pc*N + n][d/2][h/2][w/2][c]
  const bool walk = compact || psplit;
  EvenWalk ew;
  if (walk) ew.init(v0 + vl, cd, ch, cw);
  for (int64_t v = v0 + vl; v < v1; v += vstep) {
    Vec<T> a, o;
    a.load(x + off + v * C);
#pragma unroll
    for (int i = 0; i < VN; ++i) o.v[i] = fmaxf(fmaf(a.v[i], sc[i], sh[i]), 0.f);
    if (psplit) {
      const int pc = ((ew.d & 1) << 2) | ((ew.h & 1) << 1) | (ew.w & 1);
      o.store(y + ((static_cast<int64_t>(pc) * gridDim.y + n) * ew.sp2 + ew.index()) * C + cv * VN);
    } else {
      o.store(y + off + v * C);
    }
    if (DUAL) {
      if (!compact || ew.even()) {
#pragma unroll
        for (int i = 0; i < VN; ++i) o.v[i] = fmaxf(fmaf(a.v[i], sc2[i], sh2[i]), 0.f);
        o.store(compact ? y2 + (n * ew.sp2 + ew.index()) * C + cv * VN : y2 + off + v * C);
      }
    }
    if (walk) ew.step(vstep);
  }
}

// Workspace of the backward: double ws[N][C][6] (+ one trailing double used as the last-block ticket).  Per head h:
//   ws[n][c][2h]   = S1 = sum_v g            (g = dy * [relu gate])
//   ws[n][c][2h+1] = Q  = gamma_c * sum_v g*xhat
//   ws[n][c][4+h]  = sum_v g*xhat, accumulated by the apply pass only for channels with gamma_c == 0 (Q is useless there)
// The sums come either from pass 1 below or from the epilogue of the kernel that produced dy (conv_tc.cu, cls_bwd).
constexpr int kWs = 6;

// Pass 1 of the backward: per (n, c) sums S1 and Q for each head -> ws (fp64).
template <typename T, bool DUAL>
__global__ void __launch_bounds__(kThreads)
gn_relu_bwd_reduce_kernel(const T* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                          const float* __restrict__ beta, const T* __restrict__ dy, const float* __restrict__ gamma2,
                          const float* __restrict__ beta2, const T* __restrict__ dy2, double* __restrict__ ws,
                          int64_t spatial, int C, int groups, int cnt_cpg, float eps, int64_t vox_per_block, int cd, int ch,
                          int cw) {
  using V = VecH<T>;   // 8-byte vectors: half the per-thread channel state -> higher occupancy
  constexpr int VN = V::N;
  constexpr int NH = DUAL ? 2 : 1;
  const int vpv = C / VN;
  const int cv = threadIdx.x % vpv, vl = threadIdx.x / vpv, vstep = kThreads / vpv;
  const int n = blockIdx.y;
  const int cpg = C / groups;
  const double m = static_cast<double>(cnt_cpg) * static_cast<double>(spatial);   // elements that count per group
  float mu[VN], rs[VN], ga[NH][VN], be[NH][VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    const int c = cv * VN + i;
    mean_rstd(stats + (static_cast<int64_t>(n) * groups + c / cpg) * 2, m, eps, mu[i], rs[i]);
    ga[0][i] = gamma[c], be[0][i] = beta[c];
    if (DUAL) ga[1][i] = gamma2[c], be[1][i] = beta2[c];
  }
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * vox_per_block;
  const int64_t v1 = min(v0 + vox_per_block, spatial);
  const int64_t off = (static_cast<int64_t>(n) * spatial) * C + cv * VN;
  float d1[NH][VN], d2[NH][VN];   // fp32 over this thread's short voxel run; fp64 across threads below
#pragma unroll
  for (int hh = 0; hh < NH; ++hh)
#pragma unroll
    for (int i = 0; i < VN; ++i) d1[hh][i] = d2[hh][i] = 0.f;
  const bool compact = DUAL && cw > 0;
  EvenWalk ew;
  if (compact) ew.init(v0 + vl, cd, ch, cw);
#pragma unroll 2
  for (int64_t v = v0 + vl; v < v1; v += vstep) {
    V a, g;
    a.load(x + off + v * C);
    g.load(dy + off + v * C);
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      const float xh = (a.v[i] - mu[i]) * rs[i];
      const float gg = relu_gate(a.v[i], mu[i], rs[i], ga[0][i], be[0][i]) ? g.v[i] : 0.f;
      d1[0][i] += gg;
      d2[0][i] = fmaf(gg, xh, d2[0][i]);
    }
    if (DUAL) {
      if (!compact || ew.even()) {
        g.load(compact ? dy2 + (n * ew.sp2 + ew.index()) * C + cv * VN : dy2 + off + v * C);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const float xh = (a.v[i] - mu[i]) * rs[i];
          const float gg = relu_gate(a.v[i], mu[i], rs[i], ga[NH - 1][i], be[NH - 1][i]) ? g.v[i] : 0.f;
          d1[NH - 1][i] += gg;
          d2[NH - 1][i] = fmaf(gg, xh, d2[NH - 1][i]);
        }
      }
      if (compact) ew.step(vstep);
    }
  }
  __shared__ double sc[kMaxC][4];
  for (int i = threadIdx.x; i < C * 4; i += kThreads) (&sc[0][0])[i] = 0.0;
  __syncthreads();
  const bool leader = vpv >= 32 ? true : ((threadIdx.x & 31) < vpv);
#pragma unroll
  for (int hh = 0; hh < NH; ++hh)
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      double a = static_cast<double>(d1[hh][i]), b = static_cast<double>(d2[hh][i]);
      for (int o = 16; o >= vpv && o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if (leader) {
        atomicAdd(&sc[cv * VN + i][hh * 2 + 0], a);
        atomicAdd(&sc[cv * VN + i][hh * 2 + 1], b);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2 * NH; i += kThreads) {
    const int c = i / (2 * NH), k = i % (2 * NH);
    const double gam = (k & 1) ? static_cast<double>((k >> 1) == 0 ? gamma[c] : gamma2[c]) : 1.0;
    atomicAdd(&ws[(static_cast<int64_t>(n) * C + c) * kWs + k], sc[c][k] * gam);
  }
}

// Pass 2: dx = sum_heads rstd*(gamma*g - m1 - xhat*m2) (+ addend), m1 = mean_group(gamma*S1), m2 = mean_group(Q).
// The last block to finish emits dbeta = sum_n S1 and dgamma = sum_n Q / gamma.
template <typename T, bool DUAL, bool ADD>
__global__ void __launch_bounds__(kThreads)
gn_relu_bwd_apply_kernel(const T* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const T* __restrict__ dy, const float* __restrict__ gamma2,
                         const float* __restrict__ beta2, const T* __restrict__ dy2, const T* __restrict__ addend,
                         T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ dgamma2, float* __restrict__ dbeta2, double* __restrict__ ws,
                         int N, int64_t spatial, int C, int groups, int cnt_cpg, float eps, int64_t vox_per_block, int cd,
                         int ch, int cw) {
  using V = VecH<T>;
  constexpr int VN = V::N;
  constexpr int NH = DUAL ? 2 : 1;
  const int vpv = C / VN;
  const int cv = threadIdx.x % vpv, vl = threadIdx.x / vpv, vstep = kThreads / vpv;
  const int n = blockIdx.y;
  const int cpg = C / groups;
  const double m = static_cast<double>(cnt_cpg) * static_cast<double>(spatial);   // elements that count per group
  __shared__ float gm[32][4];  // per group: m1,m2 per head (already divided by m)
  __shared__ bool s_last;
  if (threadIdx.x < groups * NH) {
    const int g = threadIdx.x / NH, hh = threadIdx.x % NH;
    const float* gam = hh == 0 ? gamma : gamma2;
    double a = 0, b = 0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const double* w = ws + (static_cast<int64_t>(n) * C + c) * kWs + hh * 2;
      a += static_cast<double>(gam[c]) * w[0];
      b += w[1];
    }
    gm[g][hh * 2 + 0] = static_cast<float>(a / m);
    gm[g][hh * 2 + 1] = static_cast<float>(b / m);
  }
  __syncthreads();
  // Per channel and head the whole update folds into   dx += cA * g + cB * x + cC   with
  //   cA = rstd*gamma,  cB = -rstd^2 * m2,  cC = rstd * (rstd*mean*m2 - m1),
  // and the ReLU gate is the forward's own fused multiply-add (sc = rstd*gamma = cA, sh = beta - mean*sc):
  // 2 FMAs for the gate/select and 2 for the update per element and head.
  float mu[VN], rs[VN], ga[NH][VN], sh[NH][VN], cB[NH][VN], cC[NH][VN];
  float cA[NH][VN];
  bool any_zero_gamma = false;
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    const int c = cv * VN + i, g = c / cpg;
    mean_rstd(stats + (static_cast<int64_t>(n) * groups + g) * 2, m, eps, mu[i], rs[i]);
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      const float gam = (hh == 0 ? gamma : gamma2)[c], bet = (hh == 0 ? beta : beta2)[c];
      const float m1 = gm[g][hh * 2 + 0], m2 = gm[g][hh * 2 + 1];
      ga[hh][i] = gam;
      cA[hh][i] = rs[i] * gam;                      // == the forward's scale
      sh[hh][i] = bet - mu[i] * cA[hh][i];          // == the forward's shift
      cB[hh][i] = -rs[i] * rs[i] * m2;
      cC[hh][i] = rs[i] * (rs[i] * mu[i] * m2 - m1);
      any_zero_gamma |= gam == 0.f;
    }
  }
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * vox_per_block;
  const int64_t v1 = min(v0 + vox_per_block, spatial);
  const int64_t off = (static_cast<int64_t>(n) * spatial) * C + cv * VN;
  // The streaming loop exists twice: the plain one, and (block-uniform choice, practically never taken) one that also
  // accumulates sum g*xhat for the channels of this block whose gamma is exactly 0.
  auto stream = [&](auto exc_tag) {
    constexpr bool EXC = decltype(exc_tag)::value;
    float ex[NH][VN];
#pragma unroll
    for (int hh = 0; hh < NH; ++hh)
#pragma unroll
      for (int i = 0; i < VN; ++i) ex[hh][i] = 0.f;
    const bool compact = DUAL && cw > 0;
    EvenWalk ew;
    if (compact) ew.init(v0 + vl, cd, ch, cw);
#pragma unroll 2
    for (int64_t v = v0 + vl; v < v1; v += vstep) {
      V a, g, g2, ad, o;
      a.load(x + off + v * C);
      g.load(dy + off + v * C);
      if (DUAL) {
        if (!compact) {
          g2.load(dy2 + off + v * C);
        } else {
          // the gradient of the compact head exists on the even voxels only; the m1 / m2 terms of that head still reach
          // every voxel (its statistics are those of the whole tensor)
          if (ew.even()) {
            g2.load(dy2 + (n * ew.sp2 + ew.index()) * C + cv * VN);
          } else {
#pragma unroll
            for (int i = 0; i < VN; ++i) g2.v[i] = 0.f;
          }
          ew.step(vstep);
        }
      }
      if (ADD) ad.load(addend + off + v * C);
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        float r = fmaf(cB[0][i], a.v[i], cC[0][i]);
        const float gg = fmaf(a.v[i], cA[0][i], sh[0][i]) > 0.f ? g.v[i] : 0.f;
        r = fmaf(cA[0][i], gg, r);
        if (EXC) ex[0][i] = fmaf(gg, (a.v[i] - mu[i]) * rs[i], ex[0][i]);
        if (DUAL) {
          r += fmaf(cB[NH - 1][i], a.v[i], cC[NH - 1][i]);
          const float gg2 = fmaf(a.v[i], cA[NH - 1][i], sh[NH - 1][i]) > 0.f ? g2.v[i] : 0.f;
          r = fmaf(cA[NH - 1][i], gg2, r);
          if (EXC) ex[NH - 1][i] = fmaf(gg2, (a.v[i] - mu[i]) * rs[i], ex[NH - 1][i]);
        }
        if (ADD) r += ad.v[i];
        o.v[i] = r;
      }
      o.store(dx + off + v * C);
    }
    if (EXC) {
#pragma unroll
      for (int hh = 0; hh < NH; ++hh)
#pragma unroll
        for (int i = 0; i < VN; ++i)
          if (ga[hh][i] == 0.f)
            atomicAdd(&ws[(static_cast<int64_t>(n) * C + cv * VN + i) * kWs + 4 + hh], static_cast<double>(ex[hh][i]));
    }
  };
  if (__syncthreads_or(any_zero_gamma ? 1 : 0))
    stream(std::true_type{});
  else
    stream(std::false_type{});
  // last block: parameter gradients
  unsigned int* ticket = reinterpret_cast<unsigned int*>(ws + static_cast<int64_t>(N) * C * kWs);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    const volatile double* vws = ws;
    for (int i = threadIdx.x; i < C * NH; i += kThreads) {
      const int c = i / NH, hh = i % NH;
      const float gam = (hh == 0 ? gamma : gamma2)[c];
      double a = 0, b = 0;
      for (int nn = 0; nn < N; ++nn) {
        const volatile double* w = vws + (static_cast<int64_t>(nn) * C + c) * kWs;
        a += w[hh * 2];
        b += gam != 0.f ? w[hh * 2 + 1] / static_cast<double>(gam) : w[4 + hh];
      }
      (hh == 0 ? dbeta : dbeta2)[c] = static_cast<float>(a);
      (hh == 0 ? dgamma : dgamma2)[c] = static_cast<float>(b);
    }
    // Leave the workspace clean (sums and ticket): producers of a later backward through the same node accumulate into
    // it again.  Every other block has passed its ticket, i.e. is done with the workspace -- no separate memset node.
    __syncthreads();
    for (int64_t i = threadIdx.x; i < static_cast<int64_t>(N) * C * kWs; i += kThreads) ws[i] = 0.0;
    if (threadIdx.x == 0) *ticket = 0u;
  }
}

int check_shape(int c, int groups, int dtype, bool half = false) {
  const int vn = (dtype == MMPL_BF16 ? 8 : 4) / (half ? 2 : 1);
  MMPL_REQUIRE(groups > 0 && groups <= 32 && c % groups == 0, MMPL_E_SHAPE, "GroupNorm: C=%d groups=%d", c, groups);
  MMPL_REQUIRE(c % vn == 0 && c <= kMaxC && kThreads % (c / vn) == 0, MMPL_E_SHAPE,
               "GroupNorm: C=%d must be a power-of-two multiple of %d and <= %d", c, vn, kMaxC);
  const int cpg = c / groups;
  MMPL_REQUIRE(cpg >= vn ? cpg % vn == 0 : vn % cpg == 0, MMPL_E_SHAPE, "GroupNorm: C/groups=%d vs vector %d", cpg, vn);
  return MMPL_OK;
}

// Blocks per sample: enough to give every SM several blocks, at least 64 voxel-rows per block.
void plan(int n, int64_t spatial, int c, int dtype, int& blocks_x, int64_t& vox_per_block, bool half = false) {
  const int vn = (dtype == MMPL_BF16 ? 8 : 4) / (half ? 2 : 1);
  const int vstep = kThreads / (c / vn);
  int64_t want = static_cast<int64_t>(num_sms()) * 8 / (n > 0 ? n : 1);
  if (want < 1) want = 1;
  int64_t vpb = (spatial + want - 1) / want;
  const int64_t min_vpb = static_cast<int64_t>(vstep) * 16;
  if (vpb < min_vpb) vpb = min_vpb;
  vpb = (vpb + vstep - 1) / vstep * vstep;
  vox_per_block = vpb;
  blocks_x = ceil_div(spatial, vpb);
}

template <typename T>
int launch_gn_bwd(const void* x, const double* stats, const float* gamma, const float* beta, const void* dy,
                  const float* gamma2, const float* beta2, const void* dy2, const void* addend, void* dx,
                  float* dgamma, float* dbeta, float* dgamma2, float* dbeta2, double* workspace, int reduced, int n,
                  int64_t spatial, int c, int groups, int cnt_cpg, float eps, int bx, int64_t vpb, int cd, int ch, int cw,
                  cudaStream_t s) {
  const bool dual = gamma2 != nullptr;
  const bool add = addend != nullptr;
  const T* xx = static_cast<const T*>(x);
  const T* g1 = static_cast<const T*>(dy);
  const T* g2 = static_cast<const T*>(dy2);
  const T* ad = static_cast<const T*>(addend);
  T* out = static_cast<T*>(dx);
  const dim3 grid(bx, n);
  if (!reduced) {
    if (dual)
      gn_relu_bwd_reduce_kernel<T, true><<<grid, kThreads, 0, s>>>(xx, stats, gamma, beta, g1, gamma2, beta2, g2,
                                                                  workspace, spatial, c, groups, cnt_cpg, eps, vpb, cd, ch, cw);
    else
      gn_relu_bwd_reduce_kernel<T, false><<<grid, kThreads, 0, s>>>(xx, stats, gamma, beta, g1, nullptr, nullptr,
                                                                   nullptr, workspace, spatial, c, groups, cnt_cpg, eps, vpb, 0, 0, 0);
    MMPL_CHECK_LAUNCH("gn_relu_bwd_reduce");
  }
  if (dual && add)
    gn_relu_bwd_apply_kernel<T, true, true><<<grid, kThreads, 0, s>>>(xx, stats, gamma, beta, g1, gamma2, beta2, g2,
        ad, out, dgamma, dbeta, dgamma2, dbeta2, workspace, n, spatial, c, groups, cnt_cpg, eps, vpb, cd, ch, cw);
  else if (dual)
    gn_relu_bwd_apply_kernel<T, true, false><<<grid, kThreads, 0, s>>>(xx, stats, gamma, beta, g1, gamma2, beta2, g2,
        ad, out, dgamma, dbeta, dgamma2, dbeta2, workspace, n, spatial, c, groups, cnt_cpg, eps, vpb, cd, ch, cw);
  else if (add)
    gn_relu_bwd_apply_kernel<T, false, true><<<grid, kThreads, 0, s>>>(xx, stats, gamma, beta, g1, gamma2, beta2, g2,
        ad, out, dgamma, dbeta, dgamma2, dbeta2, workspace, n, spatial, c, groups, cnt_cpg, eps, vpb, cd, ch, cw);
  else
    gn_relu_bwd_apply_kernel<T, false, false><<<grid, kThreads, 0, s>>>(xx, stats, gamma, beta, g1, gamma2, beta2, g2,
        ad, out, dgamma, dbeta, dgamma2, dbeta2, workspace, n, spatial, c, groups, cnt_cpg, eps, vpb, cd, ch, cw);
  return MMPL_OK;
}

}  // namespace
}  // namespace mmpl

using namespace mmpl;

extern "C" int mmpl_gn_stats(const void* x, double* stats, int n, int64_t spatial, int c, int groups, int dtype,
                             mmpl_stream_t stream) {
  if (int e = check_shape(c, groups, dtype)) return e;
  int bx;
  int64_t vpb;
  plan(n, spatial, c, dtype, bx, vpb);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  MMPL_DISPATCH_DTYPE(dtype, T, (gn_stats_kernel<T><<<dim3(bx, n), kThreads, 0, s>>>(
                                    static_cast<const T*>(x), stats, spatial, c, groups, vpb)));
  MMPL_CHECK_LAUNCH("gn_stats");
  return MMPL_OK;
}

extern "C" int mmpl_gn_relu_fwd(const void* x, const double* stats, const float* gamma, const float* beta, void* y,
                                const float* gamma2, const float* beta2, void* y2, int n, int64_t spatial, int c,
                                int groups, int real_cpg, int vol_d, int vol_h, int vol_w, int layout, float eps,
                                int dtype, mmpl_stream_t stream) {
  if (int e = check_shape(c, groups, dtype)) return e;
  MMPL_REQUIRE((layout & ~3) == 0, MMPL_E_SHAPE, "GroupNorm: layout=%d", layout);
  MMPL_REQUIRE(layout == 0 || static_cast<int64_t>(vol_d) * vol_h * vol_w == spatial, MMPL_E_SHAPE,
               "GroupNorm: layout %d needs the volume extents (%d,%d,%d) of the %lld voxels", layout, vol_d, vol_h, vol_w,
               (long long)spatial);
  MMPL_REQUIRE(!(layout & 2) || y2 != nullptr, MMPL_E_SHAPE, "GroupNorm: compact second head without a second head");
  MMPL_REQUIRE(!(layout & 1) || ((vol_d | vol_h | vol_w) & 1) == 0, MMPL_E_SHAPE,
               "GroupNorm: the parity-split layout needs even extents, got (%d,%d,%d)", vol_d, vol_h, vol_w);
  MMPL_REQUIRE(real_cpg >= 0 && real_cpg <= c / groups, MMPL_E_SHAPE, "GroupNorm: real_cpg=%d of %d", real_cpg, c / groups);
  const int cnt_cpg = real_cpg > 0 ? real_cpg : c / groups;
  int bx;
  int64_t vpb;
  plan(n, spatial, c, dtype, bx, vpb);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool dual = gamma2 != nullptr;
  MMPL_DISPATCH_DTYPE(dtype, T, {
    if (dual)
      gn_relu_fwd_kernel<T, true><<<dim3(bx, n), kThreads, 0, s>>>(static_cast<const T*>(x), stats, gamma, beta,
                                                                  static_cast<T*>(y), gamma2, beta2,
                                                                  static_cast<T*>(y2), spatial, c, groups, cnt_cpg, eps, vpb,
                                                                  vol_d, vol_h, vol_w, layout);
    else
      gn_relu_fwd_kernel<T, false><<<dim3(bx, n), kThreads, 0, s>>>(static_cast<const T*>(x), stats, gamma, beta,
                                                                   static_cast<T*>(y), nullptr, nullptr, nullptr,
                                                                   spatial, c, groups, cnt_cpg, eps, vpb, vol_d, vol_h, vol_w,
                                                                   layout);
  });
  MMPL_CHECK_LAUNCH("gn_relu_fwd");
  return MMPL_OK;
}

extern "C" int mmpl_gn_relu_bwd(const void* x, const double* stats, const float* gamma, const float* beta,
                                const void* dy, const float* gamma2, const float* beta2, const void* dy2,
                                const void* addend, void* dx, float* dgamma, float* dbeta, float* dgamma2,
                                float* dbeta2, double* workspace, int reduced, int n, int64_t spatial, int c, int groups,
                                int real_cpg, int vol_d, int vol_h, int vol_w, int layout, float eps, int dtype,
                                mmpl_stream_t stream) {
  if (int e = check_shape(c, groups, dtype, true)) return e;
  MMPL_REQUIRE((layout & ~2) == 0, MMPL_E_SHAPE, "GroupNorm backward: layout=%d (only the compact second head, 2)", layout);
  MMPL_REQUIRE(layout == 0 || (dy2 != nullptr && static_cast<int64_t>(vol_d) * vol_h * vol_w == spatial), MMPL_E_SHAPE,
               "GroupNorm: compact second head needs the volume extents (%d,%d,%d) of the %lld voxels", vol_d, vol_h, vol_w,
               (long long)spatial);
  const int y2_d = layout ? vol_d : 0, y2_h = layout ? vol_h : 0, y2_w = layout ? vol_w : 0;
  MMPL_REQUIRE(real_cpg >= 0 && real_cpg <= c / groups, MMPL_E_SHAPE, "GroupNorm: real_cpg=%d of %d", real_cpg, c / groups);
  const int cnt_cpg = real_cpg > 0 ? real_cpg : c / groups;
  int bx;
  int64_t vpb;
  plan(n, spatial, c, dtype, bx, vpb, true);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t ws_bytes = sizeof(double) * (static_cast<size_t>(kWs) * n * c + 1);
  if (!reduced) MMPL_CUDA(cudaMemsetAsync(workspace, 0, ws_bytes, s));
  int rc = MMPL_OK;
  MMPL_DISPATCH_DTYPE(dtype, T, rc = (launch_gn_bwd<T>(x, stats, gamma, beta, dy, gamma2, beta2, dy2, addend, dx, dgamma,
                                                     dbeta, dgamma2, dbeta2, workspace, reduced, n, spatial, c, groups,
                                                     cnt_cpg, eps, bx, vpb, y2_d, y2_h, y2_w, s)));
  if (rc) return rc;
  MMPL_CHECK_LAUNCH("gn_relu_bwd_apply");      // its last block leaves the workspace zeroed
  return MMPL_OK;
}
