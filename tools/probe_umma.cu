// Hardware probe for sm_100a: settles the tcgen05 shared-memory-descriptor and TMA semantics the conv kernels
// rely on (row-shifted views of a halo block, strides that are not a multiple of the swizzle repeat,
// MN-major operands with overlapping LBO chunks, TMA traversal strides).  It is a development tool, not product.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I multimodal-pl_b200/csrc tools/probe_umma.cu -o tools/probe_umma
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <string>
#include <vector>

#include "ptx.cuh"

using namespace mmpl::ptx;

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

struct Op {
  uint64_t adesc, bdesc;
  uint32_t idesc, tmem_col, acc, pad;
};

__device__ bool wait_capped(uint64_t* bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

__global__ void __launch_bounds__(128, 1)
playground(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int a0, int a1, int a2,
           int b0, int b1, int b2, uint32_t bytesA, uint32_t bytesB, uint32_t offB, const Op* ops, int nops,
           float* dump, int ncols, uint8_t* smem_dump, int smem_dump_bytes, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar_tma, bytesA + bytesB);
    tma_load_3d(smem, &tmA, &bar_tma, a0, a1, a2);
    if (bytesB) tma_load_3d(smem + offB, &tmB, &bar_tma, b0, b1, b2);
    if (!wait_capped(&bar_tma, 0)) status[0] = 1;
    // give over-delivering loads (traversal-stride probe) time to land
    for (int i = 0; i < 20000; ++i) __nanosleep(100);
  }
  __syncthreads();
  if (smem_dump) {
    for (int i = threadIdx.x; i < smem_dump_bytes; i += blockDim.x) smem_dump[i] = smem[i];
  }
  __syncthreads();
  if (nops > 0) {
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t base = smem_u32(smem);
      for (int i = 0; i < nops; ++i) {
        Op op = ops[i];
        uint64_t ad = op.adesc + static_cast<uint64_t>((base >> 4) & 0x3FFF);
        uint64_t bd = op.bdesc + static_cast<uint64_t>((base >> 4) & 0x3FFF);
        umma_f16(tmem_base + op.tmem_col, ad, bd, op.idesc, op.acc);
      }
      umma_commit(&bar_mma);
      if (!wait_capped(&bar_mma, 0)) status[0] = 2;
    }
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) dump[(warp * 32 + lane) * ncols + c0 + j] = __uint_as_float(r[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
  if (threadIdx.x == 0) status[1] = static_cast<int>(smem_u32(smem));
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;

static CUtensorMap make_map3d(void* ptr, const int dims[3], const int box[3], int swz, const int es[3]) {
  CUtensorMap m;
  cuuint64_t gd[3] = {(cuuint64_t)dims[0], (cuuint64_t)dims[1], (cuuint64_t)dims[2]};
  cuuint64_t gs[2] = {(cuuint64_t)dims[0] * 2, (cuuint64_t)dims[0] * dims[1] * 2};
  cuuint32_t bx[3] = {(cuuint32_t)box[0], (cuuint32_t)box[1], (cuuint32_t)box[2]};
  cuuint32_t st[3] = {(cuuint32_t)es[0], (cuuint32_t)es[1], (cuuint32_t)es[2]};
  CUtensorMapSwizzle s = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, gd, gs, bx, st, CU_TENSOR_MAP_INTERLEAVE_NONE, s,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("  cuTensorMapEncodeTiled failed: %d\n", (int)r);
    memset(&m, 0, sizeof(m));
  }
  return m;
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Tensor3 {  // global bf16 tensor [R2][R1][C], C innermost
  int C, R1, R2;
  std::vector<float> h;
  __nv_bfloat16* d = nullptr;
  void init(int c, int r1, int r2, unsigned seed) {
    C = c, R1 = r1, R2 = r2;
    h.resize((size_t)c * r1 * r2);
    std::vector<__nv_bfloat16> hb(h.size());
    srand(seed);
    for (size_t i = 0; i < h.size(); ++i) {
      float v = bf16r((float)(rand() % 2001 - 1000) / 1000.0f);
      h[i] = v;
      hb[i] = __float2bfloat16(v);
    }
    CK(cudaMalloc(&d, hb.size() * 2));
    CK(cudaMemcpy(d, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  }
  float at(int c, int r1, int r2) const {
    if (c < 0 || c >= C || r1 < 0 || r1 >= R1 || r2 < 0 || r2 >= R2) return 0.f;
    return h[((size_t)r2 * R1 + r1) * C + c];
  }
};

// Logical (un-swizzled) image of a box as TMA lays it out: [b2][b1][b0] with b0 innermost.
struct BoxImage {
  int b0, b1, b2;
  std::vector<float> v;
  float at_byte(long off) const {
    long e = off / 2;
    if (e < 0 || e >= (long)v.size()) return NAN;
    return v[e];
  }
};
static BoxImage box_image(const Tensor3& t, const int box[3], int c0, int c1, int c2, const int es[3]) {
  BoxImage b;
  b.b0 = box[0], b.b1 = box[1], b.b2 = box[2];
  b.v.resize((size_t)box[0] * box[1] * box[2]);
  for (int k = 0; k < box[2]; ++k)
    for (int j = 0; j < box[1]; ++j)
      for (int i = 0; i < box[0]; ++i)
        b.v[((size_t)k * box[1] + j) * box[0] + i] = t.at(c0 + i * es[0], c1 + j * es[1], c2 + k * es[2]);
  return b;
}

struct Runner {
  Op* d_ops;
  float* d_dump;
  uint8_t* d_sdump;
  int* d_status;
  std::vector<float> dump;
  std::vector<uint8_t> sdump;
  int status[2];
  Runner() {
    CK(cudaMalloc(&d_ops, sizeof(Op) * 4096));
    CK(cudaMalloc(&d_dump, sizeof(float) * 128 * 512));
    CK(cudaMalloc(&d_sdump, 200 * 1024));
    CK(cudaMalloc(&d_status, 8));
    CK(cudaFuncSetAttribute(playground, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
  }
  bool run(const CUtensorMap& mA, const CUtensorMap& mB, const int ca[3], const int cb[3], uint32_t bytesA,
           uint32_t bytesB, uint32_t offB, const std::vector<Op>& ops, int ncols, int sdump_bytes) {
    CK(cudaMemset(d_status, 0, 8));
    CK(cudaMemset(d_dump, 0xFF, sizeof(float) * 128 * 512));
    if (!ops.empty()) CK(cudaMemcpy(d_ops, ops.data(), sizeof(Op) * ops.size(), cudaMemcpyHostToDevice));
    playground<<<1, 128, 200 * 1024 + 1024>>>(mA, mB, ca[0], ca[1], ca[2], cb[0], cb[1], cb[2], bytesA, bytesB, offB,
                                              d_ops, (int)ops.size(), d_dump, ncols, sdump_bytes ? d_sdump : nullptr,
                                              sdump_bytes, d_status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("  kernel error: %s\n", cudaGetErrorString(e));
      exit(3);
    }
    CK(cudaMemcpy(status, d_status, 8, cudaMemcpyDeviceToHost));
    dump.resize(128 * (size_t)ncols);
    if (ncols) CK(cudaMemcpy(dump.data(), d_dump, sizeof(float) * 128 * ncols, cudaMemcpyDeviceToHost));
    sdump.resize(sdump_bytes);
    if (sdump_bytes) CK(cudaMemcpy(sdump.data(), d_sdump, sdump_bytes, cudaMemcpyDeviceToHost));
    if (status[0]) printf("  [timeout code %d]\n", status[0]);
    return status[0] == 0;
  }
};

static float bf16_from_bytes(const uint8_t* p) {
  uint16_t u = (uint16_t)p[0] | ((uint16_t)p[1] << 8);
  uint32_t w = (uint32_t)u << 16;
  float f;
  memcpy(&f, &w, 4);
  return f;
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  g_encode = (EncodeFn)fn;
  if (!g_encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 2;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  Runner R;
  const int one[3] = {1, 1, 1};

  // ------------------------------------------------------------------ T0: TMA swizzle + traversal-stride models
  for (int swz : {128, 64}) {
    const int KC = swz / 2;
    Tensor3 X;
    X.init(KC, 40, 40, 11);
    int dims[3] = {KC, 40, 40}, box[3] = {KC, 10, 4};
    CUtensorMap m = make_map3d(X.d, dims, box, swz, one);
    int ca[3] = {0, 3, 5}, cb[3] = {0, 0, 0};
    uint32_t bytes = KC * 2 * 10 * 4;
    R.run(m, m, ca, cb, bytes, 0, 0, {}, 0, bytes);
    BoxImage img = box_image(X, box, 0, 3, 5, one);
    int bad = 0;
    for (uint32_t off = 0; off < bytes; off += 2) {
      uint32_t phys = swz == 128 ? (off ^ (((off >> 7) & 7) << 4)) : (off ^ (((off >> 7) & 3) << 4));
      if (bf16_from_bytes(&R.sdump[phys]) != img.at_byte(off)) ++bad;
    }
    printf("T0 TMA swizzle model swz=%d (smem base & 1023 = %d): mismatches=%d -> %s\n", swz, R.status[1] & 1023, bad,
           bad ? "MODEL WRONG" : "ok");
  }
  {
    Tensor3 X;
    X.init(64, 40, 40, 12);
    int dims[3] = {64, 40, 40}, box[3] = {64, 16, 8}, es[3] = {1, 2, 2};
    CUtensorMap m = make_map3d(X.d, dims, box, 128, es);
    int ca[3] = {0, 1, 3}, cb[3] = {0, 0, 0};
    const uint32_t small = 64 * 2 * 8 * 4, big = 64 * 2 * 16 * 8;
    R.run(m, m, ca, cb, small, 0, 0, {}, 0, big);
    // H1: box counts source elements -> smem holds ceil(16/2) x ceil(8/2) rows
    int boxH1[3] = {64, 8, 4};
    BoxImage i1 = box_image(X, boxH1, 0, 1, 3, es);
    BoxImage i2 = box_image(X, box, 0, 1, 3, es);  // H2: box counts output elements
    int bad1 = 0, bad2 = 0;
    for (uint32_t off = 0; off < small; off += 2) {
      uint32_t phys = off ^ (((off >> 7) & 7) << 4);
      if (bf16_from_bytes(&R.sdump[phys]) != i1.at_byte(off)) ++bad1;
    }
    for (uint32_t off = 0; off < big; off += 2) {
      uint32_t phys = off ^ (((off >> 7) & 7) << 4);
      if (bf16_from_bytes(&R.sdump[phys]) != i2.at_byte(off)) ++bad2;
    }
    printf("T0 TMA elementStrides=(1,2,2) box(64,16,8): H1[box in source elems, 8x4 rows] mismatches=%d ; "
           "H2[box in output elems, 16x8 rows] mismatches=%d\n", bad1, bad2);
  }

  // ------------------------------------------------------------------ T1: K-major views (fprop / dgrad A operand)
  for (int swz : {128, 64}) {
    const int KC = swz / 2, RB = swz, N = 32;
    Tensor3 X, W;
    X.init(KC, 40, 40, 21);
    W.init(KC, N, 1, 22);
    int dimsA[3] = {KC, 40, 40}, boxA[3] = {KC, 20, 20};
    int dimsB[3] = {KC, N, 1}, boxB[3] = {KC, N, 1};
    CUtensorMap mA = make_map3d(X.d, dimsA, boxA, swz, one), mB = make_map3d(W.d, dimsB, boxB, swz, one);
    int ca[3] = {0, -1, 2}, cb[3] = {0, 0, 0};
    BoxImage ia = box_image(X, boxA, ca[0], ca[1], ca[2], one), ib = box_image(W, boxB, 0, 0, 0, one);
    const uint32_t bytesA = 400 * RB, bytesB = N * RB, offB = ((bytesA + 1023) / 1024) * 1024;
    const uint32_t lay = swz == 128 ? SWZ_128B : SWZ_64B;
    for (int sbo_rows : {8, 10, 16})
      for (int shift : {0, 1, 2, 3, 5, 8})
        for (int policy : {0, 1}) {
          std::vector<Op> ops;
          for (int ks = 0; ks < KC / 16; ++ks) {
            uint32_t startA = shift * RB + ks * 32, startB = offB + ks * 32;
            // policy 1: base offset from the absolute address bits (smem base is 1024-aligned)
            uint32_t boA = policy ? ((startA >> 7) & 7) : 0;
            Op op;
            op.adesc = make_smem_desc(startA, 16, sbo_rows * RB, lay, boA);
            op.bdesc = make_smem_desc(startB, 16, 8 * RB, lay, 0);
            op.idesc = make_idesc_bf16(128, N, 0, 0);
            op.tmem_col = 0;
            op.acc = ks > 0;
            op.pad = 0;
            ops.push_back(op);
          }
          R.run(mA, mB, ca, cb, bytesA, bytesB, offB, ops, N, 0);
          double maxerr = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
              double acc = 0;
              for (int k = 0; k < KC; ++k)
                acc += (double)ia.at_byte((long)shift * RB + (m / 8) * (long)sbo_rows * RB + (m % 8) * RB + k * 2) *
                       ib.at_byte((long)n * RB + k * 2);
              double e = fabs(acc - R.dump[m * N + n]);
              if (!(e <= maxerr)) maxerr = e;
            }
          printf("T1 K-major swz=%d sbo_rows=%d shift=%d base_offset_policy=%d : maxerr=%.4g %s\n", swz, sbo_rows,
                 shift, policy, maxerr, maxerr < 1e-2 ? "OK" : "BAD");
        }
  }

  // ------------------------------------------------------------------ T2: MN-major operands (wgrad)
  for (int swz : {128, 64}) {
    const int CH = swz / 2, RB = swz, N = CH;  // N = one swizzle span of channels
    Tensor3 X, Y;
    X.init(CH, 40, 40, 31);
    Y.init(CH, 40, 40, 32);
    int dims[3] = {CH, 40, 40}, box[3] = {CH, 20, 20};
    CUtensorMap mA = make_map3d(X.d, dims, box, swz, one), mB = make_map3d(Y.d, dims, box, swz, one);
    int ca[3] = {0, 0, 1}, cb[3] = {0, 2, 0};
    BoxImage ia = box_image(X, box, ca[0], ca[1], ca[2], one), ib = box_image(Y, box, cb[0], cb[1], cb[2], one);
    const uint32_t bytesA = 400 * RB, offB = ((bytesA + 1023) / 1024) * 1024;
    const uint32_t lay = swz == 128 ? SWZ_128B : SWZ_64B;
    const int MCH = 128 / CH;  // M chunks per MMA
    struct Cfg {
      const char* name;
      int lbo_rows, sbo_rowsA, shiftA, sbo_rowsB, shiftB, policy;
    };
    std::vector<Cfg> cfgs = {
        {"standard: chunks 64 rows apart, dense k-groups", 64, 8, 0, 8, 0, 0},
        {"chunks 64 rows apart, A shifted 1 row, bo=0", 64, 8, 1, 8, 0, 0},
        {"chunks 64 rows apart, A shifted 1 row, bo=addr", 64, 8, 1, 8, 0, 1},
        {"chunks 64 rows apart, A shifted 3 rows, bo=addr", 64, 8, 3, 8, 0, 1},
        {"overlapping chunks LBO=1 row (kw-packed M), bo=0", 1, 8, 0, 8, 0, 0},
        {"overlapping chunks LBO=1 row, sboA=10 rows, bo=0", 1, 10, 0, 8, 0, 0},
        {"overlapping chunks LBO=1 row, sboA=10 rows, shiftA=10, bo=addr", 1, 10, 10, 8, 0, 1},
        {"chunks 64 rows apart, sboA=10 rows, bo=0", 64, 10, 0, 8, 0, 0},
        {"chunks 64 rows apart, sboA=10 rows, shiftA=11, sboB=16, bo=addr", 64, 10, 11, 16, 0, 1},
        {"chunks 64 rows apart, sboA=16, sboB=16 shiftB=2 bo=addr", 64, 16, 0, 16, 2, 1},
    };
    for (const Cfg& c : cfgs) {
      const int KSTEPS = 4;
      std::vector<Op> ops;
      for (int ks = 0; ks < KSTEPS; ++ks) {
        uint32_t startA = (c.shiftA + ks * 2 * c.sbo_rowsA) * RB;
        uint32_t startB = offB + (c.shiftB + ks * 2 * c.sbo_rowsB) * RB;
        Op op;
        op.adesc = make_smem_desc(startA, c.lbo_rows * RB, c.sbo_rowsA * RB, lay, c.policy ? ((startA >> 7) & 7) : 0);
        op.bdesc = make_smem_desc(startB, 64 * RB, c.sbo_rowsB * RB, lay, c.policy ? ((startB >> 7) & 7) : 0);
        op.idesc = make_idesc_bf16(128, N, 1, 1);
        op.tmem_col = 0;
        op.acc = ks > 0;
        op.pad = 0;
        ops.push_back(op);
      }
      R.run(mA, mB, ca, cb, bytesA, bytesA, offB, ops, N, 0);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double acc = 0;
          for (int k = 0; k < 16 * KSTEPS; ++k) {
            long offA = ((long)c.shiftA + (m / CH) * (long)c.lbo_rows + (k / 8) * (long)c.sbo_rowsA + (k % 8)) * RB +
                        (m % CH) * 2;
            long offBb = ((long)c.shiftB + (k / 8) * (long)c.sbo_rowsB + (k % 8)) * RB + (n % CH) * 2;
            acc += (double)ia.at_byte(offA) * ib.at_byte(offBb);
          }
          double e = fabs(acc - R.dump[m * N + n]);
          if (!(e <= maxerr)) maxerr = e;
        }
      (void)MCH;
      printf("T2 MN-major swz=%d [%s] : maxerr=%.4g %s\n", swz, c.name, maxerr, maxerr < 1e-2 ? "OK" : "BAD");
    }
  }
  printf("probe done\n");
  return 0;
}
