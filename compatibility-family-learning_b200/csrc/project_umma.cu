// tcgen05 (3xTF32) projection kernel + the single-tile self-test of the UMMA building blocks.
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"

namespace cfl {

using namespace umma;

// ---- self-test: D[128,N] = A[128,Kd] * B[N,Kd]^T on one CTA ---------------------------------
// 160 threads: warps 0-3 stage the operands (split hi/lo, canonical layout) and read TMEM back,
// warp 4 owns TMEM allocation and issues the MMAs.
__global__ void __launch_bounds__(160)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D,
                     int N, int Kd, int variant) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nks = (Kd + 7) / 8;
  const uint32_t a_step = 4u * 128u * 16u;          // [hl][chunk][128][16B] = 8 KB
  const uint32_t b_step = 4u * (uint32_t)N * 16u;   // [hl][chunk][N][16B]
  unsigned char* a_img = smem;
  unsigned char* b_img = smem + (size_t)nks * a_step;

  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 128) {
    for (int ks = 0; ks < nks; ++ks) {
      for (int c = 0; c < 2; ++c) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        int k0 = ks * 8 + c * 4;
        float* xp = &x.x;
        for (int i = 0; i < 4; ++i) if (k0 + i < Kd) xp[i] = A[(size_t)tid * Kd + k0 + i];
        float4 hi, lo;
        split_tf32x4(x, hi, lo);
        *(float4*)(a_img + (size_t)ks * a_step + ((0 * 2 + c) * 128 + tid) * 16) = hi;
        *(float4*)(a_img + (size_t)ks * a_step + ((1 * 2 + c) * 128 + tid) * 16) = lo;
      }
      for (int n = tid; n < N; n += 128) {
        for (int c = 0; c < 2; ++c) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          int k0 = ks * 8 + c * 4;
          float* xp = &x.x;
          for (int i = 0; i < 4; ++i) if (k0 + i < Kd) xp[i] = Bm[(size_t)n * Kd + k0 + i];
          float4 hi, lo;
          split_tf32x4(x, hi, lo);
          *(float4*)(b_img + (size_t)ks * b_step + ((size_t)(0 * 2 + c) * N + n) * 16) = hi;
          *(float4*)(b_img + (size_t)ks * b_step + ((size_t)(1 * 2 + c) * N + n) * 16) = lo;
        }
      }
    }
    fence_proxy_async();
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 4) tmem_alloc(&tmem_base_s, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc_tf32(128, (uint32_t)N);
    for (int ks = 0; ks < nks; ++ks) {
      uint32_t a_addr = smem_u32(a_img) + ks * a_step;
      uint32_t b_addr = smem_u32(b_img) + ks * b_step;
      if (variant == 0) {
        mma_step_3xtf32(tmem_base, a_addr, b_addr, (uint32_t)N, idesc, ks == 0);
      } else {
        // variant 1: LBO/SBO roles swapped (diagnostic only)
        uint64_t a_hi = make_smem_desc(a_addr, 128u, 128u * 16u);
        uint64_t a_lo = make_smem_desc(a_addr + 2u * 128u * 16u, 128u, 128u * 16u);
        uint64_t b_hi = make_smem_desc(b_addr, 128u, (uint32_t)N * 16u);
        uint64_t b_lo = make_smem_desc(b_addr + 2u * N * 16u, 128u, (uint32_t)N * 16u);
        mma_tf32(tmem_base, a_lo, b_hi, idesc, ks == 0 ? 0u : 1u);
        mma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
        mma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (warp < 4) {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) D[(size_t)row * N + c0 + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, ncols);
}

// placeholders until the projection kernel lands
int project_fwd_umma(const float*, int64_t, int, int64_t, const float*, int, int64_t, const float*,
                     const float*, float, int, float*, int64_t, float*, float*, void*, size_t,
                     cudaStream_t) { return CFL_ERR_UNSUPPORTED; }
size_t project_fwd_umma_workspace(int64_t, int, int) { return 0; }
bool project_fwd_umma_supported(const float*, int64_t, int, int64_t, int) { return false; }

}  // namespace cfl

using namespace cfl;

extern "C" int cfl_selftest_umma(const float* A, const float* Bm, float* D, int N, int Kd, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(A && Bm && D, CFL_ERR_INVALID, "selftest_umma: NULL argument");
  CFL_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, CFL_ERR_INVALID, "selftest_umma: N must be a multiple of 16 in [16,256]");
  CFL_REQUIRE(Kd >= 1 && Kd <= 64, CFL_ERR_INVALID, "selftest_umma: Kd must be in [1,64]");
  int nks = (Kd + 7) / 8;
  size_t smem = (size_t)nks * (4 * 128 * 16 + 4 * (size_t)N * 16);
  const char* v = getenv("CFL_UMMA_VARIANT");
  int variant = v ? atoi(v) : 0;
  CFL_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(A, Bm, D, N, Kd, variant);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}
