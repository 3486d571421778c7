// Microbenchmarks that decide the design of the lower-bound pass (score_lb_kernel):
//   1. tcgen05.ld throughput per SM for the shapes / warp counts an epilogue can use;
//   2. tcgen05.mma.kind::f16 accumulation error with fp32 and with fp16 accumulators (and the TMEM layout of
//      fp16 accumulators), relative to |a||b| -- the quantity the rigorous margin of the bound is written in.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I compatibility-family-learning_b200/csrc \
//              tools/tmem_probe.cu -o compatibility-family-learning_b200/build/tmem_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "umma.cuh"

using namespace cfl::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

#define R8(v, o)  "=r"(v[o+0]), "=r"(v[o+1]), "=r"(v[o+2]), "=r"(v[o+3]), "=r"(v[o+4]), "=r"(v[o+5]), "=r"(v[o+6]), "=r"(v[o+7])
#define P8(o) "%" #o
__device__ __forceinline__ void ld_x8(uint32_t a, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : R8(v, 0) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_x16(uint32_t a, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : R8(v, 0), R8(v, 8) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_x16p(uint32_t a, uint32_t (&v)[16]) {   // 32 columns, two halfs per register
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : R8(v, 0), R8(v, 8) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_x32(uint32_t a, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : R8(v, 0), R8(v, 8), R8(v, 16), R8(v, 24) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_x64(uint32_t a, uint32_t (&v)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
               "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
               "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
               : R8(v, 0), R8(v, 8), R8(v, 16), R8(v, 24), R8(v, 32), R8(v, 40), R8(v, 48), R8(v, 56) : "r"(a) : "memory");
}

// MODE: 8, 16, 32, 64 = 32x32b.xN ; 17 = x16 with pack::16b (32 columns)
template <int MODE>
__global__ void __launch_bounds__(512, 1)
ldtm_kernel(int iters, int reduce, unsigned long long* cyc, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  constexpr int NREG = MODE == 17 ? 16 : MODE;
  constexpr int NCOL = MODE == 17 ? 32 : MODE;
  uint32_t acc = 0;
  uint32_t col = (uint32_t)((warp >> 2) * NCOL) & 511u;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[NREG];
    if constexpr (MODE == 8) ld_x8(base + col, v);
    else if constexpr (MODE == 16) ld_x16(base + col, v);
    else if constexpr (MODE == 17) ld_x16p(base + col, v);
    else if constexpr (MODE == 32) ld_x32(base + col, v);
    else ld_x64(base + col, v);
    tmem_ld_wait();
    if (reduce) {
#pragma unroll
      for (int j = 0; j < NREG; ++j) acc ^= v[j];
    } else {
      acc ^= v[0] ^ v[NREG - 1];
    }
    col = (col + NCOL * 4) & (512u - NCOL);
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) atomicMax(&cyc[blockIdx.x], (unsigned long long)(t1 - t0));
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int MODE>
static void run_ldtm(int nwarps, int reduce) {
  const int iters = 4096, grid = 148;
  unsigned long long* cyc; uint32_t* sink;
  CK(cudaMalloc(&cyc, grid * 8)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(cyc, 0, grid * 8));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  ldtm_kernel<MODE><<<grid, nwarps * 32>>>(iters, reduce, cyc, sink);      // warm-up
  CK(cudaMemset(cyc, 0, grid * 8));
  cudaEventRecord(e0);
  ldtm_kernel<MODE><<<grid, nwarps * 32>>>(iters, reduce, cyc, sink);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(grid);
  CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
  const int ncol = MODE == 17 ? 32 : MODE;
  const int nreg = MODE == 17 ? 16 : MODE;
  const double cols = (double)nwarps * iters * ncol;                       // 32-lane column reads per CTA
  printf("ldtm mode=%2d%s warps=%2d reduce=%d : %8llu clk/CTA  %6.1f B/clk/SM (TMEM bytes)  %6.1f B/clk/SM (register bytes)  %.3f ms\n",
         MODE == 17 ? 16 : MODE, MODE == 17 ? "p" : " ", nwarps, reduce, mx, cols * 128.0 / mx,
         (double)nwarps * iters * nreg * 128.0 / mx, ms);
  cudaFree(cyc); cudaFree(sink);
}

// ---- MMA accumulation precision: D[128, N] = A[128, Kd] B[N, Kd]^T, fp16 operands ----------------------------
// dfmt: 1 = fp32 accumulators, 0 = fp16 accumulators.  Raw 32-bit TMEM columns are written to D[128][N].
__global__ void __launch_bounds__(160)
mma_f16_kernel(const __half* __restrict__ A, const __half* __restrict__ Bm, uint32_t* __restrict__ D, int N, int Kd, int dfmt) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nks = Kd / 16;
  const uint32_t a_step = 2u * 128u * 16u, b_step = 2u * (uint32_t)N * 16u;
  unsigned char* a_img = smem;
  unsigned char* b_img = smem + (size_t)nks * a_step;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 128) {
    for (int ks = 0; ks < nks; ++ks)
      for (int c = 0; c < 2; ++c) {
        *(uint4*)(a_img + (size_t)ks * a_step + (c * 128 + tid) * 16) = *(const uint4*)(A + (size_t)tid * Kd + ks * 16 + c * 8);
        for (int n = tid; n < N; n += 128)
          *(uint4*)(b_img + (size_t)ks * b_step + ((size_t)c * N + n) * 16) = *(const uint4*)(Bm + (size_t)n * Kd + ks * 16 + c * 8);
      }
    fence_proxy_async();
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 4) tmem_alloc(&slot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 4) {
    if (elect_one()) {
      const uint32_t idesc = ((uint32_t)dfmt << 4) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t ad = make_smem_desc(smem_u32(a_img) + ks * a_step, 128u * 16u, 128u);
        const uint64_t bd = make_smem_desc(smem_u32(b_img) + ks * b_step, (uint32_t)N * 16u, 128u);
        mma_f16(tb, ad, bd, idesc, ks ? 1u : 0u);
      }
      mma_commit(&bar);
    }
  } else {
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      ld_x8(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, ncols);
}

static double half_bits_to_double(uint16_t h) {
  __half_raw r; r.x = h;
  return (double)__half2float(__half(r));
}

static void run_precision(int Kd, double scale, unsigned seed) {
  const int N = 64;
  std::vector<__half> A(128 * Kd), B(N * Kd);
  srand(seed);
  auto gauss = [&]() { double u = (rand() + 1.0) / (RAND_MAX + 2.0), v = (rand() + 1.0) / (RAND_MAX + 2.0); return sqrt(-2 * log(u)) * cos(6.283185307179586 * v); };
  for (auto& x : A) x = __float2half_rn((float)(gauss() * scale));
  for (auto& x : B) x = __float2half_rn((float)(gauss() * scale));
  __half *dA, *dB; uint32_t* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)(Kd / 16) * (2 * 128 * 16 + 2 * N * 16) + 1024;
  CK(cudaFuncSetAttribute(mma_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<double> na(128), nb(N);
  for (int i = 0; i < 128; ++i) { double s = 0; for (int j = 0; j < Kd; ++j) { double x = __half2float(A[i * Kd + j]); s += x * x; } na[i] = sqrt(s); }
  for (int i = 0; i < N; ++i) { double s = 0; for (int j = 0; j < Kd; ++j) { double x = __half2float(B[i * Kd + j]); s += x * x; } nb[i] = sqrt(s); }
  for (int dfmt = 1; dfmt >= 0; --dfmt) {
    CK(cudaMemset(dD, 0xff, 128 * N * 4));
    mma_f16_kernel<<<1, 160, smem>>>(dA, dB, dD, N, Kd, dfmt);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> D(128 * N);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    if (dfmt == 0) {
      printf("  fp16-accumulator raw words row 0: %08x %08x %08x %08x | row 1: %08x %08x\n", D[0], D[1], D[2], D[3], D[N], D[N + 1]);
    }
    double worst = 0, bias = 0, worst_lo = 0, worst_hi = 0;
    for (int i = 0; i < 128; ++i)
      for (int n = 0; n < N; ++n) {
        double ex = 0;
        for (int j = 0; j < Kd; ++j) ex += (double)__half2float(A[i * Kd + j]) * (double)__half2float(B[n * Kd + j]);
        const uint32_t w = D[i * N + n];
        if (dfmt == 1) {
          float f; memcpy(&f, &w, 4);
          const double e = ((double)f - ex) / (na[i] * nb[n]);
          worst = fmax(worst, fabs(e)); bias += e;
        } else {
          const double elo = (half_bits_to_double((uint16_t)(w & 0xffff)) - ex) / (na[i] * nb[n]);
          const double ehi = (half_bits_to_double((uint16_t)(w >> 16)) - ex) / (na[i] * nb[n]);
          worst_lo = fmax(worst_lo, fabs(elo)); worst_hi = fmax(worst_hi, fabs(ehi)); bias += elo;
        }
      }
    if (dfmt == 1)
      printf("  Kd=%3d scale=%g fp32 acc: max |err|/(|a||b|) = %.3e (= %.2f x 2^-24), mean signed %.3e\n", Kd, scale, worst,
             worst * 16777216.0, bias / (128 * N));
    else
      printf("  Kd=%3d scale=%g fp16 acc: max |err|/(|a||b|) low half %.3e (= %.2f x 2^-11), high half %.3e, mean signed %.3e\n",
             Kd, scale, worst_lo, worst_lo * 2048.0, worst_hi, bias / (128 * N));
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main(int argc, char** argv) {
  printf("== tcgen05.ld throughput (148 CTAs, 4096 loads per warp, wait::ld after every load) ==\n");
  for (int reduce = 0; reduce < 2; ++reduce)
    for (int nw : {4, 8, 16}) {
      run_ldtm<8>(nw, reduce); run_ldtm<16>(nw, reduce); run_ldtm<17>(nw, reduce); run_ldtm<32>(nw, reduce); run_ldtm<64>(nw, reduce);
    }
  printf("== tcgen05.mma.kind::f16 accumulation error, fp16 operands exact on the host ==\n");
  run_precision(64, 1.0, 1);
  run_precision(64, 0.05, 2);
  run_precision(128, 1.0, 3);
  run_precision(16, 1.0, 4);
  return 0;
}
