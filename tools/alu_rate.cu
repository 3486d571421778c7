// Issue-rate probe for the packed FP32x2 ops of the scoring epilogues: clk per warp instruction per SM sub-partition
// (1 or 4 warps per sub-partition, 8 independent accumulator chains per thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/alu_rate.cu -o compatibility-family-learning_b200/build/alu_rate
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) { f2_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2_t add2(f2_t a, f2_t b) { f2_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float max2(float a, float b) { float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int OP>
__global__ void k(int iters, unsigned long long* cyc, float* sink, float seed) {
  f2_t a[8]; float s[8];
  for (int i = 0; i < 8; ++i) { s[i] = seed + i + threadIdx.x; a[i] = ((f2_t)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] * 0.5f); }
  const f2_t m = a[3];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) a[i] = fma2(a[i], m, a[i]);
      if (OP == 1) a[i] = add2(a[i], m);
      if (OP == 2) s[i] = fma1(s[i], seed, s[i]);
      if (OP == 3) s[i] = max3(s[i], s[(i + 1) & 7], seed);
      if (OP == 4) s[i] = max2(s[i], seed);
      if (OP == 5) a[i] = fma2(a[i], a[i], m);          // the epilogue's shape: c*c + acc
    }
  }
  const long long t1 = clock64();
  float r = 0; for (int i = 0; i < 8; ++i) r += s[i] + __uint_as_float((unsigned)a[i]);
  if (r == 1234.5f) sink[0] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name) {
  unsigned long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  for (int warps : {4, 16}) {
    const int iters = 4096;
    k<OP><<<148, warps * 32>>>(iters, cyc, sink, 1.0001f); cudaDeviceSynchronize();
    k<OP><<<148, warps * 32>>>(iters, cyc, sink, 1.0001f); cudaDeviceSynchronize();
    unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
    printf("%-10s %2d warps/SM: %.2f clk per warp-instruction per sub-partition\n", name, warps, (double)mx / (iters * 8.0 * (warps / 4)));
  }
}
int main() { run<0>("FFMA2"); run<5>("FFMA2 c*c+a"); run<1>("FADD2"); run<2>("FFMA"); run<3>("FMNMX3"); run<4>("FMNMX"); return 0; }
