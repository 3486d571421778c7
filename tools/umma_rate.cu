// tcgen05.mma.kind::f16 issue-rate probe: clk per MMA (M = 128, K = 16) for N in {64..256}, operands in shared memory in
// (a) the K-major SWIZZLE_NONE layout the scoring kernels use, (b) K-major SWIZZLE_128B; free-running issue from one
// thread per CTA, one CTA per SM, data = whatever shared memory holds (timing only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I compatibility-family-learning_b200/csrc \
//             tools/umma_rate.cu -o compatibility-family-learning_b200/build/umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "umma.cuh"

using namespace cfl::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;            // SBO = 8 rows x 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

// mode 0: SWIZZLE_NONE, 4 K-steps per "tile" each with its own 4 KB A block and 2*N*16 B block (as score_lb_kernel)
// mode 1: SWIZZLE_128B, 4 K-steps = 32-byte advances inside one 128 B row
// nbuf: accumulator buffers cycled per tile (1 or 2); tf32: kind::tf32 with K = 8 instead
__global__ void __launch_bounds__(128, 1)
rate_kernel(int N, int mode, int tiles, int nbuf, int tf32, unsigned long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a_base = smem_u32(smem), b_base = a_base + 65536;
      const uint32_t idesc = tf32 ? make_idesc_tf32(128, (uint32_t)N) : make_idesc_f16(128, (uint32_t)N);
      const long long t0 = clock64();
      for (int t = 0; t < tiles; ++t) {
        const uint32_t d_tmem = tb + (uint32_t)((t % nbuf) * N);
        const uint32_t a_t = a_base + (uint32_t)((t & 3) * 16384);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ad, bd;
          if (mode == 0) {
            ad = make_smem_desc(a_t + ks * 4096, 128u * 16u, 128u);
            bd = make_smem_desc(b_base + ks * (2u * N * 16u), (uint32_t)N * 16u, 128u);
          } else {
            ad = desc_sw128(a_t + ks * 32);
            bd = desc_sw128(b_base + ks * 32);
          }
          if (tf32) mma_tf32(d_tmem, ad, bd, idesc, ks ? 1u : 0u);
          else mma_f16(d_tmem, ad, bd, idesc, ks ? 1u : 0u);
        }
      }
      mma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// issue timeline of 12 back-to-back MMAs from an idle pipe: clock after every issue, then after completion
__global__ void __launch_bounds__(128, 1)
issue_kernel(int N, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a_base = smem_u32(smem), b_base = a_base + 65536;
      const uint32_t idesc = make_idesc_f16(128, (uint32_t)N);
      unsigned long long ts[16];
      ts[0] = clock64();
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const uint64_t ad = make_smem_desc(a_base + (i & 3) * 4096, 128u * 16u, 128u);
        const uint64_t bd = make_smem_desc(b_base + (i & 3) * (2u * N * 16u), (uint32_t)N * 16u, 128u);
        mma_f16(tb + (uint32_t)((i & 1) * N), ad, bd, idesc, 0u);
        ts[i + 1] = clock64();
      }
      mma_commit(&bar);
      ts[13] = clock64();
      mbar_wait(&bar, 0);
      ts[14] = clock64();
      for (int i = 0; i < 15; ++i) out[i] = ts[i] - ts[0];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// The MMA warp's loop of a scoring kernel, piece by piece: 4 MMAs per tile, then
//   bit 0: tcgen05.commit to a "stage empty" barrier      bit 1: tcgen05.commit to the tile's "tfull" barrier
//   bit 2: an epilogue warp waits on tfull and arrives on tempty; the MMA warp waits on tempty before re-using the
//          accumulator buffer (nbuf buffers)              bit 3: that wait is issued before the tile's LAST MMA
//   bit 4: whole-warp loop with an elected lane per instruction (else one elected thread runs the loop)
__global__ void __launch_bounds__(128, 1)
loop_kernel(int N, int tiles, int nbuf, int what, unsigned long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[16];
  __shared__ uint32_t slot;
  uint64_t* tfull = bars; uint64_t* tempty = bars + 4; uint64_t* sempty = bars + 8; uint64_t* done = bars + 12;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 1); mbar_init(&sempty[i], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const uint32_t a_base = smem_u32(smem), b_base = a_base + 65536;
  const uint32_t idesc = make_idesc_f16(128, (uint32_t)N);
  const uint64_t a_desc = make_smem_desc(a_base, 128u * 16u, 128u), b_desc = make_smem_desc(b_base, (uint32_t)N * 16u, 128u);
  const uint32_t a_lo0 = (uint32_t)a_desc, a_hi = (uint32_t)(a_desc >> 32), b_lo0 = (uint32_t)b_desc, b_hi = (uint32_t)(b_desc >> 32);
  if (warp == 1) {
    const bool whole = what & 16;
    if (whole || elect_one()) {
      const long long t0 = clock64();
      int buf = 0; uint32_t bph = 0;
      for (int t = 0; t < tiles; ++t) {
        const uint32_t d_tmem = tb + (uint32_t)(buf * N);
        int nb = buf + 1; uint32_t nph = bph; if (nb == nbuf) { nb = 0; nph ^= 1u; }
        if ((what & 4) && !(what & 8)) { mbar_wait(&tempty[buf], bph ^ 1u); tc_fence_after(); }
        uint32_t a_lo = a_lo0 + (uint32_t)(t & 3) * 1024u, b_lo = b_lo0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if ((what & 12) == 12 && ks == 3 && t + 1 < tiles) { mbar_wait(&tempty[nb], nph ^ 1u); tc_fence_after(); }
          if (!whole || elect_one()) mma_f16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, ks ? 1u : 0u);
          a_lo += 256u; b_lo += 2u * (uint32_t)N;
        }
        if (what & 1) { if (!whole || elect_one()) mma_commit(&sempty[t & 3]); }
        if (what & 2) { if (!whole || elect_one()) mma_commit(&tfull[buf]); }
        buf = nb; bph = nph;
      }
      if (!whole || elect_one()) mma_commit(done);
      mbar_wait(done, 0);
      const long long t1 = clock64();
      if (lane == 0 || !whole) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
  } else if (warp == 2 && (what & 4)) {
    int buf = 0; uint32_t bph = 0;
    for (int t = 0; t < tiles; ++t) {
      mbar_wait(&tfull[buf], bph);
      tc_fence_after();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
      if (++buf == nbuf) { buf = 0; bph ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  {
    unsigned long long* c; CK(cudaMalloc(&c, 148 * 8));
    const int sm = 65536 + 32768 + 1024;
    CK(cudaFuncSetAttribute(loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    struct { int what; const char* name; } cases[] = {
      {0, "MMAs only"}, {1, "+ commit(stage)"}, {3, "+ commit(stage) + commit(tfull)"},
      {7, "+ hand-shake with an epilogue warp, wait at tile start"}, {15, "+ hand-shake, wait before the last MMA"},
      {16 + 3, "whole-warp loop: commits"}, {16 + 7, "whole-warp loop: hand-shake, wait at tile start"},
      {16 + 15, "whole-warp loop: hand-shake, wait before the last MMA"}};
    for (int N : {144, 192})
      for (int nbuf = 2; nbuf <= (N == 144 ? 3 : 2); ++nbuf)
        for (auto& cs : cases) {
          for (int rep = 0; rep < 2; ++rep) { loop_kernel<<<148, 128, sm>>>(N, 1000, nbuf, cs.what, c); CK(cudaDeviceSynchronize()); }
          unsigned long long h[148]; CK(cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost));
          unsigned long long mx = 0; for (auto x : h) mx = x > mx ? x : mx;
          printf("loop N=%d nbuf=%d %-58s: %6.0f clk/tile (MMA law %3.0f)\n", N, nbuf, cs.name, (double)mx / 1000.0, 4 * (43.0 + N / 2.0));
        }
  }
  {
    unsigned long long* o; CK(cudaMalloc(&o, 16 * 8));
    CK(cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 32768 + 1024));
    for (int rep = 0; rep < 2; ++rep) {
      issue_kernel<<<1, 128, 65536 + 32768 + 1024>>>(192, o);
      CK(cudaDeviceSynchronize());
    }
    unsigned long long h[16]; CK(cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost));
    printf("issue timeline N=192 (clk after issue of MMA 1..12, after commit, after completion):");
    for (int i = 1; i < 15; ++i) printf(" %llu", h[i]);
    printf("\n");
  }

  const int grid = 148, tiles = 2000;
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, grid * 8));
  const int smem = 65536 + 32768 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int tf32 = 0; tf32 < 2; ++tf32)
    for (int mode = 0; mode < 2; ++mode)
      for (int nbuf = 1; nbuf <= 2; ++nbuf)
        for (int N : {64, 96, 128, 192, 256}) {
          if (nbuf * N > 512) continue;
          rate_kernel<<<grid, 128, smem>>>(N, mode, tiles, nbuf, tf32, cyc);
          CK(cudaDeviceSynchronize());
          rate_kernel<<<grid, 128, smem>>>(N, mode, tiles, nbuf, tf32, cyc);
          CK(cudaDeviceSynchronize());
          std::vector<unsigned long long> h(grid);
          CK(cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost));
          unsigned long long mx = 0; for (auto c : h) mx = c > mx ? c : mx;
          const double per = (double)mx / (tiles * 4.0);
          const double kk = tf32 ? 8.0 : 16.0;
          printf("%s %-12s nbuf=%d N=%3d : %6.1f clk/MMA  (nominal %5.1f)  %6.0f flop/clk/SM  operand bytes/clk %5.1f\n",
                 tf32 ? "tf32" : "f16 ", mode ? "SWIZZLE_128B" : "SWIZZLE_NONE", nbuf, N, per, 128.0 * N / 256.0,
                 2.0 * 128 * N * kk / per, (128.0 + N) * 32.0 / per);
        }
  return 0;
}
