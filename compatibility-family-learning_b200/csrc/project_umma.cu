// tcgen05 (3xTF32) projection kernel + the single-tile self-test of the UMMA building blocks.
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"
#include "tmap.cuh"

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

// ---- projection kernel ------------------------------------------------------------------------
// C[B, N] = x[B, F] * V[F, N]:  M = 128 items per tile (TMEM lanes), N = N_out padded to 16,
// MMA-K = the feature dimension F (8 per K-step).  Persistent CTAs (grid = min(tiles, SMs)).
//   warps 0-3   epilogue: tcgen05.ld 16 columns at a time for the thread's item -> in_scale,
//               weight-norm scaler, bias, activation -> y (+ pre, z)
//   warps 4-11  A producers: thread (row r, half h) streams its row of x with 4 batches (64 B each)
//               of register prefetch, splits hi/lo and writes the canonical layout
//   warp 12     one lane issues tcgen05.mma (3 per K-step) / tcgen05.commit
//   warp 13     one lane drives the TMA engine for the pre-split weight image (B operand)
// Bytes per item: 4F (read once) + 4N (written once); flops 2*F*N (x3 MMAs).
constexpr int PU_THREADS = 14 * 32;
// A-operand block of one K-step: 4 planes [hi c0][hi c1][lo c0][lo c1] of 128 rows x 16 B.  The plane
// stride (= the descriptor's LBO) and the block stride are padded so that the producers' coalesced
// mapping (4 lanes per row) stores without shared-memory bank conflicts: 8 lanes of a store phase
// write 2 rows x 4 planes, and 2080 B / 8384 B shift the planes by 8 / 16 banks.
constexpr uint32_t PU_APLANE = 128u * 16u + 32u;              // 2080
constexpr uint32_t PU_ASTAGE = 4u * PU_APLANE + 64u;          // 8384
constexpr int PU_PF = 4;

struct PuArgs {
  CUtensorMap tmx;                                         // TMA-staged kernel: x [B, F], box = 128 rows x 8*kps columns
  const float* x; int64_t B; int F; int64_t ldx;
  const unsigned char* vimg; int N, Npad, nks, nst, kps, nh, nbuf;
  const float* scaler; const float* bias; float in_scale; int act;
  float* y; int64_t ldy; float* pre; float* z;
  int64_t tiles;
  int rs;                                                  // raw-tile ring stages (TMA-staged kernel)
};

__global__ void pack_weights_kernel(const float* __restrict__ V, int F, int N, int64_t ldV, int Npad,
                                    float* __restrict__ img) {
  const int nks = F / 8;
  const int64_t total = (int64_t)nks * 2 * Npad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(e % Npad);
    const int c = (int)((e / Npad) % 2);
    const int ks = (int)(e / (2 * Npad));
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    float* xp = &x.x;
    if (n < N) {
#pragma unroll
      for (int i = 0; i < 4; ++i) xp[i] = V[(int64_t)(ks * 8 + c * 4 + i) * ldV + n];
    }
    float4 hi, lo;
    split_tf32x4(x, hi, lo);
    const size_t step = (size_t)ks * 4 * Npad;
    ((float4*)img)[step + (size_t)(0 * 2 + c) * Npad + n] = hi;
    ((float4*)img)[step + (size_t)(1 * 2 + c) * Npad + n] = lo;
  }
}

// Epilogue shared by both projection kernels (warps 0-3): tcgen05.ld 16 columns at a time for the thread's item ->
// sum of the split accumulators -> in_scale, weight-norm scaler, bias, activation -> y (+ pre, z).
__device__ __forceinline__ void project_epilogue(const PuArgs& A, int warp, int lane, int my_tiles, int64_t first, int64_t stride,
                                                 uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty, const float* scs,
                                                 const float* bis, int NH, int NBUF, int acc_cols, int nks) {
  const int Npad = A.Npad;
    // ------------------------------------ epilogue ------------------------------------
    const int lrow = warp * 32 + lane;
    const bool vec = ((A.ldy & 3) == 0) && (((uintptr_t)A.y & 15u) == 0) &&
                     (!A.pre || ((uintptr_t)A.pre & 15u) == 0) && (!A.z || ((uintptr_t)A.z & 15u) == 0);
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t % NBUF;
      const int64_t row = (first + (int64_t)t * stride) * 128 + lrow;
      mbar_wait(&tfull[buf], (uint32_t)(t / NBUF) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * acc_cols);
      const bool two_hi = NH == 2 && nks >= 2;                  // second hi accumulator was written
      for (int c0 = 0; c0 < Npad; c0 += 16) {
        float acc[16], part[16];
        tmem_ld16(taddr + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (two_hi) {
          tmem_ld16(taddr + (uint32_t)(Npad + c0), part);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += part[j];
        }
        tmem_ld16(taddr + (uint32_t)(NH * Npad + c0), part);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += part[j];
        if (row < A.B) {
          float yv[16], pv[16], zv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            zv[j] = acc[j] * A.in_scale;
            pv[j] = fmaf(zv[j], scs[c0 + j], bis[c0 + j]);
            yv[j] = apply_act(pv[j], A.act);
          }
          float* yr = A.y + row * A.ldy + c0;
          if (vec && c0 + 16 <= A.N) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *(float4*)(yr + j) = make_float4(yv[j], yv[j + 1], yv[j + 2], yv[j + 3]);
            if (A.pre) {
              float* pr = A.pre + row * A.ldy + c0;
#pragma unroll
              for (int j = 0; j < 16; j += 4) *(float4*)(pr + j) = make_float4(pv[j], pv[j + 1], pv[j + 2], pv[j + 3]);
            }
            if (A.z) {
              float* zr = A.z + row * A.ldy + c0;
#pragma unroll
              for (int j = 0; j < 16; j += 4) *(float4*)(zr + j) = make_float4(zv[j], zv[j + 1], zv[j + 2], zv[j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (c0 + j < A.N) {
                yr[j] = yv[j];
                if (A.pre) A.pre[row * A.ldy + c0 + j] = pv[j];
                if (A.z) A.z[row * A.ldy + c0 + j] = zv[j];
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
    }
  }

__global__ void __launch_bounds__(PU_THREADS, 1)
project_umma_kernel(PuArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Npad = A.Npad, nks = A.nks, NST = A.nst, KPS = A.kps;   // KPS K-steps per ring stage (2 or 4)
  const uint32_t bbytes = 4u * (uint32_t)Npad * 16u;
  const uint32_t stage_bytes = (uint32_t)KPS * (PU_ASTAGE + bbytes);
  const uint32_t b_off = (uint32_t)KPS * PU_ASTAGE;               // B blocks follow the A blocks of a stage
  const int spt = (nks + KPS - 1) / KPS;                         // stages per tile
  unsigned char* ring = smem;
  float* scs = (float*)(smem + (size_t)NST * stage_bytes);
  float* bis = scs + Npad;
  uint64_t* full = (uint64_t*)(bis + Npad);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  // The tensor core's fp32 accumulate loses about half an ulp per MMA (measured as a systematic
  // low bias that grows with F), so one output is NOT accumulated in a single TMEM column: the
  // hi*hi products of alternating K-steps go to NH accumulators and the two small cross terms
  // (lo*hi, hi*lo) to their own one; the epilogue adds the NACC partials in registers (RN).
  const int NH = A.nh, NACC = NH + 1, NBUF = A.nbuf;
  const int acc_cols = NACC * Npad;                              // TMEM columns per tile buffer
  uint32_t ncols = 32;
  while ((int)ncols < NBUF * acc_cols) ncols <<= 1;
  if (warp == 12) {
    if (lane == 0) {
      // producers: with 4 K-steps per stage both halves (256 threads) fill a stage, with 2 only one half
      const uint32_t arrivals = (KPS == 4 ? 256u : 128u) + 1u;
      for (int s = 0; s < NST; ++s) { mbar_init(&full[s], arrivals); mbar_init(&empty[s], 1); }
      mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], 128); mbar_init(&tempty[1], 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ncols);
  }
  for (int i = tid; i < Npad; i += PU_THREADS) {
    scs[i] = (A.scaler && i < A.N) ? A.scaler[i] : 1.0f;
    bis[i] = (A.bias && i < A.N) ? A.bias[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int my_tiles = (int)((A.tiles - first + stride - 1) / stride);

  if (warp == 12) {
    if (elect_one()) {
      Step3Desc sd = make_step3((uint32_t)Npad, make_idesc_tf32(128, (uint32_t)Npad));
      sd.a_hi = make_smem_desc(0, PU_APLANE, 128u);
      sd.a_lo = make_smem_desc(2u * PU_APLANE, PU_APLANE, 128u);
      const uint32_t ring_u = smem_u32(ring);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int buf = t % NBUF;
        mbar_wait(&tempty[buf], ((uint32_t)(t / NBUF) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(buf * acc_cols);
        const uint32_t d_lo = d_base + (uint32_t)(NH * Npad);
        int kidx = 0;
        for (int s = 0; s < spt; ++s) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = ring_u + stage * stage_bytes;
          const int nv = (nks - s * KPS) < KPS ? (nks - s * KPS) : KPS;
          for (int j = 0; j < nv; ++j, ++kidx) {
            const uint64_t ao = (uint64_t)((sa + j * PU_ASTAGE) >> 4), bo = (uint64_t)((sa + b_off + j * bbytes) >> 4);
            mma_tf32(d_lo, sd.a_lo + ao, sd.b_hi + bo, sd.idesc, kidx == 0 ? 0u : 1u);
            mma_tf32(d_lo, sd.a_hi + ao, sd.b_lo + bo, sd.idesc, 1u);
            mma_tf32(d_base + (uint32_t)((kidx % NH) * Npad), sd.a_hi + ao, sd.b_hi + bo, sd.idesc, kidx < NH ? 0u : 1u);
          }
          mma_commit(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        mma_commit(&tfull[buf]);
      }
    }
  } else if (warp == 13) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int s = 0; s < spt; ++s) {
          const int nv = (nks - s * KPS) < KPS ? (nks - s * KPS) : KPS;
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], (uint32_t)nv * bbytes);
          bulk_g2s(ring + stage * stage_bytes + b_off, A.vimg + (size_t)s * KPS * bbytes, (uint32_t)nv * bbytes, &full[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------- A producers ----------------------------------
    // Batch = 4 K-steps = 128 B of every row.  Half h (warps 4-7 / 8-11) owns K-steps 2h, 2h+1 of
    // each batch; inside a half, warp w covers rows 32w..32w+31 with 4 coalesced LDG.128: lane l of
    // load i reads 16 B chunk (l & 3) of row 32w + 8i + (l >> 2)  (4 lanes = 64 contiguous bytes).
    const int pt = tid - 128;
    const int h = pt >> 7, w = (pt >> 5) & 3, l = pt & 31;
    const int cc = l & 3;                                        // 16-byte chunk inside the half's 64 B
    const int kj = cc >> 1, kc = cc & 1;                         // K-step within the half, chunk within K-step
    const int rbase = 32 * w + (l >> 2);
    const int nb = (nks + 3) / 4;
    const int total = my_tiles * nb;
    float4 q[PU_PF][4];
    int lt = 0, lb = 0;
    auto load_item = [&](float4 (&dst)[4]) {
      const int ks = lb * 4 + 2 * h + kj;
      const int64_t row0 = (first + (int64_t)lt * stride) * 128 + rbase;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + 8 * i;
        dst[i] = (row < A.B && ks < nks) ? __ldg((const float4*)(A.x + row * A.ldx + (int64_t)ks * 8 + kc * 4))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (++lb == nb) { lb = 0; ++lt; }
    };
#pragma unroll
    for (int u = 0; u < PU_PF; ++u) { if (u < total) load_item(q[u]); }
    int ptile = 0, pb = 0;
    for (int base = 0; base < total; base += PU_PF) {
#pragma unroll
      for (int u = 0; u < PU_PF; ++u) {
        const int i = base + u;
        if (i < total) {
          float4 v[4] = {q[u][0], q[u][1], q[u][2], q[u][3]};
          if (i + PU_PF < total) load_item(q[u]);
          const int sl = (KPS == 4) ? pb : (2 * pb + h);               // stage index within the tile
          if (sl < spt) {
            const int sidx = ptile * spt + sl;
            const int stage = sidx % NST;
            const uint32_t phase = (uint32_t)((sidx / NST) & 1);
            const int slot = ((KPS == 4) ? 2 * h : 0) + kj;            // K-step slot inside the stage
            float4 hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32x4(v[e], hh[e], ll[e]);
            mbar_wait(&empty[stage], phase ^ 1u);
            unsigned char* st = ring + stage * stage_bytes + slot * PU_ASTAGE + kc * PU_APLANE;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              *(float4*)(st + (rbase + 8 * e) * 16) = hh[e];
              *(float4*)(st + 2 * PU_APLANE + (rbase + 8 * e) * 16) = ll[e];
            }
            fence_proxy_async();
            mbar_arrive(&full[stage]);
          }
          if (++pb == nb) { pb = 0; ++ptile; }
        }
      }
    }
  } else {
    project_epilogue(A, warp, lane, my_tiles, first, stride, tmem_base, tfull, tempty, scs, bis, NH, NBUF, acc_cols, nks);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, ncols);
}

// ---- projection kernel, x staged by the TMA engine ----------------------------------------------------------------------
// project_umma_kernel above is bound by its producers' own instruction stream (LDG -> registers -> split -> STS with the
// load latency on the chain: ~2300 clk per 16 KB stage against 900 clk of MMA at N_out = 64, profiles/r2_06).  Here warp 14
// streams the RAW x tile of a stage (128 items x 8*KPS features, one cp.async.bulk.tensor per stage, rows beyond B and
// columns beyond F arrive as zeros) into a deep ring, and the producers only re-lay it: each warp reads 512 contiguous bytes
// (conflict-free LDS.128: 4 items x 8 chunks), splits hi/lo with integer rounding and stores the canonical K-major layout
// (plane / K-step strides padded so that the 8 (K-step, chunk) targets of a quarter-warp fall into different banks).
//   warps 0-3 epilogue | 4-11 producers | 12 MMA issue | 13 weight image (bulk copy) | 14 raw x tiles (tensor map)
constexpr int PT_THREADS = 15 * 32;
constexpr uint32_t PT_APLANE = 128u * 16u + 64u;             // 2112: chunk planes shifted by 16 banks
constexpr uint32_t PT_ASTAGE = 4u * PT_APLANE + 16u;         // 8464: K-steps shifted by 4 banks

__global__ void __launch_bounds__(PT_THREADS, 1)
project_umma_tma_kernel(const __grid_constant__ PuArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Npad = A.Npad, nks = A.nks, NST = A.nst, KPS = A.kps, RS = A.rs;
  const uint32_t bbytes = 4u * (uint32_t)Npad * 16u;
  const uint32_t stage_bytes = ((uint32_t)KPS * (PT_ASTAGE + bbytes) + 127u) & ~127u;
  const uint32_t b_off = (uint32_t)KPS * PT_ASTAGE;               // B blocks follow the A blocks of a stage
  const uint32_t raw_bytes = 128u * 32u * (uint32_t)KPS;          // [128 items][8*KPS floats]
  const int spt = (nks + KPS - 1) / KPS;                         // stages per tile
  unsigned char* ring = smem;
  unsigned char* raw = smem + (size_t)NST * stage_bytes;
  float* scs = (float*)(raw + (size_t)RS * raw_bytes);
  float* bis = scs + Npad;
  uint64_t* full = (uint64_t*)(bis + Npad);
  uint64_t* empty = full + NST;
  uint64_t* rfull = empty + NST;
  uint64_t* rempty = rfull + RS;
  uint64_t* tfull = rempty + RS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int NH = A.nh, NACC = NH + 1, NBUF = A.nbuf;
  const int acc_cols = NACC * Npad;
  uint32_t ncols = 32;
  while ((int)ncols < NBUF * acc_cols) ncols <<= 1;
  if (warp == 12) {
    if (lane == 0) {
      for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 256u + 1u); mbar_init(&empty[s], 1); }
      for (int s = 0; s < RS; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rempty[s], 256u); }
      mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], 128); mbar_init(&tempty[1], 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ncols);
  }
  for (int i = tid; i < Npad; i += PT_THREADS) {
    scs[i] = (A.scaler && i < A.N) ? A.scaler[i] : 1.0f;
    bis[i] = (A.bias && i < A.N) ? A.bias[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int my_tiles = first < A.tiles ? (int)((A.tiles - first + stride - 1) / stride) : 0;

  if (warp == 12) {
    if (elect_one()) {
      // 32-bit descriptor words: only the start-address field moves (by constants inside a stage)
      const uint32_t idesc = make_idesc_tf32(128, (uint32_t)Npad);
      const uint32_t hiw = desc_hi_word(128u);
      const uint32_t ring_u = smem_u32(ring);
      const uint32_t a_w0 = desc_lo_word(ring_u, PT_APLANE);                       // A hi plane of K-step 0, stage 0
      const uint32_t b_w0 = desc_lo_word(ring_u + b_off, (uint32_t)Npad * 16u);     // B hi plane of K-step 0, stage 0
      const uint32_t a_lo_d = (2u * PT_APLANE) >> 4, b_lo_d = 2u * (uint32_t)Npad; // hi -> lo plane
      const uint32_t b_step4 = bbytes >> 4, st4 = stage_bytes >> 4;
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int buf = t % NBUF;
        mbar_wait(&tempty[buf], ((uint32_t)(t / NBUF) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(buf * acc_cols);
        const uint32_t d_lo = d_base + (uint32_t)(NH * Npad);
        uint32_t d_hi = d_base; int h = 0;                          // hi*hi accumulator of the K-step (rotates over NH)
        uint32_t seen = 0;                                          // K-steps issued so far (saturating at NH)
        for (int s = 0; s < spt; ++s) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const int nv = (nks - s * KPS) < KPS ? (nks - s * KPS) : KPS;
          uint32_t aw = a_w0 + (uint32_t)stage * st4, bw = b_w0 + (uint32_t)stage * st4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nv) {
              mma_tf32_w(d_lo, aw + a_lo_d, hiw, bw, hiw, idesc, seen ? 1u : 0u);
              mma_tf32_w(d_lo, aw, hiw, bw + b_lo_d, hiw, idesc, 1u);
              mma_tf32_w(d_hi, aw, hiw, bw, hiw, idesc, seen >= (uint32_t)NH ? 1u : 0u);
              aw += PT_ASTAGE >> 4; bw += b_step4;
              if (seen < (uint32_t)NH) ++seen;
              d_hi += (uint32_t)Npad;
              if (++h == NH) { h = 0; d_hi = d_base; }
            }
          }
          mma_commit(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        mma_commit(&tfull[buf]);
      }
    }
  } else if (warp == 13) {
    if (elect_one()) {                                            // weight image: one bulk copy per stage
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int s = 0; s < spt; ++s) {
          const int nv = (nks - s * KPS) < KPS ? (nks - s * KPS) : KPS;
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], (uint32_t)nv * bbytes);
          bulk_g2s(ring + stage * stage_bytes + b_off, A.vimg + (size_t)s * KPS * bbytes, (uint32_t)nv * bbytes, &full[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 14) {
    if (elect_one()) {                                            // raw x tiles: one tensor box per stage
      tma_prefetch_desc(&A.tmx);
      int rs = 0; uint32_t rph = 0;
      for (int t = 0; t < my_tiles; ++t) {
        const int row0 = (int)((first + (int64_t)t * stride) * 128);
        for (int s = 0; s < spt; ++s) {
          mbar_wait(&rempty[rs], rph ^ 1u);
          mbar_arrive_expect_tx(&rfull[rs], raw_bytes);
          tma_load_2d(raw + (size_t)rs * raw_bytes, &A.tmx, s * KPS * 8, row0, &rfull[rs]);
          if (++rs == RS) { rs = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------- producers: re-lay the raw tile ----------------------------------
    const int pt = tid - 128, w = pt >> 5;
    const int cpr = 2 * KPS;                                      // 16-byte chunks per item row of the raw tile
    uint32_t rd[4], st[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 8 + w) * 32 + lane;                      // a warp reads 512 contiguous bytes per step
      const int row = c / cpr, kc = c % cpr;
      rd[i] = (uint32_t)c * 16u;
      st[i] = (uint32_t)(kc >> 1) * PT_ASTAGE + (uint32_t)(kc & 1) * PT_APLANE + (uint32_t)row * 16u;
    }
    int rs = 0; uint32_t rph = 0;
    int stage = 0; uint32_t phase = 0;
    const int total = my_tiles * spt;
    for (int it = 0; it < total; ++it) {
      mbar_wait(&rfull[rs], rph);
      const unsigned char* src = raw + (size_t)rs * raw_bytes;
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < KPS) v[i] = *(const float4*)(src + rd[i]);
      mbar_wait(&empty[stage], phase ^ 1u);
      unsigned char* dst = ring + (size_t)stage * stage_bytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < KPS) {
          float4 hi, lo;
          split_tf32x4_fast(v[i], hi, lo);
          *(float4*)(dst + st[i]) = hi;
          *(float4*)(dst + st[i] + 2u * PT_APLANE) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[stage]);
      mbar_arrive(&rempty[rs]);
      if (++rs == RS) { rs = 0; rph ^= 1u; }
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  } else {
    project_epilogue(A, warp, lane, my_tiles, first, stride, tmem_base, tfull, tempty, scs, bis, NH, NBUF, acc_cols, nks);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, ncols);
}

static int pu_npad(int N) { return (N + 15) / 16 * 16; }
static int pu_kps(int Npad) { return Npad <= 64 ? 4 : 2; }
static int pu_stages(int Npad) {
  const size_t stage = (size_t)pu_kps(Npad) * (PU_ASTAGE + 64u * (size_t)Npad);
  size_t budget = 200 * 1024;
  int n = (int)(budget / stage);
  if (n > 8) n = 8;
  return n;
}

bool project_fwd_umma_supported(const float* x, int64_t B, int F, int64_t ldx, int N) {
  if (getenv("CFL_FORCE_SIMT")) return false;
  if (B < 1 || N < 1 || N > 256) return false;
  if (F % 8 != 0 || F < 8) return false;
  if ((ldx & 3) != 0 || ((uintptr_t)x & 15u) != 0) return false;
  return pu_stages(pu_npad(N)) >= 3;
}

size_t project_fwd_umma_workspace(int64_t B, int F, int N) {
  (void)B;
  if (N > 256 || F % 8 != 0) return 0;
  return align_up((size_t)F * pu_npad(N) * 8, 1024) + 1024;
}

int project_fwd_umma(const float* x, int64_t B, int F, int64_t ldx, const float* V, int N, int64_t ldV,
                     const float* scaler, const float* bias, float in_scale, int act, float* y, int64_t ldy,
                     float* pre, float* z, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Npad = pu_npad(N);
  CFL_REQUIRE(ws_bytes >= project_fwd_umma_workspace(B, F, N), CFL_ERR_WORKSPACE, "project_umma: workspace too small");
  unsigned char* vimg = (unsigned char*)align_up((size_t)(uintptr_t)ws, 1024);
  const int64_t total = (int64_t)(F / 8) * 2 * Npad;
  int pblocks = (int)((total + 255) / 256);
  if (pblocks > 1184) pblocks = 1184;
  pack_weights_kernel<<<pblocks, 256, 0, st>>>(V, F, N, ldV, Npad, (float*)vimg);
  CFL_LAUNCH_CHECK();
  PuArgs a;
  a.x = x; a.B = B; a.F = F; a.ldx = ldx; a.vimg = vimg; a.N = N; a.Npad = Npad; a.nks = F / 8;
  a.nst = pu_stages(Npad);
  a.kps = pu_kps(Npad);
  a.nh = Npad <= 80 ? 2 : 1;
  a.nbuf = (2 * (a.nh + 1) * Npad <= 512) ? 2 : 1;
  a.scaler = scaler; a.bias = bias; a.in_scale = in_scale; a.act = act;
  a.y = y; a.ldy = ldy; a.pre = pre; a.z = z;
  a.tiles = (B + 127) / 128;
  a.rs = 0;
  // x staged by the TMA engine when its tensor map can be built (aligned base, ldx % 4 == 0): two operand stages
  // (the producers only re-lay shared memory) and as many raw tiles in flight as fit
  // Stage = 4 K-steps (12 MMAs per barrier round trip) with two operand stages while the weight blocks leave room for
  // them, else 2 K-steps with three stages (two were measured slower than the register-prefetch kernel at N_out >= 80:
  // the MMA of a short stage ends before the next one is re-laid).
  const int kps_t = Npad <= 128 ? 4 : 2;
  if (!getenv("CFL_PROJECT_NO_TMA") && B < ((int64_t)1 << 31) - 128 &&
      make_tmap_2d_f32(&a.tmx, x, (uint64_t)F, (uint64_t)B, (uint64_t)ldx, (uint32_t)(8 * kps_t), 128)) {
    const int kps_keep = a.kps;
    a.kps = kps_t;
    const size_t stage = align_up((size_t)a.kps * (PT_ASTAGE + 64u * (size_t)Npad), 128);
    const size_t rawb = (size_t)128 * 32 * a.kps;
    a.nst = kps_t == 4 ? 2 : 3;
    int rs = (int)((200 * 1024 - a.nst * stage) / rawb);
    a.rs = rs > 8 ? 8 : rs;
    if (a.rs >= 2) {
      const size_t smem_t = a.nst * stage + a.rs * rawb + (size_t)Npad * 8 + (2u * a.nst + 2u * a.rs + 4u) * 8 + 64 + 1024;
      CFL_SMEM_LIMIT(project_umma_tma_kernel, smem_t);
      int grid_t = (int)(a.tiles < sm_count() ? a.tiles : sm_count());
      project_umma_tma_kernel<<<grid_t, PT_THREADS, smem_t, st>>>(a);
      CFL_LAUNCH_CHECK();
      return CFL_OK;
    }
    a.nst = pu_stages(Npad);
    a.kps = kps_keep;
  }
  const size_t smem = (size_t)a.nst * a.kps * (PU_ASTAGE + 64u * (size_t)Npad) + (size_t)Npad * 8 + (2u * a.nst + 4u) * 8 + 64 + 1024;
  CFL_SMEM_LIMIT(project_umma_kernel, smem);
  int grid = (int)(a.tiles < sm_count() ? a.tiles : sm_count());
  project_umma_kernel<<<grid, PU_THREADS, smem, st>>>(a);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

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
  CFL_SMEM_LIMIT(umma_selftest_kernel, smem);
  umma_selftest_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(A, Bm, D, N, Kd, variant);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}
