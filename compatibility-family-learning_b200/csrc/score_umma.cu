// Fused all-pairs scoring kernel for sm_100a: Gram-form GEMM on tcgen05 tensor cores (3xTF32,
// fp32 accumulators in TMEM), soft-min epilogue and running per-query top-k behind the MMA; the
// Q x N score matrix never reaches HBM.
//
//   D[c, k*QT + q] = e_c . p_{q,k}        M = 128 catalog rows (TMEM lanes), N = K*QT columns,
//                                         MMA-K = the embedding dimension d (8 per K-step)
//
// Both operands are consumed as pre-split (hi, lo) tf32 images in the canonical K-major
// SWIZZLE_NONE layout (umma.cuh):
//   * the CATALOG image is built once per catalog by cfl_catalog_pack (centred on mu, split,
//     one 8 KB block per (tile, K-step), plus |e|^2 per row) -- the catalog is static across
//     query batches, so no per-query work is spent on it;
//   * the QUERY image of the CTA's query tile (<= 96 KB) is packed per call and stays resident
//     in shared memory for the whole kernel.
// CTA (part, qtile) streams its contiguous range of 128-row catalog tiles.  Warp-specialised:
//   warp  NEPI    one lane issues tcgen05.mma (3 per K-step) and tcgen05.commit
//   warp  NEPI+1  one lane drives the TMA engine: cp.async.bulk of the query image, then one
//                 8 KB catalog block per ring stage (mbarrier expect_tx / complete_tx)
//   warps 0..NEPI-1  epilogue: tcgen05.ld of the K Gram values of 8 (16) queries for the thread's
//                 catalog row, soft-min distance (score.cuh), threshold test, warp-aggregated push
//                 into the per-query buffers (topk.cuh).  Accumulators are double-buffered in
//                 TMEM so the epilogue of tile t overlaps the MMAs of tile t+1.
// Work per score: 2*K*d flops (x3 MMAs); catalog bytes per row per query tile: 8*dpad + 4.
#include <cuda_fp16.h>
#include "score.cuh"
#include "umma.cuh"

namespace cfl {

using namespace umma;

#ifndef CFL_SU_NEPI
#define CFL_SU_NEPI 16
#endif
constexpr int SU_NEPI = CFL_SU_NEPI;                // epilogue warps (a multiple of 4: TMEM lane quarters)
constexpr int SU_THREADS = (SU_NEPI + 2) * 32;
constexpr int SU_EPI_THREADS = SU_NEPI * 32;
constexpr int SU_NSTAGE = 6;                        // ring stages of up to 2 K-steps (16 KB) each
constexpr uint32_t SU_ASTAGE = 4u * 128u * 16u;     // [hl][chunk][128 rows][16 B] = 8 KB
constexpr size_t SU_B_BUDGET = 96 * 1024;

int score_umma_qt(int K, int d) {
  if (K < 1 || K > CFL_MAX_K || d < 1 || d > 128) return 0;
  const int dpad = (d + 7) / 8 * 8;
  const int gq = K <= 2 ? 16 : (K <= 4 ? 8 : 4);
  const int cand[] = {128, 64, 48, 32, 16, 8};
  for (int qt : cand) {
    int nc = K * qt;
    if (nc > 256 || nc % 16 != 0 || qt % gq != 0) continue;
    if ((size_t)dpad * nc * 8 > SU_B_BUDGET) continue;
    return qt;
  }
  return 0;
}

bool score_umma_supported(int K, int d) {
  if (getenv("CFL_FORCE_SIMT")) return false;
  return score_umma_qt(K, d) != 0;
}

size_t score_umma_qimg_bytes(const ScorePlan& p, int K) {      // tf32 hi/lo query image
  return (size_t)p.nqt * p.dpad * 8 * (size_t)(K * p.qt);
}

// ---- catalog image ----------------------------------------------------------------------------
// [tile][kstep][hl][chunk][row 0..127][4 floats]  (8 KB per (tile,kstep)), then e2[tiles*128].
// ... then the fp16 plane of the lower-bound pass: [tile][kstep16][chunk 0|1][row 0..127][8 halfs] (4 KB per
// (tile, 16 columns)), then one flag word (non-zero: a centred value left the fp16 range, plane unusable).
size_t catalog_f16_offset(int64_t N, int d) {
  const int64_t tiles = (N + 127) / 128;
  const int nks = (d + 7) / 8;
  return (size_t)tiles * nks * SU_ASTAGE + (size_t)tiles * 128 * sizeof(float);
}
size_t catalog_f16_bytes(int64_t N, int d) {
  const int64_t tiles = (N + 127) / 128;
  return (size_t)tiles * ((d + 15) / 16) * (SU_ASTAGE / 2);
}
// ... then lbrow[tiles*128] float2 = (|e - mu|^2, |e - mu|) per row, (+inf, 0) for the padding rows of the last tile:
// what the lower-bound pass needs per row (its margin has a term in |e|), one 8-byte load per row and tile.
size_t catalog_lbrow_offset(int64_t N, int d) {
  return catalog_f16_offset(N, d) + catalog_f16_bytes(N, d) + 16;
}
size_t catalog_image_bytes(int64_t N, int d) {
  const int64_t tiles = (N + 127) / 128;
  return catalog_lbrow_offset(N, d) + (size_t)tiles * 128 * sizeof(float2);
}

__global__ void __launch_bounds__(128)
pack_catalog_kernel(const float* __restrict__ E, int64_t N, int d, int64_t lde,
                    const float* __restrict__ mu, unsigned char* __restrict__ img, float* __restrict__ e2g,
                    unsigned char* __restrict__ img16, int* __restrict__ flag16, float2* __restrict__ lbrow) {
  const int64_t tile = blockIdx.x;
  const int r = threadIdx.x;
  const int64_t row = tile * 128 + r;
  const int nks = (d + 7) / 8;
  float e2 = 0.0f;
  float vmax = 0.0f;
  const int nks16 = (d + 15) / 16;
  for (int ks = 0; ks < nks; ++ks) {
    float v[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j = ks * 8 + jj;
      v[jj] = (row < N && j < d) ? E[row * lde + j] - (mu ? mu[j] : 0.0f) : 0.0f;
      e2 = fmaf(v[jj], v[jj], e2);
    }
    float4 h0, l0, h1, l1;
    split_tf32x4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
    split_tf32x4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
    unsigned char* st = img + ((size_t)tile * nks + ks) * SU_ASTAGE;
    *(float4*)(st + ((0 * 2 + 0) * 128 + r) * 16) = h0;
    *(float4*)(st + ((0 * 2 + 1) * 128 + r) * 16) = h1;
    *(float4*)(st + ((1 * 2 + 0) * 128 + r) * 16) = l0;
    *(float4*)(st + ((1 * 2 + 1) * 128 + r) * 16) = l1;
    // the same 8 centred values as halfs: chunk (ks & 1) of fp16 K-step ks >> 1
    __align__(16) __half h[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) { h[jj] = __float2half_rn(v[jj]); vmax = fmaxf(vmax, fabsf(v[jj])); }
    unsigned char* st16 = img16 + ((size_t)tile * nks16 + (ks >> 1)) * (SU_ASTAGE / 2);
    *(uint4*)(st16 + ((ks & 1) * 128 + r) * 16) = *(const uint4*)h;
  }
  if (nks & 1) {                                             // odd number of 8-column steps: zero the last half K-step
    unsigned char* st16 = img16 + ((size_t)tile * nks16 + (nks >> 1)) * (SU_ASTAGE / 2);
    *(uint4*)(st16 + (1 * 128 + r) * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  e2g[row] = e2;
  lbrow[row] = (row < N) ? make_float2(e2, sqrtf(e2)) : make_float2(__int_as_float(0x7f800000), 0.0f);
  if (!(vmax < 60000.0f)) atomicOr(flag16, 1);               // also catches NaN / inf
}

int catalog_pack_launch(const float* E, int64_t N, int d, int64_t lde, const float* mu, void* image,
                        cudaStream_t st) {
  const int64_t tiles = (N + 127) / 128;
  const int nks = (d + 7) / 8;
  unsigned char* img = (unsigned char*)image;
  float* e2g = (float*)(img + (size_t)tiles * nks * SU_ASTAGE);
  unsigned char* img16 = img + catalog_f16_offset(N, d);
  int* flag16 = (int*)(img16 + catalog_f16_bytes(N, d));
  CFL_CUDA(cudaMemsetAsync(flag16, 0, 16, st));
  float2* lbrow = (float2*)(img + catalog_lbrow_offset(N, d));
  pack_catalog_kernel<<<(unsigned)tiles, 128, 0, st>>>(E, N, d, lde, mu, img, e2g, img16, flag16, lbrow);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

struct SuLayout {
  uint32_t b_img, a_ring, scratch, qpar, qpl, thr, cnt, need, bars, tmem_slot, total;
};
__host__ __device__ inline SuLayout su_layout(int K, int qt, int dpad) {
  SuLayout L;
  uint32_t off = 0;
  L.b_img = off;   off += (uint32_t)dpad * 8u * (uint32_t)(K * qt);   off = (off + 1023u) & ~1023u;
  L.a_ring = off;  off += SU_NSTAGE * 2u * SU_ASTAGE;
  L.scratch = off; off += TOPK_STRIDE * 8u;
  L.qpar = off;    off += (uint32_t)qt * (uint32_t)qpar_stride(K) * 4u;  off = (off + 15u) & ~15u;
  L.qpl = off;     off += (uint32_t)qt * (uint32_t)qplane_stride(K) * 4u; off = (off + 15u) & ~15u;
  L.thr = off;     off += (uint32_t)qt * 4u;
  L.cnt = off;     off += (uint32_t)qt * 4u;
  L.need = off;    off += 16u;                                        off = (off + 15u) & ~15u;
  L.bars = off;    off += (2u * SU_NSTAGE + 5u) * 8u;
  L.tmem_slot = off; off += 16u;
  L.total = off;
  return L;
}

// ---- query (B operand) image: [qtile][kstep][hl][chunk][row n = k*QT + ql][4 floats] -----------
__global__ void pack_queries_kernel(const float* __restrict__ Pc, int64_t Q, int K, int d, int qt,
                                    int dpad, float* __restrict__ img) {
  const int nc = K * qt;
  const int nks = dpad / 8;
  const int qtile = blockIdx.x;
  float* base = img + (size_t)qtile * dpad * 2 * nc;        // floats: dpad*8*nc bytes / 4
  for (int e = threadIdx.x; e < nks * 2 * nc; e += blockDim.x) {
    int n = e % nc;
    int c = (e / nc) % 2;
    int ks = e / (2 * nc);
    int k = n / qt, ql = n % qt;
    int64_t q = (int64_t)qtile * qt + ql;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    float* xp = &x.x;
    if (q < Q) {
      const float* src = Pc + (q * K + k) * (int64_t)d;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int j = ks * 8 + c * 4 + i;
        if (j < d) xp[i] = src[j];
      }
    }
    float4 hi, lo;
    split_tf32x4(x, hi, lo);
    size_t step = (size_t)ks * 4 * nc;                       // float4 units per K-step: 4*nc
    ((float4*)base)[step + (size_t)(0 * 2 + c) * nc + n] = hi;
    ((float4*)base)[step + (size_t)(1 * 2 + c) * nc + n] = lo;
  }
}

int score_umma_pack_queries(const ScoreArgs& a, void* qimg, cudaStream_t st) {
  pack_queries_kernel<<<a.plan.nqt, 256, 0, st>>>(a.Pc, a.Q, a.K, a.d, a.plan.qt, a.plan.dpad, (float*)qimg);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(SU_EPI_THREADS) : "memory");
}

// ---- the kernel ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(SU_THREADS, 1)
score_umma_kernel(ScoreArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int GQ = K <= 2 ? 16 : (K <= 4 ? 8 : 4);     // queries per epilogue group (register budget)
  const int QT = A.plan.qt;
  const int NC = K * QT;
  const int dpad = A.plan.dpad;
  const int nks = dpad / 8;
  const SuLayout L = su_layout(K, QT, dpad);
  unsigned char* b_img = smem + L.b_img;
  unsigned char* a_ring = smem + L.a_ring;
  tkey_t* scratch = (tkey_t*)(smem + L.scratch);
  float* qpar = (float*)(smem + L.qpar);
  float* qpl = (float*)(smem + L.qpl);
  float* thr = (float*)(smem + L.thr);
  int* cnt = (int*)(smem + L.cnt);
  int* need = (int*)(smem + L.need);
  uint64_t* full = (uint64_t*)(smem + L.bars);
  uint64_t* empty = full + SU_NSTAGE;
  uint64_t* tfull = empty + SU_NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint64_t* bfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(smem + L.tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int part = blockIdx.x, qtile = blockIdx.y;
  const bool redo = A.redo_tile != nullptr;
  if (redo && A.redo_tile[qtile] == 0) return;               // nothing to redo for this query tile
  const int64_t q0 = (int64_t)qtile * QT;
  const int nq = (int)((A.Q - q0 < QT) ? (A.Q - q0) : QT);
  const int64_t t0 = A.plan.tiles * part / A.plan.parts;
  const int64_t t1 = A.plan.tiles * (part + 1) / A.plan.parts;
  const int ts = A.tile_stride;                              // phase 1 visits every ts-th tile
  const int ntiles = (int)((t1 - t0 + ts - 1) / ts);
  const bool filter = A.phase == 2;
  const int qps = qpar_stride(K);
  const int kss = (nks % 2 == 0) ? 2 : 1;                    // K-steps per ring stage

  // ---- one-time setup ----
  uint32_t ncols = 32;
  while ((int)ncols < 2 * NC) ncols <<= 1;
  if (warp == SU_NEPI) {
    if (lane == 0) {
      for (int s = 0; s < SU_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], SU_EPI_THREADS); mbar_init(&tempty[1], SU_EPI_THREADS);
      mbar_init(bfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ncols);
  }
  // per-query blocks are stored interleaved by query PAIR: qpar[(pair*qps + j)*2 + (q&1)], with the
  // first K entries negated, ready for the packed (FP32x2) epilogue
  for (int i = tid; i < QT * qps; i += SU_THREADS) {
    const int ql = i / qps, j = i % qps;
    float v = (ql < nq) ? A.qpar[(q0 + ql) * qps + j] : 0.0f;
    if (j < K) v = -v;
    qpar[((ql >> 1) * qps + j) * 2 + (ql & 1)] = v;
  }
  {
    const int pbs = qplane_stride(K);                        // same pair interleave for the plane blocks
    for (int i = tid; i < QT * pbs; i += SU_THREADS) {
      const int ql = i / pbs, j = i % pbs;
      qpl[((ql >> 1) * pbs + j) * 2 + (ql & 1)] = (ql < nq) ? A.qplane[(q0 + ql) * pbs + j] : 0.0f;
    }
  }
  // padding queries of the last tile get thr = -inf: they can never be pushed
  // phase 2: fixed thresholds from the sample pass, bumped one ulp so that the strict compare
  // below implements dist <= tau (every candidate tied with the bound must be kept)
  for (int i = tid; i < QT; i += SU_THREADS) {
    float th = __int_as_float(i < nq ? 0x7f800000 : 0xff800000);
    if (filter && i < nq && A.thr_init != nullptr) {
      const float tau = A.thr_init[q0 + i];
      th = (tau < 3.0e38f) ? nextafterf(tau, 3.4e38f) : tau;
    }
    thr[i] = th;
    cnt[i] = 0;
  }
  if (tid < 4) need[tid] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == SU_NEPI) {
    // ================================ MMA issuer (one lane) =================================
    if (elect_one()) {
      mbar_wait(bfull, 0);
      const Step3Desc sd = make_step3((uint32_t)NC, make_idesc_tf32(128, (uint32_t)NC));
      const uint32_t a_base = smem_u32(a_ring), b_base = smem_u32(b_img);
      const uint32_t b_step = 4u * (uint32_t)NC * 16u;
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        mbar_wait(&tempty[buf], ((uint32_t)(t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NC);
        for (int ks = 0; ks < nks; ks += kss) {
          if (!(A.dbg_mode & 2)) mbar_wait(&full[stage], phase);
          tc_fence_after();
          for (int j = 0; j < kss; ++j)
            mma_step3(sd, d_tmem, a_base + (stage * 2 + j) * SU_ASTAGE, b_base + (ks + j) * b_step, ks + j == 0);
          mma_commit(&empty[stage]);
          if (++stage == SU_NSTAGE) { stage = 0; phase ^= 1u; }
        }
        mma_commit(&tfull[buf]);
      }
    }
  } else if (warp == SU_NEPI + 1) {
    // ================================ TMA producer (one lane) ===============================
    if (elect_one()) {
      const uint32_t bbytes = (uint32_t)dpad * 8u * (uint32_t)NC;
      const unsigned char* qsrc = (const unsigned char*)A.qimg + (size_t)qtile * bbytes;
      mbar_arrive_expect_tx(bfull, bbytes);
      for (uint32_t o = 0; o < bbytes; o += 32768u) {
        uint32_t n = bbytes - o < 32768u ? bbytes - o : 32768u;
        bulk_g2s(b_img + o, qsrc + o, n, bfull);
      }
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < ((A.dbg_mode & 2) ? 0 : ntiles); ++t) {
        const unsigned char* src = (const unsigned char*)A.cimg + (size_t)(t0 + (int64_t)t * ts) * nks * SU_ASTAGE;
        for (int ks = 0; ks < nks; ks += kss) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], (uint32_t)kss * SU_ASTAGE);
          bulk_g2s(a_ring + stage * 2 * SU_ASTAGE, src + (size_t)ks * SU_ASTAGE, (uint32_t)kss * SU_ASTAGE, &full[stage]);
          if (++stage == SU_NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ======================================= epilogue ========================================
    constexpr int QPS = qpar_stride(K);
    constexpr int CQ = K + qpar_tri(K);
    constexpr int WPQ = SU_NEPI / 4;                         // warps per TMEM lane quarter
    const int lq = warp & 3, sub = warp >> 2;
    const int lrow = lq * 32 + lane;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const bool dense = A.dist_out != nullptr;
    tkey_t* kbase = A.keys + ((int64_t)part * A.Q + q0) * TOPK_STRIDE;
    // dist >= squared distance from e to the affine hull of the query's prototypes (a quadratic in
    // the Gram values, score.cuh): in the usual regime it is within a few percent of dist, so almost
    // every (row, query) is rejected by ~11 packed FMAs per query pair and the soft-min runs only for
    // the queries of a group where some lane survives.  When the bound is loose for the data at hand
    // (prototypes far apart relative to the spread of the distances) each warp measures its hit rate
    // and switches the test off.
    constexpr int PBS = qplane_stride(K);
    bool bound_on = (K > 1) && !dense;
    int grp_seen = 0, grp_skipped = 0;
    unsigned dbg_seen = 0, dbg_skip = 0, dbg_sel = 0, dbg_full = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const int64_t row = (t0 + (int64_t)t * ts) * 128 + lrow;
      const bool valid = row < A.N;
      const float e2 = __ldg(A.e2 + row);                    // padded to whole tiles
      mbar_wait(&tfull[buf], (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * NC);
      const int nslot = filter ? 2 + ((t >> 2) & 1) : (t & 1);
      const int trig = filter ? TOPK_STRIDE / 2 : TOPK_TRIGGER;
      if (A.dbg_mode & 1) { tc_fence_before(); mbar_arrive(&tempty[buf]); continue; }
      // push the passing lanes of query ql (warp-uniform call) into its key buffer
      auto push = [&](int ql, bool pass, float dv) {
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (m == 0) return;
        const int leader = __ffs(m) - 1;
        int basei = 0;
        if (lane == leader) {
          basei = atomicAdd(&cnt[ql], __popc(m));
          if (basei + __popc(m) > trig) need[nslot] = 1;
        }
        basei = __shfl_sync(0xffffffffu, basei, leader);
        if (pass) kbase[(int64_t)ql * TOPK_STRIDE + basei + __popc(m & lane_lt)] = pack_key(dv, (uint32_t)row);
      };
      for (int g = sub; g * GQ < nq; g += WPQ) {
        float gk[K][GQ];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if constexpr (GQ == 16)     tmem_ld16(taddr + (uint32_t)(k * QT + g * GQ), gk[k]);
          else if constexpr (GQ == 8) tmem_ld8(taddr + (uint32_t)(k * QT + g * GQ), gk[k]);
          else                        tmem_ld4(taddr + (uint32_t)(k * QT + g * GQ), gk[k]);
        }
        tmem_ld_wait();
        const float* qg = qpar + g * GQ * QPS;                 // GQ/2 pair blocks of 2*QPS floats
        float tg[GQ];
#pragma unroll
        for (int i = 0; i < GQ; i += 4) {
          const float4 t4 = *(const float4*)(thr + g * GQ + i);
          tg[i] = t4.x; tg[i + 1] = t4.y; tg[i + 2] = t4.z; tg[i + 3] = t4.w;
        }
        if constexpr (K > 1) {
          if (bound_on) {
            const float e2s = e2 * (1.0f - CFL_PLANE_REL);
            const f2_t e2sp = pk2(e2s, e2s);
            float lb[GQ];
            bool nd = false;
#pragma unroll
            for (int pi = 0; pi < GQ / 2; ++pi) {
              f2_t pv[PBS];
              const ulonglong2* src = (const ulonglong2*)(qpl + (g * (GQ / 2) + pi) * PBS * 2);
#pragma unroll
              for (int j = 0; j < PBS; j += 2) { const ulonglong2 u = src[j >> 1]; pv[j] = u.x; pv[j + 1] = u.y; }
              float gA[K], gB[K];
#pragma unroll
              for (int k = 0; k < K; ++k) { gA[k] = gk[k][2 * pi]; gB[k] = gk[k][2 * pi + 1]; }
              plane_bound_pair<K>(gA, gB, e2sp, pv, lb[2 * pi], lb[2 * pi + 1]);
              nd |= (lb[2 * pi] < tg[2 * pi]) | (lb[2 * pi + 1] < tg[2 * pi + 1]);
            }
            ++grp_seen; ++dbg_seen;
            if (!__any_sync(0xffffffffu, nd && valid)) { ++grp_skipped; ++dbg_skip; continue; }
            uint32_t nb = 0;
#pragma unroll
            for (int i = 0; i < GQ; ++i) nb |= (lb[i] < tg[i]) ? (1u << i) : 0u;
            if (!valid) nb = 0;
            const uint32_t anyn = __reduce_or_sync(0xffffffffu, nb);
            if (__popc(anyn) <= GQ / 4) {                      // a few queries: soft-min only for those
              ++dbg_sel;
#pragma unroll
              for (int i = 0; i < GQ; ++i) {
                if (anyn & (1u << i)) {                        // warp-uniform
                  float gi[K];
#pragma unroll
                  for (int k = 0; k < K; ++k) gi[k] = gk[k][i];
                  const float dv = softmin_from_gram_il<K>(gi, e2, qg + (i >> 1) * QPS * 2 + (i & 1));
                  push(g * GQ + i, valid && dv < tg[i], dv);
                }
              }
              continue;
            }
          }
        }
        ++dbg_full;
        float dist[GQ];
#pragma unroll
        for (int pi = 0; pi < GQ / 2; ++pi) {
          f2_t qv[QPS];
          const ulonglong2* src = (const ulonglong2*)(qg + pi * QPS * 2);
#pragma unroll
          for (int j = 0; j < QPS; j += 2) { const ulonglong2 u = src[j >> 1]; qv[j] = u.x; qv[j + 1] = u.y; }
          float gA[K], gB[K];
#pragma unroll
          for (int k = 0; k < K; ++k) { gA[k] = gk[k][2 * pi]; gB[k] = gk[k][2 * pi + 1]; }
          softmin_pair<K>(gA, gB, e2, qv, dist[2 * pi], dist[2 * pi + 1]);
        }
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < GQ; ++i) bits |= (dist[i] < tg[i]) ? (1u << i) : 0u;
        if (!valid) bits = 0;
        if (dense && valid && A.phase != 1) {
#pragma unroll
          for (int i = 0; i < GQ; ++i)
            if (g * GQ + i < nq) A.dist_out[(q0 + g * GQ + i) * A.N + row] = dist[i];
        }
        // which queries of the group have at least one passing lane (usually 0-2 of them)
        const uint32_t anyq = __reduce_or_sync(0xffffffffu, bits);
        if (anyq) {
#pragma unroll
          for (int i = 0; i < GQ; ++i)
            if (anyq & (1u << i)) push(g * GQ + i, (bits >> i) & 1u, dist[i]);     // warp-uniform
        }
      }
      // hit-rate check of the plane bound: in the filter pass over tiles [0,16); in the adaptive
      // passes over tiles [16,32), once the running thresholds have tightened
      if (bound_on) {
        if (!filter && t == 15) { grp_seen = 0; grp_skipped = 0; }
        if (t == (filter ? 15 : 31) && grp_skipped * 2 < grp_seen) bound_on = false;
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
      // ---- compaction point: all epilogue warps sort one over-full buffer together ----
      // adaptive phases: checked every tile (a tile adds <= 128 keys per query, buffers hold 512);
      // filter phase: every 4th tile (<= 512 new keys, buffers hold 1024) and it almost never fires.
      // The decision must be identical in every warp although faster warps may already be pushing
      // for the next tile: pushes raise need[slot] (slot = parity of the tile / 4-tile epoch), which
      // is stable once every warp has passed the barrier below.
      if (!filter || (t & 3) == 3) {
        epi_bar_sync();
        if (need[nslot]) {
          for (int qb = 0; qb < nq; qb += 32) {
            const int ql0 = qb + lane;
            const uint32_t over = __ballot_sync(0xffffffffu, ql0 < nq && cnt[ql0] > trig);
            for (uint32_t m = over; m; m &= m - 1) {          // identical in every warp
              const int ql = qb + __ffs(m) - 1;
              if (filter) {
                coop_compact_big<SU_EPI_THREADS>(kbase + (int64_t)ql * TOPK_STRIDE, cnt[ql], A.plan.kk, scratch, tid, &cnt[ql], &thr[ql]);
              } else {
                coop_compact<SU_EPI_THREADS>(kbase + (int64_t)ql * TOPK_STRIDE, cnt[ql], A.plan.kk, scratch, tid, &cnt[ql], &thr[ql]);
              }
            }
          }
          if (tid == 0) need[nslot] = 0;
          epi_bar_sync();
        }
      }
    }
    if (A.dbg != nullptr && lane == 0) {
      atomicAdd(&A.dbg[0], (unsigned long long)dbg_seen); atomicAdd(&A.dbg[1], (unsigned long long)dbg_skip);
      atomicAdd(&A.dbg[2], (unsigned long long)dbg_sel);  atomicAdd(&A.dbg[3], (unsigned long long)dbg_full);
    }
    if (!filter) {
      for (int ql = 0; ql < nq; ++ql) {
        coop_compact<SU_EPI_THREADS>(kbase + (int64_t)ql * TOPK_STRIDE, cnt[ql], A.plan.kk, scratch, tid, &cnt[ql], nullptr);
        if (tid == 0) A.counts[(int64_t)part * A.Q + q0 + ql] = cnt[ql];
      }
    } else {
      epi_bar_sync();                                         // all pushes done; lists stay unsorted
      for (int ql = tid; ql < nq; ql += SU_EPI_THREADS) {
        // a redo launch restarts only the queries it was asked to redo (thr_init > -inf)
        if (redo && !(A.thr_init[q0 + ql] > __int_as_float(0xff800000))) continue;
        A.counts[(int64_t)part * A.Q + q0 + ql] = cnt[ql];
      }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == SU_NEPI) tmem_dealloc(tmem_base, ncols);
}

template <int K>
static int launch_umma(const ScoreArgs& a, cudaStream_t st) {
  SuLayout L = su_layout(K, a.plan.qt, a.plan.dpad);
  size_t smem = L.total + 1024;      // slack for the alignment of the dynamic segment
  CFL_SMEM_LIMIT(score_umma_kernel<K>, smem);
  dim3 grid(a.plan.parts, a.plan.nqt);
  score_umma_kernel<K><<<grid, SU_THREADS, smem, st>>>(a);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

int score_umma_launch(const ScoreArgs& a, cudaStream_t st) {
  switch (a.K) {
    case 1: return launch_umma<1>(a, st);
    case 2: return launch_umma<2>(a, st);
    case 3: return launch_umma<3>(a, st);
    case 4: return launch_umma<4>(a, st);
    case 5: return launch_umma<5>(a, st);
    case 6: return launch_umma<6>(a, st);
    case 7: return launch_umma<7>(a, st);
    case 8: return launch_umma<8>(a, st);
  }
  set_error("score_umma: K=%d unsupported", a.K);
  return CFL_ERR_UNSUPPORTED;
}

}  // namespace cfl
