// Weight gradient of the projection on the tensor cores: dV partials  C[f, n] = sum_b x[b, f] * dpre[b, n]
// (the tf.gradients of tf.matmul in fully_connected_weight_norm, cfl/layers.py:80, reached through
// AdamOptimizer.minimize -- cfl/models/cfl.py:1083-1085, cfl/models/dist.py:291-293).
//
// A contraction over the BATCH: M = 128 features per tile (TMEM lanes), N = N_out padded to 16, MMA-K = batch rows
// (8 per K-step), 3xTF32 (lo*hi + hi*lo + hi*hi, fp32 accumulators in TMEM).  Both operands are batch-major in memory
// (x[b, f], dpre[b, n]) while tcgen05 wants the contraction index innermost (K-major, 16-byte chunks of 4 batch rows):
// the producers transpose on the fly -- a thread owns one feature (or one output column) and gathers 4 consecutive batch
// rows with 4 scalar loads that are coalesced ACROSS the warp (32 consecutive features of one row), then writes the
// chunk with one conflict-free 128-bit shared store per plane (hi, lo).  dpre = dy * act'(y) is formed while loading.
//
// Persistent CTAs over jobs (feature tile, batch slab): every job's [128 x N] partial goes to Cpart[slab] and
// project_bwd_final (project.cu) adds the slabs in fixed order -- deterministic, no atomics.  One output is not
// accumulated in a single TMEM column (the tensor core's fp32 accumulate truncates: project_umma.cu): the hi*hi products
// rotate over NH accumulators, the cross terms have their own, and a slab is short (2048 rows by default).
//   warps 0-3    epilogue: drain the accumulators of a finished job (sum of the NH + 1 partials) to Cpart
//   warps 4-11   producers (256 threads): transpose + split + store one ring stage (2 K-steps = 16 batch rows) at a time,
//                PF stages of register prefetch
//   warp 12      one lane issues tcgen05.mma (3 per K-step) / tcgen05.commit
// HBM bytes: x once (4 B F), dpre once per feature tile (L2); flops 2 B F N (x3 MMAs issued).
#include <stdlib.h>
#include "common.cuh"
#include "umma.cuh"
#include "tmap.cuh"

namespace cfl {

using namespace umma;

constexpr int PB_THREADS = 13 * 32;
constexpr int PB_PROD = 256;
constexpr int PB_KPS = 2;                                  // K-steps per ring stage
constexpr int PB_ROWS = 8 * PB_KPS;                        // batch rows per stage
constexpr uint32_t PB_ASTEP = 4u * 128u * 16u;             // [hl][chunk][128 features][16 B]
constexpr int PB_PF = 3;                                   // stages of register prefetch

struct PbArgs {
  CUtensorMap tmx, tmd, tmy;                               // TMA-staged kernel: x [B, F], dy [B, N], y [B, N]
  const float* x; int64_t B; int F; int64_t ldx;
  const float* dy; int64_t lddy; const float* y; int64_t ldy; int act;
  int N, Npad, nst, nh, rs;
  int dbg;                                                 // experiments (CFL_EXPERIMENTS=1, CFL_PB_DBG): 1 = loader copies nothing, 2 = producers store nothing
  int64_t slab_rows; int slabs, ftiles;
  float* Cpart;                                            // [slabs][F][N]
};

template <int NBI, bool NEEDY>
__global__ void __launch_bounds__(PB_THREADS, 1)
project_bwd_umma_kernel(PbArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Npad = A.Npad, NST = A.nst, NH = A.nh;
  const uint32_t bbytes = 4u * (uint32_t)Npad * 16u;                // one K-step of the B operand
  const uint32_t stage_bytes = (uint32_t)PB_KPS * (PB_ASTEP + bbytes);
  const uint32_t b_off = (uint32_t)PB_KPS * PB_ASTEP;
  unsigned char* ring = smem;
  uint64_t* full = (uint64_t*)(smem + (size_t)NST * stage_bytes);
  uint64_t* empty = full + NST;
  uint64_t* tfull = empty + NST;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 1);

  uint32_t ncols = 32;
  while ((int)ncols < (NH + 1) * Npad) ncols <<= 1;
  if (warp == 12) {
    if (lane == 0) {
      for (int s = 0; s < NST; ++s) { mbar_init(&full[s], PB_PROD); mbar_init(&empty[s], 1); }
      mbar_init(tfull, 1);
      mbar_init(tempty, 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ncols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t njobs = (int64_t)A.ftiles * A.slabs;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int my_jobs = first < njobs ? (int)((njobs - first + stride - 1) / stride) : 0;
  // job -> (feature tile, slab): neighbouring CTAs share a slab (its dpre rows are served by L2)
  auto job_rows = [&](int64_t job, int64_t& b0) {
    const int64_t slab = job / A.ftiles;
    b0 = slab * A.slab_rows;
    const int64_t b1 = b0 + A.slab_rows < A.B ? b0 + A.slab_rows : A.B;
    return (int)(b1 - b0);
  };

  if (warp == 12) {
    // ================================ MMA issuer (one lane) =================================
    if (elect_one()) {
      const Step3Desc sd = make_step3((uint32_t)Npad, make_idesc_tf32(128, (uint32_t)Npad));
      const uint32_t ring_u = smem_u32(ring);
      const uint32_t d_lo = tmem_base + (uint32_t)(NH * Npad);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < my_jobs; ++t) {
        int64_t b0;
        const int rows = job_rows(first + (int64_t)t * stride, b0);
        const int spj = (rows + PB_ROWS - 1) / PB_ROWS;
        mbar_wait(tempty, ((uint32_t)t & 1u) ^ 1u);
        tc_fence_after();
        int kidx = 0;
        for (int s = 0; s < spj; ++s) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = ring_u + stage * stage_bytes;
#pragma unroll
          for (int j = 0; j < PB_KPS; ++j, ++kidx) {
            const uint64_t ao = (uint64_t)((sa + j * PB_ASTEP) >> 4), bo = (uint64_t)((sa + b_off + j * bbytes) >> 4);
            mma_tf32(d_lo, sd.a_lo + ao, sd.b_hi + bo, sd.idesc, kidx == 0 ? 0u : 1u);
            mma_tf32(d_lo, sd.a_hi + ao, sd.b_lo + bo, sd.idesc, 1u);
            mma_tf32(tmem_base + (uint32_t)((kidx % NH) * Npad), sd.a_hi + ao, sd.b_hi + bo, sd.idesc, kidx < NH ? 0u : 1u);
          }
          mma_commit(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        mma_commit(tfull);
      }
    }
  } else if (warp >= 4) {
    // ================================ producers =============================================
    const int p = tid - 128;
    const int fa = p & 127, ca = p >> 7;                       // A: feature within the tile, chunk (0/1) of both K-steps
    float qa[PB_PF][PB_KPS][4];
    float qd[PB_PF][NBI][4];
    float qy[PB_PF][NEEDY ? NBI : 1][4];
    // running position of the loader: (job index lt, stage ls within the job); the ring itself is sequential
    int lt = 0, ls = 0;
    int l_rows = 0, l_spj = 0; int64_t l_b0 = 0, l_f0 = 0;
    auto enter_job = [&](int t) {
      if (t < my_jobs) {
        const int64_t job = first + (int64_t)t * stride;
        l_rows = job_rows(job, l_b0);
        l_spj = (l_rows + PB_ROWS - 1) / PB_ROWS;
        l_f0 = (job % A.ftiles) * 128;
      } else { l_rows = 0; l_spj = 0; }
    };
    enter_job(0);
    auto load_stage = [&](float (&a)[PB_KPS][4], float (&dd)[NBI][4], float (&yy)[NEEDY ? NBI : 1][4]) {
      const int64_t sb = (int64_t)ls * PB_ROWS;                  // first row of the stage, relative to the slab
      const int64_t f = l_f0 + fa;
      const bool fok = f < A.F;
#pragma unroll
      for (int j = 0; j < PB_KPS; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t r = sb + j * 8 + ca * 4 + i;
          a[j][i] = (fok && r < l_rows) ? __ldg(A.x + (l_b0 + r) * A.ldx + f) : 0.0f;
        }
#pragma unroll
      for (int it = 0; it < NBI; ++it) {
        const int idx = p + PB_PROD * it;
        const int n = idx % Npad, ch = idx / Npad;               // ch = 2 * K-step + chunk
        const bool nok = ch < 2 * PB_KPS && n < A.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t r = sb + ch * 4 + i;
          const bool ok = nok && r < l_rows;
          dd[it][i] = ok ? __ldg(A.dy + (l_b0 + r) * A.lddy + n) : 0.0f;
          if (NEEDY) yy[it][i] = ok ? __ldg(A.y + (l_b0 + r) * A.ldy + n) : 0.0f;
        }
      }
      if (++ls == l_spj) { ls = 0; ++lt; enter_job(lt); }
    };
    int total = 0;                                               // stages of all my jobs
    for (int t = 0; t < my_jobs; ++t) { int64_t b0; total += (job_rows(first + (int64_t)t * stride, b0) + PB_ROWS - 1) / PB_ROWS; }
#pragma unroll
    for (int u = 0; u < PB_PF; ++u) { if (u < total) load_stage(qa[u], qd[u], qy[u]); }
    int sidx = 0;
    for (int base = 0; base < total; base += PB_PF) {
#pragma unroll
      for (int u = 0; u < PB_PF; ++u) {
        if (base + u >= total) break;
        float a[PB_KPS][4], dd[NBI][4];
#pragma unroll
        for (int j = 0; j < PB_KPS; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) a[j][i] = qa[u][j][i];
#pragma unroll
        for (int it = 0; it < NBI; ++it)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float gq = qd[u][it][i];
            if (NEEDY) gq *= act_grad_from_y(qy[u][it][i], A.act);
            dd[it][i] = gq;
          }
        if (base + u + PB_PF < total) load_stage(qa[u], qd[u], qy[u]);
        const int stage = sidx % NST;
        const uint32_t phase = (uint32_t)((sidx / NST) & 1);
        mbar_wait(&empty[stage], phase ^ 1u);
        unsigned char* sbase = ring + (size_t)stage * stage_bytes;
#pragma unroll
        for (int j = 0; j < PB_KPS; ++j) {
          float4 hi, lo;
          split_tf32x4(make_float4(a[j][0], a[j][1], a[j][2], a[j][3]), hi, lo);
          unsigned char* dst = sbase + j * PB_ASTEP + (ca * 128 + fa) * 16;
          *(float4*)dst = hi;
          *(float4*)(dst + 2 * 128 * 16) = lo;
        }
#pragma unroll
        for (int it = 0; it < NBI; ++it) {
          const int idx = p + PB_PROD * it;
          const int n = idx % Npad, ch = idx / Npad;
          if (ch < 2 * PB_KPS) {
            float4 hi, lo;
            split_tf32x4(make_float4(dd[it][0], dd[it][1], dd[it][2], dd[it][3]), hi, lo);
            unsigned char* dst = sbase + b_off + (ch >> 1) * bbytes + ((size_t)(ch & 1) * Npad + n) * 16;
            *(float4*)dst = hi;
            *(float4*)(dst + 2 * (size_t)Npad * 16) = lo;
          }
        }
        fence_proxy_async();
        mbar_arrive(&full[stage]);
        ++sidx;
      }
    }
  } else {
    // ================================ epilogue ==============================================
    const int lrow = warp * 32 + lane;
    for (int t = 0; t < my_jobs; ++t) {
      const int64_t job = first + (int64_t)t * stride;
      int64_t b0;
      const int rows = job_rows(job, b0);
      const int nksteps = (rows + PB_ROWS - 1) / PB_ROWS * PB_KPS;
      const int nw = nksteps < NH ? nksteps : NH;                // hi accumulators that were written
      const int64_t slab = job / A.ftiles;
      const int64_t f = (job % A.ftiles) * 128 + lrow;
      mbar_wait(tfull, (uint32_t)t & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      float* crow = A.Cpart + ((int64_t)slab * A.F + f) * A.N;
      for (int c0 = 0; c0 < Npad; c0 += 16) {
        float acc[16], part[16];
        tmem_ld16(taddr + (uint32_t)(NH * Npad + c0), acc);
        tmem_ld_wait();
        for (int h = 0; h < nw; ++h) {
          tmem_ld16(taddr + (uint32_t)(h * Npad + c0), part);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += part[j];
        }
        if (f < A.F) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < A.N) crow[c0 + j] = acc[j];
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, ncols);
}

// ---- variant with the raw tiles staged by the TMA engine ------------------------------------------------------------
// The register-prefetch kernel above keeps only PF stages of loads in flight per thread (24 KB per SM: latency-bound at
// ~1/4 of HBM).  Here warp 13 streams the RAW batch-major tiles -- 16 batch rows x 128 features of x, the same rows of
// dy (and y) -- into a deep shared-memory ring (mbarrier expect_tx); the producers then
// transpose out of shared memory: thread (feature f) reads 4 consecutive batch rows at stride 512 B (conflict-free),
// splits hi/lo and writes the K-major operand chunk as before.  The tiles are TMA tensor boxes (tmap.cuh: one instruction
// per tile, ragged edges zero-filled); needs 16-byte aligned bases and leading dimensions % 4 == 0 (x, dy, y), other
// shapes use the kernel above.
constexpr int PT_THREADS = 14 * 32;

template <int NBI, bool NEEDY>
__global__ void __launch_bounds__(PT_THREADS, 1)
project_bwd_umma_tma_kernel(const __grid_constant__ PbArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Npad = A.Npad, NST = A.nst, NH = A.nh, RS = A.rs;
  const uint32_t bbytes = 4u * (uint32_t)Npad * 16u;                // one K-step of the B operand
  const uint32_t stage_bytes = (uint32_t)PB_KPS * (PB_ASTEP + bbytes);
  const uint32_t b_off = (uint32_t)PB_KPS * PB_ASTEP;
  const uint32_t araw = (uint32_t)PB_ROWS * 128u * 4u;              // raw x tile: [16 rows][128 features]
  const uint32_t braw = (uint32_t)PB_ROWS * (uint32_t)Npad * 4u;    // raw dy (y) tile: [16 rows][Npad]
  const uint32_t raw_bytes = araw + (NEEDY ? 2u : 1u) * braw;
  unsigned char* ring = smem;
  unsigned char* raw = smem + (size_t)NST * stage_bytes;
  uint64_t* full = (uint64_t*)(raw + (size_t)RS * raw_bytes);
  uint64_t* empty = full + NST;
  uint64_t* rfull = empty + NST;
  uint64_t* rempty = rfull + RS;
  uint64_t* tfull = rempty + RS;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = (uint32_t*)(tempty + 1);

  uint32_t ncols = 32;
  while ((int)ncols < (NH + 1) * Npad) ncols <<= 1;
  if (warp == 12) {
    if (lane == 0) {
      for (int s = 0; s < NST; ++s) { mbar_init(&full[s], PB_PROD); mbar_init(&empty[s], 1); }
      for (int s = 0; s < RS; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rempty[s], PB_PROD); }
      mbar_init(tfull, 1);
      mbar_init(tempty, 128);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ncols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t njobs = (int64_t)A.ftiles * A.slabs;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int my_jobs = first < njobs ? (int)((njobs - first + stride - 1) / stride) : 0;
  auto job_rows = [&](int64_t job, int64_t& b0) {
    const int64_t slab = job / A.ftiles;
    b0 = slab * A.slab_rows;
    const int64_t b1 = b0 + A.slab_rows < A.B ? b0 + A.slab_rows : A.B;
    return (int)(b1 - b0);
  };

  if (warp == 12) {
    // ================================ MMA issuer (one lane) =================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, (uint32_t)Npad);
      const uint32_t hiw = desc_hi_word(128u);
      const uint32_t ring_u = smem_u32(ring);
      const uint32_t a_w0 = desc_lo_word(ring_u, 128u * 16u);
      const uint32_t b_w0 = desc_lo_word(ring_u + b_off, (uint32_t)Npad * 16u);
      const uint32_t a_lo_d = (2u * 128u * 16u) >> 4, b_lo_d = 2u * (uint32_t)Npad;
      const uint32_t b_step4 = bbytes >> 4, st4 = stage_bytes >> 4;
      const uint32_t d_lo = tmem_base + (uint32_t)(NH * Npad);
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < my_jobs; ++t) {
        int64_t b0;
        const int rows = job_rows(first + (int64_t)t * stride, b0);
        const int spj = (rows + PB_ROWS - 1) / PB_ROWS;
        mbar_wait(tempty, ((uint32_t)t & 1u) ^ 1u);
        tc_fence_after();
        uint32_t d_hi = tmem_base; int h = 0;
        uint32_t seen = 0;
        for (int s = 0; s < spj; ++s) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          uint32_t aw = a_w0 + (uint32_t)stage * st4, bw = b_w0 + (uint32_t)stage * st4;
#pragma unroll
          for (int j = 0; j < PB_KPS; ++j) {
            mma_tf32_w(d_lo, aw + a_lo_d, hiw, bw, hiw, idesc, seen ? 1u : 0u);
            mma_tf32_w(d_lo, aw, hiw, bw + b_lo_d, hiw, idesc, 1u);
            mma_tf32_w(d_hi, aw, hiw, bw, hiw, idesc, seen >= (uint32_t)NH ? 1u : 0u);
            aw += PB_ASTEP >> 4; bw += b_step4;
            if (seen < (uint32_t)NH) ++seen;
            d_hi += (uint32_t)Npad;
            if (++h == NH) { h = 0; d_hi = tmem_base; }
          }
          mma_commit(&empty[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        mma_commit(tfull);
      }
    }
  } else if (warp == 13) {
    // ================================ raw-tile loader (one lane) =============================
    if (elect_one()) {
      tma_prefetch_desc(&A.tmx); tma_prefetch_desc(&A.tmd);
      if (NEEDY) tma_prefetch_desc(&A.tmy);
      int rs = 0; uint32_t rph = 0;
      const uint32_t tx = araw + (NEEDY ? 2u : 1u) * braw;       // whole boxes arrive (out-of-range parts as zeros)
      for (int t = 0; t < my_jobs; ++t) {
        const int64_t job = first + (int64_t)t * stride;
        int64_t b0;
        const int rows = job_rows(job, b0);
        const int spj = (rows + PB_ROWS - 1) / PB_ROWS;
        const int f0 = (int)(job % A.ftiles) * 128;
        int r0 = (int)b0;
        for (int s = 0; s < spj; ++s, r0 += PB_ROWS) {
          mbar_wait(&rempty[rs], rph ^ 1u);
          unsigned char* dst = raw + (size_t)rs * raw_bytes;
          mbar_arrive_expect_tx(&rfull[rs], (A.dbg & 1) ? 0u : tx);
          if (!(A.dbg & 1)) {
            tma_load_2d(dst, &A.tmx, f0, r0, &rfull[rs]);
            tma_load_2d(dst + araw, &A.tmd, 0, r0, &rfull[rs]);
            if (NEEDY) tma_load_2d(dst + araw + braw, &A.tmy, 0, r0, &rfull[rs]);
          }
          if (++rs == RS) { rs = 0; rph ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ producers: transpose out of the raw ring ================
    // Everything that does not change from stage to stage is computed once: the thread's read positions in the raw
    // tiles and its write positions in the operand stage.
    const int p = tid - 128;
    const int fa = p & 127, ca = p >> 7;
    const int a_rd = ca * 4 * 128 + fa;                          // + (j * 8 + i) * 128
    const uint32_t a_st = (uint32_t)(ca * 128 + fa) * 16u;       // + j * PB_ASTEP (+ 4096 for lo)
    int b_rd[NBI], b_r0[NBI];
    uint32_t b_st[NBI];
#pragma unroll
    for (int it = 0; it < NBI; ++it) {
      const int idx = p + PB_PROD * it;
      const int n = idx % Npad, ch = idx / Npad;                 // ch = 2 * K-step + chunk
      b_r0[it] = ch < 2 * PB_KPS ? ch * 4 : -1;
      b_rd[it] = ch * 4 * Npad + n;                              // + i * Npad
      b_st[it] = b_off + (uint32_t)(ch >> 1) * bbytes + (uint32_t)((ch & 1) * Npad + n) * 16u;
    }
    const uint32_t lo_b = 2u * (uint32_t)Npad * 16u;
    int rs = 0; uint32_t rph = 0;
    int stage = 0; uint32_t phase = 0;
    for (int t = 0; t < my_jobs; ++t) {
      int64_t b0;
      const int rows = job_rows(first + (int64_t)t * stride, b0);
      const int spj = (rows + PB_ROWS - 1) / PB_ROWS;
      for (int s = 0; s < spj; ++s) {
        const int nr = rows - s * PB_ROWS;                       // rows of this stage that belong to the slab
        mbar_wait(&rfull[rs], rph);
        if (A.dbg & 2) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive(&full[stage]);
          mbar_arrive(&rempty[rs]);
          if (++rs == RS) { rs = 0; rph ^= 1u; }
          if (++stage == NST) { stage = 0; phase ^= 1u; }
          continue;
        }
        const float* ar = (const float*)(raw + (size_t)rs * raw_bytes);
        const float* dr = ar + PB_ROWS * 128;
        const float* yr = dr + PB_ROWS * Npad;
        float a[PB_KPS][4], dd[NBI][4];
#pragma unroll
        for (int j = 0; j < PB_KPS; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) a[j][i] = ar[a_rd + (j * 8 + i) * 128];
#pragma unroll
        for (int it = 0; it < NBI; ++it)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float gq = 0.0f;
            if (b_r0[it] >= 0) {
              gq = dr[b_rd[it] + i * Npad];
              if (NEEDY) gq *= act_grad_from_y(yr[b_rd[it] + i * Npad], A.act);
            }
            dd[it][i] = gq;
          }
        if (nr < PB_ROWS) {                                      // last stage of a slab: its tail rows are another slab's
#pragma unroll
          for (int j = 0; j < PB_KPS; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (j * 8 + ca * 4 + i >= nr) a[j][i] = 0.0f;
#pragma unroll
          for (int it = 0; it < NBI; ++it)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (b_r0[it] + i >= nr) dd[it][i] = 0.0f;
        }
        mbar_wait(&empty[stage], phase ^ 1u);
        unsigned char* sbase = ring + (size_t)stage * stage_bytes;
#pragma unroll
        for (int j = 0; j < PB_KPS; ++j) {
          float4 hi, lo;
          split_tf32x4_fast(make_float4(a[j][0], a[j][1], a[j][2], a[j][3]), hi, lo);
          unsigned char* dst = sbase + j * PB_ASTEP + a_st;
          *(float4*)dst = hi;
          *(float4*)(dst + 2 * 128 * 16) = lo;
        }
#pragma unroll
        for (int it = 0; it < NBI; ++it) {
          if (b_r0[it] >= 0) {
            float4 hi, lo;
            split_tf32x4_fast(make_float4(dd[it][0], dd[it][1], dd[it][2], dd[it][3]), hi, lo);
            unsigned char* dst = sbase + b_st[it];
            *(float4*)dst = hi;
            *(float4*)(dst + lo_b) = lo;
          }
        }
        fence_proxy_async();
        mbar_arrive(&full[stage]);
        mbar_arrive(&rempty[rs]);
        if (++rs == RS) { rs = 0; rph ^= 1u; }
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue ==============================================
    const int lrow = warp * 32 + lane;
    for (int t = 0; t < my_jobs; ++t) {
      const int64_t job = first + (int64_t)t * stride;
      int64_t b0;
      const int rows = job_rows(job, b0);
      const int nksteps = (rows + PB_ROWS - 1) / PB_ROWS * PB_KPS;
      const int nw = nksteps < NH ? nksteps : NH;
      const int64_t slab = job / A.ftiles;
      const int64_t f = (job % A.ftiles) * 128 + lrow;
      mbar_wait(tfull, (uint32_t)t & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
      float* crow = A.Cpart + ((int64_t)slab * A.F + f) * A.N;
      for (int c0 = 0; c0 < Npad; c0 += 16) {
        float acc[16], part[16];
        tmem_ld16(taddr + (uint32_t)(NH * Npad + c0), acc);
        tmem_ld_wait();
        for (int h = 0; h < nw; ++h) {
          tmem_ld16(taddr + (uint32_t)(h * Npad + c0), part);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += part[j];
        }
        if (f < A.F) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c0 + j < A.N) crow[c0 + j] = acc[j];
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, ncols);
}

static int pb_npad(int N) { return (N + 15) / 16 * 16; }
static int pb_stages(int Npad) {
  const size_t stage = (size_t)PB_KPS * (PB_ASTEP + 64u * (size_t)Npad);
  int n = (int)((200 * 1024) / stage);
  return n > 8 ? 8 : n;
}

bool project_bwd_umma_supported(int64_t B, int F, int N) {
  if (getenv("CFL_FORCE_SIMT")) return false;
  return B >= 256 && F >= 1 && N >= 1 && N <= 256 && pb_stages(pb_npad(N)) >= 3;
}

// rows per batch slab: 2048 (256 K-steps per accumulator set) unless that gives fewer jobs than SMs or more than 64 slabs
int64_t project_bwd_umma_slab_rows(int64_t B, int F) {
  const int64_t ftiles = (F + 127) / 128;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int64_t rows = 2048;
  while (rows > 256 && ftiles * ((B + rows - 1) / rows) < sms) rows /= 2;
  while ((B + rows - 1) / rows > 64) rows *= 2;
  return rows;
}
int project_bwd_umma_slabs(int64_t B, int F) {
  const int64_t rows = project_bwd_umma_slab_rows(B, F);
  return (int)((B + rows - 1) / rows);
}

// the TMA-staged variant: 3 operand stages + as many raw stages as fit
template <int NBI>
static int pt_launch(PbArgs a, int grid, cudaStream_t st) {
  const bool needy = a.act != CFL_ACT_LINEAR;
  const size_t stage = (size_t)PB_KPS * (PB_ASTEP + 64u * (size_t)a.Npad);
  const size_t rawb = (size_t)PB_ROWS * 128 * 4 + (needy ? 2 : 1) * (size_t)PB_ROWS * a.Npad * 4;
  a.nst = 3;
  int rs = (int)((200 * 1024 - a.nst * stage) / rawb);
  a.rs = rs > 12 ? 12 : rs;
  const size_t smem = a.nst * stage + a.rs * rawb + (2u * a.nst + 2u * a.rs + 2u) * 8 + 64 + 1024;
  if (needy) {
    CFL_SMEM_LIMIT((project_bwd_umma_tma_kernel<NBI, true>), smem);
    project_bwd_umma_tma_kernel<NBI, true><<<grid, PT_THREADS, smem, st>>>(a);
  } else {
    CFL_SMEM_LIMIT((project_bwd_umma_tma_kernel<NBI, false>), smem);
    project_bwd_umma_tma_kernel<NBI, false><<<grid, PT_THREADS, smem, st>>>(a);
  }
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

template <int NBI>
static int pb_launch(const PbArgs& a, size_t smem, int grid, cudaStream_t st) {
  if (a.act != CFL_ACT_LINEAR) {
    CFL_SMEM_LIMIT((project_bwd_umma_kernel<NBI, true>), smem);
    project_bwd_umma_kernel<NBI, true><<<grid, PB_THREADS, smem, st>>>(a);
  } else {
    CFL_SMEM_LIMIT((project_bwd_umma_kernel<NBI, false>), smem);
    project_bwd_umma_kernel<NBI, false><<<grid, PB_THREADS, smem, st>>>(a);
  }
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

int project_bwd_umma(const float* x, int64_t B, int F, int64_t ldx, const float* dy, int64_t lddy, const float* y,
                     int64_t ldy, int act, int N, float* Cpart, cudaStream_t st) {
  PbArgs a;
  a.x = x; a.B = B; a.F = F; a.ldx = ldx; a.dy = dy; a.lddy = lddy; a.y = y; a.ldy = ldy; a.act = act;
  a.N = N; a.Npad = pb_npad(N);
  a.nst = pb_stages(a.Npad);
  a.rs = 0;
  a.dbg = 0;
  { const char* e = getenv("CFL_EXPERIMENTS"); if (e && atoi(e) != 0 && (e = getenv("CFL_PB_DBG"))) a.dbg = atoi(e); }
  int nh = 512 / a.Npad - 1;                                  // accumulator sets that fit TMEM, one is the cross terms'
  a.nh = nh > 3 ? 3 : (nh < 1 ? 1 : nh);
  a.slab_rows = project_bwd_umma_slab_rows(B, F);
  a.slabs = project_bwd_umma_slabs(B, F);
  a.ftiles = (F + 127) / 128;
  a.Cpart = Cpart;
  const size_t smem = (size_t)a.nst * PB_KPS * (PB_ASTEP + 64u * (size_t)a.Npad) + (2u * a.nst + 2u) * 8 + 64 + 1024;
  const int64_t njobs = (int64_t)a.ftiles * a.slabs;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  const int grid = (int)(njobs < sms ? njobs : sms);
  const int nbi = (4 * a.Npad + PB_PROD - 1) / PB_PROD;       // B-operand chunks per producer thread and stage
  const bool needy = act != CFL_ACT_LINEAR;
  if (!getenv("CFL_PROJECT_BWD_NO_TMA")) {
    const size_t stage = (size_t)PB_KPS * (PB_ASTEP + 64u * (size_t)a.Npad);
    const size_t rawb = (size_t)PB_ROWS * 128 * 4 + (needy ? 2 : 1) * (size_t)PB_ROWS * a.Npad * 4;
    const bool fits = (200 * 1024 - 3 * stage) / rawb >= 3 && B < ((int64_t)1 << 31);
    if (fits && make_tmap_2d_f32(&a.tmx, x, (uint64_t)F, (uint64_t)B, (uint64_t)ldx, 128, PB_ROWS) &&
        make_tmap_2d_f32(&a.tmd, dy, (uint64_t)N, (uint64_t)B, (uint64_t)lddy, (uint32_t)a.Npad, PB_ROWS) &&
        (!needy || make_tmap_2d_f32(&a.tmy, y, (uint64_t)N, (uint64_t)B, (uint64_t)ldy, (uint32_t)a.Npad, PB_ROWS))) {
      if (nbi <= 1) return pt_launch<1>(a, grid, st);
      if (nbi <= 2) return pt_launch<2>(a, grid, st);
      return pt_launch<4>(a, grid, st);
    }
  }
  if (nbi <= 1) return pb_launch<1>(a, smem, grid, st);
  if (nbi <= 2) return pb_launch<2>(a, smem, grid, st);
  return pb_launch<4>(a, smem, grid, st);
}

}  // namespace cfl
