// Per-query all-candidate AUC on the tensor cores: the rank counts of cfl_rank_counts
//   counts[q,j] = ( #{c : dist(q,c) < t[q,j]},  #{c : dist(q,c) == t[q,j]} )
// taken INSIDE the epilogue of the fused scoring kernel -- the Q x N distances never reach HBM (the earlier
// tensor-core route wrote and re-read a dense 4 GB matrix).  roc_auc_score of cfl/utils.py:267-268 over every
// catalog row follows from these integers (cfl.ranking.auc_from_rank_counts).
//
// The kernel is score_umma_kernel's skeleton (score_umma.cu: packed catalog image through a TMA ring, 3xTF32
// tcgen05.mma into double-buffered TMEM accumulators, soft-min epilogue of score.cuh) with a counting epilogue:
//   * every (row, query) distance D~ (Gram form, 3xTF32) gets a rounding band
//       m = S * (REL1 + REL2 * sqrt(D~ * V)),   S = |e|^2 + max_k |p_k|^2,   V = spread of the prototypes under the
//     soft-min weights (softmin_pair_var: the sensitivity of the distance to an error of the soft-min exponents, which
//     carry an absolute error proportional to S -- zero when the soft-min is saturated or K = 1);
//   * every pair is counted at once against each of the query's J thresholds by the sign of t - D~ (per-thread packed
//     4-bit counters, one predicated add per (row, query, threshold); a thread keeps its TMEM lane = catalog row
//     position for the whole pass; flushed through a warp reduction every 15 tiles);
//   * a pair with SOME threshold inside its band (min_j |t_j - D~| <= m, two thresholds per packed add + 3-input min;
//     exact ties included) -- a few 1e-4 of all pairs -- also goes to a record list: rank_fix_kernel evaluates it in
//     fp32 direct-difference form with the SAME device function as cfl_pair_dist_rows / cfl_rank_counts (direct.cuh)
//     and, for every threshold inside the band, replaces the fast verdict by the exact < / == verdicts.
// The counts therefore equal cfl_rank_counts' whenever |D~ - direct| <= m for every pair.  REL1 / REL2 are set to >= 4x
// the largest deviation measured over the bench and test shapes (tools/rank_counts_perf.py prints the observed
// maximum of |D~ - direct| / m over all ambiguous records; rank_fix_kernel reports it in the call's statistics and
// tests/test_rank_counts.py asserts the headroom).  A query whose records do not fit the list is recounted from
// scratch by the CUDA-core kernel (rank_count_kernel, restricted by `only`).
#include "score.cuh"
#include "umma.cuh"
#include <stdlib.h>
#include "direct.cuh"

namespace cfl {

using namespace umma;

constexpr int RT_NEPI = 16;                         // epilogue warps: four per TMEM lane quarter
constexpr int RT_THREADS = (RT_NEPI + 2) * 32;
constexpr int RT_EPI_THREADS = RT_NEPI * 32;
constexpr int RT_WPQ = RT_NEPI / 4;
constexpr int RT_NSTAGE = 6;
constexpr uint32_t RT_ASTAGE = 4u * 128u * 16u;     // one K-step of the catalog image: [hl][chunk][128 rows][16 B]
constexpr int RT_JC = 8;                            // thresholds per launch (J > 8: several launches)
constexpr int RT_QT_MAX = 64;
constexpr float RT_REL1_DEFAULT = 1.0f / 262144.0f;     // 2^-18: Gram-form (3xTF32) + direct-form rounding, relative to S
constexpr float RT_REL2_DEFAULT = 1.0f / 2097152.0f;    // 2^-21: exponent error relative to S (sensitivity term)

int rank_tc_qt(int K, int d) {
  int qt = score_umma_qt(K, d);
  return qt > RT_QT_MAX ? RT_QT_MAX : qt;
}

struct RtArgs {
  int K, d, J, j0;                // this launch counts thresholds [j0, j0 + min(RT_JC, J - j0))
  int64_t Q, N;
  int qt, nqt, parts, dpad;
  int64_t tiles;
  const float* qpar;
  const float* qplane;            // per query: affine-hull bound block (score.cuh)
  const void* qimg;
  const void* cimg;
  const float* e2;
  const float* thr;               // [Q, J]
  unsigned long long* counts;     // [Q, J, 2]
  uint4* recs;                    // ambiguous (query, row) records
  unsigned int rec_slice;         // records per CTA slice of the list
  unsigned int* rec_n;            // [CTAs] records written per slice
  int* only;                      // [Q]: 1 = records dropped, recount this query from scratch
  float rel1, rel2;
};

struct RtLayout { uint32_t b_img, a_ring, qpar, qpl, thr, qm, skp, cs, rcn, bars, tmem_slot, total; };
__host__ __device__ inline RtLayout rt_layout(int K, int qt, int dpad) {
  RtLayout L;
  uint32_t off = 0;
  L.b_img = off;   off += (uint32_t)dpad * 8u * (uint32_t)(K * qt);  off = (off + 1023u) & ~1023u;
  L.a_ring = off;  off += RT_NSTAGE * 2u * RT_ASTAGE;
  L.qpar = off;    off += (uint32_t)qt * (uint32_t)qpar_stride(K) * 4u;  off = (off + 15u) & ~15u;
  L.qpl = off;     off += (uint32_t)qt * (uint32_t)qplane_stride(K) * 4u; off = (off + 15u) & ~15u;
  L.thr = off;     off += (uint32_t)qt * RT_JC * 4u;
  L.qm = off;      off += (uint32_t)qt * 4u;                           off = (off + 15u) & ~15u;
  L.skp = off;     off += (uint32_t)qt * 16u;                          // per query: (tmax, kappa0, kappa1, -)
  L.cs = off;      off += (uint32_t)qt * RT_JC * 4u;
  L.rcn = off;     off += 16u;
  L.bars = off;    off += (2u * RT_NSTAGE + 5u) * 8u;
  L.tmem_slot = off; off += 16u;
  L.total = off;
  return L;
}

// cnt += inc when x > 0 (x = t - D~: the pair is closer than the threshold), without a branch.
__device__ __forceinline__ void count_pos(uint32_t& cnt, float x, uint32_t inc) {
  asm("{\n\t.reg .pred p;\n\t"
      "setp.gt.f32 p, %1, 0f00000000;\n\t"
      "@p add.u32 %0, %0, %2;\n\t}"
      : "+r"(cnt)
      : "f"(x), "r"(inc));
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int K>
__global__ void __launch_bounds__(RT_THREADS, 1)     // 18 warps = 5 on one sub-partition: 5 x 32 x 96 registers is its whole file
rank_count_umma_kernel(RtArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int GQ = K <= 2 ? 16 : (K <= 4 ? 8 : 4);     // queries per epilogue group
  constexpr int NGW = K <= 2 ? 1 : (K <= 4 ? 2 : 3);     // query groups per warp: ceil((qt / GQ) / RT_WPQ), qt <= 64 (48 at K = 5)
  constexpr int QPS = qpar_stride(K);
  const int QT = A.qt;
  const int NC = K * QT;
  const int dpad = A.dpad;
  const int nks = dpad / 8;
  const RtLayout L = rt_layout(K, QT, dpad);
  unsigned char* b_img = smem + L.b_img;
  unsigned char* a_ring = smem + L.a_ring;
  float* qpar = (float*)(smem + L.qpar);
  float* thr = (float*)(smem + L.thr);
  float* qm = (float*)(smem + L.qm);
  float* qpl = (float*)(smem + L.qpl);
  float4* skp = (float4*)(smem + L.skp);
  unsigned int* cs = (unsigned int*)(smem + L.cs);
  unsigned int* rcn = (unsigned int*)(smem + L.rcn);
  uint64_t* full = (uint64_t*)(smem + L.bars);
  uint64_t* empty = full + RT_NSTAGE;
  uint64_t* tfull = empty + RT_NSTAGE;
  uint64_t* tempty = tfull + 2;
  uint64_t* bfull = tempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(smem + L.tmem_slot);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int part = blockIdx.x, qtile = blockIdx.y;
  const int64_t q0 = (int64_t)qtile * QT;
  const int nq = (int)((A.Q - q0 < QT) ? (A.Q - q0) : QT);
  const int64_t t0 = A.tiles * part / A.parts;
  const int64_t t1 = A.tiles * (part + 1) / A.parts;
  const int ntiles = (int)(t1 - t0);
  const int kss = (nks % 2 == 0) ? 2 : 1;                    // K-steps per ring stage
  const int jn = (A.J - A.j0 < RT_JC) ? (A.J - A.j0) : RT_JC;

  uint32_t ncols = 32;
  while ((int)ncols < 2 * NC) ncols <<= 1;
  if (warp == RT_NEPI) {
    if (lane == 0) {
      for (int s = 0; s < RT_NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
      mbar_init(&tempty[0], RT_EPI_THREADS); mbar_init(&tempty[1], RT_EPI_THREADS);
      mbar_init(bfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, ncols);
  }
  // per-query soft-min blocks, interleaved by query pair with the first K entries negated (softmin_pair)
  for (int i = tid; i < QT * QPS; i += RT_THREADS) {
    const int ql = i / QPS, j = i % QPS;
    float v = (ql < nq) ? A.qpar[(q0 + ql) * QPS + j] : 0.0f;
    if (j < K) v = -v;
    qpar[((ql >> 1) * QPS + j) * 2 + (ql & 1)] = v;
  }
  // thresholds: NaN (no positive in the slot), padding slots and padding queries become -inf = "count nothing"
  for (int i = tid; i < QT * RT_JC; i += RT_THREADS) {
    const int ql = i / RT_JC, j = i % RT_JC;
    float t = __int_as_float(0xff800000);
    if (ql < nq && j < jn) {
      const float v = A.thr[(q0 + ql) * A.J + A.j0 + j];
      if (v == v) t = fminf(v, 3.4028234e38f);             // +inf -> FLT_MAX: t - D~ stays -inf for padding rows
    }
    thr[i] = t;
    cs[i] = 0u;
  }
  for (int ql = tid; ql < QT; ql += RT_THREADS) {
    float pm = 0.0f;
    if (ql < nq)
      for (int k = 0; k < K; ++k) pm = fmaxf(pm, A.qpar[(q0 + ql) * QPS + k]);      // log2(e) |p_k|^2
    qm[ql] = pm * CFL_LN2;                                    // max_k |p_k|^2
  }
  {
    constexpr int PBS0 = qplane_stride(K);                     // pair interleave, as the soft-min blocks
    for (int i = tid; i < QT * PBS0; i += RT_THREADS) {
      const int ql = i / PBS0, j = i % PBS0;
      qpl[((ql >> 1) * PBS0 + j) * 2 + (ql & 1)] = (ql < nq) ? A.qplane[(q0 + ql) * PBS0 + j] : 0.0f;
    }
  }
  // Skip test of a whole query group (epilogue): with LB <= D~ the affine-hull lower bound and Vq >= V,
  //   D~ - m(D~) >= LB - S (rel1 + rel2 sqrt(LB Vq)) >= LB (1 - S k1) - S k0   (AM-GM with c = sqrt(Vq / tmax)),
  // k0 = rel1 + rel2 Vq / (2c), k1 = rel2 c / 2: when that exceeds the query's largest threshold the pair is farther than
  // every threshold and outside every band -- nothing to count, nothing to record.
  for (int ql = tid; ql < QT; ql += RT_THREADS) {
    float tmax = __int_as_float(0xff800000);
    float k0 = A.rel1, k1 = 0.0f;
    if (ql < nq) {
      for (int j = 0; j < jn; ++j) {
        const float v = A.thr[(q0 + ql) * A.J + A.j0 + j];
        if (v == v) tmax = fmaxf(tmax, fminf(v, 3.4028234e38f));
      }
      if (K > 1) {
        const float cq = A.qpar[(q0 + ql) * QPS + K + qpar_tri(K)];       // >= (1 - 1/K) / 2 * max_kl |p_k - p_l|^2
        const float vq = cq / (0.5f * (1.0f - 1.0f / (float)K));          // >= the prototypes' spread under any weights
        const float c = sqrtf(vq / fmaxf(tmax, 1.0e-30f));
        k0 = A.rel1 + 0.5f * A.rel2 * vq / fmaxf(c, 1.0e-30f);
        k1 = 0.5f * A.rel2 * c;
        if (!(c == c) || !(k0 == k0) || !(k1 == k1) || k1 > 1.0e30f) { k0 = __int_as_float(0x7f800000); k1 = 0.0f; }   // never skip
      }
    }
    skp[ql] = make_float4(tmax, k0 * 1.0001f, k1 * 1.0001f, 0.0f);
  }
  if (tid == 0) *rcn = 0u;
  const unsigned int cta = blockIdx.y * gridDim.x + blockIdx.x;
  uint4* myrecs = A.recs + (size_t)cta * A.rec_slice;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == RT_NEPI) {
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
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          for (int j = 0; j < kss; ++j)
            mma_step3(sd, d_tmem, a_base + (stage * 2 + j) * RT_ASTAGE, b_base + (ks + j) * b_step, ks + j == 0);
          mma_commit(&empty[stage]);
          if (++stage == RT_NSTAGE) { stage = 0; phase ^= 1u; }
        }
        mma_commit(&tfull[buf]);
      }
    }
  } else if (warp == RT_NEPI + 1) {
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
      for (int t = 0; t < ntiles; ++t) {
        const unsigned char* src = (const unsigned char*)A.cimg + (size_t)(t0 + t) * nks * RT_ASTAGE;
        for (int ks = 0; ks < nks; ks += kss) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], (uint32_t)kss * RT_ASTAGE);
          bulk_g2s(a_ring + stage * 2 * RT_ASTAGE, src + (size_t)ks * RT_ASTAGE, (uint32_t)kss * RT_ASTAGE, &full[stage]);
          if (++stage == RT_NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ======================================= epilogue ========================================
    const int lq = warp & 3, sub = warp >> 2;
    const int lrow = lq * 32 + lane;
    const uint32_t lane_lt = (1u << lane) - 1u;
    const float INF = __int_as_float(0x7f800000);
    // c[gi][i]: eight packed 4-bit counters (threshold j in nibble j) for query i of this warp's gi-th group
    uint32_t c[NGW][GQ];
#pragma unroll
    for (int gi = 0; gi < NGW; ++gi)
#pragma unroll
      for (int i = 0; i < GQ; ++i) c[gi][i] = 0u;

    auto flush = [&]() {
#pragma unroll
      for (int gi = 0; gi < NGW; ++gi) {
        const int g = sub + gi * RT_WPQ;
        if (g * GQ < nq) {                                    // warp-uniform
#pragma unroll
          for (int i = 0; i < GQ; ++i) {
            const uint32_t x = c[gi][i];
            c[gi][i] = 0u;
            const uint32_t ev = x & 0x0f0f0f0fu, od = (x >> 4) & 0x0f0f0f0fu;      // thresholds 0,2,4,6 | 1,3,5,7
            const uint32_t r0 = __reduce_add_sync(0xffffffffu, ev & 0x00ff00ffu);          // j = 0 | 4
            const uint32_t r1 = __reduce_add_sync(0xffffffffu, od & 0x00ff00ffu);          // j = 1 | 5
            const uint32_t r2 = __reduce_add_sync(0xffffffffu, (ev >> 8) & 0x00ff00ffu);   // j = 2 | 6
            const uint32_t r3 = __reduce_add_sync(0xffffffffu, (od >> 8) & 0x00ff00ffu);   // j = 3 | 7
            if (lane < RT_JC) {                               // lane j adds threshold j's sum
              const uint32_t lo2 = (lane & 1) ? r1 : r0, hi2 = (lane & 1) ? r3 : r2;
              const uint32_t r = (lane & 2) ? hi2 : lo2;
              const uint32_t v = (lane & 4) ? (r >> 16) : (r & 0xffffu);
              if (v) atomicAdd(cs + (g * GQ + i) * RT_JC + lane, v);
            }
          }
        }
      }
    };

    int since_flush = 0;
    // the skip test pays when most groups are far from every threshold (labelled positives near the top of the ranking:
    // the usual evaluation); each warp measures its hit rate over the first 16 tiles and switches the test off otherwise
    bool bound_on = true;
    int grp_seen = 0, grp_skipped = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const int64_t row = (t0 + t) * 128 + lrow;
      const bool valid = row < A.N;
      const float e2 = __ldg(A.e2 + row);                    // padded to whole tiles
      mbar_wait(&tfull[buf], (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * NC);
#pragma unroll
      for (int gi = 0; gi < NGW; ++gi) {
        const int g = sub + gi * RT_WPQ;
        if (g * GQ >= nq) continue;                           // warp-uniform
        float gk[K][GQ];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if constexpr (GQ == 16)     tmem_ld16(taddr + (uint32_t)(k * QT + g * GQ), gk[k]);
          else if constexpr (GQ == 8) tmem_ld8(taddr + (uint32_t)(k * QT + g * GQ), gk[k]);
          else                        tmem_ld4(taddr + (uint32_t)(k * QT + g * GQ), gk[k]);
        }
        tmem_ld_wait();
        if (bound_on) {
          constexpr int PBS = qplane_stride(K);
          const float e2s = e2 * (1.0f - CFL_PLANE_REL);
          const f2_t e2sp = pk2(e2s, e2s);
          bool need = false;
#pragma unroll
          for (int pi = 0; pi < GQ / 2; ++pi) {
            f2_t pv[PBS];
            const ulonglong2* src = (const ulonglong2*)(qpl + (g * (GQ / 2) + pi) * PBS * 2);
#pragma unroll
            for (int j = 0; j < PBS; j += 2) { const ulonglong2 u = src[j >> 1]; pv[j] = u.x; pv[j + 1] = u.y; }
            float gA[K], gB[K], lbA, lbB;
#pragma unroll
            for (int k = 0; k < K; ++k) { gA[k] = gk[k][2 * pi]; gB[k] = gk[k][2 * pi + 1]; }
            plane_bound_pair<K>(gA, gB, e2sp, pv, lbA, lbB);
            const int ql = g * GQ + 2 * pi;
            const float4 sa = skp[ql], sb = skp[ql + 1];
            const float SA = e2 + qm[ql], SB = e2 + qm[ql + 1];
            const float fa = fmaf(-SA, sa.z, 1.0f), fb = fmaf(-SB, sb.z, 1.0f);     // 1 - S k1: must stay positive
            const float va = fmaf(lbA, fa, -SA * sa.y);
            const float vb = fmaf(lbB, fb, -SB * sb.y);
            need |= !(va > sa.x) | !(vb > sb.x) | !(fminf(fa, fb) > 0.0f);
          }
          ++grp_seen;
          if (!__any_sync(0xffffffffu, need && valid)) { ++grp_skipped; continue; }
        }
        const float* qg = qpar + g * GQ * QPS;
        float dist[GQ], var[GQ];
#pragma unroll
        for (int pi = 0; pi < GQ / 2; ++pi) {
          f2_t qv[QPS];
          const ulonglong2* src = (const ulonglong2*)(qg + pi * QPS * 2);
#pragma unroll
          for (int j = 0; j < QPS; j += 2) { const ulonglong2 u = src[j >> 1]; qv[j] = u.x; qv[j + 1] = u.y; }
          float gA[K], gB[K];
#pragma unroll
          for (int k = 0; k < K; ++k) { gA[k] = gk[k][2 * pi]; gB[k] = gk[k][2 * pi + 1]; }
          softmin_pair_var<K>(gA, gB, e2, qv, dist[2 * pi], dist[2 * pi + 1], var[2 * pi], var[2 * pi + 1]);
        }
        uint32_t amb = 0u;
#pragma unroll
        for (int i = 0; i < GQ; ++i) {
          const int ql = g * GQ + i;
          // band m = S * (rel1 + rel2 * sqrt(D~ * V)); var[] keeps it for the record
          const float S = e2 + qm[ql];
          const float sens = K == 1 ? 0.0f : sqrt_approx(fmaxf(dist[i], 0.0f) * fmaxf(var[i], 0.0f));
          const float m = S * fmaf(A.rel2, sens, A.rel1);
          var[i] = m;
          if (!valid) dist[i] = INF;                          // padding rows: t - inf = -inf, never counted, never in a band
          const f2_t nd = pk2(-dist[i], -dist[i]);
          const ulonglong2 ta = *(const ulonglong2*)(thr + ql * RT_JC);
          const ulonglong2 tb = *(const ulonglong2*)(thr + ql * RT_JC + 4);
          const f2_t tp[RT_JC / 2] = {ta.x, ta.y, tb.x, tb.y};
          float near = INF;
#pragma unroll
          for (int jp = 0; jp < RT_JC / 2; ++jp) {
            float x, y;
            upk2(add2(tp[jp], nd), x, y);                     // t_j - D~ for two thresholds
            near = fminf(fminf(fabsf(x), fabsf(y)), near);
            count_pos(c[gi][i], x, 1u << (8 * jp));
            count_pos(c[gi][i], y, 1u << (8 * jp + 4));
          }
          if (!(near > m)) amb |= 1u << i;                    // some threshold inside the band (or NaN)
        }
        if (__any_sync(0xffffffffu, amb != 0u)) {
#pragma unroll
          for (int i = 0; i < GQ; ++i) {
            const bool mine = (amb >> i) & 1u;
            const uint32_t mk = __ballot_sync(0xffffffffu, mine);
            if (mk == 0u) continue;                           // warp-uniform
            const int ql = g * GQ + i;
            const int leader = __ffs(mk) - 1;
            unsigned int basei = 0;
            if (lane == leader) basei = atomicAdd(rcn, (unsigned int)__popc(mk));
            basei = __shfl_sync(0xffffffffu, basei, leader);
            if (mine) {
              const unsigned int at = basei + (unsigned int)__popc(mk & lane_lt);
              if (at < A.rec_slice) {
                myrecs[at] = make_uint4((uint32_t)(q0 + ql), (uint32_t)row, __float_as_uint(dist[i]), __float_as_uint(var[i]));
              } else {
                A.only[q0 + ql] = 1;
              }
            }
          }
        }
      }
      if (t == 15 && grp_skipped * 2 < grp_seen) bound_on = false;
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
      if (++since_flush == 15) { flush(); since_flush = 0; }
    }
    flush();
    asm volatile("bar.sync 1, %0;" ::"n"(RT_EPI_THREADS) : "memory");
    for (int i = tid; i < nq * RT_JC; i += RT_EPI_THREADS) {
      const int ql = i / RT_JC, j = i % RT_JC;
      if (j < jn && cs[i]) atomicAdd(&A.counts[((q0 + ql) * A.J + A.j0 + j) * 2], (unsigned long long)cs[i]);
    }
    if (tid == 0) A.rec_n[cta] = *rcn < A.rec_slice ? *rcn : A.rec_slice;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RT_NEPI) tmem_dealloc(tmem_base, ncols);
}

// Exact verdicts of the ambiguous pairs: one thread per record, the direct-form device function of direct.cuh on the raw
// rows (bit-identical to cfl_pair_dist_rows / cfl_rank_counts; the operands are staged in shared memory -- the slice's
// query tile once per block, 32 catalog rows per warp and step with coalesced loads -- which changes no arithmetic).  stats: [0] records, [1] max |D~ - direct| / m in units
// of 2^-20, [2] records whose deviation exceeded m / 2.
constexpr int FX_THREADS = 128;
constexpr int FX_BLOCKS = 4;                        // blocks per list slice
template <int K>
__global__ void __launch_bounds__(FX_THREADS)
rank_fix_kernel(const uint4* __restrict__ recs_all, const unsigned int* __restrict__ rec_n, unsigned int slice,
                const float* __restrict__ Pq, int64_t ldq, const float* __restrict__ E, int64_t lde, int64_t N, int64_t Q,
                int d, int qt, int parts, const float* __restrict__ thr, int J, int j0, const int* __restrict__ only,
                unsigned long long* __restrict__ counts, unsigned long long* __restrict__ stats) {
  extern __shared__ __align__(16) float fx_smem[];
  const unsigned int n = rec_n[blockIdx.y];                  // grid.y = the scoring kernel's CTAs, one list slice each
  if (n == 0 || (unsigned int)blockIdx.x * FX_THREADS >= n) return;
  const uint4* recs = recs_all + (size_t)blockIdx.y * slice;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int KD = K * d, ldp = KD + 1, lder = d + 1;
  float* pq = fx_smem;                                       // [qt][K*d + 1]: the slice's query tile
  float* er = pq + (size_t)qt * ldp;                         // [FX_THREADS][d + 1]: one catalog row per thread
  const int64_t q0 = (int64_t)(blockIdx.y / parts) * qt;     // CTA (part, qtile) of the scoring kernel = qtile * parts + part
  const int nq = (int)((Q - q0 < qt) ? (Q - q0) : qt);
  for (int i = tid; i < nq * KD; i += FX_THREADS) {
    const int ql = i / KD, e = i - ql * KD;
    pq[ql * ldp + e] = Pq[(q0 + ql) * ldq + e];
  }
  __syncthreads();
  unsigned long long worst = 0, over = 0;
  for (unsigned int c0 = blockIdx.x * FX_THREADS; c0 < n; c0 += gridDim.x * FX_THREADS) {
    const unsigned int i = c0 + tid;
    const bool have = i < n;
    const uint4 r = have ? recs[i] : make_uint4(0u, 0u, 0u, 0u);
    const int64_t q = r.x, row = r.y;
    const bool live = have && !only[q] && row < N;
    // the warp gathers its 32 rows with coalesced loads; every thread then walks its own row in shared memory
    __syncwarp();
    for (int rr = 0; rr < 32; ++rr) {
      const int64_t rw = __shfl_sync(0xffffffffu, (unsigned int)r.y, rr);
      const bool ok = __shfl_sync(0xffffffffu, live ? 1 : 0, rr) != 0;
      if (!ok) continue;                                     // warp-uniform
      float* dst = er + (wid * 32 + rr) * lder;
      for (int j = lane; j < d; j += 32) dst[j] = E[rw * lde + j];
    }
    __syncwarp();
    if (!live) continue;
    const float* ev = er + tid * lder;
    const float* qv = pq + (int)(q - q0) * ldp;
    const float dist = pcd_direct<K>([&](int j) { return ev[j]; }, [&](int k, int j) { return qv[k * d + j]; }, d);
    const float Dg = __uint_as_float(r.z), m = __uint_as_float(r.w);
    for (int j = 0; j < RT_JC && j0 + j < J; ++j) {
      const float t = thr[q * J + j0 + j];
      if (!(t == t)) continue;                               // NaN: no positive in this slot
      const float x = fminf(t, 3.4028234e38f) + (-Dg);       // the scoring kernel's arithmetic, bit for bit
      if (fabsf(x) > m) continue;                            // outside the band: the fast verdict stands
      const int fast = x > 0.0f ? 1 : 0;                     // what the scoring kernel counted
      const int lt = dist < t ? 1 : 0;
      if (lt != fast) atomicAdd(&counts[((q * J) + j0 + j) * 2], (unsigned long long)(long long)(lt - fast));
      if (dist == t) atomicAdd(&counts[((q * J) + j0 + j) * 2 + 1], 1ull);
    }
    const float dev = fabsf(Dg - dist);
    const float ratio = m > 0.0f ? dev / m : (dev > 0.0f ? 4096.0f : 0.0f);
    const unsigned long long u = (unsigned long long)fminf(ratio * 1048576.0f, 4.0e9f);
    if (u > worst) worst = u;
    if (ratio > 0.5f) ++over;
  }
  if (stats) {
    if (worst) atomicMax(&stats[1], worst);
    if (over) atomicAdd(&stats[2], over);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[0], (unsigned long long)n);
  }
}

// queries whose records were dropped are recounted from scratch: clear what the fast path added
__global__ void rank_zero_flagged_kernel(const int* __restrict__ only, int64_t Q, int J, unsigned long long* counts,
                                         unsigned long long* stats) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q || !only[q]) return;
  for (int i = 0; i < 2 * J; ++i) counts[q * 2 * J + i] = 0ull;
  if (stats) atomicAdd(&stats[3], 1ull);
}

static void rank_tc_rel(float* rel1, float* rel2) {
  *rel1 = RT_REL1_DEFAULT; *rel2 = RT_REL2_DEFAULT;
  const char* x = getenv("CFL_EXPERIMENTS");
  if (!x || atoi(x) == 0) return;
  const char* e;
  if ((e = getenv("CFL_RANK_REL1"))) *rel1 = (float)atof(e);
  if ((e = getenv("CFL_RANK_REL2"))) *rel2 = (float)atof(e);
}

static unsigned int rank_tc_rec_cap(int64_t Q, int64_t N) {
  const char* e = getenv("CFL_EXPERIMENTS");
  if (e && atoi(e) != 0 && (e = getenv("CFL_RANK_REC_CAP")) && atoi(e) > 0) {     // tests: force the overflow path
    const unsigned int c = (unsigned int)atoi(e);
    return c < 1048576u ? c : 1048576u;
  }
  double pairs = (double)Q * (double)N;
  double cap = pairs / 256.0;
  if (cap < 1048576.0) cap = 1048576.0;
  if (cap > 268435456.0) cap = 268435456.0;
  return (unsigned int)cap;
}

struct RtWs { size_t pc, qpar, qimg, recs, recn, only, stats, total; };
static RtWs rank_tc_ws(int64_t Q, int K, int d, int64_t N) {
  RtWs w;
  const int qt = rank_tc_qt(K, d);
  const int nqt = qt > 0 ? (int)((Q + qt - 1) / qt) : 0;
  const int dpad = (d + 7) / 8 * 8;
  size_t off = 0;
  w.pc = off;    off = align_up(off + (size_t)Q * K * d * 4, 256);
  w.qpar = off;  off = align_up(off + (size_t)Q * (qpar_stride(K) + qplane_stride(K)) * 4, 256);
  w.qimg = off;  off = align_up(off + (size_t)nqt * dpad * 8 * (size_t)(K * qt), 1024);
  w.recs = off;  off = align_up(off + (size_t)rank_tc_rec_cap(Q, N) * sizeof(uint4), 256);
  w.recn = off;  off = align_up(off + 4096 * sizeof(unsigned int), 256);
  w.only = off;  off = align_up(off + (size_t)Q * sizeof(int), 256);
  w.stats = off; off = align_up(off + CFL_RANK_NSTATS * sizeof(unsigned long long), 256);
  w.total = off + 1024;
  return w;
}

template <int K>
static int rank_tc_launch(const RtArgs& a, const float* Pq, int64_t ldq, const float* E, int64_t lde,
                          unsigned long long* stats, cudaStream_t cs) {
  const RtLayout L = rt_layout(K, a.qt, a.dpad);
  const size_t smem = L.total + 1024;
  CFL_SMEM_LIMIT(rank_count_umma_kernel<K>, smem);
  dim3 grid((unsigned)a.parts, (unsigned)a.nqt);
  rank_count_umma_kernel<K><<<grid, RT_THREADS, smem, cs>>>(a);
  CFL_LAUNCH_CHECK();
  const size_t fx_smem = ((size_t)a.qt * (K * a.d + 1) + (size_t)FX_THREADS * (a.d + 1)) * sizeof(float);
  CFL_SMEM_LIMIT(rank_fix_kernel<K>, fx_smem);
  rank_fix_kernel<K><<<dim3(FX_BLOCKS, (unsigned)(a.parts * a.nqt)), FX_THREADS, fx_smem, cs>>>(
      a.recs, a.rec_n, a.rec_slice, Pq, ldq, E, lde, a.N, a.Q, a.d, a.qt, a.parts, a.thr, a.J, a.j0, a.only, a.counts, stats);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

}  // namespace cfl

using namespace cfl;

extern "C" {

size_t cfl_rank_counts_packed_workspace_bytes(int64_t Q, int K, int d, int64_t N) {
  if (Q <= 0 || N <= 0 || K < 1 || K > CFL_MAX_K || d < 1 || d > 128) return 4096;
  return rank_tc_ws(Q, K, d, N).total;
}

int cfl_rank_counts_packed(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* E,
                           const void* image, int64_t N, int64_t lde, const float* mu, const float* pos_dist, int J,
                           int64_t* counts, void* ws, size_t ws_bytes, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(mode == CFL_PCD || mode == CFL_SIAMESE, CFL_ERR_UNSUPPORTED,
              "rank_counts_packed: mode %d not supported (pcd, siamese)", mode);
  CFL_REQUIRE(mode != CFL_SIAMESE || K == 1, CFL_ERR_INVALID, "rank_counts_packed: siamese needs K=1");
  CFL_REQUIRE(K >= 1 && K <= CFL_MAX_K, CFL_ERR_UNSUPPORTED, "rank_counts_packed: K=%d outside [1,%d]", K, CFL_MAX_K);
  CFL_REQUIRE(d >= 1 && d <= 128, CFL_ERR_UNSUPPORTED, "rank_counts_packed: d=%d outside [1,128]", d);
  CFL_REQUIRE(J >= 1 && J <= CFL_MAX_RANK_J, CFL_ERR_UNSUPPORTED, "rank_counts_packed: J=%d outside [1,%d]", J, CFL_MAX_RANK_J);
  CFL_REQUIRE(Q >= 0 && N >= 0 && N < ((int64_t)1 << 32) && Q < ((int64_t)1 << 31), CFL_ERR_INVALID, "rank_counts_packed: bad Q/N");
  CFL_REQUIRE(ldq >= (int64_t)K * d && lde >= d, CFL_ERR_INVALID, "rank_counts_packed: leading dimension too small");
  if (Q == 0) return CFL_OK;
  CFL_REQUIRE(Pq && pos_dist && counts, CFL_ERR_INVALID, "rank_counts_packed: NULL argument");
  cudaStream_t cs = (cudaStream_t)stream;
  CFL_CUDA(cudaMemsetAsync(counts, 0, (size_t)Q * J * 2 * sizeof(int64_t), cs));
  if (N == 0) return CFL_OK;
  CFL_REQUIRE(E && image, CFL_ERR_INVALID, "rank_counts_packed: NULL catalog / image");
  const int qt = rank_tc_qt(K, d);
  CFL_REQUIRE(qt > 0 && score_umma_supported(K, d), CFL_ERR_UNSUPPORTED,
              "rank_counts_packed: shape K=%d d=%d has no tcgen05 tiling", K, d);
  const RtWs w = rank_tc_ws(Q, K, d, N);
  CFL_REQUIRE(ws && ws_bytes >= w.total, CFL_ERR_WORKSPACE, "rank_counts_packed: workspace too small (%zu < %zu)",
              ws_bytes, w.total);
  char* base = (char*)ws;
  float* Pc = (float*)(base + w.pc);
  float* qpar = (float*)(base + w.qpar);
  float* qplane = qpar + (size_t)Q * qpar_stride(K);
  st = score_prep_queries_launch(Pq, Q, K, d, ldq, mu, Pc, qpar, qplane, cs);
  if (st != CFL_OK) return st;
  ScoreArgs sa = {};
  sa.K = K; sa.d = d; sa.Q = Q; sa.Pc = Pc;
  sa.plan.qt = qt; sa.plan.nqt = (int)((Q + qt - 1) / qt); sa.plan.dpad = (d + 7) / 8 * 8;
  st = score_umma_pack_queries(sa, base + w.qimg, cs);
  if (st != CFL_OK) return st;

  RtArgs a = {};
  a.K = K; a.d = d; a.J = J; a.Q = Q; a.N = N;
  a.qt = qt; a.nqt = sa.plan.nqt; a.dpad = sa.plan.dpad;
  a.tiles = (N + 127) / 128;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int64_t parts = sms / a.nqt;
  if (parts < 1) parts = 1;
  if (parts > a.tiles) parts = a.tiles;
  a.parts = (int)parts;
  a.qpar = qpar; a.qplane = qplane; a.qimg = base + w.qimg; a.cimg = image;
  a.e2 = (const float*)((const char*)image + (size_t)a.tiles * (a.dpad / 8) * 8192);
  a.thr = pos_dist; a.counts = (unsigned long long*)counts;
  CFL_REQUIRE((int64_t)a.parts * a.nqt <= 4096, CFL_ERR_UNSUPPORTED, "rank_counts_packed: Q=%lld needs more than 4096 CTAs",
              (long long)Q);
  a.recs = (uint4*)(base + w.recs); a.rec_slice = rank_tc_rec_cap(Q, N) / (unsigned int)(a.parts * a.nqt);
  a.rec_n = (unsigned int*)(base + w.recn);
  a.only = (int*)(base + w.only);
  rank_tc_rel(&a.rel1, &a.rel2);
  unsigned long long* stats = (unsigned long long*)(base + w.stats);
  CFL_CUDA(cudaMemsetAsync(a.only, 0, (size_t)Q * sizeof(int), cs));
  CFL_CUDA(cudaMemsetAsync(stats, 0, CFL_RANK_NSTATS * sizeof(unsigned long long), cs));
  timer_record(0, cs);
  for (int j0 = 0; j0 < J; j0 += RT_JC) {
    a.j0 = j0;
    CFL_CUDA(cudaMemsetAsync(a.rec_n, 0, 4096 * sizeof(unsigned int), cs));
    switch (K) {
      case 1: st = rank_tc_launch<1>(a, Pq, ldq, E, lde, stats, cs); break;
      case 2: st = rank_tc_launch<2>(a, Pq, ldq, E, lde, stats, cs); break;
      case 3: st = rank_tc_launch<3>(a, Pq, ldq, E, lde, stats, cs); break;
      case 4: st = rank_tc_launch<4>(a, Pq, ldq, E, lde, stats, cs); break;
      case 5: st = rank_tc_launch<5>(a, Pq, ldq, E, lde, stats, cs); break;
      case 6: st = rank_tc_launch<6>(a, Pq, ldq, E, lde, stats, cs); break;
      case 7: st = rank_tc_launch<7>(a, Pq, ldq, E, lde, stats, cs); break;
      default: st = rank_tc_launch<8>(a, Pq, ldq, E, lde, stats, cs); break;
    }
    if (st != CFL_OK) return st;
  }
  timer_record(1, cs);
  // queries whose ambiguous records overflowed the list (the flags are final only now: a later threshold chunk may
  // have raised one): clear them and recount them from scratch on the CUDA cores
  rank_zero_flagged_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, cs>>>(a.only, Q, J, a.counts, stats);
  CFL_LAUNCH_CHECK();
  return rank_counts_only_launch(mode, Pq, Q, K, d, ldq, E, N, lde, pos_dist, J, a.counts, a.only, cs);
}

int cfl_rank_counts_packed_stats(int64_t Q, int K, int d, int64_t N, const void* ws, size_t ws_bytes, int64_t* out,
                                 void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(out, CFL_ERR_INVALID, "rank_counts_packed_stats: NULL argument");
  for (int i = 0; i < CFL_RANK_NSTATS; ++i) out[i] = 0;
  if (Q <= 0 || N <= 0) return CFL_OK;
  const RtWs w = rank_tc_ws(Q, K, d, N);
  CFL_REQUIRE(ws && ws_bytes >= w.total, CFL_ERR_WORKSPACE, "rank_counts_packed_stats: workspace too small");
  cudaStream_t cs = (cudaStream_t)stream;
  CFL_CUDA(cudaMemcpyAsync(out, (const char*)ws + w.stats, CFL_RANK_NSTATS * sizeof(int64_t), cudaMemcpyDeviceToHost, cs));
  CFL_CUDA(cudaStreamSynchronize(cs));
  return CFL_OK;
}

}  // extern "C"
