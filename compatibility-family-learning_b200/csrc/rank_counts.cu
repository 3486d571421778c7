// Per-query all-candidate AUC (SURVEY 8d C3 "per-query all-candidate AUC with J planted positives",
// 8e "AUC across ranks"): the rank statistic of roc_auc_score (cfl/utils.py:267-268) for ONE query
// against EVERY catalog row, without materialising the Q x N scores.
//
//   cfl_pair_dist_rows : t[q,j] = dist(query q, catalog row pos_idx[q,j])        (the labelled positives)
//   cfl_rank_counts    : counts[q,j] = ( #{c : dist(q,c) < t[q,j]},  #{c : dist(q,c) == t[q,j]} )  over a shard
//
// Both evaluate the pair scorer (DistBase.build_dist, cfl/models/base.py:107-146) in fp32
// direct-difference form through the SAME device function, operation for operation, so a positive
// compares equal to itself and the integer counts are exact for the fp32 distances; they add over
// catalog shards (int64 all-reduce), which makes the AUC independent of the sharding.  The pcd
// arithmetic is that of merge_rescore_kernel (score.cu), i.e. of the values cfl_score_topk reports;
// the monomer arithmetic is that of score_monomer_kernel.
#include "score.cuh"
#include "direct.cuh"

namespace cfl {

constexpr int RC_QT = 8;          // queries per CTA
constexpr int RC_THREADS = 128;   // one thread per catalog row of a 128-row tile
constexpr int RC_MAX_J = 32;

struct RcArgs {
  int mode, K, d, J;
  int64_t Q, N, ldq, lde;
  const float* Pq;         // pcd: [Q, ldq] K prototypes per query; monomer: [Q, ldq] embedding a_q
  const float* Wq;         // monomer: [Q, K]
  const float* E;          // pcd: [N, lde] embeddings; monomer: [N, lde] K prototypes per row
  const float* thr;        // [Q, J] distances of the positives (NaN = no positive)
  const int64_t* pos_idx;  // [Q, J] gather kernel
  float* pos_dist;         // [Q, J] gather kernel output
  unsigned long long* counts;   // [Q, J, 2]
  const int* only;         // optional [Q]: count only the flagged queries (fallback of cfl_rank_counts_packed)
  int parts;
  int64_t tiles;
};

// Four queries per pass through a catalog row: every shared-memory load of the row serves four distance
// evaluations and the query operands arrive as one 128-bit broadcast load (the first version issued two
// 32-bit loads per FMA and sat at 66 % of the LSU pipe).  Per query the operation sequence is exactly that of
// pcd_direct / monomer_direct above, so the values -- and the counts -- are bit-identical to theirs.
template <int K>
__device__ __forceinline__ void pcd_direct4(const float* __restrict__ er, const float4* __restrict__ q4, int d,
                                            float (&dist)[4]) {
  float dk[4][K];
  float mn[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int j = 0; j < d; ++j) {
      const float e = er[j];
      const float4 p = q4[(k * d + j) * 2];
      float df;
      df = e - p.x; acc[0] = fmaf(df, df, acc[0]);
      df = e - p.y; acc[1] = fmaf(df, df, acc[1]);
      df = e - p.z; acc[2] = fmaf(df, df, acc[2]);
      df = e - p.w; acc[3] = fmaf(df, df, acc[3]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { dk[c][k] = acc[c]; mn[c] = fminf(mn[c], acc[c]); }
  }
  if (K == 1) {
#pragma unroll
    for (int c = 0; c < 4; ++c) dist[c] = dk[c][0];
    return;
  }
  float inv[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) { dk[c][k] = expf(mn[c] - dk[c][k]); sum += dk[c][k]; }
    inv[c] = 1.0f / sum;
    dist[c] = 0.0f;
  }
#pragma unroll 2
  for (int j = 0; j < d; ++j) {
    const float e = er[j];
    float m[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4 p = q4[(k * d + j) * 2];
      m[0] = fmaf(dk[0][k] * inv[0], p.x, m[0]);
      m[1] = fmaf(dk[1][k] * inv[1], p.y, m[1]);
      m[2] = fmaf(dk[2][k] * inv[2], p.z, m[2]);
      m[3] = fmaf(dk[3][k] * inv[3], p.w, m[3]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { const float r = e - m[c]; dist[c] = fmaf(r, r, dist[c]); }
  }
}

template <int K>
__device__ __forceinline__ void monomer_direct4(const float* __restrict__ er, const float4* __restrict__ a4,
                                                const float* __restrict__ w /* [4][K] */, int d, float (&dist)[4]) {
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float dk[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int j = 0; j < d; ++j) {
      const float ne = -er[k * d + j];
      const float4 a = a4[j * 2];
      float df;
      df = a.x + ne; dk[0] = fmaf(df, df, dk[0]);
      df = a.y + ne; dk[1] = fmaf(df, df, dk[1]);
      df = a.z + ne; dk[2] = fmaf(df, df, dk[2]);
      df = a.w + ne; dk[3] = fmaf(df, df, dk[3]);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = fmaf(w[c * K + k], dk[c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) dist[c] = acc[c];
}

// VW = floats per staging load (4 when the row length, the row stride and the base pointer allow 128-bit loads).
template <int K, bool MONO, int VW>
__global__ void __launch_bounds__(RC_THREADS)
rank_count_kernel(RcArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = A.d, J = A.J;
  const int ew = MONO ? K * d : d;            // floats per catalog row
  const int qw = MONO ? d : K * d;            // floats per query
  const int ldt = ew + 1;
  float4* q4 = (float4*)smem_raw;                             // [qw][2]: element i of queries 0-3 | 4-7 (zero padded)
  float* et = (float*)(q4 + 2 * qw);                          // [128][ew+1]
  float* ws = et + 128 * ldt;                                 // [QT][K] (monomer)
  float* thr = ws + RC_QT * K;                                // [QT][J]
  int* cnt = (int*)(thr + RC_QT * J);                         // [4 warps][QT][J][2]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int part = blockIdx.x;
  const int64_t q0 = (int64_t)blockIdx.y * RC_QT;
  const int nq = (int)((A.Q - q0 < RC_QT) ? (A.Q - q0) : RC_QT);
  if (A.only != nullptr) {                                   // restricted recount: leave unless a query here is flagged
    int flagged = 0;
    for (int ql = 0; ql < nq; ++ql) flagged |= A.only[q0 + ql];
    if (!flagged) return;
  }
  for (int i = tid; i < RC_QT * qw; i += RC_THREADS) {
    const int ql = i / qw, e = i % qw;
    ((float*)q4)[e * RC_QT + ql] = ql < nq ? A.Pq[(q0 + ql) * A.ldq + e] : 0.0f;
  }
  if (MONO) for (int i = tid; i < RC_QT * K; i += RC_THREADS) ws[i] = i < nq * K ? A.Wq[q0 * K + i] : 0.0f;
  for (int i = tid; i < nq * J; i += RC_THREADS) thr[i] = A.thr[q0 * J + i];
  for (int i = tid; i < (RC_THREADS / 32) * RC_QT * J * 2; i += RC_THREADS) cnt[i] = 0;
  int* mycnt = cnt + wid * RC_QT * J * 2;
  const int64_t t0 = A.tiles * part / A.parts;
  const int64_t t1 = A.tiles * (part + 1) / A.parts;
  const int dv = ew / VW;
  const int sr0 = tid / dv, sj0 = tid % dv, sdr = RC_THREADS / dv, sdj = RC_THREADS % dv;

  for (int64_t tile = t0; tile < t1; ++tile) {
    const int64_t r0 = tile * 128;
    __syncthreads();
    for (int i = tid, r = sr0, j = sj0; i < 128 * dv; i += RC_THREADS) {
      const int64_t row = r0 + r;
      const float* src = A.E + row * A.lde + VW * j;
      float* dst = et + r * ldt + VW * j;
      if (VW == 4) {
        const float4 v = row < A.N ? *(const float4*)src : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      } else {
        dst[0] = row < A.N ? *src : 0.0f;
      }
      r += sdr; j += sdj;
      if (j >= dv) { j -= dv; ++r; }
    }
    __syncthreads();
    const bool valid = r0 + tid < A.N;
    const float* er = et + tid * ldt;
    for (int qh = 0; qh * 4 < nq; ++qh) {
      float dist[4];
      if (MONO) monomer_direct4<K>(er, q4 + qh, ws + qh * 4 * K, d, dist);
      else      pcd_direct4<K>(er, q4 + qh, d, dist);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ql = qh * 4 + c;
        if (ql < nq) {                                         // warp-uniform
          for (int j = 0; j < J; ++j) {
            const float t = thr[ql * J + j];
            const unsigned lt = __ballot_sync(0xffffffffu, valid && dist[c] < t);
            const unsigned eq = __ballot_sync(0xffffffffu, valid && dist[c] == t);
            if (lane == 0) {
              mycnt[(ql * J + j) * 2] += __popc(lt);
              mycnt[(ql * J + j) * 2 + 1] += __popc(eq);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nq * J * 2; i += RC_THREADS) {
    unsigned long long s = 0;
    for (int w = 0; w < RC_THREADS / 32; ++w) s += (unsigned long long)cnt[w * RC_QT * J * 2 + i];
    if (A.only != nullptr && !A.only[q0 + i / (J * 2)]) continue;
    if (s) atomicAdd(&A.counts[q0 * J * 2 + i], s);
  }
}

// One thread per (query, positive): the same device functions on rows gathered from global memory.
template <int K, bool MONO>
__global__ void pair_dist_rows_kernel(RcArgs A) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.Q * A.J) return;
  const int64_t q = i / A.J;
  const int64_t row = A.pos_idx[i];
  float dist = __int_as_float(0x7fc00000);                   // NaN: no positive here
  if (row >= 0 && row < A.N) {
    const float* er = A.E + row * A.lde;
    const float* qv = A.Pq + q * A.ldq;
    const int d = A.d;
    if (MONO) {
      const float* wv = A.Wq + q * K;
      dist = monomer_direct<K>([&](int k, int j) { return er[k * d + j]; }, [&](int j) { return qv[j]; },
                               [&](int k) { return wv[k]; }, d);
    } else {
      dist = pcd_direct<K>([&](int j) { return er[j]; }, [&](int k, int j) { return qv[k * d + j]; }, d);
    }
  }
  A.pos_dist[i] = dist;
}

// Counts over a DENSE distance matrix (the tensor-core route: cfl_score_topk's dist_out holds the Gram-form
// distance of every (query, row); thresholds taken from the same matrix make the counts self-consistent).
// grid (chunks, Q): each CTA streams a slice of one query's row with 128-bit loads.
constexpr int DRC_THREADS = 256;
template <int JM>
__global__ void __launch_bounds__(DRC_THREADS)
dense_rank_count_kernel(const float* __restrict__ dense, int64_t N, int64_t ldn, const float* __restrict__ thr, int J,
                        unsigned long long* __restrict__ counts) {
  __shared__ float st[JM];
  __shared__ unsigned int sc[JM * 2];
  const int64_t q = blockIdx.y;
  const int t = threadIdx.x;
  if (t < JM) st[t] = t < J ? thr[q * J + t] : __int_as_float(0x7fc00000);   // NaN pads: count nothing
  if (t < 2 * JM) sc[t] = 0u;
  __syncthreads();
  const float* row = dense + q * ldn;
  const int64_t per = (N + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per;
  const int64_t hi = lo + per < N ? lo + per : N;
  unsigned int lt[JM], eq[JM];
  float tj[JM];
#pragma unroll
  for (int j = 0; j < JM; ++j) { lt[j] = 0u; eq[j] = 0u; tj[j] = st[j]; }
  for (int64_t i = lo + t; i < hi; i += DRC_THREADS) {
    const float v = row[i];
#pragma unroll
    for (int j = 0; j < JM; ++j) { lt[j] += v < tj[j]; eq[j] += v == tj[j]; }
  }
#pragma unroll
  for (int j = 0; j < JM; ++j) {
    unsigned int a = lt[j], b = eq[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((t & 31) == 0) { if (a) atomicAdd(&sc[2 * j], a); if (b) atomicAdd(&sc[2 * j + 1], b); }
  }
  __syncthreads();
  if (t < 2 * J && sc[t]) atomicAdd(&counts[q * J * 2 + t], (unsigned long long)sc[t]);
}

static size_t rc_smem_bytes(int K, int d, int J, bool mono) {
  const int ew = mono ? K * d : d, qw = mono ? d : K * d;
  return (size_t)RC_QT * qw * 4 + (size_t)128 * (ew + 1) * 4 + (size_t)RC_QT * K * 4 + (size_t)RC_QT * J * 4 +
         (size_t)(RC_THREADS / 32) * RC_QT * J * 2 * 4;
}

template <int K, bool MONO>
static int rc_launch(const RcArgs& a, bool gather, cudaStream_t cs) {
  if (gather) {
    const int64_t n = a.Q * a.J;
    pair_dist_rows_kernel<K, MONO><<<(unsigned)((n + 127) / 128), 128, 0, cs>>>(a);
  } else {
    const size_t smem = rc_smem_bytes(K, a.d, a.J, MONO);
    const int ew = MONO ? K * a.d : a.d;
    const bool v4 = ew % 4 == 0 && a.lde % 4 == 0 && ((uintptr_t)a.E & 15u) == 0;
    auto kern = v4 ? rank_count_kernel<K, MONO, 4> : rank_count_kernel<K, MONO, 1>;
    CFL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one wave: as many catalog parts as keep every CTA resident at once (only resident warps hide the
    // shared-memory latency; a partial second wave idles most SMs)
    int resident = 0;
    CFL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, RC_THREADS, smem));
    if (resident < 1) resident = 1;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int64_t nqt = (a.Q + RC_QT - 1) / RC_QT;
    int64_t parts = ((int64_t)resident * sms) / nqt;
    if (parts > a.tiles) parts = a.tiles;
    if (parts < 1) parts = 1;
    RcArgs b = a;
    b.parts = (int)parts;
    dim3 grid((unsigned)parts, (unsigned)nqt);
    kern<<<grid, RC_THREADS, smem, cs>>>(b);
  }
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

static int rc_dispatch(const RcArgs& a, bool gather, cudaStream_t cs) {
  const bool mono = a.mode == CFL_MONOMER;
#define CFL_RC_CASE(KK) case KK: return mono ? rc_launch<KK, true>(a, gather, cs) : rc_launch<KK, false>(a, gather, cs);
  switch (a.K) {
    CFL_RC_CASE(1) CFL_RC_CASE(2) CFL_RC_CASE(3) CFL_RC_CASE(4)
    CFL_RC_CASE(5) CFL_RC_CASE(6) CFL_RC_CASE(7) CFL_RC_CASE(8)
  }
#undef CFL_RC_CASE
  return CFL_ERR_UNSUPPORTED;
}

static int rc_check(const char* what, int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* Wq,
                    const float* E, int64_t N, int64_t lde, int J) {
  CFL_REQUIRE(mode == CFL_PCD || mode == CFL_SIAMESE || mode == CFL_MONOMER, CFL_ERR_INVALID, "%s: bad mode %d", what, mode);
  CFL_REQUIRE(mode != CFL_SIAMESE || K == 1, CFL_ERR_INVALID, "%s: siamese needs K=1", what);
  CFL_REQUIRE(K >= 1 && K <= CFL_MAX_K, CFL_ERR_UNSUPPORTED, "%s: K=%d outside [1,%d]", what, K, CFL_MAX_K);
  CFL_REQUIRE(d >= 1 && d <= 128, CFL_ERR_UNSUPPORTED, "%s: d=%d outside [1,128]", what, d);
  CFL_REQUIRE(J >= 1 && J <= RC_MAX_J, CFL_ERR_UNSUPPORTED, "%s: J=%d outside [1,%d]", what, J, RC_MAX_J);
  CFL_REQUIRE(Q >= 0 && N >= 0, CFL_ERR_INVALID, "%s: bad Q/N", what);
  const bool mono = mode == CFL_MONOMER;
  CFL_REQUIRE(ldq >= (mono ? d : (int64_t)K * d) && lde >= (mono ? (int64_t)K * d : d), CFL_ERR_INVALID,
              "%s: leading dimension too small", what);
  CFL_REQUIRE(Q == 0 || (Pq && (!mono || Wq)), CFL_ERR_INVALID, "%s: NULL query argument", what);
  CFL_REQUIRE(Q == 0 || N == 0 || E, CFL_ERR_INVALID, "%s: NULL catalog", what);
  // the tile of a monomer catalog (K*d floats per row) must fit the shared memory of one CTA
  CFL_REQUIRE(rc_smem_bytes(K, d, J, mono) <= 200 * 1024, CFL_ERR_UNSUPPORTED,
              "%s: K*d = %d too large for the monomer rank-count tile", what, K * d);
  return CFL_OK;
}

int rank_counts_only_launch(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* E, int64_t N,
                            int64_t lde, const float* pos_dist, int J, unsigned long long* counts, const int* only,
                            cudaStream_t st) {
  RcArgs a = {};
  a.mode = mode; a.K = K; a.d = d; a.J = J; a.Q = Q; a.N = N; a.ldq = ldq; a.lde = lde;
  a.Pq = Pq; a.E = E; a.thr = pos_dist; a.counts = counts; a.only = only;
  a.tiles = (N + 127) / 128;
  a.parts = 1;
  return rc_dispatch(a, false, st);
}

}  // namespace cfl

using namespace cfl;

extern "C" {

int cfl_pair_dist_rows(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* Wq,
                       const float* E, int64_t N, int64_t lde, const int64_t* pos_idx, int J,
                       float* pos_dist, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  st = rc_check("pair_dist_rows", mode, Pq, Q, K, d, ldq, Wq, E, N, lde, J);
  if (st != CFL_OK) return st;
  if (Q == 0) return CFL_OK;
  CFL_REQUIRE(pos_idx && pos_dist, CFL_ERR_INVALID, "pair_dist_rows: NULL argument");
  RcArgs a = {};
  a.mode = mode; a.K = K; a.d = d; a.J = J; a.Q = Q; a.N = N; a.ldq = ldq; a.lde = lde;
  a.Pq = Pq; a.Wq = Wq; a.E = E; a.pos_idx = pos_idx; a.pos_dist = pos_dist;
  return rc_dispatch(a, true, (cudaStream_t)stream);
}

int cfl_rank_counts(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* Wq,
                    const float* E, int64_t N, int64_t lde, const float* pos_dist, int J,
                    int64_t* counts, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  st = rc_check("rank_counts", mode, Pq, Q, K, d, ldq, Wq, E, N, lde, J);
  if (st != CFL_OK) return st;
  if (Q == 0) return CFL_OK;
  CFL_REQUIRE(pos_dist && counts, CFL_ERR_INVALID, "rank_counts: NULL argument");
  cudaStream_t cs = (cudaStream_t)stream;
  CFL_CUDA(cudaMemsetAsync(counts, 0, (size_t)Q * J * 2 * sizeof(int64_t), cs));
  if (N == 0) return CFL_OK;
  RcArgs a = {};
  a.mode = mode; a.K = K; a.d = d; a.J = J; a.Q = Q; a.N = N; a.ldq = ldq; a.lde = lde;
  a.Pq = Pq; a.Wq = Wq; a.E = E; a.thr = pos_dist; a.counts = (unsigned long long*)counts;
  a.tiles = (N + 127) / 128;
  a.parts = 1;                                   // set per kernel instance in rc_launch (one wave)
  return rc_dispatch(a, false, cs);
}

int cfl_dense_rank_counts(const float* dense, int64_t Q, int64_t N, int64_t ldn, const float* pos_dist, int J,
                          int64_t* counts, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(J >= 1 && J <= RC_MAX_J, CFL_ERR_UNSUPPORTED, "dense_rank_counts: J=%d outside [1,%d]", J, RC_MAX_J);
  CFL_REQUIRE(Q >= 0 && N >= 0 && ldn >= N && Q <= 65535, CFL_ERR_INVALID, "dense_rank_counts: bad Q/N/ldn");
  if (Q == 0) return CFL_OK;
  CFL_REQUIRE(pos_dist && counts && (N == 0 || dense), CFL_ERR_INVALID, "dense_rank_counts: NULL argument");
  cudaStream_t cs = (cudaStream_t)stream;
  CFL_CUDA(cudaMemsetAsync(counts, 0, (size_t)Q * J * 2 * sizeof(int64_t), cs));
  if (N == 0) return CFL_OK;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int64_t chunks = (8 * (int64_t)sms + Q - 1) / Q;                 // ~8 CTAs per SM in flight
  const int64_t max_chunks = (N + 4 * DRC_THREADS - 1) / (4 * DRC_THREADS);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)Q);
  unsigned long long* c64 = (unsigned long long*)counts;
  if (J <= 8) dense_rank_count_kernel<8><<<grid, DRC_THREADS, 0, cs>>>(dense, N, ldn, pos_dist, J, c64);
  else if (J <= 16) dense_rank_count_kernel<16><<<grid, DRC_THREADS, 0, cs>>>(dense, N, ldn, pos_dist, J, c64);
  else dense_rank_count_kernel<32><<<grid, DRC_THREADS, 0, cs>>>(dense, N, ldn, pos_dist, J, c64);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

}  // extern "C"
