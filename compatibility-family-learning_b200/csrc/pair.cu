// Stage 2 on paired rows: distance (pcd soft-min / monomer / siamese), Thresholder score and
// the per-batch loss statistics, forward and backward.  Replaces the ~14 TF ops of
// DistBase.build_dist (cfl/models/base.py:107-146) + Thresholder (cfl/models/blocks.py:21-22)
// + the reductions of _build_dist_losses (cfl/models/cfl.py:879-902,932-937) with ONE pass
// over (a, P): HBM-bound, 4d(K+1)+4 algorithmic bytes per pair, direct-difference form, fp32.
//
// Mapping: a group of G = 2^j lanes owns one pair (G*MAXE >= d); lane g holds elements
// g, g+G, ... of the embedding and of each prototype in registers, so a warp reads whole
// contiguous rows.  Reductions over d are xor-shuffles inside the group.  Batch statistics
// are reduced in a fixed order (thread -> warp -> block -> one final block), in double.
#include "common.cuh"

namespace cfl {

constexpr int PAIR_THREADS = 256;
constexpr int NSTAT = 6;   // softplus sum, correct, dist sum, sqrt sum, hinge sum, dtheta sum

__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int K, int MAXE>
struct PairRegs {
  float av[MAXE];
  float pv[K][MAXE];
  float dk[K];
  float s[K];
  float r[MAXE];
  float dist;
};

// Loads one pair and computes d_k, the softmax weights, the residual r = a - m and dist.
template <int K, int MAXE>
__device__ __forceinline__ void pair_forward(PairRegs<K, MAXE>& R, int mode,
                                             const float* __restrict__ a, int64_t lda,
                                             const float* __restrict__ P, int64_t ldP,
                                             const float* __restrict__ w, int64_t pair, int d,
                                             int G, int g) {
  const float* ar = a + pair * lda;
  const float* pr = P + pair * ldP;
#pragma unroll
  for (int e = 0; e < MAXE; ++e) {
    int j = g + G * e;
    R.av[e] = (j < d) ? __ldg(ar + j) : 0.0f;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float acc = 0.0f;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      int j = g + G * e;
      float p = (j < d) ? __ldg(pr + (int64_t)k * d + j) : 0.0f;
      R.pv[k][e] = p;
      float df = R.av[e] - p;
      acc = fmaf(df, df, acc);
    }
    R.dk[k] = group_sum(acc, G);
  }
  if (mode == CFL_MONOMER) {
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      R.s[k] = __ldg(w + pair * K + k);
      acc = fmaf(R.s[k], R.dk[k], acc);
    }
    R.dist = acc;
    return;
  }
  if (K == 1) {
    R.s[0] = 1.0f;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) R.r[e] = R.av[e] - R.pv[0][e];
    R.dist = R.dk[0];
    return;
  }
  float mn = R.dk[0];
#pragma unroll
  for (int k = 1; k < K; ++k) mn = fminf(mn, R.dk[k]);
  float sum = 0.0f;
#pragma unroll
  for (int k = 0; k < K; ++k) { R.s[k] = expf(mn - R.dk[k]); sum += R.s[k]; }
  float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < K; ++k) R.s[k] *= inv;
  float acc = 0.0f;
#pragma unroll
  for (int e = 0; e < MAXE; ++e) {
    float m = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) m = fmaf(R.s[k], R.pv[k][e], m);
    R.r[e] = R.av[e] - m;
    acc = fmaf(R.r[e], R.r[e], acc);
  }
  R.dist = group_sum(acc, G);
}

__device__ __forceinline__ void block_reduce_stats(double (&acc)[NSTAT], double* __restrict__ part) {
  __shared__ double sm[PAIR_THREADS / 32][NSTAT];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NSTAT; ++i) {
    double v = warp_sum(acc[i]);
    if (lane == 0) sm[wid][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < NSTAT) {
    double v = 0.0;
    for (int wq = 0; wq < PAIR_THREADS / 32; ++wq) v += sm[wq][threadIdx.x];
    part[(int64_t)blockIdx.x * NSTAT + threadIdx.x] = v;
  }
}

template <int K, int MAXE>
__global__ void __launch_bounds__(PAIR_THREADS)
pair_fwd_kernel(int mode, const float* __restrict__ a, int64_t lda, const float* __restrict__ P,
                int64_t ldP, const float* __restrict__ w, int64_t B, int d, int G,
                const float* __restrict__ theta, int label, float margin,
                float* __restrict__ dist, float* __restrict__ score, float* __restrict__ s_out,
                double* __restrict__ part) {
  const int g = threadIdx.x % G;
  const int64_t groups_per_block = PAIR_THREADS / G;
  const int64_t stride = groups_per_block * gridDim.x;
  const float th = theta ? fmaxf(__ldg(theta), 1e-6f) : 0.0f;   // blocks.py:21
  double acc[NSTAT];
#pragma unroll
  for (int i = 0; i < NSTAT; ++i) acc[i] = 0.0;
  PairRegs<K, MAXE> R;
  // warp-uniform trip count: every lane takes part in the shuffles
  int64_t first = (int64_t)blockIdx.x * groups_per_block;
  for (int64_t base = first; base < B; base += stride) {
    int64_t pair = base + threadIdx.x / G;
    bool valid = pair < B;
    int64_t pc = valid ? pair : B - 1;
    pair_forward<K, MAXE>(R, mode, a, lda, P, ldP, w, pc, d, G, g);
    if (valid && g == 0) {
      float sc = th - R.dist;                                       // blocks.py:22
      if (dist) dist[pair] = R.dist;
      if (score) score[pair] = sc;
      if (s_out) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_out[pair * K + k] = R.s[k];
      }
      if (label >= 0) {
        acc[0] += (double)softplusf(label ? -sc : sc);              // cfl.py:879-886
        acc[1] += (label ? (sc > 0.0f) : (sc <= 0.0f)) ? 1.0 : 0.0; // cfl.py:932-934
        acc[2] += (double)R.dist;
        acc[3] += (double)sqrtf(R.dist + 1e-7f);                    // cfl.py:897-900
        acc[4] += (double)fmaxf(0.0f, margin - R.dist);             // cfl.py:919-920
      }
    }
  }
  if (part) block_reduce_stats(acc, part);
}

__global__ void pair_stats_final(const double* __restrict__ part, int nblocks, int64_t B,
                                 double* __restrict__ stats, double* __restrict__ dtheta) {
  int i = threadIdx.x;
  if (i < NSTAT) {
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += part[(int64_t)b * NSTAT + i];
    if (stats && i < 5) stats[i] = v;
    if (dtheta && i == 5) *dtheta = v;
  }
  if (stats && i == 5) { stats[5] = (double)B; stats[6] = 0.0; stats[7] = 0.0; }
}

template <int K, int MAXE>
__global__ void __launch_bounds__(PAIR_THREADS)
pair_bwd_kernel(int mode, const float* __restrict__ a, int64_t lda, const float* __restrict__ P,
                int64_t ldP, const float* __restrict__ w, int64_t B, int d, int G,
                const float* __restrict__ theta, int label, float margin, float c_ce, float c_lin,
                float c_margin, const float* __restrict__ ddist_in, float* __restrict__ da,
                int64_t ldda, float* __restrict__ dP, int64_t lddP, float* __restrict__ dw,
                double* __restrict__ part) {
  const int g = threadIdx.x % G;
  const int64_t groups_per_block = PAIR_THREADS / G;
  const int64_t stride = groups_per_block * gridDim.x;
  const float th = theta ? fmaxf(__ldg(theta), 1e-6f) : 0.0f;
  double acc[NSTAT];
#pragma unroll
  for (int i = 0; i < NSTAT; ++i) acc[i] = 0.0;
  PairRegs<K, MAXE> R;
  int64_t first = (int64_t)blockIdx.x * groups_per_block;
  for (int64_t base = first; base < B; base += stride) {
    int64_t pair = base + threadIdx.x / G;
    bool valid = pair < B;
    int64_t pc = valid ? pair : B - 1;
    pair_forward<K, MAXE>(R, mode, a, lda, P, ldP, w, pc, d, G, g);
    // upstream dL/ddist
    float dd, ce = 0.0f;
    if (ddist_in) {
      dd = __ldg(ddist_in + pc);
    } else if (label == 1) {
      ce = c_ce * sigmoidf(R.dist - th);
      dd = ce + c_lin;
    } else {
      ce = -c_ce * sigmoidf(th - R.dist);
      dd = ce - ((R.dist < margin) ? c_margin : 0.0f);
    }
    if (valid && g == 0) acc[5] -= (double)ce;            // dL/dtheta+ = -sum(CE part)
    float* dar = da ? da + pair * ldda : nullptr;
    float* dpr = dP ? dP + pair * lddP : nullptr;
    if (mode == CFL_MONOMER) {
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        int j = g + G * e;
        float ga = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float df = R.av[e] - R.pv[k][e];
          float t = 2.0f * R.s[k] * df * dd;
          ga += t;
          if (valid && dpr && j < d) dpr[(int64_t)k * d + j] = -t;
        }
        if (valid && dar && j < d) dar[j] = ga;
      }
      if (valid && dw && g == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) dw[pair * K + k] = R.dk[k] * dd;
      }
    } else if (K == 1) {
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        int j = g + G * e;
        float t = 2.0f * R.r[e] * dd;
        if (valid && j < d) {
          if (dar) dar[j] = t;
          if (dpr) dpr[j] = -t;
        }
      }
    } else {
      // u_k = -2 r.p_k ; c_k = -s_k (u_k - sum_l s_l u_l)     
      float c[K];
      float ubar = 0.0f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float t = 0.0f;
#pragma unroll
        for (int e = 0; e < MAXE; ++e) t = fmaf(R.r[e], R.pv[k][e], t);
        c[k] = -2.0f * group_sum(t, G);
        ubar = fmaf(R.s[k], c[k], ubar);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) c[k] = -R.s[k] * (c[k] - ubar);
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        int j = g + G * e;
        float ga = 2.0f * R.r[e];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float df = R.av[e] - R.pv[k][e];
          ga = fmaf(2.0f * c[k], df, ga);
          float gp = (-2.0f * R.s[k] * R.r[e] - 2.0f * c[k] * df) * dd;
          if (valid && dpr && j < d) dpr[(int64_t)k * d + j] = gp;
        }
        if (valid && dar && j < d) dar[j] = ga * dd;
      }
    }
  }
  if (part) block_reduce_stats(acc, part);
}

static int pick_group(int d, int* G, int* maxe) {
  if (d <= 0 || d > CFL_MAX_D) return -1;
  if (d > 128) { *G = 32; *maxe = 8; return 0; }
  *maxe = 4;
  int g = 1;
  while (g * 4 < d) g <<= 1;
  *G = g;
  return 0;
}

static int pair_blocks(int64_t B, int G) {
  int64_t gpb = PAIR_THREADS / G;
  int64_t nb = (B + gpb - 1) / gpb;
  int64_t cap = (int64_t)sm_count() * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int)nb;
}

template <int K>
static int launch_fwd(int maxe, int nb, cudaStream_t st, int mode, const float* a, int64_t lda,
                      const float* P, int64_t ldP, const float* w, int64_t B, int d, int G,
                      const float* theta, int label, float margin, float* dist, float* score,
                      float* s, double* part) {
  if (maxe == 4)
    pair_fwd_kernel<K, 4><<<nb, PAIR_THREADS, 0, st>>>(mode, a, lda, P, ldP, w, B, d, G, theta,
                                                        label, margin, dist, score, s, part);
  else
    pair_fwd_kernel<K, 8><<<nb, PAIR_THREADS, 0, st>>>(mode, a, lda, P, ldP, w, B, d, G, theta,
                                                        label, margin, dist, score, s, part);
  return 0;
}

template <int K>
static int launch_bwd(int maxe, int nb, cudaStream_t st, int mode, const float* a, int64_t lda,
                      const float* P, int64_t ldP, const float* w, int64_t B, int d, int G,
                      const float* theta, int label, float margin, float c_ce, float c_lin,
                      float c_margin, const float* ddist_in, float* da, int64_t ldda, float* dP,
                      int64_t lddP, float* dw, double* part) {
  if (maxe == 4)
    pair_bwd_kernel<K, 4><<<nb, PAIR_THREADS, 0, st>>>(mode, a, lda, P, ldP, w, B, d, G, theta,
                                                        label, margin, c_ce, c_lin, c_margin,
                                                        ddist_in, da, ldda, dP, lddP, dw, part);
  else
    pair_bwd_kernel<K, 8><<<nb, PAIR_THREADS, 0, st>>>(mode, a, lda, P, ldP, w, B, d, G, theta,
                                                        label, margin, c_ce, c_lin, c_margin,
                                                        ddist_in, da, ldda, dP, lddP, dw, part);
  return 0;
}

#define DISPATCH_K(K, CALL)                                                               \
  switch (K) {                                                                            \
    case 1: { constexpr int KK = 1; CALL; } break;                                        \
    case 2: { constexpr int KK = 2; CALL; } break;                                        \
    case 3: { constexpr int KK = 3; CALL; } break;                                        \
    case 4: { constexpr int KK = 4; CALL; } break;                                        \
    case 5: { constexpr int KK = 5; CALL; } break;                                        \
    case 6: { constexpr int KK = 6; CALL; } break;                                        \
    case 7: { constexpr int KK = 7; CALL; } break;                                        \
    case 8: { constexpr int KK = 8; CALL; } break;                                        \
  }

static int check_pair_args(int mode, const float* a, const float* P, const float* w, int64_t B,
                           int K, int d, int64_t lda, int64_t ldP) {
  CFL_REQUIRE(mode >= CFL_PCD && mode <= CFL_SIAMESE, CFL_ERR_INVALID, "pair: bad mode %d", mode);
  CFL_REQUIRE(B >= 0 && ((a && P) || B == 0), CFL_ERR_INVALID, "pair: NULL input");
  CFL_REQUIRE(K >= 1 && K <= CFL_MAX_K, CFL_ERR_UNSUPPORTED, "pair: K=%d outside [1,%d]", K, CFL_MAX_K);
  CFL_REQUIRE(d >= 1 && d <= CFL_MAX_D, CFL_ERR_UNSUPPORTED, "pair: d=%d outside [1,%d]", d, CFL_MAX_D);
  CFL_REQUIRE(lda >= d && ldP >= (int64_t)K * d, CFL_ERR_INVALID, "pair: leading dimension too small");
  CFL_REQUIRE(mode != CFL_MONOMER || w, CFL_ERR_INVALID, "pair: monomer mode needs w");
  CFL_REQUIRE(mode != CFL_SIAMESE || K == 1, CFL_ERR_INVALID, "pair: siamese mode needs K=1");
  return CFL_OK;
}

}  // namespace cfl

using namespace cfl;

extern "C" {

size_t cfl_pair_workspace_bytes(int64_t B) {
  (void)B;
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  return align_up((size_t)sms * 8 * NSTAT * sizeof(double), 256) + 256;
}

int cfl_pair_loss_fwd(int mode, const float* a, int64_t lda, const float* P, int64_t ldP,
                      const float* w, int64_t B, int K, int d, const float* theta, int label,
                      float margin, float* dist, float* score, float* s, double* stats, void* ws,
                      size_t ws_bytes, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  st = check_pair_args(mode, a, P, w, B, K, d, lda, ldP);
  if (st != CFL_OK) return st;
  CFL_REQUIRE(!(score || stats) || theta, CFL_ERR_INVALID, "pair_fwd: score/stats need theta");
  CFL_REQUIRE(label >= -1 && label <= 1, CFL_ERR_INVALID, "pair_fwd: label must be -1, 0 or 1");
  cudaStream_t cs = (cudaStream_t)stream;
  if (B == 0) {   // empty batch: zero statistics, nothing else to do
    if (stats) CFL_CUDA(cudaMemsetAsync(stats, 0, CFL_PAIR_STATS * sizeof(double), cs));
    return CFL_OK;
  }
  int G, maxe;
  CFL_REQUIRE(pick_group(d, &G, &maxe) == 0, CFL_ERR_UNSUPPORTED, "pair_fwd: unsupported d=%d", d);
  int nb = pair_blocks(B, G);
  double* part = nullptr;
  if (stats) {
    CFL_REQUIRE(label == 0 || label == 1, CFL_ERR_INVALID, "pair_fwd: stats need label 0 or 1");
    CFL_REQUIRE(ws && ws_bytes >= cfl_pair_workspace_bytes(B), CFL_ERR_WORKSPACE,
                "pair_fwd: workspace too small");
    part = (double*)ws;
  } else {
    label = -1;
  }
  DISPATCH_K(K, launch_fwd<KK>(maxe, nb, cs, mode, a, lda, P, ldP, w, B, d, G, theta, label,
                               margin, dist, score, s, part));
  CFL_LAUNCH_CHECK();
  if (stats) {
    pair_stats_final<<<1, 32, 0, cs>>>(part, nb, B, stats, nullptr);
    CFL_LAUNCH_CHECK();
  }
  return CFL_OK;
}

int cfl_pair_loss_bwd(int mode, const float* a, int64_t lda, const float* P, int64_t ldP,
                      const float* w, int64_t B, int K, int d, const float* theta, int label,
                      float margin, float c_ce, float c_lin, float c_margin,
                      const float* ddist_in, float* da, int64_t ldda, float* dP, int64_t lddP,
                      float* dw, double* dtheta_sum, void* ws, size_t ws_bytes, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  st = check_pair_args(mode, a, P, w, B, K, d, lda, ldP);
  if (st != CFL_OK) return st;
  CFL_REQUIRE(ddist_in || (theta && (label == 0 || label == 1)), CFL_ERR_INVALID,
              "pair_bwd: need ddist_in or (theta, label in {0,1})");
  CFL_REQUIRE(!da || ldda >= d, CFL_ERR_INVALID, "pair_bwd: ldda too small");
  CFL_REQUIRE(!dP || lddP >= (int64_t)K * d, CFL_ERR_INVALID, "pair_bwd: lddP too small");
  cudaStream_t cs = (cudaStream_t)stream;
  if (B == 0) {
    if (dtheta_sum) CFL_CUDA(cudaMemsetAsync(dtheta_sum, 0, sizeof(double), cs));
    return CFL_OK;
  }
  int G, maxe;
  CFL_REQUIRE(pick_group(d, &G, &maxe) == 0, CFL_ERR_UNSUPPORTED, "pair_bwd: unsupported d=%d", d);
  int nb = pair_blocks(B, G);
  double* part = nullptr;
  if (dtheta_sum) {
    CFL_REQUIRE(ws && ws_bytes >= cfl_pair_workspace_bytes(B), CFL_ERR_WORKSPACE,
                "pair_bwd: workspace too small");
    part = (double*)ws;
  }
  DISPATCH_K(K, launch_bwd<KK>(maxe, nb, cs, mode, a, lda, P, ldP, w, B, d, G, theta, label,
                               margin, c_ce, c_lin, c_margin, ddist_in, da, ldda, dP, lddP, dw,
                               part));
  CFL_LAUNCH_CHECK();
  if (dtheta_sum) {
    pair_stats_final<<<1, 32, 0, cs>>>(part, nb, B, nullptr, dtheta_sum);
    CFL_LAUNCH_CHECK();
  }
  return CFL_OK;
}

}  // extern "C"
