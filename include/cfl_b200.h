/* cfl_b200.h -- C ABI of the B200-native CFL compatibility-scoring hot path.
 *
 * The reference (appier/compatibility-family-learning) has no FFI: the path is reached
 * through Python constructors that emit stock TensorFlow ops.  Each entry point below
 * replaces the TF op sequence of one reference call site (cited per function, paths
 * relative to the reference root); the Python mirror in
 * compatibility-family-learning_b200/cfl/ binds them with ctypes (INTEGRATION.md shows
 * the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in
 *     _host; fp32 row-major; `ld*` are leading dimensions in ELEMENTS;
 *   - nothing is allocated inside: cfl_*_workspace_bytes() tells the caller what to pass;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and
 *     re-entrant per stream; no global mutable state except a per-device immutable
 *     property cache;
 *   - return value: CFL_OK (0) or a negative cfl_status; cfl_last_error() returns a
 *     thread-local message.  There is NO CPU fallback: without an sm_100 device every
 *     compute call returns CFL_ERR_DEVICE.
 */
#ifndef CFL_B200_H_
#define CFL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CFL_OK = 0,
  CFL_ERR_INVALID = -1,     /* bad argument (shape, NULL, alignment)          */
  CFL_ERR_CUDA = -2,        /* a CUDA runtime call / launch failed            */
  CFL_ERR_UNSUPPORTED = -3, /* shape outside the compiled range               */
  CFL_ERR_WORKSPACE = -4,   /* workspace too small                            */
  CFL_ERR_DEVICE = -5       /* no sm_100 device                               */
} cfl_status;

/* dist_type -- cfl/models/base.py:109 (monomer), :119 (siamese), :125 (pcd*) */
typedef enum { CFL_PCD = 0, CFL_MONOMER = 1, CFL_SIAMESE = 2 } cfl_mode;
/* act_type -- cfl/models/cfl.py:579-586 (+ lrelu, cfl/ops.py:10, for hidden layers) */
typedef enum { CFL_ACT_LINEAR = 0, CFL_ACT_TANH = 1, CFL_ACT_SIGMOID = 2, CFL_ACT_RELU = 3,
               CFL_ACT_LRELU = 4 } cfl_act;

#define CFL_MAX_K 8        /* num_components                                   */
#define CFL_MAX_D 256      /* latent_size (paired kernels); all-pairs: <= 128  */
#define CFL_MAX_TOPK 128   /* k of the all-pairs ranking                       */
#define CFL_PAIR_STATS 8   /* doubles written by cfl_pair_loss_fwd             */

const char* cfl_last_error(void);
int cfl_version(void);
/* Fills SM count and compute capability of the current device. */
int cfl_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------
 * Stage 1 -- projection.  Replaces fully_connected_weight_norm (cfl/layers.py:80-94:
 * tf.matmul, g/sqrt(reduce_sum(square(V))), bias_add, activation) and, with
 * weight_norm = 0, tf.contrib.layers.fully_connected (cfl/models/dist.py:45-65).  The
 * input normaliser x/scale (cfl/ops.py:198-199; normalize_v2 with a single norm,
 * cfl/ops.py:107-110) is folded in as in_scale.
 *   y[b,j]  = act( (sum_i in_scale*x[b,i]*V[i,j]) * s_j + bias[j] ),  s_j = g_j/|V_:j| (or 1)
 *   z[b,j]  = sum_i in_scale*x[b,i]*V[i,j]          (optional, saved for the backward)
 *   pre[b,j]= value before the activation          (optional)
 * Arithmetic: 3xTF32 error-compensated tcgen05 MMA, fp32 accumulate (fp32-equivalent).
 * ------------------------------------------------------------------------------------- */
size_t cfl_project_fwd_workspace_bytes(int64_t B, int F, int N);
int cfl_project_fwd(const float* x, int64_t B, int F, int64_t ldx,
                    const float* V, int N, int64_t ldV,
                    const float* g, const float* bias, int weight_norm,
                    float in_scale, int act,
                    float* y, int64_t ldy, float* pre, float* z,
                    void* ws, size_t ws_bytes, void* stream);

/* Backward of the above w.r.t. (V, g, bias) -- the tf.gradients the reference gets
 * implicitly from AdamOptimizer.minimize (cfl/models/cfl.py:1083-1085,
 * cfl/models/dist.py:291-293).  dy is dL/dy (post-activation); y the saved output.
 *   accumulate != 0 adds into dV/dg/dbias (pos and neg batches share one encoder).
 *   reg_c adds the l2_regularizer gradient c*V and c*bias (cfl/models/cfl.py:870-874). */
size_t cfl_project_bwd_workspace_bytes(int64_t B, int F, int N);
int cfl_project_bwd(const float* x, int64_t B, int F, int64_t ldx,
                    const float* V, int N, int64_t ldV,
                    const float* g, const float* bias, int weight_norm, float in_scale, int act,
                    const float* y, int64_t ldy, const float* z, const float* dy, int64_t lddy,
                    float* dV, float* dg, float* dbias, int accumulate, float reg_c,
                    void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stage 2, paired rows -- DistBase.build_dist (cfl/models/base.py:107-146, identical copy
 * cfl/models/dist.py:70-89) + Thresholder (cfl/models/blocks.py:21-22) + the per-batch
 * reductions of _build_dist_losses (cfl/models/cfl.py:879-902,932-937).
 *   pcd:     a = v (target e0) [B,d];  P = source prototypes [B,K,d]
 *   monomer: a = source e0;  P = TARGET prototypes;  w = source gate softmax [B,K]
 *   siamese: a, P = the two e0 rows (K = 1)
 *   label: 1 positive pairs, 0 negative pairs, -1 no loss statistics
 * Outputs (each optional): dist[B], score[B] = max(theta,1e-6) - dist, s[B,K] softmax,
 * stats[8] doubles = { sum softplus(-/+score), #correct, sum dist, sum sqrt(dist+1e-7),
 *                      sum max(0,margin-dist), B, 0, 0 } reduced in a fixed order.
 * Direct-difference form, fp32.
 * ------------------------------------------------------------------------------------- */
size_t cfl_pair_workspace_bytes(int64_t B);
int cfl_pair_loss_fwd(int mode, const float* a, int64_t lda, const float* P, int64_t ldP,
                      const float* w, int64_t B, int K, int d,
                      const float* theta, int label, float margin,
                      float* dist, float* score, float* s, double* stats,
                      void* ws, size_t ws_bytes, void* stream);

/* Backward of the paired distance + loss (its own kernel; recomputes dist/softmax).
 * Upstream: ddist_in[B] if non-NULL, else the fused loss gradient
 *   label 1:  c_ce*sigmoid(dist-theta+) + c_lin        (cfl.py:879-882, 912-929)
 *   label 0: -c_ce*sigmoid(theta+-dist) - c_margin*[dist<margin]
 * dtheta_sum (double, optional) = -sum of the sigmoid-CE part (apply the theta>=1e-6
 * mask on the host, cfl/models/blocks.py:21).  Outputs da[B,d], dP[B,K,d], dw[B,K]. */
int cfl_pair_loss_bwd(int mode, const float* a, int64_t lda, const float* P, int64_t ldP,
                      const float* w, int64_t B, int K, int d,
                      const float* theta, int label, float margin,
                      float c_ce, float c_lin, float c_margin, const float* ddist_in,
                      float* da, int64_t ldda, float* dP, int64_t lddP, float* dw,
                      double* dtheta_sum, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stage 2, all pairs (extension, SURVEY App. A.6: the pair scorer on the QxN cross
 * product; no reference call site).  pcd (K>=1) and siamese (K=1); monomer: cfl_score_topk_monomer.
 *   Pq[Q,K,d] query prototypes, E[N,d] catalog embeddings, mu[d] optional centring vector
 *   (distance is translation invariant; centring keeps the Gram form accurate).
 * Gram form d_k = |e|^2+|p_k|^2-2 p_k.e on tensor cores (3xTF32), soft-min epilogue and a
 * running per-query top-(k+slack) fused behind the MMA; the slack winners are rescored in
 * direct-difference form and the best k returned:
 *   top_val[Q,k] ascending exact distances, top_idx[Q,k] = idx_base + row (int64);
 *   ties -> lower index.  dist_out[Q,N] (optional, parity/debug) = un-rescored Gram values.
 * ------------------------------------------------------------------------------------- */
size_t cfl_score_topk_workspace_bytes(int64_t Q, int K, int d, int64_t N, int k);
int cfl_score_topk(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq,
                   const float* E, int64_t N, int64_t lde, const float* mu,
                   int k, int64_t idx_base,
                   float* top_val, int64_t* top_idx, float* dist_out,
                   void* ws, size_t ws_bytes, void* stream);

/* Catalog image for the tensor-core kernel: the catalog is static across query batches, so its MMA
 * operand form (centred on mu, hi/lo tf32 split, canonical shared-memory layout, |e-mu|^2 per row)
 * is built once.  cfl_catalog_pack_bytes returns 0 when (K, d) has no tcgen05 tiling (use
 * cfl_score_topk then).  `image` must be 1024-byte aligned; mu must be the same vector later passed
 * to cfl_score_topk_packed, which is cfl_score_topk minus the per-call packing (E is still needed
 * for the direct-form rescoring of the winners). */
size_t cfl_catalog_pack_bytes(int64_t N, int K, int d);
int cfl_catalog_pack(const float* E, int64_t N, int K, int d, int64_t lde, const float* mu,
                     void* image, size_t image_bytes, void* stream);
size_t cfl_score_topk_packed_workspace_bytes(int64_t Q, int K, int d, int64_t N, int k);
int cfl_score_topk_packed(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq,
                          const void* image, const float* E, int64_t N, int64_t lde, const float* mu,
                          int k, int64_t idx_base,
                          float* top_val, int64_t* top_idx, float* dist_out,
                          void* ws, size_t ws_bytes, void* stream);

/* Statistics of the LAST cfl_score_topk / cfl_score_topk_packed call that used workspace `ws` with this shape
 * (packed = 1 for the _packed entry point): CFL_SCORE_NSTATS unsigned 64-bit counters copied to device memory
 * `stats_out` on `stream`: [0] keys that survived the full filter pass, summed over queries (each is rescored
 * exactly), [1] queries that used their spill list, [2] queries the probe took out of the lower-bound pass,
 * [3] queries that failed the first verification (second lower-bound round under the safe threshold, or the exact
 * redo), [4] 1 when the lower-bound pass ran, [5] queries redone by the exact kernel under the safe threshold.  All zero
 * for catalogs short enough for the single adaptive pass.  thr_out (optional, device, 3*Q floats): the per-query
 * thresholds of that call -- safe threshold tau, optimistic threshold tau_opt, redo threshold (-inf = not redone).
 * (bench.py prints the counters so that a regression to the slow paths is visible; no reference call site.) */
#define CFL_SCORE_NSTATS 8
int cfl_score_topk_stats(int64_t Q, int K, int d, int64_t N, int k, int packed, const void* ws, size_t ws_bytes,
                         unsigned long long* stats_out, float* thr_out, void* stream);

/* Monomer mode on the cross product (SURVEY App. A.6 applied to DistBase.build_dist, monomer branch,
 * cfl/models/base.py:109-117; gate cfl/models/base.py:94-105).  The roles follow the reference:
 *   Aq[Q,d]  = act(e0) of the SOURCE (query) items,  Wq[Q,K] = their gate softmax (dense, row-major),
 *   Pc[N,K,d] = the K prototypes of the TARGET (catalog) items (prototype k of row c at columns
 *   [k*d,(k+1)*d) of a row of length ldp), so the catalog side carries K*d floats per row;
 *   dist(q,c) = sum_k Wq[q,k] * |Aq[q] - Pc[c,k]|^2   (direct-difference form, fp32: values are final).
 * Outputs as cfl_score_topk: top_val ascending, top_idx = idx_base + row, ties -> lower index;
 * dist_out[Q,N] optional (parity/debug).  CUDA-core kernel (packed FP32x2), no tensor-core image. */
size_t cfl_score_topk_monomer_workspace_bytes(int64_t Q, int K, int d, int64_t N, int k);
int cfl_score_topk_monomer(const float* Aq, int64_t lda, const float* Wq, int64_t Q, int K, int d,
                           const float* Pc, int64_t N, int64_t ldp, int k, int64_t idx_base,
                           float* top_val, int64_t* top_idx, float* dist_out,
                           void* ws, size_t ws_bytes, void* stream);

/* Monomer mode, tensor-core path.  With the augmented vectors v_q = [2 w_qk a_q]_k || [-w_qk]_k and
 * c'_c = [P'_ck]_k || [|P'_ck|^2]_k (both sides centred on mu) the distance is (sum_k w_qk)|a_q|^2 - v_q.c'_c: ONE Gram
 * matrix of inner dimension K(d+1) <= 128.  cfl_monomer_pack builds the catalog's fp16 image once (1024-byte aligned,
 * cfl_monomer_pack_bytes; 0 = no tensor-core path for this shape: use cfl_score_topk_monomer).
 * cfl_score_topk_monomer_packed: thresholds from the exact CUDA-core kernel on a 1/32 sample of the tiles, the full
 * pass as a single-product fp16 tcgen05 Gram filter with a rigorous rounding margin, exact fp32 direct-form rescoring
 * of every survivor (the arithmetic of cfl_score_topk_monomer: identical bits), verification, exact redo of the
 * queries that fail it.  Outputs as cfl_score_topk_monomer.  stats_out: optional CFL_SCORE_NSTATS counters (device)
 * as cfl_score_topk_stats.  Pc / mu must be the arrays the image was built from. */
size_t cfl_monomer_pack_bytes(int64_t N, int K, int d);
int cfl_monomer_pack(const float* Pc, int64_t N, int K, int d, int64_t ldp, const float* mu, void* image,
                     size_t image_bytes, void* stream);
size_t cfl_score_topk_monomer_packed_workspace_bytes(int64_t Q, int K, int d, int64_t N, int k);
int cfl_score_topk_monomer_packed(const float* Aq, int64_t lda, const float* Wq, int64_t Q, int K, int d,
                                  const void* image, const float* Pc, int64_t N, int64_t ldp, const float* mu, int k,
                                  int64_t idx_base, float* top_val, int64_t* top_idx,
                                  unsigned long long* stats_out, void* ws, size_t ws_bytes, void* stream);

/* Merge R sorted lists per query (the per-rank results after ncclAllGather) into one:
 * vals[R,Q,k], idx[R,Q,k] -> top_val[Q,k], top_idx[Q,k]; order (value, index). */
int cfl_topk_merge(const float* vals, const int64_t* idx, int R, int64_t Q, int k,
                   float* top_val, int64_t* top_idx, void* stream);

/* The same merge straight from the exchange buffer of the sharded ranking: every rank packs its [Q,k] lists into ONE
 * record (cfl_topk_pack_records: [idx: Q*k int64][val: Q*k float], cfl_topk_record_bytes(Q,k) bytes, a multiple of 16),
 * one all-gather of the records, cfl_topk_merge_records over the R gathered records -- no repacking copies on either
 * side of the collective (the exchange is latency-bound: host time per step matters). */
size_t cfl_topk_record_bytes(int64_t Q, int k);
int cfl_topk_pack_records(const float* vals, const int64_t* idx, int64_t Q, int k, void* rec, void* stream);
int cfl_topk_merge_records(const void* recs, int R, int64_t Q, int k, float* top_val, int64_t* top_idx, void* stream);

/* Column mean of a catalog E[N,d] (the centring vector mu). */
int cfl_col_mean(const float* E, int64_t N, int d, int64_t lde, float* mu,
                 void* ws, size_t ws_bytes, void* stream);
size_t cfl_col_mean_workspace_bytes(int64_t N, int d);

/* ---------------------------------------------------------------------------------------
 * AUC -- replaces sklearn.metrics.roc_auc_score at cfl/utils.py:267-268 and the accuracy
 * counts at cfl/utils.py:247,264.  Exact integers: out[0] = 2*#{s+>s-} + #{s+==s-},
 * out[1] = n_pos, out[2] = n_neg, out[3] = #{s+>0} + #{s-<=0}.  AUC = out[0]/(2 n+ n-).
 * Radix sort of the negatives + rank/count of every positive.
 * ------------------------------------------------------------------------------------- */
size_t cfl_auc_workspace_bytes(int64_t n_pos, int64_t n_neg);
int cfl_auc(const float* pos_scores, int64_t n_pos, const float* neg_scores, int64_t n_neg,
            int64_t* out4, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Per-query all-candidate AUC -- the rank statistic of roc_auc_score (cfl/utils.py:267-268) for one
 * query against EVERY catalog row, without materialising the Q x N scores.
 * Operands per mode (the pair scorer's roles, cfl/models/base.py:107-146):
 *   pcd / siamese: Pq[Q,K,d] query prototypes (ldq >= K*d), Wq NULL, E[N,d] candidate embeddings;
 *   monomer:       Pq[Q,d] query embeddings (ldq >= d), Wq[Q,K] gate softmax, E[N,K,d] candidate
 *                  prototypes (lde >= K*d); the 128-row tile of K*d floats per row must fit one CTA's shared
 *                  memory: K*d <= ~380, else CFL_ERR_UNSUPPORTED.
 * cfl_pair_dist_rows: pos_dist[q,j] = dist(q, row pos_idx[q,j]) for the J labelled positives of each
 *   query; pos_idx is LOCAL to this shard, entries outside [0,N) (e.g. -1: positive lives on another
 *   shard, or padding) give NaN.
 * cfl_rank_counts: counts[q,j,0] = #{c < N : dist(q,c) <  pos_dist[q,j]},
 *                  counts[q,j,1] = #{c < N : dist(q,c) == pos_dist[q,j]}   (int64; NaN counts nothing).
 * Both evaluate the distance in fp32 direct-difference form through one device function, so a positive
 * compares equal to itself and the counts are exact integers that ADD over catalog shards.  With n
 * candidates in total and the positives' own contributions removed, AUC_q = 2U_q / (2 J (n - J)).
 * ------------------------------------------------------------------------------------- */
#define CFL_MAX_RANK_J 32
int cfl_pair_dist_rows(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* Wq,
                       const float* E, int64_t N, int64_t lde, const int64_t* pos_idx, int J,
                       float* pos_dist, void* stream);
int cfl_rank_counts(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* Wq,
                    const float* E, int64_t N, int64_t lde, const float* pos_dist, int J,
                    int64_t* counts, void* stream);

/* The same counts over a DENSE distance matrix dense[Q, ldn] (first N columns), e.g. the dist_out of
 * cfl_score_topk: the tensor-core route for pcd -- the Gram-form distances of every (query, row) are written
 * once and counted at HBM speed; thresholds read from the same matrix (pos_dist[q,j] = dense[q, pos]) keep the
 * counts self-consistent (they are exact for the Gram-form fp32 values, which carry the 3xTF32 error of
 * cfl_score_topk's un-rescored distances instead of the direct form's). */
int cfl_dense_rank_counts(const float* dense, int64_t Q, int64_t N, int64_t ldn, const float* pos_dist, int J,
                          int64_t* counts, void* stream);

/* The same counts on the tensor cores, taken inside the fused scoring kernel's epilogue (no Q x N matrix anywhere):
 * image = cfl_catalog_pack of the same E / mu.  Every (row, query) Gram-form distance is classified against the
 * thresholds with a rounding band; pairs inside a band (exact ties included) are re-evaluated in the direct form of
 * cfl_pair_dist_rows, so counts == cfl_rank_counts(...) for pos_dist produced by cfl_pair_dist_rows.  pcd / siamese.
 * Replaces roc_auc_score over all candidates (cfl/utils.py:267-268); see csrc/rank_counts_tc.cu.  One call takes up
 * to 4096 scoring CTAs' worth of queries (>= 131072 queries; CFL_ERR_UNSUPPORTED beyond: split the query batch).
 * cfl_rank_counts_packed_stats (synchronises the stream): out[CFL_RANK_NSTATS] = {ambiguous records, largest
 * |Gram - direct| / band in units of 2^-20, records beyond half the band, queries recounted on the CUDA cores}. */
#define CFL_RANK_NSTATS 4
size_t cfl_rank_counts_packed_workspace_bytes(int64_t Q, int K, int d, int64_t N);
int cfl_rank_counts_packed(int mode, const float* Pq, int64_t Q, int K, int d, int64_t ldq, const float* E,
                           const void* image, int64_t N, int64_t lde, const float* mu, const float* pos_dist, int J,
                           int64_t* counts, void* ws, size_t ws_bytes, void* stream);
int cfl_rank_counts_packed_stats(int64_t Q, int K, int d, int64_t N, const void* ws, size_t ws_bytes, int64_t* out,
                                 void* stream);

/* TF-1.x Adam (cfl/models/cfl.py:1083-1085): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
 * p -= lr_t*m/(sqrt(v)+eps).  grad_scale multiplies g first (1/world for DP averaging). */
int cfl_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int step,
                  float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* Same update with the step count t read from DEVICE memory (*step_dev >= 1) when the kernel runs, so
 * a CUDA graph of the whole train step can be replayed while the count advances on the device. */
int cfl_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const int* step_dev,
                      float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Measurement hook: when both events are non-NULL, cfl_score_topk records them (cudaEventRecord
 * on the call's stream) immediately before and after its dominant scoring kernel, so bench.py can
 * time that kernel alone inside the timed region.  Thread-local; pass NULLs to clear. */
int cfl_set_kernel_timer(void* start_event, void* stop_event);

/* Self-test of the tcgen05 3xTF32 GEMM core on one tile: D[128,N] = A[128,Kd] * B[N,Kd]^T.
 * Test-only; used by tests/ to validate descriptors and layouts in isolation. */
int cfl_selftest_umma(const float* A, const float* Bm, float* D, int N, int Kd, void* stream);

/* Self-test of the lower-bound pass's MMA: operands rounded to fp16, one tcgen05.mma.kind::f16 per 16 dimensions, fp32
 * accumulators; D[128,N] = A[128,Kd] * B[N,Kd]^T, Kd <= 128.  Test-only: tests/ measures the rounding the error margin
 * of the bound has to cover (operand rounding + accumulation inside the tensor core). */
int cfl_selftest_umma_f16(const float* A, const float* Bm, float* D, int N, int Kd, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFL_B200_H_ */
