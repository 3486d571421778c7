// Standalone driver of the C ABI for compute-sanitizer (python + torch do not survive under the sanitizer on this image):
// projection -> catalog image -> all-pairs scoring through the sampled two-pass path (exact 3xTF32 kernel, lower-bound
// tensor-core pass, rescoring, second round, redo) -> monomer tensor-core path -> fused per-query rank counts (against the
// CUDA-core route) -> projection forward (TMA-staged) and weight gradient (tensor-map TMA + transposing producers).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include tools/sanitize_driver.cu \
//        -o compatibility-family-learning_b200/build/sanitize_driver -L compatibility-family-learning_b200/cfl/_lib -lcfl_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../cfl/_lib'
// Run:   CFL_SCORE_MIN_TILES=2 compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck ./sanitize_driver
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "cfl_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
#define CF(x) do { int s_ = (x); if (s_ != 0) { printf("cfl error %d (%s) at line %d\n", s_, cfl_last_error(), __LINE__); exit(3); } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.0f - 1.0f; }
template <class T> static T* dalloc(size_t n) { T* p; CK(cudaMalloc(&p, n * sizeof(T) + 1024)); return p; }
static float* upload(const std::vector<float>& h) { float* p = dalloc<float>(h.size()); CK(cudaMemcpy(p, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); return p; }
static void* aligned(size_t bytes) { char* p; CK(cudaMalloc(&p, bytes + 2048)); return (void*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023); }


// ---- remaining entry points (end of round 2): paired loss forward / backward in every mode, Adam (host and device step
// count), the pair AUC (radix sort + rank / count, against a host count), the sharded ranking's exchange kernels
// (cfl_topk_pack_records -> cfl_topk_merge_records against cfl_topk_merge), the unpacked scoring entry points on a
// short catalog (single adaptive pass) and the dense rank counts.  `sanitize_driver extra` runs this section alone.
static int run_extra() {
  int bad = 0;
  {
    struct Shape { int mode, K, d; };
    const Shape shapes[] = {{CFL_PCD, 3, 64}, {CFL_PCD, 1, 20}, {CFL_MONOMER, 4, 20}, {CFL_SIAMESE, 1, 64}, {CFL_PCD, 8, 128}};
    const int64_t B = 1003;                                      // ragged last block
    for (const Shape& sh : shapes) {
      const int K = sh.K, d = sh.d;
      std::vector<float> ha((size_t)B * d), hP((size_t)B * K * d), hw((size_t)B * K), hth(1, 1.5f);
      for (auto& x : ha) x = frand();
      for (auto& x : hP) x = frand();
      for (int64_t b = 0; b < B; ++b) {
        float s = 0; for (int kk = 0; kk < K; ++kk) { hw[b * K + kk] = 0.1f + fabsf(frand()); s += hw[b * K + kk]; }
        for (int kk = 0; kk < K; ++kk) hw[b * K + kk] /= s;
      }
      float *a = upload(ha), *P = upload(hP), *w = upload(hw), *th = upload(hth);
      float *dist = dalloc<float>(B), *score = dalloc<float>(B), *sm = dalloc<float>((size_t)B * K);
      double* stats = dalloc<double>(CFL_PAIR_STATS);
      float *da = dalloc<float>((size_t)B * d), *dP = dalloc<float>((size_t)B * K * d), *dw = dalloc<float>((size_t)B * K);
      double* dth = dalloc<double>(1);
      const size_t pw = cfl_pair_workspace_bytes(B);
      void* pws = aligned(pw);
      const float* wq = sh.mode == CFL_MONOMER ? w : nullptr;
      double hsum = 0; int nan = 0;
      for (int label = 1; label >= 0; --label) {
        CF(cfl_pair_loss_fwd(sh.mode, a, d, P, (int64_t)K * d, wq, B, K, d, th, label, 2.0f, dist, score, sm, stats, pws, pw, nullptr));
        CF(cfl_pair_loss_bwd(sh.mode, a, d, P, (int64_t)K * d, wq, B, K, d, th, label, 2.0f, 1.0f / B, 0.25f / B, 0.5f / B, nullptr,
                             da, d, dP, (int64_t)K * d, sh.mode == CFL_MONOMER ? dw : nullptr, dth, pws, pw, nullptr));
        CK(cudaDeviceSynchronize());
        std::vector<float> hd(B), hda((size_t)B * d);
        CK(cudaMemcpy(hd.data(), dist, B * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hda.data(), da, hda.size() * 4, cudaMemcpyDeviceToHost));
        for (float v : hd) { hsum += v; nan += !(v == v) || v < 0; }
        for (float v : hda) nan += !(v == v);
      }
      // siamese / pcd K = 1: dist = |a - P|^2, checked on the host
      if (K == 1) {
        std::vector<float> hd(B);
        CK(cudaMemcpy(hd.data(), dist, B * 4, cudaMemcpyDeviceToHost));
        for (int64_t b = 0; b < B; ++b) {
          double r = 0; for (int j = 0; j < d; ++j) { double t = (double)ha[b * d + j] - hP[b * d + j]; r += t * t; }
          nan += fabs(r - hd[b]) > 1e-4 * (1 + r);
        }
      }
      printf("pair mode %d K %d d %d: sum dist %.6g, bad values %d\n", sh.mode, K, d, hsum, nan);
      bad += nan;
    }
  }
  {
    const int64_t n = 10007;
    std::vector<float> hp(n), hg(n);
    for (auto& x : hp) x = frand();
    for (auto& x : hg) x = frand();
    float *p = upload(hp), *g = upload(hg), *m = dalloc<float>(n), *v = dalloc<float>(n);
    CK(cudaMemset(m, 0, n * 4)); CK(cudaMemset(v, 0, n * 4));
    int one = 2; int* step_dev; CK(cudaMalloc(&step_dev, 4)); CK(cudaMemcpy(step_dev, &one, 4, cudaMemcpyHostToDevice));
    CF(cfl_adam_step(p, g, m, v, n, 1, 1e-3f, 0.9f, 0.999f, 1e-8f, 1.0f, nullptr));
    CF(cfl_adam_step_dev(p, g, m, v, n, step_dev, 1e-3f, 0.9f, 0.999f, 1e-8f, 0.5f, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> out(n);
    CK(cudaMemcpy(out.data(), p, n * 4, cudaMemcpyDeviceToHost));
    int nan = 0; double mv = 0;
    for (int64_t i = 0; i < n; ++i) { nan += !(out[i] == out[i]); mv = fmax(mv, fabs((double)out[i] - hp[i])); }
    nan += !(mv > 1e-4 && mv < 3e-3);                            // two steps of at most lr each
    printf("adam: largest parameter move %.3g, bad %d\n", mv, nan);
    bad += nan;
  }
  {
    const int64_t np_ = 3001, nn = 50021;                        // ragged radix tiles; scores rounded so that ties occur
    std::vector<float> hpos(np_), hneg(nn);
    for (auto& x : hpos) x = roundf((frand() + 0.3f) * 200.0f) / 200.0f;
    for (auto& x : hneg) x = roundf(frand() * 200.0f) / 200.0f;
    float *pos = upload(hpos), *neg = upload(hneg);
    int64_t* out4 = dalloc<int64_t>(4);
    const size_t aw = cfl_auc_workspace_bytes(np_, nn);
    void* aws = aligned(aw);
    CF(cfl_auc(pos, np_, neg, nn, out4, aws, aw, nullptr));
    CK(cudaDeviceSynchronize());
    long long h4[4];
    CK(cudaMemcpy(h4, out4, 32, cudaMemcpyDeviceToHost));
    std::vector<float> sn(hneg);
    std::sort(sn.begin(), sn.end());
    long long two_u = 0, correct = 0;
    for (float s : hpos) {
      two_u += (std::lower_bound(sn.begin(), sn.end(), s) - sn.begin()) + (std::upper_bound(sn.begin(), sn.end(), s) - sn.begin());
      correct += s > 0.0f;
    }
    for (float s : hneg) correct += s <= 0.0f;
    const int diff = (h4[0] != two_u) + (h4[1] != np_) + (h4[2] != nn) + (h4[3] != correct);
    printf("auc: twoU %lld (host %lld), correct %lld (host %lld), differences %d\n", h4[0], two_u, h4[3], correct, diff);
    bad += diff;
  }
  {
    const int R = 3, k = 20; const int64_t Q = 37;
    std::vector<float> hv((size_t)R * Q * k); std::vector<long long> hi((size_t)R * Q * k);
    for (int r = 0; r < R; ++r)
      for (int64_t q = 0; q < Q; ++q) {
        float run = fabsf(frand());
        for (int i = 0; i < k; ++i) {                            // ascending lists with ties inside and across the ranks
          if (rand() % 3) run += roundf(fabsf(frand()) * 4.0f) / 4.0f;
          hv[((size_t)r * Q + q) * k + i] = roundf(run * 4.0f) / 4.0f;
          hi[((size_t)r * Q + q) * k + i] = (long long)r * 100000 + i * 7 + (long long)q;
        }
      }
    float* v = upload(hv);
    int64_t* ix = dalloc<int64_t>(hi.size()); CK(cudaMemcpy(ix, hi.data(), hi.size() * 8, cudaMemcpyHostToDevice));
    float *mv = dalloc<float>((size_t)Q * k), *rv = dalloc<float>((size_t)Q * k);
    int64_t *mi = dalloc<int64_t>((size_t)Q * k), *ri = dalloc<int64_t>((size_t)Q * k);
    CF(cfl_topk_merge(v, ix, R, Q, k, mv, mi, nullptr));
    const size_t rb = cfl_topk_record_bytes(Q, k);
    char* recs = (char*)aligned(rb * R);
    for (int r = 0; r < R; ++r)
      CF(cfl_topk_pack_records(v + (size_t)r * Q * k, ix + (size_t)r * Q * k, Q, k, recs + (size_t)r * rb, nullptr));
    CF(cfl_topk_merge_records(recs, R, Q, k, rv, ri, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> a((size_t)Q * k), b((size_t)Q * k); std::vector<long long> ai((size_t)Q * k), bi((size_t)Q * k);
    CK(cudaMemcpy(a.data(), mv, a.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), rv, b.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ai.data(), mi, ai.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(bi.data(), ri, bi.size() * 8, cudaMemcpyDeviceToHost));
    int diff = 0;
    for (size_t i = 0; i < a.size(); ++i) diff += a[i] != b[i] || ai[i] != bi[i];
    for (int64_t q = 0; q < Q; ++q)                              // order (value, index)
      for (int i = 1; i < k; ++i)
        diff += a[q * k + i] < a[q * k + i - 1] || (a[q * k + i] == a[q * k + i - 1] && ai[q * k + i] <= ai[q * k + i - 1]);
    printf("exchange: merge over records vs merge over lists differences %d (record %zu bytes)\n", diff, rb);
    bad += diff;
  }
  {
    const int K = 3, d = 32, k = 10, J = 3; const int64_t N = 3000, Q = 33;   // short catalog: the single adaptive pass
    std::vector<float> hE((size_t)N * d), hP((size_t)Q * K * d);
    for (auto& x : hE) x = frand();
    for (auto& x : hP) x = frand();
    float *E = upload(hE), *Pq = upload(hP);
    float *tv = dalloc<float>((size_t)Q * k), *dense = dalloc<float>((size_t)Q * N);
    int64_t* ti = dalloc<int64_t>((size_t)Q * k);
    const size_t wb = cfl_score_topk_workspace_bytes(Q, K, d, N, k);
    void* ws = aligned(wb);
    CF(cfl_score_topk(CFL_PCD, Pq, Q, K, d, (int64_t)K * d, E, N, d, nullptr, k, 1000, tv, ti, dense, ws, wb, nullptr));
    std::vector<float> hpd((size_t)Q * J);
    CK(cudaDeviceSynchronize());
    for (int64_t q = 0; q < Q; ++q)
      for (int j = 0; j < J; ++j) CK(cudaMemcpy(&hpd[q * J + j], dense + q * N + (q * 31 + j * 97) % N, 4, cudaMemcpyDeviceToHost));
    float* pd = upload(hpd);
    int64_t* cnt = dalloc<int64_t>((size_t)Q * J * 2);
    CF(cfl_dense_rank_counts(dense, Q, N, N, pd, J, cnt, nullptr));
    CF(cfl_score_topk(CFL_SIAMESE, Pq, Q, 1, d, (int64_t)K * d, E, N, d, nullptr, k, 0, tv, ti, nullptr, ws, wb, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<long long> hc((size_t)Q * J * 2), hti((size_t)Q * k); std::vector<float> htv((size_t)Q * k);
    CK(cudaMemcpy(hc.data(), cnt, hc.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(htv.data(), tv, htv.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hti.data(), ti, hti.size() * 8, cudaMemcpyDeviceToHost));
    int diff = 0;
    for (int64_t q = 0; q < Q; ++q) {
      for (int j = 0; j < J; ++j) diff += hc[(q * J + j) * 2 + 1] < 1 || hc[(q * J + j) * 2] < 0 || hc[(q * J + j) * 2] >= N;
      double best = 1e30; long long arg = -1;                    // siamese: nearest row on the host
      for (int64_t c = 0; c < N; ++c) {
        double r = 0; for (int j = 0; j < d; ++j) { double t = (double)hP[q * K * d + j] - hE[c * d + j]; r += t * t; }
        if (r < best) { best = r; arg = c; }
      }
      diff += hti[q * k] != arg || fabs(htv[q * k] - best) > 1e-4 * (1 + best);
    }
    // monomer, CUDA-core kernel
    const int Km = 4, dm = 20;
    std::vector<float> hPt((size_t)N * Km * dm), ha((size_t)Q * dm), hw((size_t)Q * Km, 0.25f);
    for (auto& x : hPt) x = frand();
    for (auto& x : ha) x = frand();
    float *Pt = upload(hPt), *a = upload(ha), *w = upload(hw);
    const size_t mwb = cfl_score_topk_monomer_workspace_bytes(Q, Km, dm, N, k);
    void* mws = aligned(mwb);
    CF(cfl_score_topk_monomer(a, dm, w, Q, Km, dm, Pt, N, (int64_t)Km * dm, k, 0, tv, ti, nullptr, mws, mwb, nullptr));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(htv.data(), tv, htv.size() * 4, cudaMemcpyDeviceToHost));
    for (int64_t q = 0; q < Q; ++q)
      for (int i = 1; i < k; ++i) diff += !(htv[q * k + i] >= htv[q * k + i - 1]);
    printf("unpacked scoring (pcd + dense counts, siamese vs host, monomer): problems %d\n", diff);
    bad += diff;
  }
  printf("sanitize_driver extra done\n");
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "extra")) return run_extra();
  srand(633);
  const int K = 3, d = 64, Q = 70, k = 20;
  const int64_t N = 40000;
  std::vector<float> hE(N * d), hP((size_t)Q * K * d);
  for (auto& x : hE) x = frand();
  for (int q = 0; q < Q; ++q)
    for (int j = 0; j < K * d; ++j) hP[(size_t)q * K * d + j] = hE[(size_t)(q * 37 % N) * d + j % d] + 0.3f * frand();
  float *E = upload(hE), *Pq = upload(hP);
  // catalog mean, image, scoring (two-pass path is forced by CFL_SCORE_MIN_TILES=2 in the environment)
  float* mu = dalloc<float>(d);
  void* ws_mean = aligned(cfl_col_mean_workspace_bytes(N, d));
  CF(cfl_col_mean(E, N, d, d, mu, ws_mean, cfl_col_mean_workspace_bytes(N, d), nullptr));
  const size_t ib = cfl_catalog_pack_bytes(N, K, d);
  void* img = aligned(ib);
  CF(cfl_catalog_pack(E, N, K, d, d, mu, img, ib, nullptr));
  const size_t wb = cfl_score_topk_packed_workspace_bytes(Q, K, d, N, k);
  void* ws = aligned(wb);
  float* tv = dalloc<float>((size_t)Q * k); int64_t* ti = dalloc<int64_t>((size_t)Q * k);
  for (int rep = 0; rep < 2; ++rep)
    CF(cfl_score_topk_packed(CFL_PCD, Pq, Q, K, d, (int64_t)K * d, img, E, N, d, mu, k, 0, tv, ti, nullptr, ws, wb, nullptr));
  unsigned long long* st = dalloc<unsigned long long>(CFL_SCORE_NSTATS);
  CF(cfl_score_topk_stats(Q, K, d, N, k, 1, ws, wb, st, nullptr, nullptr));
  CK(cudaDeviceSynchronize());
  std::vector<float> hv((size_t)Q * k); std::vector<long long> hi((size_t)Q * k); unsigned long long hs[CFL_SCORE_NSTATS];
  CK(cudaMemcpy(hv.data(), tv, hv.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hi.data(), ti, hi.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hs, st, sizeof(hs), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int q = 0; q < Q; ++q)
    for (int i = 1; i < k; ++i) bad += !(hv[(size_t)q * k + i] >= hv[(size_t)q * k + i - 1]) || hi[(size_t)q * k + i] < 0 || hi[(size_t)q * k + i] >= N;
  printf("pcd: survivors %llu, redo %llu, lower-bound pass %llu, order violations %d, best dist of query 0 = %g (row %lld)\n",
         hs[0], hs[3], hs[4], bad, hv[0], hi[0]);

  // monomer, tensor-core path
  const int Km = 4, dm = 20;
  std::vector<float> hPt((size_t)N * Km * dm), ha((size_t)Q * dm), hw((size_t)Q * Km);
  for (auto& x : hPt) x = frand();
  for (int q = 0; q < Q; ++q) {
    for (int j = 0; j < dm; ++j) ha[(size_t)q * dm + j] = hPt[(size_t)(q * 53 % N) * Km * dm + j] + 0.2f * frand();
    float s = 0; for (int kk = 0; kk < Km; ++kk) { hw[(size_t)q * Km + kk] = 0.1f + fabsf(frand()); s += hw[(size_t)q * Km + kk]; }
    for (int kk = 0; kk < Km; ++kk) hw[(size_t)q * Km + kk] /= s;
  }
  float *Pt = upload(hPt), *a = upload(ha), *w = upload(hw);
  const size_t mib = cfl_monomer_pack_bytes(N, Km, dm);
  void* mimg = aligned(mib);
  CF(cfl_monomer_pack(Pt, N, Km, dm, (int64_t)Km * dm, nullptr, mimg, mib, nullptr));
  const size_t mwb = cfl_score_topk_monomer_packed_workspace_bytes(Q, Km, dm, N, k);
  void* mws = aligned(mwb);
  CF(cfl_score_topk_monomer_packed(a, dm, w, Q, Km, dm, mimg, Pt, N, (int64_t)Km * dm, nullptr, k, 0, tv, ti, st, mws, mwb, nullptr));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(hs, st, sizeof(hs), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hv.data(), tv, hv.size() * 4, cudaMemcpyDeviceToHost));
  printf("monomer: survivors %llu, redo %llu, best dist of query 0 = %g\n", hs[0], hs[3], hv[0]);
  // ---- round-2 kernels: fused rank counts, TMA-staged projection forward, dV GEMM (tensor-map TMA) ----
  {
    const int J = 9;                                             // two threshold chunks
    std::vector<long long> hpos((size_t)Q * J);
    for (auto& x : hpos) x = rand() % N;
    long long* pos; CK(cudaMalloc(&pos, hpos.size() * 8)); CK(cudaMemcpy(pos, hpos.data(), hpos.size() * 8, cudaMemcpyHostToDevice));
    float* tpos = dalloc<float>((size_t)Q * J);
    CF(cfl_pair_dist_rows(CFL_PCD, Pq, Q, K, d, (int64_t)K * d, nullptr, E, N, d, (const int64_t*)pos, J, tpos, nullptr));
    int64_t *cf = dalloc<int64_t>((size_t)Q * J * 2), *cd = dalloc<int64_t>((size_t)Q * J * 2);
    const size_t rwb = cfl_rank_counts_packed_workspace_bytes(Q, K, d, N);
    void* rws = aligned(rwb);
    CF(cfl_rank_counts_packed(CFL_PCD, Pq, Q, K, d, (int64_t)K * d, E, img, N, d, mu, tpos, J, cf, rws, rwb, nullptr));
    CF(cfl_rank_counts(CFL_PCD, Pq, Q, K, d, (int64_t)K * d, nullptr, E, N, d, tpos, J, cd, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<long long> hcf((size_t)Q * J * 2), hcd((size_t)Q * J * 2);
    CK(cudaMemcpy(hcf.data(), cf, hcf.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hcd.data(), cd, hcd.size() * 8, cudaMemcpyDeviceToHost));
    int diff = 0;
    for (size_t i = 0; i < hcf.size(); ++i) diff += hcf[i] != hcd[i];
    int64_t rst[CFL_RANK_NSTATS];
    CF(cfl_rank_counts_packed_stats(Q, K, d, N, rws, rwb, rst, nullptr));
    printf("rank counts: fused vs direct differences %d, ambiguous records %lld\n", diff, (long long)rst[0]);
    bad += diff;
  }
  {
    const int64_t B = 2500; const int F = 512, No = 80;          // ragged last tile, two batch slabs of the dV kernel
    std::vector<float> hx((size_t)B * F), hV((size_t)F * No), hdy((size_t)B * No), hg(No, 1.0f), hb(No, 0.1f);
    for (auto& x : hx) x = fabsf(frand());
    for (auto& x : hV) x = 0.05f * frand();
    for (auto& x : hdy) x = frand();
    float *x = upload(hx), *V = upload(hV), *dy = upload(hdy), *g = upload(hg), *bb = upload(hb);
    float *y = dalloc<float>((size_t)B * No), *z = dalloc<float>((size_t)B * No);
    float *dV = dalloc<float>((size_t)F * No), *dg = dalloc<float>(No), *db = dalloc<float>(No);
    const size_t fw = cfl_project_fwd_workspace_bytes(B, F, No), bw = cfl_project_bwd_workspace_bytes(B, F, No);
    void *fws = aligned(fw), *bws = aligned(bw);
    CF(cfl_project_fwd(x, B, F, F, V, No, No, g, bb, 1, 0.5f, CFL_ACT_TANH, y, No, nullptr, z, fws, fw, nullptr));
    CF(cfl_project_bwd(x, B, F, F, V, No, No, g, bb, 1, 0.5f, CFL_ACT_TANH, y, No, z, dy, No, dV, dg, db, 0, 0.0f, bws, bw, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> hy((size_t)B * No), hdV((size_t)F * No);
    CK(cudaMemcpy(hy.data(), y, hy.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdV.data(), dV, hdV.size() * 4, cudaMemcpyDeviceToHost));
    double sy = 0, sv = 0; int nan = 0;
    for (float v : hy) { sy += v; nan += !(v == v); }
    for (float v : hdV) { sv += v; nan += !(v == v); }
    printf("projection: sum y %.6g, sum dV %.6g, NaNs %d\n", sy, sv, nan);
    bad += nan;
  }
  bad += run_extra();
  printf("sanitize_driver done\n");
  return bad ? 1 : 0;
}
