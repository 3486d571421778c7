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
#include <vector>
#include "cfl_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)
#define CF(x) do { int s_ = (x); if (s_ != 0) { printf("cfl error %d (%s) at line %d\n", s_, cfl_last_error(), __LINE__); exit(3); } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.0f - 1.0f; }
template <class T> static T* dalloc(size_t n) { T* p; CK(cudaMalloc(&p, n * sizeof(T) + 1024)); return p; }
static float* upload(const std::vector<float>& h) { float* p = dalloc<float>(h.size()); CK(cudaMemcpy(p, h.data(), h.size() * 4, cudaMemcpyHostToDevice)); return p; }
static void* aligned(size_t bytes) { char* p; CK(cudaMalloc(&p, bytes + 2048)); return (void*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023); }

int main() {
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
  printf("sanitize_driver done\n");
  return bad ? 1 : 0;
}
