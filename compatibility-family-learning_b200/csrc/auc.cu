// AUC as exact integers (replaces sklearn.metrics.roc_auc_score at cfl/utils.py:267-268 and
// the accuracy counts at cfl/utils.py:247,264).
//
//   twoU = sum over positives p of ( #{neg < p} + #{neg <= p} )
// Negatives are sorted with a hand-written LSD radix sort (4 passes x 8 bits over the
// order-preserving u32 image of the float); every positive then binary-searches the sorted
// negatives (lower/upper bound) and the integer counts are added with 64-bit atomics
// (integer addition: the result does not depend on the order).  HBM-bound: the sort moves
// ~ n_neg * 4 B * (2 reads + 1 write) per pass.
#include "common.cuh"

namespace cfl {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 keys per block
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void keys_from_scores(const float* __restrict__ s, int64_t n, uint32_t* __restrict__ k) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) k[i] = f2ord(s[i]);
}

// per-block digit histogram, stored digit-major: hist[digit * nblocks + block]
__global__ void __launch_bounds__(RS_THREADS)
radix_hist(const uint32_t* __restrict__ keys, int64_t n, int shift, int nblocks,
           uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
  for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
    int64_t g = base + i;
    if (g < n) atomicAdd(&h[(keys[g] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of the digit-major histogram, two levels: block d scans row d (the per-block
// counts of digit d) in place and publishes the digit total; one small block then scans the 256
// totals.  The scatter kernel adds the two.
__global__ void __launch_bounds__(256)
radix_scan_rows(uint32_t* __restrict__ hist, int nblocks, uint32_t* __restrict__ digit_total) {
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t carry;
  uint32_t* row = hist + (int64_t)blockIdx.x * nblocks;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int c0 = 0; c0 < nblocks; c0 += 256) {
    const int i = c0 + threadIdx.x;
    const uint32_t v = (i < nblocks) ? row[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) woff += (w < wid) ? wsum[w] : 0u;
    const uint32_t excl = carry + woff + (x - v);
    if (i < nblocks) row[i] = excl;
    __syncthreads();
    if (threadIdx.x == 255) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) digit_total[blockIdx.x] = carry;
}
__global__ void __launch_bounds__(256)
radix_scan_digits(uint32_t* __restrict__ digit_total) {
  __shared__ uint32_t s[256];
  s[threadIdx.x] = digit_total[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int d = 0; d < 256; ++d) { uint32_t c = s[d]; s[d] = run; run += c; }
  }
  __syncthreads();
  digit_total[threadIdx.x] = s[threadIdx.x];
}

// stable scatter: warp w owns the contiguous slice [w*256, (w+1)*256) of the tile and walks
// it 32 keys at a time, so (warp, step, lane) order == tile order.
__global__ void __launch_bounds__(RS_THREADS)
radix_scatter(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n, int shift,
              int nblocks, const uint32_t* __restrict__ offs, const uint32_t* __restrict__ digit_base) {
  __shared__ uint32_t wcnt[RS_WARPS][256];     // per-warp running digit counts
  __shared__ uint32_t dbase[256];              // global offset of (digit, this block)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
  dbase[threadIdx.x] = digit_base[threadIdx.x] + offs[(int64_t)threadIdx.x * nblocks + blockIdx.x];
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)wid * (RS_TILE / RS_WARPS);
  uint32_t key[RS_ITEMS];
  uint32_t rank[RS_ITEMS];
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    int64_t g = wbase + it * 32 + lane;
    bool valid = g < n;
    key[it] = valid ? in[g] : 0xffffffffu;
    uint32_t dg = (key[it] >> shift) & 255u;
    // lanes with the same digit (invalid lanes form their own group via the 9th bit)
    uint32_t peers = __match_any_sync(0xffffffffu, valid ? dg : 256u + 0u);
    uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t basecnt = valid ? wcnt[wid][dg] : 0u;
    rank[it] = basecnt + before;
    __syncwarp();
    if (valid && before == 0) wcnt[wid][dg] = basecnt + __popc(peers);
    __syncwarp();
  }
  __syncthreads();
  // exclusive scan over warps per digit
  {
    uint32_t run = 0;
    int dg = threadIdx.x;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) { uint32_t c = wcnt[w][dg]; wcnt[w][dg] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < RS_ITEMS; ++it) {
    int64_t g = wbase + it * 32 + lane;
    if (g < n) {
      uint32_t dg = (key[it] >> shift) & 255u;
      out[(int64_t)dbase[dg] + wcnt[wid][dg] + rank[it]] = key[it];
    }
  }
}

__global__ void __launch_bounds__(256)
auc_count_kernel(const float* __restrict__ pos, int64_t n_pos, const uint32_t* __restrict__ neg_sorted,
                 int64_t n_neg, unsigned long long* __restrict__ out) {
  unsigned long long two_u = 0, cpos = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pos;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s = pos[i];
    uint32_t kq = f2ord(s);
    int64_t lo = 0, hi = n_neg;                  // lower bound: #neg < s
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (neg_sorted[m] < kq) lo = m + 1; else hi = m; }
    int64_t lb = lo;
    hi = n_neg;                                  // upper bound: #neg <= s
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (neg_sorted[m] <= kq) lo = m + 1; else hi = m; }
    two_u += (unsigned long long)(lb + lo);
    cpos += (s > 0.0f) ? 1ull : 0ull;            // cfl/utils.py:247
  }
  // warp then block reduction of integers (exact), one atomic per block
  for (int o = 16; o > 0; o >>= 1) {
    two_u += __shfl_xor_sync(0xffffffffu, two_u, o);
    cpos += __shfl_xor_sync(0xffffffffu, cpos, o);
  }
  __shared__ unsigned long long su[8], sc[8];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { su[wid] = two_u; sc[wid] = cpos; }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0, b = 0;
    for (int w = 0; w < 8; ++w) { a += su[w]; b += sc[w]; }
    atomicAdd(&out[0], a);
    atomicAdd(&out[3], b);
  }
}

__global__ void auc_finish_kernel(const uint32_t* __restrict__ neg_sorted, int64_t n_pos,
                                  int64_t n_neg, unsigned long long* __restrict__ out) {
  // #{neg <= 0}: upper bound of key(+0.0) in the sorted negatives (cfl/utils.py:264)
  uint32_t kz = f2ord(0.0f);
  int64_t lo = 0, hi = n_neg;
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (neg_sorted[m] <= kz) lo = m + 1; else hi = m; }
  out[1] = (unsigned long long)n_pos;
  out[2] = (unsigned long long)n_neg;
  out[3] += (unsigned long long)lo;
}

}  // namespace cfl

using namespace cfl;

extern "C" {

size_t cfl_auc_workspace_bytes(int64_t n_pos, int64_t n_neg) {
  (void)n_pos;
  int64_t nb = (n_neg + RS_TILE - 1) / RS_TILE;
  if (nb < 1) nb = 1;
  return align_up((size_t)(n_neg > 0 ? n_neg : 1) * 4, 256) * 2 + align_up((size_t)nb * 256 * 4, 256) + 4096;
}

int cfl_auc(const float* pos, int64_t n_pos, const float* neg, int64_t n_neg, int64_t* out4,
            void* ws, size_t ws_bytes, void* stream) {
  int st = device_check();
  if (st != CFL_OK) return st;
  CFL_REQUIRE(out4 && n_pos >= 0 && n_neg >= 0, CFL_ERR_INVALID, "auc: bad arguments");
  CFL_REQUIRE((n_pos == 0 || pos) && (n_neg == 0 || neg), CFL_ERR_INVALID, "auc: NULL scores");
  CFL_REQUIRE(ws && ws_bytes >= cfl_auc_workspace_bytes(n_pos, n_neg), CFL_ERR_WORKSPACE,
              "auc: workspace too small");
  cudaStream_t cs = (cudaStream_t)stream;
  CFL_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(int64_t), cs));
  Workspace W(ws, ws_bytes);
  int64_t nalloc = n_neg > 0 ? n_neg : 1;
  uint32_t* ka = W.take<uint32_t>(nalloc);
  uint32_t* kb = W.take<uint32_t>(nalloc);
  int nb = (int)((n_neg + RS_TILE - 1) / RS_TILE);
  uint32_t* hist = W.take<uint32_t>((size_t)(nb > 0 ? nb : 1) * 256);
  uint32_t* dtot = W.take<uint32_t>(256);
  if (n_neg > 0) {
    keys_from_scores<<<(unsigned)((n_neg + 255) / 256), 256, 0, cs>>>(neg, n_neg, ka);
    CFL_LAUNCH_CHECK();
    for (int pass = 0; pass < 4; ++pass) {
      int shift = pass * 8;
      radix_hist<<<nb, RS_THREADS, 0, cs>>>(ka, n_neg, shift, nb, hist);
      CFL_LAUNCH_CHECK();
      radix_scan_rows<<<256, 256, 0, cs>>>(hist, nb, dtot);
      CFL_LAUNCH_CHECK();
      radix_scan_digits<<<1, 256, 0, cs>>>(dtot);
      CFL_LAUNCH_CHECK();
      radix_scatter<<<nb, RS_THREADS, 0, cs>>>(ka, kb, n_neg, shift, nb, hist, dtot);
      CFL_LAUNCH_CHECK();
      uint32_t* t = ka; ka = kb; kb = t;
    }
  }
  if (n_pos > 0) {
    int blocks = (int)((n_pos + 255) / 256);
    int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    auc_count_kernel<<<blocks, 256, 0, cs>>>(pos, n_pos, ka, n_neg, (unsigned long long*)out4);
    CFL_LAUNCH_CHECK();
  }
  auc_finish_kernel<<<1, 1, 0, cs>>>(ka, n_pos, n_neg, (unsigned long long*)out4);
  CFL_LAUNCH_CHECK();
  return CFL_OK;
}

}  // extern "C"
