// Running per-query top-k machinery shared by the all-pairs kernels.
//
// A (distance, candidate) pair is one u64 key = (orderable(dist) << 32) | local_index, so the
// ranking order "ascending distance, ties -> lower index" is plain u64 '<'.
//
// Each CTA owns, for every query of its query tile, an append buffer of TOPK_CAP keys in the
// caller's workspace (L2 resident) plus a count and a threshold in shared memory.  Candidates
// are visited in increasing index order, so a candidate may be dropped iff dist >= thr where
// thr = the kk-th best value at the last compaction (strictly-less test keeps the tie rule
// exact).  A tile adds at most TOPK_TILE keys per query; whenever a count exceeds
// TOPK_CAP - TOPK_TILE - ... at a tile boundary, one warp sorts the buffer (bitonic network in
// a 4 KB shared scratch), keeps the best kk and tightens thr.
#pragma once
#include "common.cuh"

namespace cfl {

constexpr int TOPK_CAP = 512;     // keys per (part, query) buffer that the adaptive mode may hold
constexpr int TOPK_STRIDE = 2048; // keys reserved per (part, query) buffer (filter / lower-bound passes use all)
constexpr int TOPK_FILTER_TRIGGER = TOPK_STRIDE - 256;   // filter mode: slow path above this
constexpr int LB_SPILL = 16384;   // per-QUERY spill list of the lower-bound pass (keys beyond a full part buffer)
constexpr int TOPK_TILE = 128;    // max appends per query between two compaction points
constexpr int TOPK_TRIGGER = TOPK_CAP - TOPK_TILE;   // compact when cnt > 384 at a tile end

typedef unsigned long long tkey_t;

// Sorts 512 keys in shared memory ascending; one warp.
__device__ __forceinline__ void warp_sort512(tkey_t* s, int lane) {
  for (int size = 2; size <= TOPK_CAP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll 4
      for (int t = lane; t < TOPK_CAP / 2; t += 32) {
        int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        int j = i | stride;
        bool asc = (i & size) == 0;
        tkey_t a = s[i], b = s[j];
        if ((a > b) == asc) { s[i] = b; s[j] = a; }
      }
      __syncwarp();
    }
  }
}

// One warp: keep the best kk of the n keys in buf (global), sorted; returns new count and
// writes the tightened threshold.
__device__ __forceinline__ int warp_compact(tkey_t* __restrict__ buf, int n, int kk,
                                            tkey_t* scratch, int lane, float* thr_out) {
  for (int i = lane; i < TOPK_CAP; i += 32) scratch[i] = (i < n) ? buf[i] : CFL_KEY_INF;
  __syncwarp();
  warp_sort512(scratch, lane);
  int nk = n < kk ? n : kk;
  for (int i = lane; i < nk; i += 32) buf[i] = scratch[i];
  if (lane == 0 && thr_out)
    *thr_out = (n >= kk) ? ord2f((uint32_t)(scratch[kk - 1] >> 32)) : __int_as_float(0x7f800000);
  __syncwarp();
  return nk;
}

// ---- cooperative variant: NT threads (>= 256) sort the 512 keys together ----------------------
// Comparator t (0..255) of a stage touches elements i(t) and i(t)|stride.  Warp w owns
// comparators 32w..32w+31, i.e. elements [64w, 64w+64) whenever stride <= 32, so those stages
// only need __syncwarp; a named barrier (id 1, NT threads) is needed only where the next stage
// reads keys written by another warp.  Threads t >= 256 just take part in the barriers.
template <int NT>
__device__ __forceinline__ void coop_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}
template <int NT>
__device__ __forceinline__ void coop_sort512(tkey_t* s, int t) {
  for (int size = 2; size <= TOPK_CAP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (t < TOPK_CAP / 2) {
        const int i = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
        const int j = i | stride;
        const bool asc = (i & size) == 0;
        const tkey_t a = s[i], b = s[j];
        if ((a > b) == asc) { s[i] = b; s[j] = a; }
      }
      if (stride >= 64 || (stride == 1 && size >= 64)) coop_bar<NT>();   // next stage crosses warps
      else __syncwarp();
    }
  }
}
// NT threads: keep the best kk of buf[0..n) (global), sorted ascending; thread 0 publishes the
// new count and threshold.  Ends with a barrier so `scratch` can be reused immediately.
template <int NT>
__device__ __forceinline__ void coop_compact(tkey_t* __restrict__ buf, int n, int kk, tkey_t* scratch,
                                             int t, int* cnt_out, float* thr_out) {
  for (int i = t; i < TOPK_CAP; i += NT) scratch[i] = (i < n) ? buf[i] : CFL_KEY_INF;
  coop_bar<NT>();
  coop_sort512<NT>(scratch, t);
  const int nk = n < kk ? n : kk;
  if (t < nk) buf[t] = scratch[t];
  if (t == 0) {
    *cnt_out = nk;
    if (thr_out) *thr_out = (n >= kk) ? ord2f((uint32_t)(scratch[kk - 1] >> 32)) : __int_as_float(0x7f800000);
  }
  coop_bar<NT>();
}

// Generic cooperative sort of NKEYS (power of two) keys by NT threads, one block barrier per
// stage (used only on the rare overflow path of the filter mode).
template <int NT, int NKEYS>
__device__ __forceinline__ void coop_sort_generic(tkey_t* s, int t) {
  for (int size = 2; size <= NKEYS; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int c = t; c < NKEYS / 2; c += NT) {
        const int i = ((c & ~(stride - 1)) << 1) | (c & (stride - 1));
        const int j = i | stride;
        const bool asc = (i & size) == 0;
        const tkey_t a = s[i], b = s[j];
        if ((a > b) == asc) { s[i] = b; s[j] = a; }
      }
      coop_bar<NT>();
    }
  }
}
template <int NT>
__device__ __forceinline__ void coop_compact_big(tkey_t* __restrict__ buf, int n, int kk, tkey_t* scratch,
                                                 int t, int* cnt_out, float* thr_out) {
  for (int i = t; i < TOPK_STRIDE; i += NT) scratch[i] = (i < n) ? buf[i] : CFL_KEY_INF;
  coop_bar<NT>();
  coop_sort_generic<NT, TOPK_STRIDE>(scratch, t);
  const int nk = n < kk ? n : kk;
  if (t < nk) buf[t] = scratch[t];
  if (t == 0) {
    *cnt_out = nk;
    if (thr_out && n >= kk) *thr_out = ord2f((uint32_t)(scratch[kk - 1] >> 32));
  }
  coop_bar<NT>();
}

// ---- block-wide selection of the kk smallest keys over all catalog parts -----------------------
// 256 threads per query.  Keys stream through a 512-slot shared buffer; once kk keys are known,
// later keys must beat the current kk-th best to be admitted, so most of them never reach a sort.
constexpr int MRG_THREADS = 256;

__device__ __forceinline__ void mrg_compact(tkey_t* s, int* s_fill, tkey_t* s_thr, int kk, int t) {
  const int fill = *s_fill;
  for (int i = fill + t; i < TOPK_CAP; i += MRG_THREADS) s[i] = CFL_KEY_INF;
  __syncthreads();
  coop_sort512<MRG_THREADS>(s, t);                 // ends with a block barrier (256 == whole block)
  if (t == 0) {
    const int nf = fill < kk ? fill : kk;
    *s_fill = nf;
    if (nf >= kk) *s_thr = s[kk - 1];                // otherwise the admission bound stays as it was
  }
  __syncthreads();
}

// returns the number of keys kept (<= kk), sorted ascending in s[0..)
__device__ __forceinline__ int block_merge_topkk(const tkey_t* __restrict__ keys, const int* __restrict__ counts,
                                                 int parts, int64_t Q, int64_t q, int kk, tkey_t* s, int* s_fill,
                                                 tkey_t* s_thr, int stride = TOPK_STRIDE) {
  const int t = threadIdx.x;
  if (t == 0) { *s_fill = 0; *s_thr = CFL_KEY_INF; }
  __syncthreads();
  for (int p = 0; p < parts; ++p) {
    int c = counts[(int64_t)p * Q + q];
    if (c > stride) c = stride;                              // an over-full lower-bound buffer (flagged for redo)
    const tkey_t* src = keys + ((int64_t)p * Q + q) * stride;
    for (int base = 0; base < c; base += MRG_THREADS) {
      // every thread must take the same branch: read the fill level between two barriers, before
      // any thread of this round can append
      const int f0 = *s_fill;
      __syncthreads();
      if (f0 + MRG_THREADS > TOPK_CAP) mrg_compact(s, s_fill, s_thr, kk, t);
      const int i = base + t;
      if (i < c) {
        const tkey_t key = src[i];
        if (key < *s_thr) s[atomicAdd(s_fill, 1)] = key;
      }
      __syncthreads();
    }
  }
  mrg_compact(s, s_fill, s_thr, kk, t);
  return *s_fill;
}

}  // namespace cfl
