// TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: the library does not link libcuda)
// and the tiled bulk-tensor load they drive.  One instruction moves a whole [rows x cols] box of a strided fp32 matrix
// into shared memory (row-major, rows beyond the matrix and columns beyond its width arrive as zeros).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "umma.cuh"

namespace cfl {

// 2-D fp32 matrix [rows, cols] with leading dimension ld (floats): needs a 16-byte aligned base and ld % 4 == 0,
// box_cols <= 256, box_rows <= 256.  Returns false when the driver refuses (the caller falls back).
inline bool make_tmap_2d_f32(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint64_t ld,
                             uint32_t box_cols, uint32_t box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  if (!fn || ((uintptr_t)base & 15u) != 0 || (ld & 3u) != 0 || box_cols > 256 || box_rows > 256 || cols == 0 || rows == 0)
    return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * sizeof(float)};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// box at (col0, row0) -> dst (shared, 128-byte aligned); completes box bytes on the mbarrier
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* tmap, int col0, int row0, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          umma::smem_u32(dst_smem)),
      "l"(tmap), "r"(col0), "r"(row0), "r"(umma::smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// round-to-nearest (ties away) tf32 split in integer arithmetic: the bits of cvt.rna.tf32.f32 for finite values, 5
// instructions per element instead of 9 (the compiler guards cvt.rna against inf / NaN)
__device__ __forceinline__ void split_tf32_fast(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32x4_fast(const float4& x, float4& hi, float4& lo) {
  split_tf32_fast(x.x, hi.x, lo.x); split_tf32_fast(x.y, hi.y, lo.y);
  split_tf32_fast(x.z, hi.z, lo.z); split_tf32_fast(x.w, hi.w, lo.w);
}
#endif

}  // namespace cfl
