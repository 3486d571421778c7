// sm_100a building blocks: mbarrier, bulk async copy (TMA engine), TMEM allocation, tcgen05.mma
// (kind::tf32, cta_group::1, operands in shared memory, accumulator in TMEM), tcgen05.ld.
//
// 3xTF32 error compensation: an fp32 value x is split as hi = x with the low 13 mantissa bits
// cleared (exactly a tf32) and lo = trunc_tf32(x - hi); A.B ~= Ahi.Bhi + Ahi.Blo + Alo.Bhi with
// fp32 accumulation in TMEM, relative error ~2^-21 per product (fp32-equivalent for this path).
//
// Shared-memory operand layout (K-major, SWIZZLE_NONE canonical layout): one MMA K-step is
// 8 tf32 = 32 bytes per row = two 16-byte chunks.  For an operand with R rows the element
// (row r, chunk c) of a K-step lives at byte  c*(R*16) + r*16 : 8 consecutive rows of one
// chunk form one 128-byte core matrix, so
//     SBO (next 8-row group)   = 128 bytes
//     LBO (next K chunk)       = R*16 bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cfl {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier -----------------------------------------------------------------------------
// One lane of a converged warp (elect.sync).  Code under `if (elect_one())` is known to the compiler
// to run in a single thread, so the uniform-datapath instructions in it (UTCHMMA, UTCBAR, UBLKCP) are
// issued directly; under `if (lane == 0)` each of them is wrapped in an ELECT / BRA.U.ANY waterfall
// loop that costs more than the MMA it issues.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)      // suspend-time hint (ns): sleep, don't spin
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 28); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}

// same, barrier given by its shared-memory address (keeps issue loops in uniform registers)
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (ok) return;
  }
  __trap();
}

// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- bulk async copy global -> shared (TMA engine, completes on an mbarrier) -----------------
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------
// one full warp; ncols power of two in [32,512]; base address lands in *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x (N consecutive 32-bit columns): thread t of the warp gets lane (base_lane + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors ----------------------------------------------------------------------------
// Shared-memory matrix descriptor (SM100 format): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), base_offset=0, lbo_mode=0, layout SWIZZLE_NONE=0 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major, N>>3 at [17,23),
// M>>4 at [24,29).
__device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// kind::f16 with fp16 operands and an fp32 accumulator: D=F32 (1<<4), A=B=F16 (0<<7, 0<<10); one instruction
// contracts 16 halfs = 32 bytes per row, i.e. the SAME two 16-byte chunks per K-step as the tf32 layout above,
// so the shared-memory descriptors (LBO = R*16, SBO = 128) are unchanged.  M = 128 needs N % 16 == 0.
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Same, descriptors given as (low word, high word): only the low word (start address >> 4, LBO) changes between the
// MMAs of a kernel, so the issue loop advances one 32-bit value per operand.
__device__ __forceinline__ void mma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the descriptors as (low word, high word): between the MMAs of a kernel only the 14-bit start-address field
// of the low word changes, so an issue loop advances one 32-bit value per operand instead of rebuilding 64-bit
// descriptors (the issuing thread is bound by the latency of its own instruction stream, profiles/r2_02).
__device__ __forceinline__ void mma_tf32_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// words of a SWIZZLE_NONE K-major descriptor: low = start >> 4 | (LBO >> 4) << 16, high = SBO >> 4 | version 1 << 14
__device__ __forceinline__ uint32_t desc_lo_word(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3fffu) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi_word(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14); }

// all previously issued MMAs of this thread arrive on the mbarrier when complete
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mma_commit_addr(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- 3xTF32 split ----------------------------------------------------------------------------
// round-to-nearest tf32 (low 13 mantissa bits zero): |x-hi| <= 2^-11|x|, |x-hi-lo| <= 2^-22|x|
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - hi);
}
__device__ __forceinline__ void split_tf32x4(const float4& x, float4& hi, float4& lo) {
  split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
  split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

// Precomputed descriptor parts for a run of K-steps: only the 14-bit start-address field of an
// smem descriptor changes from K-step to K-step (addresses < 256 KB, so no carry leaves the field).
struct Step3Desc {
  uint64_t a_hi, a_lo, b_hi, b_lo;
  uint32_t idesc;
};
__device__ __forceinline__ Step3Desc make_step3(uint32_t N, uint32_t idesc) {
  Step3Desc d;
  d.a_hi = make_smem_desc(0, 128u * 16u, 128u);
  d.a_lo = make_smem_desc(2u * 128u * 16u, 128u * 16u, 128u);
  d.b_hi = make_smem_desc(0, N * 16u, 128u);
  d.b_lo = make_smem_desc(2u * N * 16u, N * 16u, 128u);
  d.idesc = idesc;
  return d;
}
// a_addr / b_addr: shared addresses of the K-step's A block and B block
__device__ __forceinline__ void mma_step3(const Step3Desc& d, uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr,
                                          bool first) {
  const uint64_t ao = (uint64_t)(a_addr >> 4), bo = (uint64_t)(b_addr >> 4);
  mma_tf32(d_tmem, d.a_lo + ao, d.b_hi + bo, d.idesc, first ? 0u : 1u);
  mma_tf32(d_tmem, d.a_hi + ao, d.b_lo + bo, d.idesc, 1u);
  mma_tf32(d_tmem, d.a_hi + ao, d.b_hi + bo, d.idesc, 1u);
}

// Issues the three MMAs of one K-step (8 tf32): small terms first.
//   a_stage: [hl][chunk][128 rows][16 B]   (hl stride 2*128*16 = 4096 B)
//   b_step : [hl][chunk][N rows][16 B]     (hl stride 2*N*16 B)
__device__ __forceinline__ void mma_step_3xtf32(uint32_t d_tmem, uint32_t a_stage, uint32_t b_step,
                                                uint32_t N, uint32_t idesc, bool first) {
  const uint32_t a_hl = 2u * 128u * 16u, b_hl = 2u * N * 16u;
  uint64_t a_hi = make_smem_desc(a_stage, 128u * 16u, 128u);
  uint64_t a_lo = make_smem_desc(a_stage + a_hl, 128u * 16u, 128u);
  uint64_t b_hi = make_smem_desc(b_step, N * 16u, 128u);
  uint64_t b_lo = make_smem_desc(b_step + b_hl, N * 16u, 128u);
  mma_tf32(d_tmem, a_lo, b_hi, idesc, first ? 0u : 1u);
  mma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
  mma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
}

}  // namespace umma
}  // namespace cfl
