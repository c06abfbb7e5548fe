// umma.cuh -- 5th-generation tensor-core building blocks (tcgen05.mma, TMEM, mbarrier), sm_100a only.
//
// Everything here is inline PTX: shared-memory matrix descriptors for the canonical K-major
// no-swizzle operand layout, the instruction descriptor of kind::tf32 / kind::f16, TMEM allocation,
// the single-thread MMA issue, completion through an mbarrier, and the TMEM -> register loads of the
// epilogue.  SASS: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit).
//
// Operand layout ("canonical K-major, SWIZZLE_NONE"): an operand is an [R x K] matrix (R = M rows of A or
// N rows of B, K contiguous within a 16-byte chunk).  Its unit is the CORE MATRIX: 8 rows x 16 bytes,
// stored as 128 contiguous bytes (row r of the core matrix at +16 r).  Core matrices that are
// neighbours along K are `lbo` bytes apart, neighbours along R are `sbo` bytes apart:
//     byte offset of element (r, k) = (r / 8) * sbo + (k / EPC) * lbo + (r % 8) * 16 + (k % EPC) * ES
// with ES = element size, EPC = 16 / ES elements per chunk (4 for tf32, 8 for bf16).
// One MMA consumes 32 bytes of K per row (K = 8 tf32 / 16 bf16 elements) = two chunks.
//
// fp32 accuracy on the tensor cores: 3xTF32.  x = hi + lo with hi = x rounded to tf32 (10 explicit
// mantissa bits) and lo = x - hi (exact in fp32; the hardware truncates it to tf32, relative error
// 2^-11 of a term that is itself <= 2^-11 |x|).  a * b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi, all three
// accumulated in fp32 in TMEM: the dropped term and the truncations are O(2^-22) relative per
// product, the same order as fp32 rounding of the products themselves.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace cal {
namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- descriptors ----
// 64-bit shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"):
// [0,14) start address >> 4 | [16,30) leading-dimension byte offset >> 4 | [32,46) stride-dimension byte
// offset >> 4 | [46,48) version = 1 | [61,64) swizzle mode (0 = none).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// 32-bit instruction descriptor, dense, fp32 accumulate, both operands K-major:
// [4,6) D format (1 = f32) | [7,10) A format | [10,13) B format (kind::f16: 0 = f16, 1 = bf16; kind::tf32: 2 = tf32)
// | [15] A major, [16] B major (0 = K) | [17,23) N >> 3 | [24,29) M >> 4.
constexpr uint32_t kFmtBF16 = 1, kFmtTF32 = 2;
__host__ __device__ constexpr uint32_t instr_desc(uint32_t fmt, int M, int N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation (one warp, all 32 lanes) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_addr(slot_in_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy writes (st.shared) -> visible to the async proxy the tensor core reads through
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP), completion counted in bytes on `bar`;
// bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_addr(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

// ---- MMA issue (ONE thread) ----
// D[tmem] (+)= A[smem] * B[smem]^T : A is [M x K] K-major, B is [N x K] K-major, D is M lanes x N fp32 columns.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_addr(bar)) : "memory");
}

// ---- TMEM -> registers.  Warp w of the CTA may touch lanes 32 (w % 4) .. +31 only; thread = lane = D row. ----
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, int lane, int col) { return base + ((uint32_t)lane << 16) + (uint32_t)col; }
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- operand preparation ----
// x = hi + lo, hi = x rounded (to nearest, ties away) to tf32; both returned as fp32 bit patterns the
// tensor core reads (it ignores the 13 low mantissa bits).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  const uint32_t u = __float_as_uint(x);
  hi = __uint_as_float((u + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
// byte offset of the 16-byte chunk holding elements (r, kc * EPC ..) of a canonical K-major operand
__device__ __forceinline__ uint32_t canon_off(int r, int kc, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)kc * lbo + (uint32_t)(r & 7) * 16u;
}

}  // namespace umma
}  // namespace cal
