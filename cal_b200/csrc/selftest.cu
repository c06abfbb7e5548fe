// selftest.cu -- cal_selftest_umma: the tensor-core building block of umma.cuh exercised in isolation.
//
// D[M x N] = A[M x K] * B[N x K]^T on ONE CTA with tcgen05.mma (accumulator in TMEM), for the operand
// layouts, instruction shapes and precisions the readout / node-transform kernels use.  The tests
// compare it with a float64 product (tests/test_gpu_umma.py); it is also the smallest reproducer when
// a descriptor field is in doubt.
#include "internal.cuh"
#include "umma.cuh"

namespace cal {
namespace {

constexpr int kKB = 32;   // K elements staged per block (3xTF32: A and B, hi and lo, 128 + 256 rows -> 96 KB)

// kind: 0 = 3xTF32, 1 = TF32, 2 = BF16.  variant bit 0: core matrices contiguous along R instead of along K;
// bit 1: instruction N rounded to a multiple of 8 instead of 16; bit 2 (with bit 0 clear): K chunks 144 bytes apart.
__global__ void __launch_bounds__(128) k_selftest_umma(int kind, int M, int N, int K, const float* __restrict__ A,
                                                        const float* __restrict__ B, float* __restrict__ D, int variant) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int MR = 128;                                   // instruction M (rows >= M are zero)
  const int NR = (variant & 2) ? (N + 7) & ~7 : (N + 15) & ~15;   // instruction N (bit 1: multiples of 8)
  const bool bf16 = kind == 2;
  const int ES = bf16 ? 2 : 4, EPC = 16 / ES;           // element size, elements per 16-byte chunk
  const int KC = kKB / EPC;                             // chunks per staged block
  // byte strides of the canonical layout
  uint32_t lboA, sboA, lboB, sboB;
  if ((variant & 1) == 0) {
    lboA = lboB = (variant & 4) ? 144 : 128;            // bit 2: chunk stride padded by 16 bytes (bank-conflict-free
    sboA = sboB = (uint32_t)KC * lboA;                  // stores when a warp's lanes walk along K)
  } else {
    sboA = sboB = 128;
    lboA = (uint32_t)(MR / 8) * 128;
    lboB = (uint32_t)(NR / 8) * 128;
  }
  const uint32_t pad = (variant & 4) ? 9 : 8;           // eighths
  const uint32_t bytesA = (uint32_t)MR * kKB * ES / 8 * pad, bytesB = (uint32_t)NR * kKB * ES / 8 * pad;
  unsigned char* sAh = smem;
  unsigned char* sAl = sAh + bytesA;
  unsigned char* sBh = sAl + bytesA;
  unsigned char* sBl = sBh + bytesB;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (t == 0) {
    umma::mbar_init(&bar, 1);
    umma::mbar_fence_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = umma::instr_desc(bf16 ? umma::kFmtBF16 : umma::kFmtTF32, MR, NR);

  uint32_t phase = 0;
  int issued = 0;
  for (int k0 = 0; k0 < K; k0 += kKB) {
    // ---- stage one K block of both operands (hi / lo) in the canonical layout ----
    for (int which = 0; which < 2; ++which) {
      const int R = which ? NR : MR, Rv = which ? N : M;
      const float* src = which ? B : A;
      unsigned char* dh = which ? sBh : sAh;
      unsigned char* dl = which ? sBl : sAl;
      const uint32_t lbo = which ? lboB : lboA, sbo = which ? sboB : sboA;
      for (int i = t; i < R * KC; i += blockDim.x) {
        const int r = i % R, kc = i / R;
        const uint32_t off = umma::canon_off(r, kc, lbo, sbo);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = k0 + kc * EPC + e;
          v[e] = (e < EPC && r < Rv && k < K) ? src[(size_t)r * K + k] : 0.f;
        }
        if (bf16) {
          __nv_bfloat162 p[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) p[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
          *reinterpret_cast<uint4*>(dh + off) = *reinterpret_cast<uint4*>(p);
        } else {
          float h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (kind == 0) umma::split_tf32(v[e], h[e], l[e]);
            else { h[e] = v[e]; l[e] = 0.f; }
          }
          *reinterpret_cast<float4*>(dh + off) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(dl + off) = make_float4(l[0], l[1], l[2], l[3]);
        }
      }
    }
    umma::fence_async_smem();
    __syncthreads();
    if (t == 0) {
      umma::fence_after_sync();
      const int steps = kKB * ES / 32;                  // MMAs per block: 32 bytes of K each
      for (int s = 0; s < steps; ++s) {
        const uint32_t ka = 2u * s * lboA, kb = 2u * s * lboB;
        const uint64_t ah = umma::smem_desc(umma::smem_addr(sAh) + ka, lboA, sboA);
        const uint64_t al = umma::smem_desc(umma::smem_addr(sAl) + ka, lboA, sboA);
        const uint64_t bh = umma::smem_desc(umma::smem_addr(sBh) + kb, lboB, sboB);
        const uint64_t bl = umma::smem_desc(umma::smem_addr(sBl) + kb, lboB, sboB);
        if (bf16) {
          umma::mma_f16(tmem, ah, bh, idesc, issued++ > 0);
        } else {
          if (kind == 0) {
            umma::mma_tf32(tmem, al, bh, idesc, issued++ > 0);
            umma::mma_tf32(tmem, ah, bl, idesc, issued++ > 0);
          }
          umma::mma_tf32(tmem, ah, bh, idesc, issued++ > 0);
        }
      }
      umma::commit(&bar);
    }
    umma::mbar_wait(&bar, phase);                       // operands may be overwritten, accumulator is current
    phase ^= 1;
    umma::fence_after_sync();
  }
  // ---- epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows of D ----
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < NR; c0 += 8) {
    float v[8];
    umma::ld8(umma::tmem_addr(tmem, warp * 32, c0), v);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (row < M && c0 + e < N) D[(size_t)row * N + c0 + e] = v[e];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace
}  // namespace cal

using namespace cal;

extern "C" int cal_selftest_umma(int kind, int M, int N, int K, const float* A, const float* B, float* D, int variant,
                                 void* stream) {
  if (A == nullptr || B == nullptr || D == nullptr) return CAL_ENULL;
  if (kind < 0 || kind > 2 || M < 1 || M > 128 || N < 1 || N > 256 || K < 1 || K > 4096) return CAL_EINVAL;
  const int NR = (variant & 2) ? (N + 7) & ~7 : (N + 15) & ~15, ES = kind == 2 ? 2 : 4;
  const size_t smem = 2 * (size_t)(128 + NR) * kKB * ES / 8 * ((variant & 4) ? 9 : 8);
  cudaError_t e = cudaFuncSetAttribute(k_selftest_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  k_selftest_umma<<<1, 128, smem, (cudaStream_t)stream>>>(kind, M, N, K, A, B, D, variant);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}
