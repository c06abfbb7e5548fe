// fsg_dev.cuh -- device helpers shared by the fused small-graph kernels (fsg.cu forward, fsg_bwd.cu backward):
// operand layouts of the tensor-core node transforms, the 3xTF32 issue loop, and the deterministic in-kernel
// all-reduce of per-CTA fp64 vectors.
#pragma once
#include "fsg.cuh"
#include "umma.cuh"

namespace cal {
namespace {

constexpr int FT = 256;                               // threads per CTA
constexpr int FH = 128;                               // hidden size this path is built for
constexpr uint32_t kALbo = 2048, kASbo = 128;         // weight operand: chunk c of row m at c * 2048 + m * 16
constexpr uint32_t kBLbo = 144, kBSbo = 32 * 144;     // node operand: chunk c of row i at (i / 8) * 4608 + c * 144 + (i % 8) * 16
constexpr int kBPart = (kFsgRows / 8) * (int)kBSbo;   // 23040 bytes
constexpr int kTmemCols = 256;                        // main accumulators at columns 0 / 64, correction terms at 128 / 192

// per-category cycle counters of CTA 0 / thread 0 (-DCAL_PHASE_TIMING builds): status[48 + category]
#ifdef CAL_PHASE_TIMING
#define FSG_TDECL long long ft_last = clock64(); long long ft_acc[16] = {0};
#define FSG_T(cat) do { const long long t_ = clock64(); ft_acc[cat] += t_ - ft_last; ft_last = t_; } while (0)
#define FSG_TDUMP(c, base) do { if (blockIdx.x == 0 && threadIdx.x == 0) for (int q_ = 0; q_ < 16 && (base) + q_ < 128; ++q_) (c).status[(base) + q_] = (int)ft_acc[q_]; } while (0)
#define FSG_TVAL(q) ((int)ft_acc[q])
#else
#define FSG_TDECL
#define FSG_T(cat)
#define FSG_TDUMP(c, base)
#define FSG_TVAL(q) 0
#endif

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- the in-kernel all-reduce of per-CTA fp64 vectors: exact fixed-point accumulation with integer atomics ----
// Integer addition is associative, so the totals do not depend on the arrival order (bit-identical runs) although
// every CTA adds straight into shared accumulators -- no group leaders, no second counter phase.  A value v is scaled
// to |v| * 2^40 (truncated; |v| < 2^55, resolution 2^-40 -- far below the fp32 results derived from the sums) and split
// into two limbs of 48 bits, each added with the sign of v into its own 64-bit word (no carries between the words:
// up to 2^15 addends).  kFxCopies accumulator copies (CTA b adds into copy b % kFxCopies) cut the same-address
// serialisation of the L2 atomic unit; the reader adds the copies (integers: still exact) and recombines
// hi * 2^48 + lo.  Latency: the atomics of a CTA are fire-and-forget (RED), then ONE fence + ONE counter atomic; a
// waiting CTA polls that counter and reads 2 * kFxCopies words per value in one round trip.
constexpr int kFxCopies = 8;
constexpr int kFxWords = 2 * kFsgVec;                                 // 64-bit words per copy per phase
constexpr int kFsgCtl = 2 * kFsgPhases * kFsgCntStride;               // control words behind the site counters: epoch of kernel 0 / 1
__device__ __forceinline__ void fsg_fx_add(long long* acc2, double v) {
  // limbs of the MAGNITUDE (every step below is exact in fp64: the operands are aligned), added with the value's sign
  const double a = fabs(v) * 1099511627776.0;                         // 2^40 (exact scaling)
  const double hi = floor(a * 3.552713678800501e-15);                 // 2^-48
  const double lo = floor(a - hi * 281474976710656.0);                // in [0, 2^48): truncates below 2^-40
  long long l0 = (long long)lo, l1 = (long long)hi;
  if (v < 0.0) {
    l0 = -l0; l1 = -l1;
  }
  atomicAdd(reinterpret_cast<unsigned long long*>(acc2), (unsigned long long)l0);
  atomicAdd(reinterpret_cast<unsigned long long*>(acc2 + 1), (unsigned long long)l1);
}
// Two accumulator SETS alternate between the launches of a kernel (its epoch word, bumped by the last CTA to arrive at
// the launch's final all-reduce site): a launch adds into set epoch & 1 and, right after its dependency wait, every CTA
// clears its slice of the OTHER set -- the one the previous launch dirtied and the next launch will use.  Nothing is
// left to do at teardown (a single CTA re-arming 256 KB of accumulators there cost 6 us per kernel on the critical path).
__device__ __forceinline__ unsigned int* fsg_cnt(const FsgWs& w, int set, int phase) {
  return w.cnt + (size_t)(set * kFsgPhases + phase) * kFsgCntStride;
}
__device__ __forceinline__ long long* fsg_acc(const FsgWs& w, int set, int phase) {
  return w.acc + ((size_t)set * kFsgPhases + phase) * kFxCopies * kFxWords;
}
// kern: 0 = forward (sites [0, L]), 1 = backward (sites [12, 12 + L]).  Call after the dependency wait, all CTAs.
__device__ __forceinline__ int fsg_epoch_begin(const FsgWs& w, int kern, int p0, int p1) {
  const int set = (int)(ld_acquire_gpu(&w.cnt[kFsgCtl + kern]) & 1u), other = set ^ 1;
  longlong2* z = reinterpret_cast<longlong2*>(fsg_acc(w, other, p0));
  const int items = (p1 - p0) * kFxCopies * kFxWords / 2;
  for (int i = blockIdx.x * FT + threadIdx.x; i < items; i += gridDim.x * FT) z[i] = make_longlong2(0, 0);
  if (blockIdx.x == 0 && (int)threadIdx.x < p1 - p0) *fsg_cnt(w, other, p0 + threadIdx.x) = 0u;
  return set;
}
// `last_site`: this is the launch's final site -- the CTA that completes it moves the kernel's epoch on
__device__ __forceinline__ void fsg_publish_fx(const FsgWs& w, int set, int phase, int G, const double* sPart, int n,
                                               int kern = -1) {
  const int t = threadIdx.x;
  long long* acc = fsg_acc(w, set, phase) + (size_t)(blockIdx.x % kFxCopies) * kFxWords;
  for (int i = t; i < n; i += FT) fsg_fx_add(acc + 2 * i, sPart[i]);
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned int old = atomicAdd(fsg_cnt(w, set, phase), 1u);
    if (kern >= 0 && old == (unsigned int)G - 1u) atomicAdd(&w.cnt[kFsgCtl + kern], 1u);
  }
}
// A live block whose graph does not fit the path's limits (fsg.cuh) arrives at every all-reduce site [first, last]
// of its launch without adding anything, so that the other blocks' waits complete.  `last` is the launch's final
// site (fsg_publish_fx: moves the epoch on).  One thread.  (Inline on purpose: a call in the kernel body puts the
// whole kernel under the function-call ABI -- 204 instead of 182 registers and an 8 % slower forward, measured.)
__device__ __forceinline__ void fsg_unfit_arrive(const FsgWs& w, int set, int G, int first, int last, int kern) {
  for (int p = first; p <= last; ++p) {
    const unsigned int old = atomicAdd(fsg_cnt(w, set, p), 1u);
    if (p == last && old == (unsigned int)G - 1u) atomicAdd(&w.cnt[kFsgCtl + kern], 1u);
  }
}

__device__ __forceinline__ void fsg_wait_total_fx(const FsgWs& w, int set, int phase, int G, int n, double* sTot) {
  const int t = threadIdx.x;
  if (t == 0) {
    while (ld_acquire_gpu(fsg_cnt(w, set, phase)) < (unsigned int)G) {
    }
    __threadfence();
  }
  __syncthreads();
  const longlong2* acc = reinterpret_cast<const longlong2*>(fsg_acc(w, set, phase));
  for (int i = t; i < n; i += FT) {
    longlong2 v[kFxCopies];
#pragma unroll
    for (int cp = 0; cp < kFxCopies; ++cp) v[cp] = __ldcg(acc + (size_t)cp * (kFxWords / 2) + i);
    long long lo = 0, hi = 0;
#pragma unroll
    for (int cp = 0; cp < kFxCopies; ++cp) {
      lo += v[cp].x;
      hi += v[cp].y;
    }
    sTot[i] = ((double)hi * 281474976710656.0 + (double)lo) * 9.094947017729282e-13;   // 2^48, 2^-40
  }
  __syncthreads();
}

__device__ __forceinline__ uint32_t b_off(int i, int kc) { return (uint32_t)(i >> 3) * kBSbo + (uint32_t)kc * kBLbo + (uint32_t)(i & 7) * 16u; }

// hi / lo split of one 16-byte chunk into the node operand
__device__ __forceinline__ void put_b(unsigned char* b_hi, unsigned char* b_lo, int i, int kc, float4 v) {
  float h0, h1, h2, h3, l0, l1, l2, l3;
  umma::split_tf32(v.x, h0, l0);
  umma::split_tf32(v.y, h1, l1);
  umma::split_tf32(v.z, h2, l2);
  umma::split_tf32(v.w, h3, l3);
  const uint32_t off = b_off(i, kc);
  *reinterpret_cast<float4*>(b_hi + off) = make_float4(h0, h1, h2, h3);
  *reinterpret_cast<float4*>(b_lo + off) = make_float4(l0, l1, l2, l3);
}

// D_main (+)= A_hi B_hi ; D_corr (+)= A_lo B_hi + A_hi B_lo over `ksteps` steps of 8 k.  One thread.
// (Two accumulators: every MMA re-rounds its accumulator, so the small correction terms are kept out of the
// main sum's rounding chain: 16 instead of 48 roundings of the main accumulator at K = 128.)
__device__ __forceinline__ void issue_3xtf32(const unsigned char* a_hi, const unsigned char* a_lo, const unsigned char* b_hi,
                                             const unsigned char* b_lo, uint32_t d_main, uint32_t d_corr, int ksteps, int npad) {
  const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, npad);
  // the descriptors of step s differ from those of step 0 only in the start-address field (bits [0, 14) = byte address
  // >> 4; shared memory is < 256 KB, so the sum never carries out of the field): one 64-bit add per operand per step
  // instead of rebuilding four descriptors -- the issue loop, not the tensor pipe, bounds these small products
  uint64_t ah = umma::smem_desc(umma::smem_addr(a_hi), kALbo, kASbo), al = umma::smem_desc(umma::smem_addr(a_lo), kALbo, kASbo);
  uint64_t bh = umma::smem_desc(umma::smem_addr(b_hi), kBLbo, kBSbo), bl = umma::smem_desc(umma::smem_addr(b_lo), kBLbo, kBSbo);
  constexpr uint64_t da = (2u * kALbo) >> 4, db = (2u * kBLbo) >> 4;
#pragma unroll 4
  for (int s = 0; s < ksteps; ++s) {
    umma::mma_tf32(d_corr, al, bh, idesc, s > 0);
    umma::mma_tf32(d_corr, ah, bl, idesc, 1u);
    umma::mma_tf32(d_main, ah, bh, idesc, s > 0);
    ah += da; al += da; bh += db; bl += db;
  }
}


}  // namespace
}  // namespace cal
