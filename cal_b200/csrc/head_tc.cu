// head_tc.cu -- the three readout MLPs (model.py:125-164) and their backward on the 5th-generation tensor
// cores: tcgen05.mma with the accumulators in TMEM, one CTA per readout head.
//
//   forward   u = xc_g | xo_g | xc_g[perm] (+,||) xo_g  ->  bn1 -> fc1 -> ReLU -> bn2 -> fc2 -> log_softmax -> loss parts
//   backward  d logits -> fc2^T, bn2', ReLU', d W1 = da1^T y1 (tensor core), d y1 = da1 W1 (tensor core), bn1' -> d u
//
// Every product is evaluated TRANSPOSED: D^T[channel][graph] = W[channel][k] * act[graph][k]^T, i.e. the
// weight matrix is the M operand (128 lanes of TMEM = 128 output channels) and the B graph rows sit on the
// N dimension (TMEM columns).  A thread of the epilogue owns one TMEM lane = one CHANNEL and walks over
// the graphs, so every BatchNorm reduction over the B rows (bn1, bn2 and their backward sums), every bias
// gradient and the fc2 weight gradient is a loop inside ONE thread: no cross-CTA partial sums, no
// cluster / DSMEM exchange, fixed summation order => deterministic.
//
// Precision: 3xTF32 (umma.cuh) -- three kind::tf32 MMAs per product on hi / lo splits, fp32 accumulate:
// the 1e-5 parity budget of the fp32 path; or, for cal_model_desc.readout_bf16 (BASELINE.json configs[4]:
// "bf16 MLP / fp32 aggregate"), one kind::f16 MMA on bf16 operands with fp32 accumulate.
//
// Operand staging: both operands of a product go through a two-slot shared-memory ring of K chunks
// (128 bytes of K per row and slot), written by the CTA's threads in the canonical K-major no-swizzle
// layout with thread = operand ROW (a warp's 32 lanes write 512 contiguous bytes: conflict-free), while
// the MMAs of the previous chunk run; a slot is handed back by tcgen05.commit on its mbarrier.
#include "internal.cuh"
#include "umma.cuh"

namespace cal {
namespace {

constexpr int kT = 256;                       // threads per CTA
constexpr int kPart = 16384;                  // one operand part of a ring slot: 128 rows x 128 bytes of K
constexpr uint32_t kLbo = 2048, kSbo = 128;   // chunk c of row r at c * 2048 + r * 16 (row groups contiguous)
constexpr int kTmemCols = 512;

__device__ __forceinline__ int clampB(const Ctx& c) { return imin(imax(c.dims[2], 0), c.Bm); }

struct Pipe {
  uint32_t uses[2];
  uint32_t n;
};

// One 16-byte chunk of an operand row: 4 (tf32: hi | lo parts) or 8 (bf16) consecutive K elements.
template <bool BF16>
__device__ __forceinline__ void put_chunk(unsigned char* part_hi, unsigned char* part_lo, int r, int c, const float (&v)[8]) {
  const uint32_t off = (uint32_t)c * kLbo + (uint32_t)r * 16u;
  if constexpr (BF16) {
    __nv_bfloat162 p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) p[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(part_hi + off) = *reinterpret_cast<uint4*>(p);
  } else {
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) umma::split_tf32(v[e], h[e], l[e]);
    *reinterpret_cast<float4*>(part_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(part_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// D^T[128 lanes][128 columns at d_tmem] = sum_k A[row][k] * B[col][k], k < K.
// ga(r, k, v) / gb(r, k, v): v[0..3] = elements k .. k+3 of row r of the operand (zero outside the matrix;
// k % 4 == 0).  All kT threads call.  Returns after the LAST chunk has been issued (gemm_wait completes it).
template <bool BF16, class GA, class GB>
__device__ __forceinline__ void gemm_tn(unsigned char* ring, uint64_t* bars, Pipe& ps, uint32_t d_tmem, int K, GA ga, GB gb) {
  constexpr int EPC = BF16 ? 8 : 4;            // elements per 16-byte chunk
  constexpr int KSLOT = 8 * EPC;               // K elements per ring slot
  constexpr int PARTS = BF16 ? 2 : 4;
  const int t = threadIdx.x, r = t & 127, c0 = t >> 7;
  const uint32_t idesc = umma::instr_desc(BF16 ? umma::kFmtBF16 : umma::kFmtTF32, 128, 128);
  uint32_t issued = 0;
  for (int k0 = 0; k0 < K; k0 += KSLOT) {
    const int s = (int)(ps.n & 1u);
    unsigned char* slot = ring + (size_t)s * PARTS * kPart;
    unsigned char* pAh = slot;
    unsigned char* pAl = slot + kPart;                       // (unused for bf16)
    unsigned char* pBh = slot + (BF16 ? 1 : 2) * kPart;
    unsigned char* pBl = slot + 3 * kPart;
    if (ps.uses[s] > 0) umma::mbar_wait(&bars[s], (ps.uses[s] - 1u) & 1u);     // the MMAs that read this slot are done
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + 2 * j;
      const int k = k0 + c * EPC;
      float va[8], vb[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) va[e] = vb[e] = 0.f;
      if (k < K) {
        float q[4];
        ga(r, k, q);
#pragma unroll
        for (int e = 0; e < 4; ++e) va[e] = q[e];
        gb(r, k, q);
#pragma unroll
        for (int e = 0; e < 4; ++e) vb[e] = q[e];
        if constexpr (BF16) {
          if (k + 4 < K) {
            ga(r, k + 4, q);
#pragma unroll
            for (int e = 0; e < 4; ++e) va[4 + e] = q[e];
            gb(r, k + 4, q);
#pragma unroll
            for (int e = 0; e < 4; ++e) vb[4 + e] = q[e];
          }
        }
      }
      put_chunk<BF16>(pAh, pAl, r, c, va);
      put_chunk<BF16>(pBh, pBl, r, c, vb);
    }
    umma::fence_async_smem();
    __syncthreads();
    if (t == 0) {
      umma::fence_after_sync();
      const int ksteps = imin(4, (K - k0 + 2 * EPC - 1) / (2 * EPC));     // MMAs of 32 bytes of K each
      for (int st = 0; st < ksteps; ++st) {
        const uint32_t adv = (uint32_t)st * 2u * kLbo;
        const uint64_t ah = umma::smem_desc(umma::smem_addr(pAh) + adv, kLbo, kSbo);
        const uint64_t bh = umma::smem_desc(umma::smem_addr(pBh) + adv, kLbo, kSbo);
        if constexpr (BF16) {
          umma::mma_f16(d_tmem, ah, bh, idesc, issued++ > 0);
        } else {
          const uint64_t al = umma::smem_desc(umma::smem_addr(pAl) + adv, kLbo, kSbo);
          const uint64_t bl = umma::smem_desc(umma::smem_addr(pBl) + adv, kLbo, kSbo);
          umma::mma_tf32(d_tmem, al, bh, idesc, issued++ > 0);
          umma::mma_tf32(d_tmem, ah, bl, idesc, issued++ > 0);
          umma::mma_tf32(d_tmem, ah, bh, idesc, issued++ > 0);
        }
      }
      umma::commit(&bars[s]);
    }
    ps.uses[s] += 1u;
    ps.n += 1u;
  }
}
// every MMA issued so far has completed: its accumulator may be read, both ring slots are free
__device__ __forceinline__ void gemm_wait(uint64_t* bars, const Pipe& ps) {
  if (ps.n == 0u) return;
  const int s = (int)((ps.n - 1u) & 1u);
  umma::mbar_wait(&bars[s], (ps.uses[s] - 1u) & 1u);
  umma::fence_after_sync();
}

struct HeadIn {                       // the readout input u_h[b][k] (model.py:119-121,152-157)
  const float* gc;
  const float* go;
  const int* perm;
  int h, H, cat;
  __device__ __forceinline__ float at(int b, int k) const {
    if (h == 0) return gc[(size_t)b * H + k];
    if (h == 1) return go[(size_t)b * H + k];
    if (cat) return k < H ? gc[(size_t)perm[b] * H + k] : go[(size_t)b * H + (k - H)];
    return gc[(size_t)perm[b] * H + k] + go[(size_t)b * H + k];
  }
  __device__ __forceinline__ void at4(int b, int k, float (&v)[4]) const {       // k % 4 == 0
    float4 a;
    if (h == 0) a = *reinterpret_cast<const float4*>(gc + (size_t)b * H + k);
    else if (h == 1) a = *reinterpret_cast<const float4*>(go + (size_t)b * H + k);
    else if (cat) a = k < H ? *reinterpret_cast<const float4*>(gc + (size_t)perm[b] * H + k)
                            : *reinterpret_cast<const float4*>(go + (size_t)b * H + (k - H));
    else {
      const float4 x = *reinterpret_cast<const float4*>(gc + (size_t)perm[b] * H + k);
      const float4 y = *reinterpret_cast<const float4*>(go + (size_t)b * H + k);
      a = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};

// shared-memory carve-up (bytes), common to both kernels
struct TcSmem {
  size_t ring, lg, misc, total;
};
__host__ __device__ inline TcSmem tc_smem(int Bm, int C, bool bf16) {
  TcSmem s;
  size_t o = 0;
  s.ring = o;  o += (size_t)2 * (bf16 ? 2 : 4) * kPart;          // 128 KB (tf32) / 64 KB (bf16); >= the 66 KB y2 tile
  if (o < (size_t)128 * 129 * 4) o = (size_t)128 * 129 * 4;
  o = (o + 15) & ~(size_t)15;
  s.lg = o;    o += (size_t)imax(Bm, 1) * C * 4;                  // logits / d logits [B][C]
  o = (o + 15) & ~(size_t)15;
  s.misc = o;  o += 4 * 256 * 8 + 12 * 256 * 4 + (size_t)32 * 128 * 4 + (size_t)imax(Bm, 1) * 4;
  s.total = o;
  return s;
}

// ---------------------------------------------------------------------------------------------
// Forward.  grid = 3 (head c, o, co), 256 threads.
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kT, 1) k_readout_tc_fwd(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3, half = warp >> 2;
  const int h = blockIdx.x;
  const int H = c.H, C = c.C;
  const int K1 = (h == 2 && c.cat) ? 2 * H : H;
  const int bn1 = c.L + 3 + h, bn2 = c.L + 6 + h;
  const TcSmem L = tc_smem(c.Bm, C, BF16);
  unsigned char* ring = smem + L.ring;
  float* sLg = reinterpret_cast<float*>(smem + L.lg);
  double* dscr = reinterpret_cast<double*>(smem + L.misc);          // [4][256]
  float* fscr = reinterpret_cast<float*>(dscr + 4 * 256);           // [12][256]
  float *sc1 = fscr, *sh1 = fscr + 256, *sc2 = fscr + 512, *sh2 = fscr + 768, *b1s = fscr + 1024, *red = fscr + 1280;
  float* sW2 = fscr + 12 * 256;                                      // [C][H] (C <= 32, H <= 128)
  int* sPerm = reinterpret_cast<int*>(sW2 + 32 * 128);

  // ---- before the dependency wait: parameters only ----
  if (warp == 0) umma::tmem_alloc(&tmem_slot, kTmemCols);
  if (t == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::mbar_fence_init();
  }
  float g1 = 1.f, be1 = 0.f, rm1 = 0.f, rv1 = 1.f, g2 = 1.f, be2 = 0.f, rm2 = 0.f, rv2 = 1.f;
  const bool has_run = c.bn_buffers != nullptr && c.bn_rm[bn1] >= 0;
  if (t < K1) {
    g1 = c.params[c.bn_gamma[bn1] + t];
    be1 = c.params[c.bn_beta[bn1] + t];
    if (has_run) {
      rm1 = c.bn_buffers[c.bn_rm[bn1] + t];
      rv1 = c.bn_buffers[c.bn_rv[bn1] + t];
    }
  }
  if (t < H) {
    g2 = c.params[c.bn_gamma[bn2] + t];
    be2 = c.params[c.bn_beta[bn2] + t];
    b1s[t] = c.params[c.po.fc1_b[h] + t];
    if (has_run) {
      rm2 = c.bn_buffers[c.bn_rm[bn2] + t];
      rv2 = c.bn_buffers[c.bn_rv[bn2] + t];
    }
  }
  for (int i = t; i < C * H; i += kT) sW2[i] = c.params[c.po.fc2_w[h] + i];
  const float* W1 = c.params + c.po.fc1_w[h];                         // [H][K1]
  umma::fence_before_sync();
  PT_DECL
  pdl_sync();                                                         // everything below may read the predecessor's output
  PT_MARK();                                                          // 0: dependency wait
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int B = clampB(c);
  for (int i = t; i < B; i += kT) sPerm[i] = c.perm[i];
  __syncthreads();
  HeadIn in = {c.pooled, c.pooled + (size_t)c.Bm * H, sPerm, h, H, c.cat};
  PT_MARK();                                                          // 1: tmem + perm

  // ---- bn1: thread = input channel; statistics over the B graph rows in a fixed order ----
  if (t < K1) {
    float sc, sh;
    if (c.train) {
      double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
      int b = 0;
      for (; b + 8 <= B; b += 8) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = in.at(b + e, t);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          s0 += (double)v[e];     q0 += (double)v[e] * (double)v[e];
          s1 += (double)v[e + 1]; q1 += (double)v[e + 1] * (double)v[e + 1];
        }
      }
      for (; b < B; ++b) {
        const float v = in.at(b, t);
        s0 += (double)v;
        q0 += (double)v * (double)v;
      }
      const double sum = s0 + s1, sq = q0 + q1;
      double mean = B > 0 ? sum / B : 0.0, var = B > 0 ? sq / B - mean * mean : 0.0;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
      sc = g1 * rstd;
      sh = be1 - (float)mean * sc;
      c.bnf(bn1, BN_SCALE)[t] = sc;
      c.bnf(bn1, BN_SHIFT)[t] = sh;
      c.bnf(bn1, BN_MEAN)[t] = (float)mean;
      c.bnf(bn1, BN_RSTD)[t] = rstd;
      if (has_run) {
        const double unb = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
        c.bn_buffers[c.bn_rm[bn1] + t] = (1.f - c.momentum) * rm1 + c.momentum * (float)mean;
        c.bn_buffers[c.bn_rv[bn1] + t] = (1.f - c.momentum) * rv1 + c.momentum * (float)unb;
      }
    } else {
      sc = c.bnf(bn1, BN_SCALE)[t];
      sh = c.bnf(bn1, BN_SHIFT)[t];
    }
    sc1[t] = sc;
    sh1[t] = sh;
  }
  if (c.train && t == 0 && c.nbt != nullptr) {
    c.nbt[bn1] += 1;
    c.nbt[bn2] += 1;
  }
  __syncthreads();
  PT_MARK();                                                          // 2: bn1 statistics

  // ---- fc1 on the tensor cores: a1^T[m][b] = sum_k W1[m][k] * y1[b][k], 128 graph rows per N block ----
  Pipe ps = {{0u, 0u}, 0u};
  const int nblk = (B + 127) / 128;
  for (int nb = 0; nb < nblk; ++nb) {
    const int b0 = nb * 128;
    gemm_tn<BF16>(
        ring, bars, ps, tmem + (uint32_t)(nb * 128), K1,
        [&](int r, int k, float (&v)[4]) {                      // A: W1 rows (output channels)
          if (r < H) {
            const float4 w = *reinterpret_cast<const float4*>(W1 + (size_t)r * K1 + k);
            v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
          } else {
            v[0] = v[1] = v[2] = v[3] = 0.f;
          }
        },
        [&](int r, int k, float (&v)[4]) {                      // B: bn1(u) rows (graphs)
          const int b = b0 + r;
          if (b < B) {
            in.at4(b, k, v);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = fmaf(v[e], sc1[k + e], sh1[k + e]);
          } else {
            v[0] = v[1] = v[2] = v[3] = 0.f;
          }
        });
  }
  PT_MARK();                                                          // 3: fc1 staging + issue
  gemm_wait(bars, ps);
  PT_MARK();                                                          // 4: fc1 MMA tail

  // ---- epilogue 1: thread = (hidden channel m, column half): bias + ReLU, save h1, bn2 statistics ----
  const int m = q * 32 + lane;
  float* H1 = c.H1 + (size_t)h * c.Bm * H;
  double es = 0.0, eq = 0.0;
  const float bias1 = m < H ? b1s[m] : 0.f;
  for (int nb = 0; nb < nblk; ++nb) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = half * 64 + cc * 32;
      float v[32];
      umma::ld32(umma::tmem_addr(tmem, q * 32, nb * 128 + col), v);
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int b = nb * 128 + col + e;
        if (b < B && m < H) {
          const float x = fmaxf(v[e] + bias1, 0.f);
          H1[(size_t)b * H + m] = x;
          es += (double)x;
          eq += (double)x * (double)x;
        }
      }
    }
  }
  if (c.train) {
    dscr[half * 256 + m] = es;
    dscr[512 + half * 256 + m] = eq;
  }
  __syncthreads();
  if (t < H) {
    float sc, sh;
    if (c.train) {
      const double sum = dscr[t] + dscr[256 + t], sq = dscr[512 + t] + dscr[768 + t];
      double mean = B > 0 ? sum / B : 0.0, var = B > 0 ? sq / B - mean * mean : 0.0;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
      sc = g2 * rstd;
      sh = be2 - (float)mean * sc;
      c.bnf(bn2, BN_SCALE)[t] = sc;
      c.bnf(bn2, BN_SHIFT)[t] = sh;
      c.bnf(bn2, BN_MEAN)[t] = (float)mean;
      c.bnf(bn2, BN_RSTD)[t] = rstd;
      if (has_run) {
        const double unb = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
        c.bn_buffers[c.bn_rm[bn2] + t] = (1.f - c.momentum) * rm2 + c.momentum * (float)mean;
        c.bn_buffers[c.bn_rv[bn2] + t] = (1.f - c.momentum) * rv2 + c.momentum * (float)unb;
      }
    } else {
      sc = c.bnf(bn2, BN_SCALE)[t];
      sh = c.bnf(bn2, BN_SHIFT)[t];
    }
    sc2[t] = sc;
    sh2[t] = sh;
  }
  __syncthreads();
  PT_MARK();                                                          // 5: epilogue 1 + bn2

  // ---- epilogue 2: y2 = bn2(h1) as a [m][b] tile in shared memory (the ring is free), then
  //      fc2: logits[b][cls] = sum_m y2[m][b] * W2[cls][m] + b2[cls]  (C <= 32: 128 x C x 128 FMAs) ----
  float* sY2 = reinterpret_cast<float*>(ring);                      // [128][129]
  const float s2c = m < H ? sc2[m] : 0.f, s2h = m < H ? sh2[m] : 0.f;
  for (int nb = 0; nb < nblk; ++nb) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = half * 64 + cc * 32;
      float v[32];
      umma::ld32(umma::tmem_addr(tmem, q * 32, nb * 128 + col), v);
#pragma unroll
      for (int e = 0; e < 32; ++e) sY2[m * 129 + col + e] = m < H ? fmaf(fmaxf(v[e] + bias1, 0.f), s2c, s2h) : 0.f;
    }
    __syncthreads();
    for (int i = t; i < 128 * C; i += kT) {
      const int bl = i & 127, cls = i >> 7;
      const int b = nb * 128 + bl;
      float s = 0.f;
      const float* w = sW2 + cls * H;
#pragma unroll 8
      for (int k = 0; k < H; ++k) s = fmaf(sY2[k * 129 + bl], w[k], s);
      if (b < B) sLg[b * C + cls] = s + c.params[c.po.fc2_b[h] + cls];
    }
    __syncthreads();
  }

  PT_MARK();                                                          // 6: y2 tile + fc2
  // ---- log_softmax, outputs, loss parts (train_causal.py:178-186) ----
  float loss_part = 0.f, correct_part = 0.f;
  for (int b = t; b < B; b += kT) {
    float mx = -INFINITY;
    int am = 0;
    for (int cls = 0; cls < C; ++cls) {
      const float v = sLg[b * C + cls];
      if (v > mx) {
        mx = v;
        am = cls;
      }
    }
    float se = 0.f;
    for (int cls = 0; cls < C; ++cls) se += expf(sLg[b * C + cls] - mx);
    const float lse = logf(se);
    float slp = 0.f, picked = 0.f;
    const long long yb = (c.with_loss && c.y != nullptr) ? c.y[b] : -1;
    for (int cls = 0; cls < C; ++cls) {
      const float lp = sLg[b * C + cls] - mx - lse;
      c.logp[((size_t)h * c.Bm + b) * C + cls] = lp;
      slp += lp;
      if ((long long)cls == yb) picked = lp;
    }
    if (c.with_loss) {
      loss_part += h == 0 ? -logf((float)C) - slp / (float)C : -picked;      // KL(uniform || .) row / NLL row
      correct_part += (long long)am == yb ? 1.f : 0.f;
    }
  }
  if (c.with_loss) {
    // deterministic block sums: xor-shuffle tree inside each warp, then the 8 warp totals in order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
      correct_part += __shfl_xor_sync(0xffffffffu, correct_part, o);
    }
    if (lane == 0) {
      red[warp] = loss_part;
      red[8 + warp] = correct_part;
    }
    __syncthreads();
    if (t == 0) {
      float ls = 0.f, cs = 0.f;
      for (int w = 0; w < 8; ++w) {
        ls += red[w];
        cs += red[8 + w];
      }
      c.loss[1 + h] = B > 0 ? ls / (float)B : 0.f;
      c.loss[4 + h] = cs;
    }
    if (grid_last_block(&c.counters[CNT_HEAD2], 3)) {
      if (t == 0) {
        const volatile float* lv = c.loss;
        c.loss[0] = c.w_c * lv[1] + c.w_o * lv[2] + c.w_co * lv[3];
        c.loss[7] = 0.f;
      }
    }
  }
  PT_MARK();                                                          // 7: log_softmax + loss
  PT_DUMP(c, 64);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// Backward.  grid = 3, 256 threads.  Writes the readout parameter gradients straight into the flat
// gradient buffer and d u (gradient w.r.t. the readout inputs) into CAL_WS_DU.
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kT, 1) k_readout_tc_bwd(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3, half = warp >> 2;
  const int h = blockIdx.x;
  const int H = c.H, C = c.C;
  const int K1 = (h == 2 && c.cat) ? 2 * H : H;
  const int bn1 = c.L + 3 + h, bn2 = c.L + 6 + h;
  const TcSmem L = tc_smem(c.Bm, C, BF16);
  unsigned char* ring = smem + L.ring;
  float* sDl = reinterpret_cast<float*>(smem + L.lg);                 // [B][C]
  double* dscr = reinterpret_cast<double*>(smem + L.misc);
  float* fscr = reinterpret_cast<float*>(dscr + 4 * 256);
  float *sc1 = fscr, *sh1 = fscr + 256, *mean1 = fscr + 512, *rstd1 = fscr + 768;
  float* sW2 = fscr + 12 * 256;
  int* sPerm = reinterpret_cast<int*>(sW2 + 32 * 128);

  if (warp == 0) umma::tmem_alloc(&tmem_slot, kTmemCols);
  if (t == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::mbar_fence_init();
  }
  for (int i = t; i < C * H; i += kT) sW2[i] = c.params[c.po.fc2_w[h] + i];
  const float* W1 = c.params + c.po.fc1_w[h];                         // [H][K1]
  umma::fence_before_sync();
  PT_DECL
  pdl_sync();
  PT_MARK();                                                          // 0: dependency wait
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int B = clampB(c);
  for (int i = t; i < B; i += kT) sPerm[i] = c.perm[i];
  if (t < K1) {
    sc1[t] = c.bnf(bn1, BN_SCALE)[t];
    sh1[t] = c.bnf(bn1, BN_SHIFT)[t];
    mean1[t] = c.bnf(bn1, BN_MEAN)[t];
    rstd1[t] = c.bnf(bn1, BN_RSTD)[t];
  }
  // ---- d logits (log_softmax backward) for all rows ----
  for (int b = t; b < B; b += kT) {
    const long long yb = c.y != nullptr ? c.y[b] : -1;
    float sd = 0.f;
    for (int cls = 0; cls < C; ++cls) {
      float dlp;
      if (c.grad_logp != nullptr) dlp = c.grad_logp[((size_t)h * B + b) * C + cls];
      else if (h == 0) dlp = -c.w_c / ((float)C * (float)B);
      else dlp = (long long)cls == yb ? -(h == 1 ? c.w_o : c.w_co) / (float)B : 0.f;
      sDl[b * C + cls] = dlp;
      sd += dlp;
    }
    for (int cls = 0; cls < C; ++cls) {
      const float lp = c.logp[((size_t)h * c.Bm + b) * C + cls];
      sDl[b * C + cls] -= expf(lp) * sd;
    }
  }
  __syncthreads();
  HeadIn in = {c.pooled, c.pooled + (size_t)c.Bm * H, sPerm, h, H, c.cat};
  PT_MARK();                                                          // 1: d logits

  // ---- fc2 / bn2 backward: thread = (hidden channel m, half of the graph rows), everything local ----
  const int m = q * 32 + lane;
  const bool live = m < H;
  const float* H1 = c.H1 + (size_t)h * c.Bm * H;
  float* DH = c.dh + (size_t)h * c.Bm * H;                            // d a1 [b][m]
  const int Bh = (B + 1) / 2;
  const int bb = half == 0 ? 0 : Bh, be = half == 0 ? Bh : B;
  const float s2c = live ? c.bnf(bn2, BN_SCALE)[m] : 0.f, s2h = live ? c.bnf(bn2, BN_SHIFT)[m] : 0.f;
  const float mean2 = live ? c.bnf(bn2, BN_MEAN)[m] : 0.f, rstd2 = live ? c.bnf(bn2, BN_RSTD)[m] : 0.f;
  double a1s = 0.0, a2s = 0.0;
  if (live) {
    for (int b = bb; b < be; ++b) {
      const float x = H1[(size_t)b * H + m];
      float dy = 0.f;
      for (int cls = 0; cls < C; ++cls) dy = fmaf(sDl[b * C + cls], sW2[cls * H + m], dy);
      a1s += (double)dy;
      a2s += (double)dy * (double)((x - mean2) * rstd2);
    }
  }
  dscr[half * 256 + m] = a1s;
  dscr[512 + half * 256 + m] = a2s;
  // d W2[cls][m] = sum_b dl[b][cls] * y2[b][m]: four classes per pass over this thread's rows
  for (int c0 = 0; c0 < C; c0 += 4) {
    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
    if (live) {
      for (int b = bb; b < be; ++b) {
        const float y2 = fmaf(H1[(size_t)b * H + m], s2c, s2h);
        const float* d = sDl + b * C + c0;
        w0 = fmaf(d[0], y2, w0);
        if (c0 + 1 < C) w1 = fmaf(d[1], y2, w1);
        if (c0 + 2 < C) w2 = fmaf(d[2], y2, w2);
        if (c0 + 3 < C) w3 = fmaf(d[3], y2, w3);
      }
    }
    __syncthreads();                                   // (also orders the dscr writes above before their readers)
    float* wr = fscr + 1024;                           // [2][4][128]
    wr[(half * 4 + 0) * 128 + m] = w0;
    wr[(half * 4 + 1) * 128 + m] = w1;
    wr[(half * 4 + 2) * 128 + m] = w2;
    wr[(half * 4 + 3) * 128 + m] = w3;
    __syncthreads();
    if (half == 0 && live)
      for (int i = 0; i < 4 && c0 + i < C; ++i)
        c.grads[c.po.fc2_w[h] + (size_t)(c0 + i) * H + m] = wr[i * 128 + m] + wr[(4 + i) * 128 + m];
  }
  if (t < C) {                                         // d b2
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += sDl[b * C + t];
    c.grads[c.po.fc2_b[h] + t] = s;
  }
  __syncthreads();
  const double inv = B > 0 ? 1.0 / B : 0.0;
  const double t1 = dscr[m] + dscr[256 + m], t2 = dscr[512 + m] + dscr[768 + m];
  const float c1 = (float)(t1 * inv), c2 = (float)(t2 * inv);
  if (half == 0 && live) {
    c.grads[c.bn_gamma[bn2] + m] = (float)t2;
    c.grads[c.bn_beta[bn2] + m] = (float)t1;
  }
  // d a1 = relu'(h1) * bn2'(d y2): to global [b][m] (the operand rows of the two products below), d b1 on the fly
  float db1 = 0.f;
  if (live) {
    for (int b = bb; b < be; ++b) {
      const float x = H1[(size_t)b * H + m];
      float dy = 0.f;
      for (int cls = 0; cls < C; ++cls) dy = fmaf(sDl[b * C + cls], sW2[cls * H + m], dy);
      const float xh = (x - mean2) * rstd2;
      const float u = x > 0.f ? s2c * (dy - c1 - xh * c2) : 0.f;
      DH[(size_t)b * H + m] = u;
      db1 += u;
    }
  }
  __syncthreads();
  {
    float* wr = fscr + 1024;
    wr[half * 128 + m] = db1;
    __syncthreads();
    if (half == 0 && live) c.grads[c.po.fc1_b[h] + m] = wr[m] + wr[128 + m];
  }
  __syncthreads();                                     // DH (global, written by this CTA) is visible to all its threads
  PT_MARK();                                                          // 2: fc2 / bn2 backward, d a1

  // ---- d W1[m][k] = sum_b da1[b][m] * y1[b][k]: M = hidden channels, N = input channels, K = graph rows ----
  Pipe ps = {{0u, 0u}, 0u};
  const int nkb = (K1 + 127) / 128;
  for (int kb = 0; kb < nkb; ++kb) {
    gemm_tn<BF16>(
        ring, bars, ps, tmem + (uint32_t)(kb * 128), B,
        [&](int r, int k, float (&v)[4]) {              // A: rows = hidden channel r, K = graph rows k .. k+3
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (r < H && k + e < B) ? DH[(size_t)(k + e) * H + r] : 0.f;
        },
        [&](int r, int k, float (&v)[4]) {              // B: rows = input channel kb * 128 + r
          const int kk = kb * 128 + r;
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = (kk < K1 && k + e < B) ? fmaf(in.at(k + e, kk), sc1[kk], sh1[kk]) : 0.f;
        });
  }
  PT_MARK();                                                          // 3: d W1 staging + issue
  gemm_wait(bars, ps);
  PT_MARK();                                                          // 4: d W1 MMA tail
  for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = half * 64 + cc * 32;
      float v[32];
      umma::ld32(umma::tmem_addr(tmem, q * 32, kb * 128 + col), v);
      const int k0 = kb * 128 + col;
      if (live && k0 < K1) {
        float* dst = c.grads + c.po.fc1_w[h] + (size_t)m * K1 + k0;
#pragma unroll
        for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
    }
  }
  umma::fence_before_sync();                           // the accumulator columns are reused below
  __syncthreads();
  PT_MARK();                                                          // 5: d W1 store

  // ---- d y1[b][k] = sum_m da1[b][m] * W1[m][k]: M = input channels (128 per pass), N = graph rows, K = hidden ----
  const int nblk = (B + 127) / 128;
  for (int kb = 0; kb < nkb; ++kb) {
    for (int nb = 0; nb < nblk; ++nb) {
      gemm_tn<BF16>(
          ring, bars, ps, tmem + (uint32_t)(nb * 128), H,
          [&](int r, int k, float (&v)[4]) {            // A: rows = input channel, K = hidden channels k .. k+3
            const int kk = kb * 128 + r;
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = kk < K1 ? W1[(size_t)(k + e) * K1 + kk] : 0.f;
          },
          [&](int r, int k, float (&v)[4]) {            // B: rows = graph nb * 128 + r
            const int b = nb * 128 + r;
            if (b < B) {
              const float4 d = *reinterpret_cast<const float4*>(DH + (size_t)b * H + k);
              v[0] = d.x; v[1] = d.y; v[2] = d.z; v[3] = d.w;
            } else {
              v[0] = v[1] = v[2] = v[3] = 0.f;
            }
          });
    }
    PT_MARK();                                                        // 6: d y1 staging + issue
    gemm_wait(bars, ps);
    PT_MARK();                                                        // 7: d y1 MMA tail
    // bn1 backward: thread = (input channel kk, column half), two passes over the accumulator
    const int kk = kb * 128 + m;
    const bool lk = kk < K1;
    const float m1 = lk ? mean1[kk] : 0.f, r1 = lk ? rstd1[kk] : 0.f, s1c = lk ? sc1[kk] : 0.f;
    double d1 = 0.0, d2 = 0.0;
    for (int nb = 0; nb < nblk; ++nb) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int col = half * 64 + cc * 32;
        float v[32];
        umma::ld32(umma::tmem_addr(tmem, q * 32, nb * 128 + col), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int b = nb * 128 + col + e;
          if (lk && b < B) {
            d1 += (double)v[e];
            d2 += (double)v[e] * (double)((in.at(b, kk) - m1) * r1);
          }
        }
      }
    }
    __syncthreads();
    dscr[half * 256 + m] = d1;
    dscr[512 + half * 256 + m] = d2;
    __syncthreads();
    const double u1 = dscr[m] + dscr[256 + m], u2 = dscr[512 + m] + dscr[768 + m];
    const float e1 = (float)(u1 * inv), e2 = (float)(u2 * inv);
    if (half == 0 && lk) {
      c.grads[c.bn_gamma[bn1] + kk] = (float)u2;
      c.grads[c.bn_beta[bn1] + kk] = (float)u1;
    }
    for (int nb = 0; nb < nblk; ++nb) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int col = half * 64 + cc * 32;
        float v[32];
        umma::ld32(umma::tmem_addr(tmem, q * 32, nb * 128 + col), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int b = nb * 128 + col + e;
          if (lk && b < B) {
            const float xh = (in.at(b, kk) - m1) * r1;
            c.du[((size_t)h * c.Bm + b) * 2 * H + kk] = s1c * (v[e] - e1 - xh * e2);
          }
        }
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    PT_MARK();                                                        // 8: bn1 backward + d u
  }
  PT_DUMP(c, 80);
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

template <typename K>
int set_smem_tc(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

bool readout_tc_supported(const Ctx& c) { return c.Bm <= kTmemCols && tc_smem(c.Bm, c.C, false).total <= 220 * 1024; }

int launch_readout_tc_forward(const Ctx& c, cudaStream_t s) {
  const bool bf16 = c.readout_bf16 != 0;
  const size_t smem = tc_smem(c.Bm, c.C, bf16).total;
  int rc = bf16 ? set_smem_tc(k_readout_tc_fwd<true>, smem) : set_smem_tc(k_readout_tc_fwd<false>, smem);
  if (rc) return rc;
  if (bf16) launch_k(k_readout_tc_fwd<true>, dim3(3), dim3(kT), smem, s, c);
  else launch_k(k_readout_tc_fwd<false>, dim3(3), dim3(kT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_readout_tc_backward(const Ctx& c, cudaStream_t s) {
  const bool bf16 = c.readout_bf16 != 0;
  const size_t smem = tc_smem(c.Bm, c.C, bf16).total;
  int rc = bf16 ? set_smem_tc(k_readout_tc_bwd<true>, smem) : set_smem_tc(k_readout_tc_bwd<false>, smem);
  if (rc) return rc;
  if (bf16) launch_k(k_readout_tc_bwd<true>, dim3(3), dim3(kT), smem, s, c);
  else launch_k(k_readout_tc_bwd<false>, dim3(3), dim3(kT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
