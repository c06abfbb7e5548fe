// head_tc2.cu -- the tensor-core readout (see head_tc.cu for the formulation) for the common envelope
// B <= 128 graphs, cat_or_add = "add", C <= 8 classes: everything a CTA touches more than once is RESIDENT in
// shared memory.
//
// head_tc.cu streams its operands from global memory inside per-thread loops over the graph rows; measured
// on B200 (profiles/r02_c_phases.txt) those loops are latency-bound -- one L2 round trip per row -- and the
// kernel is 3x slower than the FFMA cluster kernels it replaces.  Here the readout input u [B, H] and
// the hidden activations h1 [B, H] are loaded ONCE with fully coalesced, fully overlapped float4 loads
// into two padded shared-memory tiles (row stride H + 4 floats: row-per-lane float4 accesses are
// conflict-free), and every later pass -- BatchNorm statistics, operand staging of all three tensor-core
// products, the BatchNorm backward sums -- reads shared memory.  d a1 overwrites h1 in place, so the
// backward products read their operands from the same two tiles.
//
//   forward : u tile -> bn1 -> [tcgen05] a1^T = W1 y1^T -> +b1, ReLU -> h1 tile -> bn2 -> fc2 (FFMA, C <= 8) -> loss
//   backward: d logits -> fc2 / bn2 backward in the h1 tile -> [tcgen05] dW1 = da1^T y1, [tcgen05] dy1^T = W1^T da1^T -> bn1'
#include "internal.cuh"
#include "umma.cuh"

namespace cal {
namespace {

constexpr int kT = 256;
constexpr int kRows = 128;                    // graph rows (tile rows, MMA N)
constexpr int kLd = 132;                      // tile row stride in floats (H <= 128, + 4)
constexpr int kPart = 8192;                   // operand part of a ring slot: 128 rows x 64 bytes of K (4 chunks)
constexpr uint32_t kLbo = 2048, kSbo = 128;   // chunk c of row r at c * 2048 + r * 16
constexpr int kMaxC = 8;

__device__ __forceinline__ int clampB(const Ctx& c) { return imin(imax(c.dims[2], 0), c.Bm); }

struct Pipe {
  uint32_t uses[2];
  uint32_t n;
};

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

// D^T[128 lanes][128 columns at d_tmem] = sum_{k < K} A[lane][k] * B[column][k].  ga / gb(r, k, v): v[0..3] =
// elements k .. k+3 of operand row r (zero outside; k % 4 == 0).  The operands of slot i + 1 are fetched
// into registers right after slot i has been written, so global-memory getters overlap the MMAs.
template <bool BF16, class GA, class GB>
__device__ __forceinline__ void gemm_tn(unsigned char* ring, uint64_t* bars, Pipe& ps, uint32_t d_tmem, int K, GA ga, GB gb) {
  constexpr int EPC = BF16 ? 8 : 4;            // elements per 16-byte chunk
  constexpr int KSLOT = 4 * EPC;               // K elements per ring slot (64 bytes per row)
  constexpr int PARTS = BF16 ? 2 : 4;
  constexpr int NQ = BF16 ? 2 : 1;             // getter calls per chunk
  const int t = threadIdx.x, r = t & 127, c0 = t >> 7;
  const uint32_t idesc = umma::instr_desc(BF16 ? umma::kFmtBF16 : umma::kFmtTF32, 128, 128);
  float va[2][8], vb[2][8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) va[j][e] = vb[j][e] = 0.f;
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq) {
        const int k = k0 + (c0 + 2 * j) * EPC + 4 * qq;
        if (k < K) {
          float q4[4];
          ga(r, k, q4);
#pragma unroll
          for (int e = 0; e < 4; ++e) va[j][4 * qq + e] = q4[e];
          gb(r, k, q4);
#pragma unroll
          for (int e = 0; e < 4; ++e) vb[j][4 * qq + e] = q4[e];
        }
      }
    }
  };
  fetch(0);
  uint32_t issued = 0;
  for (int k0 = 0; k0 < K; k0 += KSLOT) {
    const int s = (int)(ps.n & 1u);
    unsigned char* slot = ring + (size_t)s * PARTS * kPart;
    unsigned char* pAh = slot;
    unsigned char* pAl = slot + kPart;
    unsigned char* pBh = slot + (BF16 ? 1 : 2) * kPart;
    unsigned char* pBl = slot + 3 * kPart;
    if (ps.uses[s] > 0) umma::mbar_wait(&bars[s], (ps.uses[s] - 1u) & 1u);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      put_chunk<BF16>(pAh, pAl, r, c0 + 2 * j, va[j]);
      put_chunk<BF16>(pBh, pBl, r, c0 + 2 * j, vb[j]);
    }
    if (k0 + KSLOT < K) fetch(k0 + KSLOT);
    umma::fence_async_smem();
    __syncthreads();
    if (t == 0) {
      umma::fence_after_sync();
      const int ksteps = imin(2, (K - k0 + 2 * EPC - 1) / (2 * EPC));
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
__device__ __forceinline__ void gemm_wait(uint64_t* bars, const Pipe& ps) {
  if (ps.n == 0u) return;
  const int s = (int)((ps.n - 1u) & 1u);
  umma::mbar_wait(&bars[s], (ps.uses[s] - 1u) & 1u);
  umma::fence_after_sync();
}

// shared-memory carve-up (bytes)
struct Smem2 {
  size_t ring, t1, t2, lg, dscr, fscr, w2, perm, total;
};
__host__ __device__ inline Smem2 smem2(bool bf16) {
  Smem2 s;
  size_t o = 0;
  s.ring = o;  o += (size_t)2 * (bf16 ? 2 : 4) * kPart;       // 64 KB (tf32) / 32 KB (bf16)
  s.t1 = o;    o += (size_t)kRows * kLd * 4;                   // u tile
  s.t2 = o;    o += (size_t)kRows * kLd * 4;                   // h1 / d a1 tile
  s.lg = o;    o += (size_t)kRows * kMaxC * 4;                 // logits / d logits
  s.dscr = o;  o += 4 * 256 * 8;
  s.fscr = o;  o += 10 * 256 * 4;
  s.w2 = o;    o += (size_t)kMaxC * 128 * 4;
  s.perm = o;  o += (size_t)kRows * 4;
  s.total = o;
  return s;
}

// the readout input rows u_h[b][:] -> tile (model.py:119-121,152-157; "add" variant), zero rows beyond B
__device__ __forceinline__ void load_input_tile(const Ctx& c, int h, int B, int H, const int* sPerm, float* T1) {
  const float* gc = c.pooled;
  const float* go = c.pooled + (size_t)c.Bm * H;
  const int q4 = H / 4;
  for (int i0 = 0; i0 < kRows * q4; i0 += 4 * kT) {
    float4 v[4], w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kT + (int)threadIdx.x;
      const int b = i / q4, k = (i - b * q4) * 4;
      v[u] = w[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < kRows * q4 && b < B) {
        if (h == 0) v[u] = *reinterpret_cast<const float4*>(gc + (size_t)b * H + k);
        else if (h == 1) v[u] = *reinterpret_cast<const float4*>(go + (size_t)b * H + k);
        else {
          v[u] = *reinterpret_cast<const float4*>(gc + (size_t)sPerm[b] * H + k);
          w[u] = *reinterpret_cast<const float4*>(go + (size_t)b * H + k);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kT + (int)threadIdx.x;
      const int b = i / q4, k = (i - b * q4) * 4;
      if (i < kRows * q4)
        *reinterpret_cast<float4*>(T1 + b * kLd + k) = make_float4(v[u].x + w[u].x, v[u].y + w[u].y, v[u].z + w[u].z, v[u].w + w[u].w);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward.  grid = 3 (head c, o, co), 256 threads.
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kT, 1) k_readout_tc2_fwd(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3, half = warp >> 2;
  const int h = blockIdx.x;
  const int H = c.H, C = c.C;
  const int bn1 = c.L + 3 + h, bn2 = c.L + 6 + h;
  const Smem2 L = smem2(BF16);
  unsigned char* ring = smem + L.ring;
  float* T1 = reinterpret_cast<float*>(smem + L.t1);
  float* T2 = reinterpret_cast<float*>(smem + L.t2);
  float* sLg = reinterpret_cast<float*>(smem + L.lg);
  double* dscr = reinterpret_cast<double*>(smem + L.dscr);
  float* fscr = reinterpret_cast<float*>(smem + L.fscr);
  float *sc1 = fscr, *sh1 = fscr + 128, *sc2 = fscr + 256, *sh2 = fscr + 384, *b1s = fscr + 512, *b2f = fscr + 640, *red = fscr + 768;
  float* sW2 = reinterpret_cast<float*>(smem + L.w2);               // [C][128]: W2 (later W2 * sc2)
  int* sPerm = reinterpret_cast<int*>(smem + L.perm);

  // ---- before the dependency wait: parameters only ----
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 128);
  if (t == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::mbar_fence_init();
  }
  const int ch = t & 127, hf = t >> 7;                 // (channel, row half) of the statistics passes
  float g1 = 1.f, be1 = 0.f, rm1 = 0.f, rv1 = 1.f, g2 = 1.f, be2 = 0.f, rm2 = 0.f, rv2 = 1.f;
  const bool has_run = c.bn_buffers != nullptr && c.bn_rm[bn1] >= 0;
  if (t < H) {
    g1 = c.params[c.bn_gamma[bn1] + t];
    be1 = c.params[c.bn_beta[bn1] + t];
    g2 = c.params[c.bn_gamma[bn2] + t];
    be2 = c.params[c.bn_beta[bn2] + t];
    b1s[t] = c.params[c.po.fc1_b[h] + t];
    if (has_run) {
      rm1 = c.bn_buffers[c.bn_rm[bn1] + t];
      rv1 = c.bn_buffers[c.bn_rv[bn1] + t];
      rm2 = c.bn_buffers[c.bn_rm[bn2] + t];
      rv2 = c.bn_buffers[c.bn_rv[bn2] + t];
    }
  }
  for (int i = t; i < C * 128; i += kT) sW2[i] = (i & 127) < H ? c.params[c.po.fc2_w[h] + (size_t)(i >> 7) * H + (i & 127)] : 0.f;
  const float* W1 = c.params + c.po.fc1_w[h];                         // [H][H]
  umma::fence_before_sync();
  PT_DECL
  pdl_sync();                                                         // everything below may read the predecessor's output
  PT_MARK();                                                          // 0: dependency wait
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int B = clampB(c);
  long long ylab = -1;                                                // label of graph row t (input data), fetched early
  if (t < B) {
    sPerm[t] = c.perm[t];
    if (c.with_loss && c.y != nullptr) ylab = c.y[t];
  }
  __syncthreads();
  load_input_tile(c, h, B, H, sPerm, T1);
  __syncthreads();
  PT_MARK();                                                          // 1: input tile

  // ---- bn1: two threads per input channel (row halves), fixed order ----
  if (c.train) {
    double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
    const int rb = hf * 64, re = imin(B, rb + 64);
    for (int b = rb; b + 1 < re; b += 2) {
      const double x0 = (double)T1[b * kLd + ch], x1 = (double)T1[(b + 1) * kLd + ch];
      s0 += x0; q0 += x0 * x0;
      s1 += x1; q1 += x1 * x1;
    }
    if (re > rb && ((re - rb) & 1)) {
      const double x0 = (double)T1[(re - 1) * kLd + ch];
      s0 += x0; q0 += x0 * x0;
    }
    dscr[hf * 256 + ch] = s0 + s1;
    dscr[512 + hf * 256 + ch] = q0 + q1;
  }
  __syncthreads();
  if (t < H) {
    float sc, sh;
    if (c.train) {
      const double sum = dscr[t] + dscr[256 + t], sq = dscr[512 + t] + dscr[768 + t];
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
  } else if (t < 128) {
    sc1[t] = sh1[t] = 0.f;
  }
  if (c.train && t == 0 && c.nbt != nullptr) {
    c.nbt[bn1] += 1;
    c.nbt[bn2] += 1;
  }
  __syncthreads();
  PT_MARK();                                                          // 2: bn1

  // ---- fc1 on the tensor cores: a1^T[m][b] = sum_k W1[m][k] * y1[b][k] ----
  Pipe ps = {{0u, 0u}, 0u};
  gemm_tn<BF16>(
      ring, bars, ps, tmem, H,
      [&](int r, int k, float (&v)[4]) {                        // A: W1 rows (output channels), global
        if (r < H) {
          const float4 w = *reinterpret_cast<const float4*>(W1 + (size_t)r * H + k);
          v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
        } else {
          v[0] = v[1] = v[2] = v[3] = 0.f;
        }
      },
      [&](int r, int k, float (&v)[4]) {                        // B: bn1(u) rows (graphs), from the tile
        if (r < B) {
          const float4 x = *reinterpret_cast<const float4*>(T1 + r * kLd + k);
          const float4 a = *reinterpret_cast<const float4*>(sc1 + k), bq = *reinterpret_cast<const float4*>(sh1 + k);
          v[0] = fmaf(x.x, a.x, bq.x); v[1] = fmaf(x.y, a.y, bq.y); v[2] = fmaf(x.z, a.z, bq.z); v[3] = fmaf(x.w, a.w, bq.w);
        } else {
          v[0] = v[1] = v[2] = v[3] = 0.f;
        }
      });
  PT_MARK();                                                          // 3: fc1 staging + issue
  gemm_wait(bars, ps);
  PT_MARK();                                                          // 4: fc1 MMA tail

  // ---- epilogue: thread = (hidden channel m, column half): + bias, ReLU -> h1 (global + tile), bn2 statistics ----
  const int m = q * 32 + lane;
  float* H1 = c.H1 + (size_t)h * c.Bm * H;
  const float bias1 = m < H ? b1s[m] : 0.f;
  {
    double es0 = 0.0, es1 = 0.0, eq0 = 0.0, eq1 = 0.0;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = half * 64 + cc * 32;
      float v[32];
      umma::ld32(umma::tmem_addr(tmem, q * 32, col), v);
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const int b = col + e;
        const float x0 = (b < B && m < H) ? fmaxf(v[e] + bias1, 0.f) : 0.f;
        const float x1 = (b + 1 < B && m < H) ? fmaxf(v[e + 1] + bias1, 0.f) : 0.f;
        T2[b * kLd + m] = x0;
        T2[(b + 1) * kLd + m] = x1;
        if (b < B && m < H) H1[(size_t)b * H + m] = x0;
        if (b + 1 < B && m < H) H1[(size_t)(b + 1) * H + m] = x1;
        es0 += (double)x0; eq0 += (double)x0 * (double)x0;
        es1 += (double)x1; eq1 += (double)x1 * (double)x1;
      }
    }
    if (c.train) {
      dscr[half * 256 + m] = es0 + es1;
      dscr[512 + half * 256 + m] = eq0 + eq1;
    }
  }
  __syncthreads();
  if (t < 128) {
    float sc = 0.f, sh = 0.f;
    if (t < H) {
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
    }
    sc2[t] = sc;
    sh2[t] = sh;
  }
  __syncthreads();
  PT_MARK();                                                          // 5: epilogue + bn2
  // ---- fc2 with bn2 folded into the weights: logits[b][cls] = sum_m h1[b][m] (sc2[m] W2[cls][m]) + (b2[cls] + sum_m sh2[m] W2[cls][m]) ----
  if (warp < C) {                                                     // warp cls: folded bias, then scale its weight row
    float s = 0.f;
    for (int k = lane; k < 128; k += 32) s = fmaf(sh2[k], sW2[warp * 128 + k], s);
    s = warp_sum(s);
    if (lane == 0) b2f[warp] = s + c.params[c.po.fc2_b[h] + warp];
    __syncwarp();
    for (int k = lane; k < 128; k += 32) sW2[warp * 128 + k] *= sc2[k];
  }
  __syncthreads();
  for (int i = t; i < kRows * C; i += kT) {
    const int b = i & 127, cls = i >> 7;
    const float4* xr = reinterpret_cast<const float4*>(T2 + b * kLd);
    const float4* wr = reinterpret_cast<const float4*>(sW2 + cls * 128);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
    for (int k4 = 0; k4 < 32; ++k4) {
      const float4 x = xr[k4], w = wr[k4];
      s0 = fmaf(x.x, w.x, s0); s1 = fmaf(x.y, w.y, s1);
      s0 = fmaf(x.z, w.z, s0); s1 = fmaf(x.w, w.w, s1);
    }
    sLg[b * kMaxC + cls] = s0 + s1 + b2f[cls];
  }
  __syncthreads();
  PT_MARK();                                                          // 6: fc2

  // ---- log_softmax, outputs, loss parts (train_causal.py:178-186): thread = graph row ----
  float loss_part = 0.f, correct_part = 0.f;
  if (t < B) {
    const int b = t;
    float mx = -INFINITY;
    int am = 0;
    for (int cls = 0; cls < C; ++cls) {
      const float v = sLg[b * kMaxC + cls];
      if (v > mx) {
        mx = v;
        am = cls;
      }
    }
    float se = 0.f;
    for (int cls = 0; cls < C; ++cls) se += expf(sLg[b * kMaxC + cls] - mx);
    const float lse = logf(se);
    float slp = 0.f, picked = 0.f;
    for (int cls = 0; cls < C; ++cls) {
      const float lp = sLg[b * kMaxC + cls] - mx - lse;
      c.logp[((size_t)h * c.Bm + b) * C + cls] = lp;
      slp += lp;
      if ((long long)cls == ylab) picked = lp;
    }
    if (c.with_loss) {
      loss_part = h == 0 ? -logf((float)C) - slp / (float)C : -picked;      // KL(uniform || .) row / NLL row
      correct_part = (long long)am == ylab ? 1.f : 0.f;
    }
  }
  if (c.with_loss) {
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
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------------------------
// Backward.  grid = 3, 256 threads.
// ---------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(kT, 1) k_readout_tc2_bwd(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int q = warp & 3, half = warp >> 2;
  const int h = blockIdx.x;
  const int H = c.H, C = c.C;
  const int bn1 = c.L + 3 + h, bn2 = c.L + 6 + h;
  const Smem2 L = smem2(BF16);
  unsigned char* ring = smem + L.ring;
  float* T1 = reinterpret_cast<float*>(smem + L.t1);                 // u [b][k]
  float* T2 = reinterpret_cast<float*>(smem + L.t2);                 // h1 [b][m], then d a1 [b][m]
  float* sDl = reinterpret_cast<float*>(smem + L.lg);                // [b][kMaxC]
  double* dscr = reinterpret_cast<double*>(smem + L.dscr);
  float* fscr = reinterpret_cast<float*>(smem + L.fscr);
  float *sc1 = fscr, *sh1 = fscr + 128, *wr = fscr + 256;            // wr: [2][kMaxC + 1][128] partial sums
  float* sW2 = reinterpret_cast<float*>(smem + L.w2);
  int* sPerm = reinterpret_cast<int*>(smem + L.perm);

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (t == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::mbar_fence_init();
  }
  for (int i = t; i < C * 128; i += kT) sW2[i] = (i & 127) < H ? c.params[c.po.fc2_w[h] + (size_t)(i >> 7) * H + (i & 127)] : 0.f;
  const float* W1 = c.params + c.po.fc1_w[h];                         // [H][H]
  umma::fence_before_sync();
  PT_DECL
  pdl_sync();
  PT_MARK();                                                          // 0: dependency wait
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int B = clampB(c);
  const int m = t & 127, hf = t >> 7;                  // (channel, row half) of the per-channel passes
  const bool live = m < H;
  // per-channel BatchNorm records (written by the forward kernel)
  const float s1c = live ? c.bnf(bn1, BN_SCALE)[m] : 0.f, s1h = live ? c.bnf(bn1, BN_SHIFT)[m] : 0.f;
  const float mean1 = live ? c.bnf(bn1, BN_MEAN)[m] : 0.f, rstd1 = live ? c.bnf(bn1, BN_RSTD)[m] : 0.f;
  const float s2c = live ? c.bnf(bn2, BN_SCALE)[m] : 0.f, s2h = live ? c.bnf(bn2, BN_SHIFT)[m] : 0.f;
  const float mean2 = live ? c.bnf(bn2, BN_MEAN)[m] : 0.f, rstd2 = live ? c.bnf(bn2, BN_RSTD)[m] : 0.f;
  if (hf == 0) {
    sc1[m] = s1c;
    sh1[m] = s1h;
  }
  if (t < B) sPerm[t] = c.perm[t];
  // ---- d logits (log_softmax backward): thread = graph row ----
  if (t < B) {
    const int b = t;
    const long long yb = c.y != nullptr ? c.y[b] : -1;
    float dl[kMaxC], lp[kMaxC];
    float sd = 0.f;
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls) {
      dl[cls] = 0.f;
      lp[cls] = 0.f;
      if (cls < C) {
        lp[cls] = c.logp[((size_t)h * c.Bm + b) * C + cls];
        if (c.grad_logp != nullptr) dl[cls] = c.grad_logp[((size_t)h * B + b) * C + cls];
        else if (h == 0) dl[cls] = -c.w_c / ((float)C * (float)B);
        else dl[cls] = (long long)cls == yb ? -(h == 1 ? c.w_o : c.w_co) / (float)B : 0.f;
        sd += dl[cls];
      }
    }
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls) sDl[b * kMaxC + cls] = cls < C ? dl[cls] - expf(lp[cls]) * sd : 0.f;
  } else if (t < kRows) {
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls) sDl[t * kMaxC + cls] = 0.f;
  }
  __syncthreads();
  load_input_tile(c, h, B, H, sPerm, T1);
  {                                                    // h1 tile (coalesced float4 loads, all in flight)
    const float* H1 = c.H1 + (size_t)h * c.Bm * H;
    const int q4 = H / 4;
    for (int i0 = 0; i0 < kRows * q4; i0 += 4 * kT) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kT + t;
        const int b = i / q4, k = (i - b * q4) * 4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < kRows * q4 && b < B) v[u] = *reinterpret_cast<const float4*>(H1 + (size_t)b * H + k);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kT + t;
        const int b = i / q4, k = (i - b * q4) * 4;
        if (i < kRows * q4) *reinterpret_cast<float4*>(T2 + b * kLd + k) = v[u];
      }
    }
  }
  __syncthreads();
  PT_MARK();                                                          // 1: d logits + tiles

  // ---- fc2 / bn2 backward in the h1 tile: thread = (hidden channel m, row half) ----
  const int rb = hf * 64, re = imin(B, rb + 64);
  float w2r[kMaxC];
#pragma unroll
  for (int cls = 0; cls < kMaxC; ++cls) w2r[cls] = cls < C ? sW2[cls * 128 + m] : 0.f;
  double a1s = 0.0, a2s = 0.0;
  float dw2[kMaxC];
#pragma unroll
  for (int cls = 0; cls < kMaxC; ++cls) dw2[cls] = 0.f;
  for (int b = rb; b < re; ++b) {
    const float x = T2[b * kLd + m];
    const float4 d0 = *reinterpret_cast<const float4*>(sDl + b * kMaxC), d1 = *reinterpret_cast<const float4*>(sDl + b * kMaxC + 4);
    const float dl[kMaxC] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    const float y2 = fmaf(x, s2c, s2h);
    float dy = 0.f;
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls) {
      dy = fmaf(dl[cls], w2r[cls], dy);
      dw2[cls] = fmaf(dl[cls], y2, dw2[cls]);
    }
    a1s += (double)dy;
    a2s += (double)dy * (double)((x - mean2) * rstd2);
  }
  dscr[hf * 256 + m] = a1s;
  dscr[512 + hf * 256 + m] = a2s;
#pragma unroll
  for (int cls = 0; cls < kMaxC; ++cls) wr[(hf * (kMaxC + 1) + cls) * 128 + m] = dw2[cls];
  __syncthreads();
  const double inv = B > 0 ? 1.0 / B : 0.0;
  const double t1 = dscr[m] + dscr[256 + m], t2 = dscr[512 + m] + dscr[768 + m];
  const float c1 = (float)(t1 * inv), c2 = (float)(t2 * inv);
  if (hf == 0 && live) {
    c.grads[c.bn_gamma[bn2] + m] = (float)t2;
    c.grads[c.bn_beta[bn2] + m] = (float)t1;
    for (int cls = 0; cls < C; ++cls)
      c.grads[c.po.fc2_w[h] + (size_t)cls * H + m] = wr[cls * 128 + m] + wr[((kMaxC + 1) + cls) * 128 + m];
  }
  if (hf == 1 && m < C) {                              // d b2
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += sDl[b * kMaxC + m];
    c.grads[c.po.fc2_b[h] + m] = s;
  }
  // d a1 = relu'(h1) * bn2'(d y2), in place; d b1 on the fly
  float db1 = 0.f;
  for (int b = rb; b < re; ++b) {
    const float x = T2[b * kLd + m];
    const float4 d0 = *reinterpret_cast<const float4*>(sDl + b * kMaxC), d1 = *reinterpret_cast<const float4*>(sDl + b * kMaxC + 4);
    const float dl[kMaxC] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
    float dy = 0.f;
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls) dy = fmaf(dl[cls], w2r[cls], dy);
    const float xh = (x - mean2) * rstd2;
    const float u = (live && x > 0.f) ? s2c * (dy - c1 - xh * c2) : 0.f;
    T2[b * kLd + m] = u;
    db1 += u;
  }
  wr[(hf * (kMaxC + 1) + kMaxC) * 128 + m] = db1;
  __syncthreads();
  if (hf == 0 && live) c.grads[c.po.fc1_b[h] + m] = wr[kMaxC * 128 + m] + wr[((kMaxC + 1) + kMaxC) * 128 + m];
  PT_MARK();                                                          // 2: fc2 / bn2 backward, d a1

  // ---- d W1[m][k] = sum_b da1[b][m] y1[b][k]  (TMEM columns 0..127);  d y1^T[k][b] = sum_m W1[m][k] da1[b][m]  (128..255) ----
  Pipe ps = {{0u, 0u}, 0u};
  gemm_tn<BF16>(
      ring, bars, ps, tmem, B,
      [&](int r, int k, float (&v)[4]) {                // A: rows = hidden channel r, K = graph rows k .. k+3 (tile column reads)
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = r < H ? T2[(k + e) * kLd + r] : 0.f;
      },
      [&](int r, int k, float (&v)[4]) {                // B: rows = input channel r: y1 = bn1(u)
        const float a = sc1[r], bq = sh1[r];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (k + e < B && r < H) ? fmaf(T1[(k + e) * kLd + r], a, bq) : 0.f;
      });
  gemm_tn<BF16>(
      ring, bars, ps, tmem + 128u, H,
      [&](int r, int k, float (&v)[4]) {                // A: rows = input channel r, K = hidden channels k .. k+3 (global, coalesced)
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = r < H ? W1[(size_t)(k + e) * H + r] : 0.f;
      },
      [&](int r, int k, float (&v)[4]) {                // B: rows = graph r: d a1 row (tile rows beyond B are zero)
        const float4 d = *reinterpret_cast<const float4*>(T2 + r * kLd + k);
        v[0] = d.x; v[1] = d.y; v[2] = d.z; v[3] = d.w;
      });
  PT_MARK();                                                          // 3: staging + issue of both products
  gemm_wait(bars, ps);
  PT_MARK();                                                          // 4: MMA tail
  {                                                    // d W1 rows: thread = (hidden channel, column half)
    const int mm = q * 32 + lane;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = half * 64 + cc * 32;
      float v[32];
      umma::ld32(umma::tmem_addr(tmem, q * 32, col), v);
      if (mm < H && col < H) {
        float* dst = c.grads + c.po.fc1_w[h] + (size_t)mm * H + col;
#pragma unroll
        for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(dst + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
      }
    }
  }
  PT_MARK();                                                          // 5: d W1 store
  // ---- bn1 backward: thread = (input channel kk, column half), the accumulator stays in registers ----
  {
    const int kk = q * 32 + lane;
    const bool lk = kk < H;
    const float m1 = lk ? c.bnf(bn1, BN_MEAN)[kk] : 0.f, r1 = lk ? c.bnf(bn1, BN_RSTD)[kk] : 0.f, sck = lk ? sc1[kk] : 0.f;
    float v[2][32];
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = half * 64 + cc * 32;
      umma::ld32(umma::tmem_addr(tmem, q * 32, 128 + col), v[cc]);
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int b = col + e;
        if (lk && b < B) {
          d1 += (double)v[cc][e];
          d2 += (double)v[cc][e] * (double)((T1[b * kLd + kk] - m1) * r1);
        }
      }
    }
    __syncthreads();
    dscr[half * 256 + kk] = d1;
    dscr[512 + half * 256 + kk] = d2;
    __syncthreads();
    const double u1 = dscr[kk] + dscr[256 + kk], u2 = dscr[512 + kk] + dscr[768 + kk];
    const float e1 = (float)(u1 * inv), e2 = (float)(u2 * inv);
    if (half == 0 && lk) {
      c.grads[c.bn_gamma[bn1] + kk] = (float)u2;
      c.grads[c.bn_beta[bn1] + kk] = (float)u1;
    }
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int col = half * 64 + cc * 32;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int b = col + e;
        if (lk && b < B) {
          const float xh = (T1[b * kLd + kk] - m1) * r1;
          c.du[((size_t)h * c.Bm + b) * 2 * H + kk] = sck * (v[cc][e] - e1 - xh * e2);
        }
      }
    }
  }
  PT_MARK();                                                          // 6: bn1 backward + d u
  PT_DUMP(c, 80);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

template <typename K>
int set_smem2(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace

bool readout_tc2_supported(const Ctx& c) { return c.Bm <= kRows && !c.cat && c.C <= kMaxC; }

int launch_readout_tc2_forward(const Ctx& c, cudaStream_t s) {
  const bool bf16 = c.readout_bf16 != 0;
  const size_t smem = smem2(bf16).total;
  int rc = bf16 ? set_smem2(k_readout_tc2_fwd<true>, smem) : set_smem2(k_readout_tc2_fwd<false>, smem);
  if (rc) return rc;
  if (bf16) launch_k(k_readout_tc2_fwd<true>, dim3(3), dim3(kT), smem, s, c);
  else launch_k(k_readout_tc2_fwd<false>, dim3(3), dim3(kT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_readout_tc2_backward(const Ctx& c, cudaStream_t s) {
  const bool bf16 = c.readout_bf16 != 0;
  const size_t smem = smem2(bf16).total;
  int rc = bf16 ? set_smem2(k_readout_tc2_bwd<true>, smem) : set_smem2(k_readout_tc2_bwd<false>, smem);
  if (rc) return rc;
  if (bf16) launch_k(k_readout_tc2_bwd<true>, dim3(3), dim3(kT), smem, s, c);
  else launch_k(k_readout_tc2_bwd<false>, dim3(3), dim3(kT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
