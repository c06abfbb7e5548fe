// head_ro.cu -- the readout MLPs of the fused small-graph path (model.py:125-164, train_causal.py:178-186 and
// their backward), B <= 128 graphs, hidden 128, cat_or_add = "add", C <= 8: one latency chain per head, built to be
// short.  512 threads per CTA so that every per-row / per-channel pass is one round; the fc1 products run on the
// tensor cores (tcgen05.mma, 3xTF32, accumulators in TMEM) with the pre-split weight images of k_fsg_prep streamed
// through a two-stage ring by cp.async.bulk (the first two K slices land before the dependency wait).
//
//   forward  (one CTA per head):   u -> bn1 -> [tcgen05] a1^T = W1 y1^T -> +b1, ReLU -> bn2 -> fc2 (FFMA) -> log-softmax,
//                                  KL / NLL loss parts, correct counts
//   backward (two CTAs per head):  both redo the cheap chain  d logits -> fc2 / bn2 backward -> d a1  from the saved
//                                  activations; CTA "input"  : [tcgen05] d y1^T = W1^T d a1^T -> bn1 backward -> d u
//                                               CTA "weight" : [tcgen05] d W1 = d a1^T y1 (K = graph rows), d b1
// so that the two 128 x 128 x 128 products of the backward run side by side instead of back to back.
#include "fsg_dev.cuh"

namespace cal {
namespace {

constexpr int RT = 512;                               // threads per CTA
constexpr int kRoB = 128;                             // graph rows
constexpr uint32_t kYLbo = 144, kYSbo = 32 * 144;     // row operand: chunk c of row i at (i / 8) * 4608 + c * 144 + (i % 8) * 16
constexpr int kYPart = (kRoB / 8) * (int)kYSbo;       // 73728
constexpr int kStage = 32768;                         // ring stage: 8 K-chunks of the image, hi (16 KB) | lo (16 KB)
constexpr int kLdT = FH + 4;                          // row stride of the h1 tile
constexpr int kMaxC = 8;

__device__ __forceinline__ int ro_clampB(const Ctx& c) { return imin(imax(c.dims[2], 0), imin(c.Bm, kRoB)); }
__device__ __forceinline__ uint32_t y_off(int i, int kc) { return (uint32_t)(i >> 3) * kYSbo + (uint32_t)kc * kYLbo + (uint32_t)(i & 7) * 16u; }

// one K slice (32 k = 8 chunks) of a pre-split image -> ring stage
__device__ __forceinline__ void ring_load(unsigned char* stage, const float* img, int slice, uint64_t* bar) {
  umma::mbar_expect_tx(bar, (uint32_t)kStage);
  umma::bulk_g2s(stage, img + (size_t)slice * 4096, 16384u, bar);
  umma::bulk_g2s(stage + 16384, img + kFsgImgPart + (size_t)slice * 4096, 16384u, bar);
}

// D[128 lanes][N columns] (+)= image(M = 128, K = 128) x rows(N, K = 128): the image streams through the ring
// (slices 0 and 1 have been requested by the caller), the row operand is resident.  One thread.
__device__ __forceinline__ void ro_gemm_image(unsigned char* ring, const float* img, uint64_t* bar_full, uint64_t* bar_empty,
                                              const unsigned char* y_hi, const unsigned char* y_lo, uint32_t d_main, uint32_t d_corr,
                                              int npad) {
  const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, npad);
  const uint32_t yh = umma::smem_addr(y_hi), yl = umma::smem_addr(y_lo);
  auto slice_mmas = [&](int s) {
    const uint32_t st = umma::smem_addr(ring + (s & 1) * kStage);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int gk = s * 4 + ks;
      const uint64_t ah = umma::smem_desc(st + (uint32_t)ks * 2u * kALbo, kALbo, kASbo);
      const uint64_t al = umma::smem_desc(st + 16384u + (uint32_t)ks * 2u * kALbo, kALbo, kASbo);
      const uint64_t bh = umma::smem_desc(yh + (uint32_t)gk * 2u * kYLbo, kYLbo, kYSbo);
      const uint64_t bl = umma::smem_desc(yl + (uint32_t)gk * 2u * kYLbo, kYLbo, kYSbo);
      umma::mma_tf32(d_main, ah, bh, idesc, gk > 0);
      umma::mma_tf32(d_corr, ah, bl, idesc, gk > 0);
      umma::mma_tf32(d_corr, al, bh, idesc, 1u);
    }
  };
  umma::mbar_wait(&bar_full[0], 0);
  umma::fence_after_sync();
  slice_mmas(0);
  umma::commit(&bar_empty[0]);
  umma::mbar_wait(&bar_full[1], 0);
  umma::fence_after_sync();
  slice_mmas(1);
  umma::commit(&bar_empty[1]);
  umma::mbar_wait(&bar_empty[0], 0);
  ring_load(ring, img, 2, &bar_full[0]);
  umma::mbar_wait(&bar_empty[1], 0);
  ring_load(ring + kStage, img, 3, &bar_full[1]);
  umma::mbar_wait(&bar_full[0], 1);
  umma::fence_after_sync();
  slice_mmas(2);
  umma::mbar_wait(&bar_full[1], 1);
  umma::fence_after_sync();
  slice_mmas(3);
}

// readout input u_h[b][4q .. 4q+3] ("add": model.py:152-160)
__device__ __forceinline__ float4 ro_input4(const Ctx& c, int h, int b, int q, const int* sPerm) {
  const float4* gc = reinterpret_cast<const float4*>(c.pooled);
  const float4* go = reinterpret_cast<const float4*>(c.pooled + (size_t)c.Bm * FH);
  if (h == 0) return __ldcg(gc + (size_t)b * (FH / 4) + q);
  if (h == 1) return __ldcg(go + (size_t)b * (FH / 4) + q);
  const float4 a = __ldcg(gc + (size_t)sPerm[b] * (FH / 4) + q), d = __ldcg(go + (size_t)b * (FH / 4) + q);
  return make_float4(a.x + d.x, a.y + d.y, a.z + d.z, a.w + d.w);
}
__device__ __forceinline__ float ro_input1(const Ctx& c, int h, int b, int k, const int* sPerm) {
  const float* gc = c.pooled;
  const float* go = c.pooled + (size_t)c.Bm * FH;
  if (h == 0) return __ldcg(gc + (size_t)b * FH + k);
  if (h == 1) return __ldcg(go + (size_t)b * FH + k);
  return __ldcg(gc + (size_t)sPerm[b] * FH + k) + __ldcg(go + (size_t)b * FH + k);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct RoFSmem {
  size_t y_hi, y_lo, ring, vec, w2, lg, perm, red, total;
};
__host__ __device__ inline RoFSmem ro_fsmem() {
  RoFSmem s;
  size_t o = 0;
  s.y_hi = o;  o += kYPart;                            // y1 operand hi; later the h1 tile [128][132]
  s.y_lo = o;  o += kYPart;                            // y1 operand lo; before that the fp64 statistics scratch
  s.ring = o;  o += 2 * kStage;
  s.vec = o;   o += 6 * FH * 4;                        // sc1 | sh1 | sc2 | sh2 | b1 | spare
  s.w2 = o;    o += kMaxC * kLdT * 4;
  s.lg = o;    o += kRoB * kMaxC * 4;
  s.perm = o;  o += kRoB * 4;
  s.red = o;   o += 64 * 4;
  s.total = o;
  return s;
}

__global__ void __launch_bounds__(RT, 1) k_ro_fwd(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_mma;
  __shared__ uint32_t tmem_slot;
  const RoFSmem S = ro_fsmem();
  unsigned char* sYh = smem + S.y_hi;
  unsigned char* sYl = smem + S.y_lo;
  unsigned char* sRing = smem + S.ring;
  float* sVec = reinterpret_cast<float*>(smem + S.vec);
  float* sW2 = reinterpret_cast<float*>(smem + S.w2);
  float* sLg = reinterpret_cast<float*>(smem + S.lg);
  int* sPerm = reinterpret_cast<int*>(smem + S.perm);
  float* sRed = reinterpret_cast<float*>(smem + S.red);
  float* sT = reinterpret_cast<float*>(sYh);            // h1 tile (after the product)
  double* sStat = reinterpret_cast<double*>(sYl);       // [2][16][128] (before the operand is written)
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int h = blockIdx.x;
  const FsgWs ws = fsg_ws(c);
  const int L = c.L, C = c.C;
  const int bn1 = L + 3 + h, bn2 = L + 6 + h;
  const float* img = fsg_img_fc1_fwd(ws, L, h);

  // ---- before the dependency wait: TMEM, barriers, the first two image slices, parameters ----
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (t == 0) {
    umma::mbar_init(&bar_full[0], 1);
    umma::mbar_init(&bar_full[1], 1);
    umma::mbar_init(&bar_empty[0], 1);
    umma::mbar_init(&bar_empty[1], 1);
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_fence_init();
    ring_load(sRing, img, 0, &bar_full[0]);
    ring_load(sRing + kStage, img, 1, &bar_full[1]);
  }
  float g1 = 1.f, be1 = 0.f, rm1 = 0.f, rv1 = 1.f, g2 = 1.f, be2 = 0.f, rm2 = 0.f, rv2 = 1.f;
  if (t < FH) {
    sVec[4 * FH + t] = c.params[c.po.fc1_b[h] + t];
    g1 = c.params[c.bn_gamma[bn1] + t];
    be1 = c.params[c.bn_beta[bn1] + t];
    g2 = c.params[c.bn_gamma[bn2] + t];
    be2 = c.params[c.bn_beta[bn2] + t];
    if (c.bn_buffers != nullptr && c.bn_rm[bn1] >= 0) {
      rm1 = c.bn_buffers[c.bn_rm[bn1] + t];
      rv1 = c.bn_buffers[c.bn_rv[bn1] + t];
      rm2 = c.bn_buffers[c.bn_rm[bn2] + t];
      rv2 = c.bn_buffers[c.bn_rv[bn2] + t];
    }
  }
  for (int i = t; i < C * FH; i += RT) sW2[(i >> 7) * kLdT + (i & 127)] = c.params[c.po.fc2_w[h] + i];
  const int B = ro_clampB(c);
  if (t < kRoB) sPerm[t] = t < B ? c.perm[t] : 0;         // (cal_prep's output: complete before the predecessor started)
  umma::fence_before_sync();
  FSG_TDECL
  pdl_sync();
  FSG_T(0);                                              // 0: dependency wait
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int npad = imax(8, (B + 7) & ~7);

  // ---- A: the input rows (thread = 4 channels x 8 rows), bn1 over the B rows ----
  const int q = t & 31, g = t >> 5;
  float4 u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = g * 8 + i;
    u[i] = r < B ? ro_input4(c, h, r, q, sPerm) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float* sc1 = sVec;
  float* sh1 = sVec + FH;
  float* sc2 = sVec + 2 * FH;
  float* sh2 = sVec + 3 * FH;
  const float* b1 = sVec + 4 * FH;
  if (c.train) {
    double s[4] = {0.0, 0.0, 0.0, 0.0}, qq[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[0] += (double)u[i].x; s[1] += (double)u[i].y; s[2] += (double)u[i].z; s[3] += (double)u[i].w;
      qq[0] += (double)u[i].x * (double)u[i].x; qq[1] += (double)u[i].y * (double)u[i].y;
      qq[2] += (double)u[i].z * (double)u[i].z; qq[3] += (double)u[i].w * (double)u[i].w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sStat[g * FH + q * 4 + e] = s[e];
      sStat[16 * FH + g * FH + q * 4 + e] = qq[e];
    }
    __syncthreads();
    if (t < FH) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int gg = 0; gg < 16; ++gg) {
        a += sStat[gg * FH + t];
        b += sStat[16 * FH + gg * FH + t];
      }
      const double mean = B > 0 ? a / B : 0.0;
      double var = B > 0 ? b / B - mean * mean : 0.0;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
      const float sc = g1 * rstd, sh = be1 - (float)mean * sc;
      sc1[t] = sc;
      sh1[t] = sh;
      c.bnf(bn1, BN_SCALE)[t] = sc;
      c.bnf(bn1, BN_SHIFT)[t] = sh;
      c.bnf(bn1, BN_MEAN)[t] = (float)mean;
      c.bnf(bn1, BN_RSTD)[t] = rstd;
      if (c.bn_buffers != nullptr && c.bn_rm[bn1] >= 0) {
        const double unb = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
        c.bn_buffers[c.bn_rm[bn1] + t] = (1.f - c.momentum) * rm1 + c.momentum * (float)mean;
        c.bn_buffers[c.bn_rv[bn1] + t] = (1.f - c.momentum) * rv1 + c.momentum * (float)unb;
      }
      if (t == 0 && c.nbt != nullptr) c.nbt[bn1] += 1;
    }
  } else if (t < FH) {
    sc1[t] = c.bnf(bn1, BN_SCALE)[t];
    sh1[t] = c.bnf(bn1, BN_SHIFT)[t];
  }
  __syncthreads();
  {
    const float4 sc = *reinterpret_cast<const float4*>(sc1 + q * 4), sh = *reinterpret_cast<const float4*>(sh1 + q * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = g * 8 + i;
      if (r >= npad) break;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < B) y = make_float4(fmaf(u[i].x, sc.x, sh.x), fmaf(u[i].y, sc.y, sh.y), fmaf(u[i].z, sc.z, sh.z), fmaf(u[i].w, sc.w, sh.w));
      float h0, h1_, h2, h3, l0, l1, l2, l3;
      umma::split_tf32(y.x, h0, l0);
      umma::split_tf32(y.y, h1_, l1);
      umma::split_tf32(y.z, h2, l2);
      umma::split_tf32(y.w, h3, l3);
      const uint32_t off = y_off(r, q);
      *reinterpret_cast<float4*>(sYh + off) = make_float4(h0, h1_, h2, h3);
      *reinterpret_cast<float4*>(sYl + off) = make_float4(l0, l1, l2, l3);
    }
  }
  umma::fence_async_smem();
  __syncthreads();
  FSG_T(1);                                              // 1: input rows, bn1, operand

  // ---- B: a1^T [out channel][graph] = W1 y1^T on the tensor cores ----
  if (t == 0) {
    umma::fence_after_sync();
    ro_gemm_image(sRing, img, bar_full, bar_empty, sYh, sYl, tmem, tmem + 128u, npad);
    umma::commit(&bar_mma);
  }
  umma::mbar_wait(&bar_mma, 0);
  umma::fence_after_sync();
  FSG_T(2);                                              // 2: fc1 product

  // ---- C: epilogue: h1 = relu(a1 + b1) -> tile [graph][channel]; bn2 sums (thread = channel, 4 column groups) ----
  double* sSt2 = reinterpret_cast<double*>(sRing);      // [2][4][128] (the ring is idle)
  {
    const int j = (warp & 3) * 32 + lane, cg = warp >> 2, c0 = cg * 32;
    double s = 0.0, qq = 0.0;
    if (c0 < npad) {
      const float bj = b1[j];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float vm[16], vc[16];
        umma::ld16(umma::tmem_addr(tmem, (warp & 3) * 32, c0 + hh * 16), vm);
        umma::ld16(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + c0 + hh * 16), vc);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int b = c0 + hh * 16 + e;
          if (b < B) {
            const float v = fmaxf((vm[e] + vc[e]) + bj, 0.f);
            sT[b * kLdT + j] = v;
            s += (double)v;
            qq += (double)v * (double)v;
          }
        }
      }
    }
    sSt2[cg * FH + j] = s;
    sSt2[4 * FH + cg * FH + j] = qq;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (c.train) {
    if (t < FH) {
      const double a = (sSt2[t] + sSt2[FH + t]) + (sSt2[2 * FH + t] + sSt2[3 * FH + t]);
      const double b = (sSt2[4 * FH + t] + sSt2[5 * FH + t]) + (sSt2[6 * FH + t] + sSt2[7 * FH + t]);
      const double mean = B > 0 ? a / B : 0.0;
      double var = B > 0 ? b / B - mean * mean : 0.0;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
      const float sc = g2 * rstd, sh = be2 - (float)mean * sc;
      sc2[t] = sc;
      sh2[t] = sh;
      c.bnf(bn2, BN_SCALE)[t] = sc;
      c.bnf(bn2, BN_SHIFT)[t] = sh;
      c.bnf(bn2, BN_MEAN)[t] = (float)mean;
      c.bnf(bn2, BN_RSTD)[t] = rstd;
      if (c.bn_buffers != nullptr && c.bn_rm[bn2] >= 0) {
        const double unb = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
        c.bn_buffers[c.bn_rm[bn2] + t] = (1.f - c.momentum) * rm2 + c.momentum * (float)mean;
        c.bn_buffers[c.bn_rv[bn2] + t] = (1.f - c.momentum) * rv2 + c.momentum * (float)unb;
      }
      if (t == 0 && c.nbt != nullptr) c.nbt[bn2] += 1;
    }
  } else if (t < FH) {
    sc2[t] = c.bnf(bn2, BN_SCALE)[t];
    sh2[t] = c.bnf(bn2, BN_SHIFT)[t];
  }
  __syncthreads();
  FSG_T(3);                                              // 3: epilogue + bn2

  // ---- D: h1 to the workspace (the backward reads it); fc2: thread per (graph, class) ----
  {
    float4* H1 = reinterpret_cast<float4*>(c.H1 + (size_t)h * c.Bm * FH);
    for (int i = t; i < B * (FH / 4); i += RT) {
      const int b = i >> 5, qd = i & 31;
      H1[(size_t)b * (FH / 4) + qd] = *reinterpret_cast<const float4*>(sT + b * kLdT + qd * 4);
    }
  }
  for (int i = t; i < B * C; i += RT) {
    const int b = i / C, cls = i - b * C;
    const float* hr = sT + b * kLdT;
    const float* wr = sW2 + cls * kLdT;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
    for (int k = 0; k < FH; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(hr + k), wv = *reinterpret_cast<const float4*>(wr + k);
      const float4 sc = *reinterpret_cast<const float4*>(sc2 + k), sh = *reinterpret_cast<const float4*>(sh2 + k);
      s0 = fmaf(fmaf(hv.x, sc.x, sh.x), wv.x, s0);
      s1 = fmaf(fmaf(hv.y, sc.y, sh.y), wv.y, s1);
      s0 = fmaf(fmaf(hv.z, sc.z, sh.z), wv.z, s0);
      s1 = fmaf(fmaf(hv.w, sc.w, sh.w), wv.w, s1);
    }
    sLg[i] = (s0 + s1) + c.params[c.po.fc2_b[h] + cls];
  }
  __syncthreads();
  FSG_T(4);                                              // 4: h1 store + fc2

  // ---- E: log-softmax, loss parts (train_causal.py:178-183), correct counts (:184-186) ----
  float loss_part = 0.f, correct_part = 0.f;
  if (t < B) {
    const int b = t;
    float m = -INFINITY;
    int am = 0;
    for (int cls = 0; cls < C; ++cls) {
      const float v = sLg[b * C + cls];
      if (v > m) {
        m = v;
        am = cls;
      }
    }
    float se = 0.f;
    for (int cls = 0; cls < C; ++cls) se += expf(sLg[b * C + cls] - m);
    const float lse = logf(se);
    const long long yb = (c.with_loss && c.y != nullptr) ? c.y[b] : -1;
    float slp = 0.f, picked = 0.f;
    for (int cls = 0; cls < C; ++cls) {
      const float lp = sLg[b * C + cls] - m - lse;
      c.logp[((size_t)h * c.Bm + b) * C + cls] = lp;
      slp += lp;
      if ((long long)cls == yb) picked = lp;
    }
    if (c.with_loss) {
      loss_part = h == 0 ? -logf((float)C) - slp / (float)C : -picked;   // KL(uniform || .) row / NLL row
      correct_part = (long long)am == yb ? 1.f : 0.f;
    }
  }
  if (c.with_loss) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
      correct_part += __shfl_xor_sync(0xffffffffu, correct_part, o);
    }
    if (lane == 0) {
      sRed[warp] = loss_part;
      sRed[16 + warp] = correct_part;
    }
    __syncthreads();
    if (t == 0) {
      float ls = 0.f, cs = 0.f;
      for (int w = 0; w < RT / 32; ++w) {
        ls += sRed[w];
        cs += sRed[16 + w];
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
  FSG_T(5);                                              // 5: log-softmax, loss
  FSG_TDUMP(c, 80);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------
// backward: blockIdx.x = 2 * head + role (0: input gradient, 1: weight gradient)
// ---------------------------------------------------------------------------------------------
struct RoBSmem {
  size_t op, ring, vec, w2, dl, perm, st, dw2, total;
};
__host__ __device__ inline RoBSmem ro_bsmem() {
  RoBSmem s;
  size_t o = 0;
  s.op = o;    o += 2 * kYPart;                        // role 0: d a1 row operand hi | lo;  role 1: two 64 KB K-slice stages
  s.dw2 = s.op;                                        // fc2 weight-gradient partials [4][C][128] (before the operands are built)
  s.st = s.op + 131072;                                // fp64 partial sums [2][4][128]: only live while the operand area is idle
  s.ring = o;  o += 2 * kStage;                        // role 0: W1^T image ring
  s.vec = o;   o += 12 * FH * 4;                       // bn2: sc sh mean rstd c1 c2 | bn1: sc sh mean rstd c1 c2
  s.w2 = o;    o += kMaxC * FH * 4;
  s.dl = o;    o += kRoB * kMaxC * 4;
  s.perm = o;  o += kRoB * 4;
  s.total = o;
  return s;
}
static_assert(2 * kYPart >= 131072 + 2 * 4 * FH * 8, "the fp64 partial sums fit behind the weight-gradient stages");

__global__ void __launch_bounds__(RT, 1) k_ro_bwd(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_mma;
  __shared__ uint32_t tmem_slot;
  const RoBSmem S = ro_bsmem();
  unsigned char* sOp = smem + S.op;
  unsigned char* sRing = smem + S.ring;
  float* sVec = reinterpret_cast<float*>(smem + S.vec);
  float* sW2 = reinterpret_cast<float*>(smem + S.w2);   // [C][128]
  float* sDl = reinterpret_cast<float*>(smem + S.dl);   // [B][C]
  int* sPerm = reinterpret_cast<int*>(smem + S.perm);
  double* sSt = reinterpret_cast<double*>(smem + S.st);
  float* sDw2 = reinterpret_cast<float*>(smem + S.dw2);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int h = blockIdx.x >> 1, role = blockIdx.x & 1;
  const FsgWs ws = fsg_ws(c);
  const int L = c.L, C = c.C;
  const int bn1 = L + 3 + h, bn2 = L + 6 + h;
  const float* img = fsg_img_fc1_bwd(ws, L, h);
  float* v2 = sVec;                                     // bn2 vectors
  float* v1 = sVec + 6 * FH;                            // bn1 vectors

  // ---- before the dependency wait ----
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (t == 0) {
    umma::mbar_init(&bar_full[0], 1);
    umma::mbar_init(&bar_full[1], 1);
    umma::mbar_init(&bar_empty[0], 1);
    umma::mbar_init(&bar_empty[1], 1);
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_fence_init();
    if (role == 0) {
      ring_load(sRing, img, 0, &bar_full[0]);
      ring_load(sRing + kStage, img, 1, &bar_full[1]);
    }
  }
  for (int i = t; i < C * FH; i += RT) sW2[i] = c.params[c.po.fc2_w[h] + i];
  const int B = ro_clampB(c);
  if (t < kRoB) sPerm[t] = t < B ? c.perm[t] : 0;
  umma::fence_before_sync();
  FSG_TDECL
  pdl_sync();
  FSG_T(0);                                              // 0: dependency wait
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int npad = imax(8, (B + 7) & ~7);

  // ---- d logits (thread per graph), the BatchNorm records of the forward ----
  if (t < B) {
    const int b = t;
    const long long yb = c.y != nullptr ? c.y[b] : -1;
    float dlp[kMaxC], sd = 0.f;
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls) {
      dlp[cls] = 0.f;
      if (cls < C) {
        if (c.grad_logp != nullptr) dlp[cls] = c.grad_logp[((size_t)h * B + b) * C + cls];
        else if (h == 0) dlp[cls] = -c.w_c / ((float)C * (float)B);
        else dlp[cls] = (long long)cls == yb ? -(h == 1 ? c.w_o : c.w_co) / (float)B : 0.f;
        sd += dlp[cls];
      }
    }
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls)
      if (cls < C) sDl[b * C + cls] = dlp[cls] - expf(c.logp[((size_t)h * c.Bm + b) * C + cls]) * sd;
  }
  if (t < FH) {
    v2[t] = c.bnf(bn2, BN_SCALE)[t];
    v2[FH + t] = c.bnf(bn2, BN_SHIFT)[t];
    v2[2 * FH + t] = c.bnf(bn2, BN_MEAN)[t];
    v2[3 * FH + t] = c.bnf(bn2, BN_RSTD)[t];
  } else if (t < 2 * FH) {
    const int k = t - FH;
    v1[k] = c.bnf(bn1, BN_SCALE)[k];
    v1[FH + k] = c.bnf(bn1, BN_SHIFT)[k];
    v1[2 * FH + k] = c.bnf(bn1, BN_MEAN)[k];
    v1[3 * FH + k] = c.bnf(bn1, BN_RSTD)[k];
  }
  __syncthreads();
  FSG_T(1);                                              // 1: d logits, records

  // ---- fc2 / bn2 backward sums: thread = (hidden channel j, row part): d y2 = dl W2, sum d y2, sum d y2 * hhat,
  //      and d W2[cls][j] = sum_b dl[b][cls] * y2[b][j] ----
  const int j = t & 127, part = t >> 7;
  const float* H1 = c.H1 + (size_t)h * c.Bm * FH;
  float w2c[kMaxC];
#pragma unroll
  for (int cls = 0; cls < kMaxC; ++cls) w2c[cls] = cls < C ? sW2[cls * FH + j] : 0.f;
  const float sc2 = v2[j], sh2 = v2[FH + j], mu2 = v2[2 * FH + j], rs2 = v2[3 * FH + j];
  {
    double s = 0.0, qq = 0.0;
    float dw[kMaxC];
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls) dw[cls] = 0.f;
    for (int b0 = part; b0 < B; b0 += 32) {              // 8 rows per batch of loads
      float hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int b = b0 + 4 * i;
        hv[i] = b < B ? __ldcg(H1 + (size_t)b * FH + j) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int b = b0 + 4 * i;
        if (b < B) {
          const float y2 = fmaf(hv[i], sc2, sh2), hh = (hv[i] - mu2) * rs2;
          float dy = 0.f;
#pragma unroll
          for (int cls = 0; cls < kMaxC; ++cls)
            if (cls < C) {
              const float d = sDl[b * C + cls];
              dy = fmaf(d, w2c[cls], dy);
              dw[cls] = fmaf(d, y2, dw[cls]);
            }
          s += (double)dy;
          qq += (double)dy * (double)hh;
        }
      }
    }
    sSt[part * FH + j] = s;
    sSt[4 * FH + part * FH + j] = qq;
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls)
      if (cls < C) sDw2[(part * kMaxC + cls) * FH + j] = dw[cls];
  }
  __syncthreads();
  if (t < FH) {
    const double a = (sSt[t] + sSt[FH + t]) + (sSt[2 * FH + t] + sSt[3 * FH + t]);
    const double b = (sSt[4 * FH + t] + sSt[5 * FH + t]) + (sSt[6 * FH + t] + sSt[7 * FH + t]);
    const double inv = B > 0 ? 1.0 / B : 0.0;
    v2[4 * FH + t] = (float)(a * inv);
    v2[5 * FH + t] = (float)(b * inv);
    if (role == 0) {
      c.bnf(bn2, BN_C1)[t] = (float)(a * inv);
      c.bnf(bn2, BN_C2)[t] = (float)(b * inv);
      c.grads[c.bn_gamma[bn2] + t] = (float)b;
      c.grads[c.bn_beta[bn2] + t] = (float)a;
    }
  }
  if (role == 0) {
    for (int i = t; i < C * FH; i += RT) {
      const int cls = i >> 7, k = i & 127;
      c.grads[c.po.fc2_w[h] + i] = (sDw2[(0 * kMaxC + cls) * FH + k] + sDw2[(1 * kMaxC + cls) * FH + k]) +
                                   (sDw2[(2 * kMaxC + cls) * FH + k] + sDw2[(3 * kMaxC + cls) * FH + k]);
    }
    if (t < C) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += sDl[b * C + t];
      c.grads[c.po.fc2_b[h] + t] = s;
    }
  }
  __syncthreads();
  FSG_T(2);                                              // 2: fc2 / bn2 backward sums
  const float c1 = v2[4 * FH + j], c2 = v2[5 * FH + j];
  // d a1[b][j] = relu'(h1) * bn2'(d y2)
  auto da1 = [&](int b, float hval) -> float {
    float dy = 0.f;
#pragma unroll
    for (int cls = 0; cls < kMaxC; ++cls)
      if (cls < C) dy = fmaf(sDl[b * C + cls], w2c[cls], dy);
    const float hh = (hval - mu2) * rs2;
    return hval > 0.f ? sc2 * (dy - c1 - hh * c2) : 0.f;
  };

  if (role == 0) {
    // ================= input gradient: d y1^T [in channel][graph] = W1^T d a1^T =================
    unsigned char* sYh = sOp;
    unsigned char* sYl = sOp + kYPart;
    for (int b0 = part; b0 < npad; b0 += 32) {
      float hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int b = b0 + 4 * i;
        hv[i] = b < B ? __ldcg(H1 + (size_t)b * FH + j) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int b = b0 + 4 * i;
        if (b < npad) {
          const float v = b < B ? da1(b, hv[i]) : 0.f;
          float hi, lo;
          umma::split_tf32(v, hi, lo);
          const uint32_t off = y_off(b, j >> 2) + (uint32_t)(j & 3) * 4u;
          *reinterpret_cast<float*>(sYh + off) = hi;
          *reinterpret_cast<float*>(sYl + off) = lo;
        }
      }
    }
    umma::fence_async_smem();
    __syncthreads();
    FSG_T(3);                                            // 3: d a1 operand
    if (t == 0) {
      umma::fence_after_sync();
      ro_gemm_image(sRing, img, bar_full, bar_empty, sYh, sYl, tmem, tmem + 128u, npad);
      umma::commit(&bar_mma);
    }
    // the input rows of this thread's channel and column group, fetched while the product runs
    const int k = (warp & 3) * 32 + lane, cg = warp >> 2, c0 = cg * 32;
    float uu[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int b = c0 + e;
      uu[e] = b < B ? ro_input1(c, h, b, k, sPerm) : 0.f;
    }
    const float sc1 = v1[k], mu1 = v1[2 * FH + k], rs1 = v1[3 * FH + k];
    umma::mbar_wait(&bar_mma, 0);
    umma::fence_after_sync();
    FSG_T(4);                                            // 4: product (+ input rows)
    float d[32];
    double s = 0.0, qq = 0.0;
    if (c0 < npad) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float vm[16], vc[16];
        umma::ld16(umma::tmem_addr(tmem, (warp & 3) * 32, c0 + hh * 16), vm);
        umma::ld16(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + c0 + hh * 16), vc);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int b = c0 + hh * 16 + e;
          const float v = b < B ? vm[e] + vc[e] : 0.f;
          d[hh * 16 + e] = v;
          uu[hh * 16 + e] = (uu[hh * 16 + e] - mu1) * rs1;      // uhat
          s += (double)v;
          qq += (double)v * (double)uu[hh * 16 + e];
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) d[e] = 0.f;
    }
    sSt[cg * FH + k] = s;
    sSt[4 * FH + cg * FH + k] = qq;
    __syncthreads();
    const double a = (sSt[k] + sSt[FH + k]) + (sSt[2 * FH + k] + sSt[3 * FH + k]);
    const double bsum = (sSt[4 * FH + k] + sSt[5 * FH + k]) + (sSt[6 * FH + k] + sSt[7 * FH + k]);
    const double inv = B > 0 ? 1.0 / B : 0.0;
    const float e1 = (float)(a * inv), e2 = (float)(bsum * inv);
    if (cg == 0) {
      c.bnf(bn1, BN_C1)[k] = e1;
      c.bnf(bn1, BN_C2)[k] = e2;
      c.grads[c.bn_gamma[bn1] + k] = (float)bsum;
      c.grads[c.bn_beta[bn1] + k] = (float)a;
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int b = c0 + e;
      if (b < B) c.du[((size_t)h * c.Bm + b) * 2 * FH + k] = sc1 * (d[e] - e1 - uu[e] * e2);
    }
    FSG_T(5);                                            // 5: bn1 backward, d u
    if (blockIdx.x == 0 && threadIdx.x == 0)
      for (int q_ = 0; q_ < 8; ++q_) c.status[112 + q_] = FSG_TVAL(q_);
  } else {
    // ================= weight gradient: d W1 [out][in] = d a1^T y1 over the graph rows; d b1 =================
    // thread groups: (t < 256: K slice 2p, t >= 256: K slice 2p + 1) x (first 128: d a1^T chunks, next 128: y1^T chunks)
    const int grp = t >> 8, sub = (t >> 7) & 1, ch = t & 127;
    const float sc1 = v1[ch], sh1 = v1[FH + ch];
    float db1 = 0.f;
    const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, 128);
    for (int p = 0; p < 2; ++p) {
      if (p == 1) {                                      // the stages are reused: the first pass's products must be done
        umma::mbar_wait(&bar_empty[0], 0);
        umma::fence_after_sync();
      }
      const int s = 2 * p + grp;                         // K slice: rows 32 s .. 32 s + 31
      unsigned char* st = sOp + grp * 65536 + sub * 32768;     // stage: d a1^T hi | lo | y1^T hi | lo (16 KB each)
      if (32 * s < npad) {
        if (sub == 0) {
          float hv[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int b = 32 * s + e;
            hv[e] = b < B ? __ldcg(H1 + (size_t)b * FH + ch) : 0.f;
          }
#pragma unroll
          for (int kc = 0; kc < 8; ++kc) {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int b = 32 * s + kc * 4 + e;
              const float v = b < B ? da1(b, hv[kc * 4 + e]) : 0.f;
              db1 += v;
              umma::split_tf32(v, hi[e], lo[e]);
            }
            const uint32_t off = (uint32_t)kc * kALbo + (uint32_t)ch * 16u;
            *reinterpret_cast<float4*>(st + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(st + 16384 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
        } else {
          float uv[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int b = 32 * s + e;
            uv[e] = b < B ? ro_input1(c, h, b, ch, sPerm) : 0.f;
          }
#pragma unroll
          for (int kc = 0; kc < 8; ++kc) {
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int b = 32 * s + kc * 4 + e;
              const float v = b < B ? fmaf(uv[kc * 4 + e], sc1, sh1) : 0.f;
              umma::split_tf32(v, hi[e], lo[e]);
            }
            const uint32_t off = (uint32_t)kc * kALbo + (uint32_t)ch * 16u;
            *reinterpret_cast<float4*>(st + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(st + 16384 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
      umma::fence_async_smem();
      __syncthreads();
      if (t == 0) {
        umma::fence_after_sync();
        for (int g2 = 0; g2 < 2; ++g2) {
          const int sl = 2 * p + g2;
          if (32 * sl >= npad) break;
          const uint32_t base = umma::smem_addr(sOp + g2 * 65536);
          const int ksteps = imin(4, (npad - 32 * sl) / 8);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t o = (uint32_t)ks * 2u * kALbo;
            const uint64_t ah = umma::smem_desc(base + o, kALbo, kASbo), al = umma::smem_desc(base + 16384u + o, kALbo, kASbo);
            const uint64_t bh = umma::smem_desc(base + 32768u + o, kALbo, kASbo), bl = umma::smem_desc(base + 49152u + o, kALbo, kASbo);
            const uint32_t first = (sl == 0 && ks == 0) ? 0u : 1u;
            umma::mma_tf32(tmem, al, bh, idesc, first);
            umma::mma_tf32(tmem, ah, bl, idesc, 1u);
            umma::mma_tf32(tmem, ah, bh, idesc, 1u);
          }
        }
        umma::commit(p == 0 ? &bar_empty[0] : &bar_mma);
      }
    }
    // d b1: the two slice groups of a channel, fixed order
    float* sDb = reinterpret_cast<float*>(sSt);
    if (sub == 0) sDb[grp * FH + ch] = db1;
    FSG_T(3);                                            // 3: operand slices + issue (both passes)
    umma::mbar_wait(&bar_mma, 0);
    umma::fence_after_sync();
    __syncthreads();
    FSG_T(4);                                            // 4: product tail
    if (t < FH) c.grads[c.po.fc1_b[h] + t] = sDb[t] + sDb[FH + t];
    // d W1 from TMEM: 16 warps = 4 lane quarters x 4 column groups, transposed through a private scratch
    {
      float* sw = reinterpret_cast<float*>(sOp) + warp * 32 * 33;     // (the operand stages are idle)
      const int row0 = (warp & 3) * 32, col0 = (warp >> 2) * 32;
      float v[32];
      umma::ld32(umma::tmem_addr(tmem, row0, col0), v);
#pragma unroll
      for (int cc = 0; cc < 32; ++cc) sw[lane * 33 + cc] = v[cc];
      __syncwarp();
      float* dst = c.grads + c.po.fc1_w[h];
      const int rr = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* qv = sw + (rr + 4 * i) * 33 + c4;
        *reinterpret_cast<float4*>(dst + (size_t)(row0 + rr + 4 * i) * FH + col0 + c4) = make_float4(qv[0], qv[1], qv[2], qv[3]);
      }
    }
    FSG_T(5);                                            // 5: d W1 drain
    if (blockIdx.x == 1 && threadIdx.x == 0)
      for (int q_ = 0; q_ < 8; ++q_) c.status[120 + q_] = FSG_TVAL(q_);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace

bool readout_ro_supported(const Ctx& c) {
  return c.fsg_on && c.H == FH && !c.cat && c.Bm <= kRoB && c.C <= kMaxC && !c.readout_bf16 && !c.readout_tc;
}

int launch_readout_ro_forward(const Ctx& c, cudaStream_t s) {
  const size_t smem = ro_fsmem().total;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ro_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_k(k_ro_fwd, dim3(3), dim3(RT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_readout_ro_backward(const Ctx& c, cudaStream_t s) {
  const size_t smem = ro_bsmem().total;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_ro_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_k(k_ro_bwd, dim3(6), dim3(RT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
