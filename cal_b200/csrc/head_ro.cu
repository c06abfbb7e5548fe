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
  constexpr uint64_t da = (2u * kALbo) >> 4, db = (2u * kYLbo) >> 4;       // descriptor step: start-address field only
  auto slice_mmas = [&](int s) {
    const uint32_t st = umma::smem_addr(ring + (s & 1) * kStage);
    uint64_t ah = umma::smem_desc(st, kALbo, kASbo), al = umma::smem_desc(st + 16384u, kALbo, kASbo);
    uint64_t bh = umma::smem_desc(yh + (uint32_t)(s * 4) * 2u * kYLbo, kYLbo, kYSbo);
    uint64_t bl = umma::smem_desc(yl + (uint32_t)(s * 4) * 2u * kYLbo, kYLbo, kYSbo);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t acc = (s * 4 + ks) > 0 ? 1u : 0u;
      umma::mma_tf32(d_main, ah, bh, idesc, acc);
      umma::mma_tf32(d_corr, ah, bl, idesc, acc);
      umma::mma_tf32(d_corr, al, bh, idesc, 1u);
      ah += da; al += da; bh += db; bl += db;
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

// CC: number of classes when it is 2, 3 or 4 (loops over the classes unrolled without predicates), 8 = generic (C <= 8)
template <int CC>
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
  float* sB2raw = sRed + 48;                              // [C] fc2 bias
  if (t < C) sB2raw[t] = c.params[c.po.fc2_b[h] + t];
  const int B = ro_clampB(c);
  if (t < kRoB) sPerm[t] = t < B ? c.perm[t] : 0;         // (cal_prep's output: complete before the predecessor started)
  umma::fence_before_sync();
  FSG_TDECL
  pdl_sync();
  if (blockIdx.x == 0) CAL_TL(c.status, 4);
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
    // per thread: mean and centred sum of squares of its <= 8 rows (fp32: 8 terms, no cancellation), combined over the
    // 16 row groups in fp64 (Chan et al.): as accurate as fp64 accumulation at a fraction of the instructions
    const int cnt = imin(imax(B - g * 8, 0), 8);
    float* sMean = reinterpret_cast<float*>(sStat);       // [16][128]
    float* sM2 = sMean + 16 * FH;
    {
      float m[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 8; ++i) {                        // rows >= B hold zeros
        m[0] += u[i].x; m[1] += u[i].y; m[2] += u[i].z; m[3] += u[i].w;
      }
      const float inv = cnt > 0 ? 1.f / (float)cnt : 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] *= inv;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < cnt) {
          const float d0 = u[i].x - m[0], d1 = u[i].y - m[1], d2 = u[i].z - m[2], d3 = u[i].w - m[3];
          m2[0] = fmaf(d0, d0, m2[0]); m2[1] = fmaf(d1, d1, m2[1]); m2[2] = fmaf(d2, d2, m2[2]); m2[3] = fmaf(d3, d3, m2[3]);
        }
      *reinterpret_cast<float4*>(sMean + g * FH + q * 4) = make_float4(m[0], m[1], m[2], m[3]);
      *reinterpret_cast<float4*>(sM2 + g * FH + q * 4) = make_float4(m2[0], m2[1], m2[2], m2[3]);
    }
    __syncthreads();
    FSG_T(6);                                            // 6: input rows loaded + per-thread statistics
    if (t < FH) {
      // Chan combination of the 16 groups, fp32: every term is a mean or a centred sum of squares (no cancellation)
      float a = 0.f;
#pragma unroll
      for (int gg = 0; gg < 16; ++gg) a = fmaf((float)imin(imax(B - gg * 8, 0), 8), sMean[gg * FH + t], a);
      const float invB = B > 0 ? 1.f / (float)B : 0.f;
      const float mean = a * invB;
      float b = 0.f;
#pragma unroll
      for (int gg = 0; gg < 16; ++gg) {
        const float dm = sMean[gg * FH + t] - mean;
        b += sM2[gg * FH + t] + (float)imin(imax(B - gg * 8, 0), 8) * dm * dm;
      }
      const float var = b * invB;
      const float rstd = 1.0f / sqrtf(var + c.eps);
      const float sc = g1 * rstd, sh = be1 - mean * sc;
      sc1[t] = sc;
      sh1[t] = sh;
      c.bnf(bn1, BN_SCALE)[t] = sc;
      c.bnf(bn1, BN_SHIFT)[t] = sh;
      c.bnf(bn1, BN_MEAN)[t] = mean;
      c.bnf(bn1, BN_RSTD)[t] = rstd;
      if (c.bn_buffers != nullptr && c.bn_rm[bn1] >= 0) {
        const float unb = B > 1 ? b / (float)(B - 1) : var;
        c.bn_buffers[c.bn_rm[bn1] + t] = (1.f - c.momentum) * rm1 + c.momentum * mean;
        c.bn_buffers[c.bn_rv[bn1] + t] = (1.f - c.momentum) * rv1 + c.momentum * unb;
      }
      if (t == 0 && c.nbt != nullptr) atomicAdd(reinterpret_cast<unsigned long long*>(c.nbt + bn1), 1ull);   // (RED: no load round trip in front of the barrier)
    }
  } else if (t < FH) {
    sc1[t] = c.bnf(bn1, BN_SCALE)[t];
    sh1[t] = c.bnf(bn1, BN_SHIFT)[t];
  }
  __syncthreads();
  FSG_T(7);                                              // 7: bn1 finalisation
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
  float* sSt2 = reinterpret_cast<float*>(sRing);        // [2][4][128]: mean | centred sum of squares per column group (the ring is idle)
  {
    const int j = (warp & 3) * 32 + lane, cg = warp >> 2, c0 = cg * 32;
    const int cnt = imin(imax(B - c0, 0), 32);
    float v[32];
    float m = 0.f, m2 = 0.f;
    if (c0 < npad) {
      const float bj = b1[j];
      float* H1 = c.H1 + (size_t)h * c.Bm * FH;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float vm[16], vc[16];
        umma::ld16(umma::tmem_addr(tmem, (warp & 3) * 32, c0 + hh * 16), vm);
        umma::ld16(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + c0 + hh * 16), vc);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int b = c0 + hh * 16 + e;
          const float x = b < B ? fmaxf((vm[e] + vc[e]) + bj, 0.f) : 0.f;
          v[hh * 16 + e] = x;
          if (b < B) {
            sT[b * kLdT + j] = x;
            H1[(size_t)b * FH + j] = x;                    // the backward reads h1 from the workspace
            m += x;
          }
        }
      }
      m *= cnt > 0 ? 1.f / (float)cnt : 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (e < cnt) m2 = fmaf(v[e] - m, v[e] - m, m2);
    }
    sSt2[cg * FH + j] = m;
    sSt2[4 * FH + cg * FH + j] = m2;
  }
  umma::fence_before_sync();
  __syncthreads();
  FSG_T(8);                                              // 8: TMEM epilogue (h1 tile, bn2 partial statistics)
  if (c.train) {
    if (t < FH) {
      float a = 0.f;
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) a = fmaf((float)imin(imax(B - cg * 32, 0), 32), sSt2[cg * FH + t], a);
      const float invB = B > 0 ? 1.f / (float)B : 0.f;
      const float mean = a * invB;
      float b = 0.f;
#pragma unroll
      for (int cg = 0; cg < 4; ++cg) {
        const float dm = sSt2[cg * FH + t] - mean;
        b += sSt2[4 * FH + cg * FH + t] + (float)imin(imax(B - cg * 32, 0), 32) * dm * dm;
      }
      const float var = b * invB;
      const float rstd = 1.0f / sqrtf(var + c.eps);
      const float sc = g2 * rstd, sh = be2 - mean * sc;
      sc2[t] = sc;
      sh2[t] = sh;
      c.bnf(bn2, BN_SCALE)[t] = sc;
      c.bnf(bn2, BN_SHIFT)[t] = sh;
      c.bnf(bn2, BN_MEAN)[t] = mean;
      c.bnf(bn2, BN_RSTD)[t] = rstd;
      if (c.bn_buffers != nullptr && c.bn_rm[bn2] >= 0) {
        const float unb = B > 1 ? b / (float)(B - 1) : var;
        c.bn_buffers[c.bn_rm[bn2] + t] = (1.f - c.momentum) * rm2 + c.momentum * mean;
        c.bn_buffers[c.bn_rv[bn2] + t] = (1.f - c.momentum) * rv2 + c.momentum * unb;
      }
      if (t == 0 && c.nbt != nullptr) atomicAdd(reinterpret_cast<unsigned long long*>(c.nbt + bn2), 1ull);   // (RED: no load round trip in front of the barrier)
    }
  } else if (t < FH) {
    sc2[t] = c.bnf(bn2, BN_SCALE)[t];
    sh2[t] = c.bnf(bn2, BN_SHIFT)[t];
  }
  __syncthreads();
  FSG_T(3);                                              // 3: epilogue + bn2

  // ---- D: h1 to the workspace (the backward reads it); fc2 with bn2 folded into its weights:
  //      logits[b][cls] = sum_k h1[b][k] (sc2[k] W2[cls][k]) + (b2[cls] + sum_k sh2[k] W2[cls][k]); thread per (graph, class) ----
  {
    float* sB2 = sRed + 32;                               // [C] folded bias
    if (warp < C) {                                        // warp cls: the folded bias from the raw weights
      float a = 0.f;
#pragma unroll
      for (int k = lane; k < FH; k += 32) a = fmaf(sh2[k], sW2[warp * kLdT + k], a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) sB2[warp] = a + sB2raw[warp];
    }
    __syncthreads();
    for (int i = t; i < C * FH; i += RT) {
      const int cls = i >> 7, k = i & 127;
      sW2[cls * kLdT + k] *= sc2[k];
    }
    __syncthreads();
    FSG_T(9);                                            // 9: weight fold + h1 store
    for (int i = t; i < B * C; i += RT) {
      const int b = i / C, cls = i - b * C;
      const float* hr = sT + b * kLdT;
      const float* wr = sW2 + cls * kLdT;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
      for (int k = 0; k < FH; k += 4) {
        const float4 hv = *reinterpret_cast<const float4*>(hr + k), wv = *reinterpret_cast<const float4*>(wr + k);
        s0 = fmaf(hv.x, wv.x, s0);
        s1 = fmaf(hv.y, wv.y, s1);
        s2 = fmaf(hv.z, wv.z, s2);
        s3 = fmaf(hv.w, wv.w, s3);
      }
      sLg[i] = ((s0 + s1) + (s2 + s3)) + sB2[cls];
    }
  }
  __syncthreads();
  FSG_T(4);                                              // 4: h1 store + fc2

  // ---- E: log-softmax, loss parts (train_causal.py:178-183), correct counts (:184-186) ----
  float loss_part = 0.f, correct_part = 0.f;
  if (t < B) {
    const int b = t;
    float m = -INFINITY;
    int am = 0;
    float lg[CC];
#pragma unroll
    for (int cls = 0; cls < CC; ++cls) {
      lg[cls] = cls < C ? sLg[b * C + cls] : -INFINITY;
      if (lg[cls] > m) {
        m = lg[cls];
        am = cls;
      }
    }
    float se = 0.f;
#pragma unroll
    for (int cls = 0; cls < CC; ++cls)
      if (cls < C) se += expf(lg[cls] - m);
    const float lse = logf(se);
    const long long yb = (c.with_loss && c.y != nullptr) ? c.y[b] : -1;
    float slp = 0.f, picked = 0.f;
#pragma unroll
    for (int cls = 0; cls < CC; ++cls)
      if (cls < C) {
        const float lp = lg[cls] - m - lse;
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
    FSG_T(10);                                           // 10: log-softmax + loss parts
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
  if (blockIdx.x == 0) CAL_TL(c.status, 5);
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

template <int CC>
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
  if (blockIdx.x == 0) CAL_TL(c.status, 6);
  FSG_T(0);                                              // 0: dependency wait
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int npad = imax(8, (B + 7) & ~7);

  // ---- d logits (thread per graph), the BatchNorm records of the forward ----
  if (t < B) {
    const int b = t;
    const long long yb = c.y != nullptr ? c.y[b] : -1;
    float dlp[CC], sd = 0.f;
#pragma unroll
    for (int cls = 0; cls < CC; ++cls) {
      dlp[cls] = 0.f;
      if (cls < C) {
        if (c.grad_logp != nullptr) dlp[cls] = c.grad_logp[((size_t)h * B + b) * C + cls];
        else if (h == 0) dlp[cls] = -c.w_c / ((float)C * (float)B);
        else dlp[cls] = (long long)cls == yb ? -(h == 1 ? c.w_o : c.w_co) / (float)B : 0.f;
        sd += dlp[cls];
      }
    }
#pragma unroll
    for (int cls = 0; cls < CC; ++cls)
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
  // thread = (hidden channel j, row part p): the 32 CONSECUTIVE graph rows 32 p .. 32 p + 31 of channel j stay in
  // registers from here to the operand build (no second pass over memory, no recomputation)
  const int j = t & 127, part = t >> 7;
  const float* H1 = c.H1 + (size_t)h * c.Bm * FH;
  float hv[32], dy[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int b = 32 * part + i;
    hv[i] = b < B ? __ldcg(H1 + (size_t)b * FH + j) : 0.f;
  }
  float w2c[CC];
#pragma unroll
  for (int cls = 0; cls < CC; ++cls) w2c[cls] = cls < C ? sW2[cls * FH + j] : 0.f;
  __syncthreads();
  FSG_T(1);                                              // 1: d logits, records, h1 rows

  // ---- fc2 / bn2 backward sums: d y2 = dl W2, sum d y2, sum d y2 * hhat (fp32 over the thread's 32 rows, fp64 over the
  //      4 parts), and d W2[cls][j] = sum_b dl[b][cls] * y2[b][j] ----
  const float sc2 = v2[j], sh2 = v2[FH + j], mu2 = v2[2 * FH + j], rs2 = v2[3 * FH + j];
  float* sStF = reinterpret_cast<float*>(sSt);            // [2][4][128]
  {
    float s = 0.f, qq = 0.f, dw[CC];
#pragma unroll
    for (int cls = 0; cls < CC; ++cls) dw[cls] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int b = 32 * part + i;
      float d = 0.f;
      if (b < B) {
        const float y2 = fmaf(hv[i], sc2, sh2);
#pragma unroll
        for (int cls = 0; cls < CC; ++cls)
          if (cls < C) {
            const float dl = sDl[b * C + cls];
            d = fmaf(dl, w2c[cls], d);
            dw[cls] = fmaf(dl, y2, dw[cls]);
          }
        s += d;
        qq = fmaf(d, (hv[i] - mu2) * rs2, qq);
      }
      dy[i] = d;
    }
    sStF[part * FH + j] = s;
    sStF[4 * FH + part * FH + j] = qq;
#pragma unroll
    for (int cls = 0; cls < CC; ++cls)
      if (cls < C) sDw2[(part * kMaxC + cls) * FH + j] = dw[cls];
  }
  __syncthreads();
  if (t < FH) {
    const double a = ((double)sStF[t] + (double)sStF[FH + t]) + ((double)sStF[2 * FH + t] + (double)sStF[3 * FH + t]);
    const double b = ((double)sStF[4 * FH + t] + (double)sStF[5 * FH + t]) + ((double)sStF[6 * FH + t] + (double)sStF[7 * FH + t]);
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
    if (warp < C) {                                      // warp cls: d b2[cls] = sum_b dl[b][cls]
      float a = 0.f;
      for (int b = lane; b < B; b += 32) a += sDl[b * C + warp];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) c.grads[c.po.fc2_b[h] + warp] = a;
    }
  }
  __syncthreads();
  FSG_T(2);                                              // 2: fc2 / bn2 backward sums
  {
    // d a1[b][j] = relu'(h1) * bn2'(d y2), in place of d y2
    const float c1 = v2[4 * FH + j], c2 = v2[5 * FH + j];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float hh = (hv[i] - mu2) * rs2;
      dy[i] = (hv[i] > 0.f && 32 * part + i < B) ? sc2 * (dy[i] - c1 - hh * c2) : 0.f;
    }
  }

  if (role == 0) {
    // ================= input gradient: d y1^T [in channel][graph] = W1^T d a1^T =================
    unsigned char* sYh = sOp;
    unsigned char* sYl = sOp + kYPart;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int b = 32 * part + i;
      if (b < npad) {
        float hi, lo;
        umma::split_tf32(dy[i], hi, lo);
        const uint32_t off = y_off(b, j >> 2) + (uint32_t)(j & 3) * 4u;
        *reinterpret_cast<float*>(sYh + off) = hi;
        *reinterpret_cast<float*>(sYl + off) = lo;
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
    float s = 0.f, qq = 0.f;
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
          dy[hh * 16 + e] = v;                             // (d a1 is dead: the registers now hold d y1)
          uu[hh * 16 + e] = (uu[hh * 16 + e] - mu1) * rs1;  // uhat
          s += v;
          qq = fmaf(v, uu[hh * 16 + e], qq);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) dy[e] = 0.f;
    }
    sStF[cg * FH + k] = s;
    sStF[4 * FH + cg * FH + k] = qq;
    __syncthreads();
    const double a = ((double)sStF[k] + (double)sStF[FH + k]) + ((double)sStF[2 * FH + k] + (double)sStF[3 * FH + k]);
    const double bsum = ((double)sStF[4 * FH + k] + (double)sStF[5 * FH + k]) + ((double)sStF[6 * FH + k] + (double)sStF[7 * FH + k]);
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
      if (b < B) c.du[((size_t)h * c.Bm + b) * 2 * FH + k] = sc1 * (dy[e] - e1 - uu[e] * e2);
    }
    FSG_T(5);                                            // 5: bn1 backward, d u
    if (blockIdx.x == 0 && threadIdx.x == 0)
      for (int q_ = 0; q_ < 8; ++q_) c.status[112 + q_] = FSG_TVAL(q_);
  } else {
    // ================= weight gradient: d W1 [out][in] = d a1^T y1 over the graph rows; d b1 =================
    // A operand d a1^T: all four 32-row K slices resident (slice p at sOp + p * 32 KB: hi 16 KB | lo 16 KB), built by
    // thread (j, p) straight from its registers.  B operand y1^T: two slices at a time in the (otherwise unused) ring
    // area; thread (k, p) builds 16 rows of a slice per pass.
    float db1 = 0.f;
    {
      unsigned char* st = sOp + part * 32768;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          db1 += dy[kc * 4 + e];
          umma::split_tf32(dy[kc * 4 + e], hi[e], lo[e]);
        }
        const uint32_t off = (uint32_t)kc * kALbo + (uint32_t)j * 16u;
        *reinterpret_cast<float4*>(st + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(st + 16384 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    const float sc1 = v1[j], sh1 = v1[FH + j];
    const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, 128);
    for (int p = 0; p < 2; ++p) {
      if (32 * 2 * p >= npad) break;
      if (p == 1) {                                        // the y1^T stages are reused: the first pass's products must be done
        umma::mbar_wait(&bar_empty[0], 0);
        umma::fence_after_sync();
      }
      const int sl = 2 * p + (part >> 1), r0 = 32 * sl + 16 * (part & 1);      // this thread: rows r0 .. r0 + 15 of slice sl
      unsigned char* st = sRing + (part >> 1) * kStage;
      float uv[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) uv[e] = r0 + e < B ? ro_input1(c, h, r0 + e, j, sPerm) : 0.f;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = r0 + kc * 4 + e < B ? fmaf(uv[kc * 4 + e], sc1, sh1) : 0.f;
          umma::split_tf32(v, hi[e], lo[e]);
        }
        const uint32_t off = (uint32_t)((part & 1) * 4 + kc) * kALbo + (uint32_t)j * 16u;
        *reinterpret_cast<float4*>(st + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(st + 16384 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      umma::fence_async_smem();
      __syncthreads();
      if (t == 0) {
        umma::fence_after_sync();
        for (int g2 = 0; g2 < 2; ++g2) {
          const int s2 = 2 * p + g2;
          if (32 * s2 >= npad) break;
          const uint32_t abase = umma::smem_addr(sOp + s2 * 32768), bbase = umma::smem_addr(sRing + g2 * kStage);
          const int ksteps = imin(4, (npad - 32 * s2) / 8);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t o = (uint32_t)ks * 2u * kALbo;
            const uint64_t ah = umma::smem_desc(abase + o, kALbo, kASbo), al = umma::smem_desc(abase + 16384u + o, kALbo, kASbo);
            const uint64_t bh = umma::smem_desc(bbase + o, kALbo, kASbo), bl = umma::smem_desc(bbase + 16384u + o, kALbo, kASbo);
            const uint32_t first = (s2 == 0 && ks == 0) ? 0u : 1u;
            umma::mma_tf32(tmem, al, bh, idesc, first);
            umma::mma_tf32(tmem, ah, bl, idesc, 1u);
            umma::mma_tf32(tmem, ah, bh, idesc, 1u);
          }
        }
        if (p == 0 && 64 < npad) umma::commit(&bar_empty[0]);
      }
    }
    if (t == 0) umma::commit(&bar_mma);
    // d b1: the four row parts of a channel, fixed order
    float* sDb = sStF;
    sDb[part * FH + j] = db1;
    FSG_T(3);                                            // 3: operand slices + issue (both passes)
    umma::mbar_wait(&bar_mma, 0);
    umma::fence_after_sync();
    __syncthreads();
    FSG_T(4);                                            // 4: product tail
    if (t < FH) c.grads[c.po.fc1_b[h] + t] = (sDb[t] + sDb[FH + t]) + (sDb[2 * FH + t] + sDb[3 * FH + t]);
    // d W1 from TMEM: 16 warps = 4 lane quarters x 4 column groups, transposed through a private scratch
    {
      float* sw = reinterpret_cast<float*>(sOp) + warp * 32 * 33;     // (the operand slices are idle)
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
  if (blockIdx.x == 0) CAL_TL(c.status, 7);
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
  auto go = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    launch_k(kern, dim3(3), dim3(RT), smem, s, c);
    return 0;
  };
  const int rc = c.C == 2 ? go(k_ro_fwd<2>) : c.C == 3 ? go(k_ro_fwd<3>) : c.C == 4 ? go(k_ro_fwd<4>) : go(k_ro_fwd<8>);
  if (rc) return rc;
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_readout_ro_backward(const Ctx& c, cudaStream_t s) {
  const size_t smem = ro_bsmem().total;
  auto go = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    launch_k(kern, dim3(6), dim3(RT), smem, s, c);
    return 0;
  };
  const int rc = c.C == 2 ? go(k_ro_bwd<2>) : c.C == 3 ? go(k_ro_bwd<3>) : c.C == 4 ? go(k_ro_bwd<4>) : go(k_ro_bwd<8>);
  if (rc) return rc;
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
