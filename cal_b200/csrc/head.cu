// head.cu -- global_add_pool, the random-intervention mix, the three readout MLPs, the loss, and
// their backward passes (model.py:115-164, train_causal.py:178-186).
//
//   pool   : xc_g[b] = sum_{batch_n = b} zc[n], xo_g likewise (CTA per graph, fixed order, no atomics)
//   readout: u_c = xc_g, u_o = xo_g, u_co = xc_g[perm] (+ or ||) xo_g;
//            logp = log_softmax(fc2(bn2(relu(fc1(bn1(u)))))); KL(uniform || c) batchmean, NLL(o), NLL(co)
//
// One thread-block CLUSTER of 8 CTAs per readout head.  CTA j of a cluster owns a 1/8 COLUMN slice
// and all B graph rows of it, so every BatchNorm over the B rows (bn1 on the input slice, bn2 on the
// hidden slice) and every bias / BatchNorm / fc2 / fc1 weight-gradient reduction over rows is local
// to one CTA -- no cross-CTA partial sums, no serial tails.  The two places where columns mix (fc1:
// K split over the input slices; fc1 backward: K split over the hidden slices) exchange their
// [rows x H] partial products through distributed shared memory (reduce-scatter by column slice).
// Parameter gradients of the readout are written straight into the flat gradient buffer.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "internal.cuh"

namespace cg = cooperative_groups;

namespace cal {

namespace {

constexpr int kRC = 8;            // CTAs per cluster = column slices
constexpr int kChunk = 128;       // graph rows per GEMM / exchange chunk

__device__ __forceinline__ int clampB(const Ctx& c) { return imin(imax(c.dims[2], 0), c.Bm); }

template <int VEC>
__global__ void __launch_bounds__(256) k_pool(const Ctx c) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  __shared__ float sRed[2][kRowWarps][H];
  const int B = clampB(c);
  const int b = blockIdx.x;
  if (b >= B) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = c.graph_ptr[b], n1 = c.graph_ptr[b + 1];
  const float* Zc = c.Z;
  const float* Zo = c.Z + (size_t)c.Nm * H;
  RowVec<VEC> ac, ao;
  ac.zero();
  ao.zero();
  for (int n = n0 + warp; n < n1; n += kRowWarps) {
    RowVec<VEC> vc, vo;
    vc.load_coherent(Zc + (size_t)n * H, lane);
    vo.load_coherent(Zo + (size_t)n * H, lane);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      ac.v[i] += vc.v[i];
      ao.v[i] += vo.v[i];
    }
  }
  ac.store(&sRed[0][warp][0], lane);
  ao.store(&sRed[1][warp][0], lane);
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * H; t += blockDim.x) {
    const int br = t / H, k = t - br * H;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) s += sRed[br][w][k];
    c.pooled[((size_t)br * c.Bm + b) * H + k] = s;
  }
}

// readout input u_h[b][k], k < K1
__device__ __forceinline__ float head_input(const Ctx& c, int h, int b, int k, int H) {
  const float* gc = c.pooled;
  const float* go = c.pooled + (size_t)c.Bm * H;
  if (h == 0) return gc[(size_t)b * H + k];
  if (h == 1) return go[(size_t)b * H + k];
  if (c.cat) return k < H ? gc[(size_t)c.perm[b] * H + k] : go[(size_t)b * H + (k - H)];
  return gc[(size_t)c.perm[b] * H + k] + go[(size_t)b * H + k];
}

// Column sums over the B rows of two per-element quantities, for `ncols` (4, 8, 16 or 32) columns,
// in fp64.  f(b, col, v0, v1) yields the two values.  256 threads = ncols columns x (256 / ncols)
// row parts; fixed order => deterministic.  Results in out0[col], out1[col] (shared memory).
template <typename F>
__device__ __forceinline__ void column_sums2(int B, int ncols, double* scratch /*[2][256]*/, double* out0,
                                             double* out1, F f) {
  const int t = threadIdx.x;
  const int nparts = 256 / ncols;
  const int col = t % ncols, part = t / ncols;
  double s0 = 0.0, s1 = 0.0;
  for (int b = part; b < B; b += nparts) {
    float v0, v1;
    f(b, col, v0, v1);
    s0 += (double)v0;
    s1 += (double)v1;
  }
  __syncthreads();
  scratch[t] = s0;
  scratch[256 + t] = s1;
  __syncthreads();
  if (t < ncols) {
    double a = 0.0, bq = 0.0;
    for (int p = 0; p < nparts; ++p) {
      a += scratch[p * ncols + t];
      bq += scratch[256 + p * ncols + t];
    }
    out0[t] = a;
    out1[t] = bq;
  }
  __syncthreads();
}

struct ReadoutSmem {
  size_t off_u, off_w, off_p, off_h, off_x, off_du, off_lg, off_w2, off_misc, total;   // bytes
};
// Shared-memory layout of the readout kernels for Bp rows (multiple of kChunk), hidden H, KSm =
// widest input slice, C classes.
__host__ __device__ inline ReadoutSmem readout_smem(int Bp, int H, int KSm, int C, bool backward) {
  const int HS = H / kRC;
  ReadoutSmem s;
  size_t o = 0;
  s.off_u = o;    o += (size_t)Bp * KSm * 4;                            // input slice (raw / normalised)
  s.off_w = o;    o += (size_t)(backward ? HS * 2 * H : KSm * H) * 4;   // fc1 weight slice
  s.off_p = o;    o += (size_t)imax(kChunk * H, Bp * C) * 4;            // partial products / partial logits / d logits
  s.off_h = o;    o += (size_t)Bp * HS * 4;                             // hidden slice
  s.off_x = o;    o += backward ? (size_t)Bp * HS * 4 : 0;              // backward: u = d(fc1 out) slice
  s.off_du = o;   o += backward ? (size_t)Bp * KSm * 4 : 0;             // backward: gathered d(bn1 out) slice
  s.off_lg = o;   o += backward ? 0 : (size_t)Bp * C * 4;               // forward: summed logits (CTA 0)
  s.off_w2 = o;   o += (size_t)C * HS * 4;                              // fc2 weight slice [C][HS]
  o = (o + 15) & ~(size_t)15;
  s.off_misc = o; o += 2 * 256 * 8 + 256 * 8 + 512 * 4;                 // fp64 scratch | fp64 sums [4][64] | float consts [16][32]
  s.total = o;
  return s;
}

// training-mode BatchNorm over the B local rows of `ncols` columns (global column k0 + t):
// scale / shift into sc / sh, the record into the workspace, running statistics updated.
// gamma / beta / old running statistics were prefetched into shared memory (g, be, rm0, rv0).
__device__ __forceinline__ void bn_local_finalize(const Ctx& c, int id, int k0, int ncols, int B, const double* sum,
                                                  const double* sq, const float* g, const float* be,
                                                  const float* rm0, const float* rv0, float* sc, float* sh) {
  const int t = threadIdx.x;
  if (t < ncols) {
    const int k = k0 + t;
    double mean = B > 0 ? sum[t] / B : 0.0, var = B > 0 ? sq[t] / B - mean * mean : 0.0;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
    const float s = g[t] * rstd;
    const float o = be[t] - (float)mean * s;
    sc[t] = s;
    sh[t] = o;
    c.bnf(id, BN_SCALE)[k] = s;
    c.bnf(id, BN_SHIFT)[k] = o;
    c.bnf(id, BN_MEAN)[k] = (float)mean;
    c.bnf(id, BN_RSTD)[k] = rstd;
    if (c.bn_buffers != nullptr && c.bn_rm[id] >= 0) {
      const double unb = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
      c.bn_buffers[c.bn_rm[id] + k] = (1.f - c.momentum) * rm0[t] + c.momentum * (float)mean;
      c.bn_buffers[c.bn_rv[id] + k] = (1.f - c.momentum) * rv0[t] + c.momentum * (float)unb;
    }
  }
}

// loads `n` elements f(idx) -> dst[idx] with 8 independent loads in flight per thread
template <typename F>
__device__ __forceinline__ void load_batched(float* dst, int n, F f) {
  for (int base = 0; base < n; base += 8 * 256) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + (int)threadIdx.x + 256 * u;
      v[u] = idx < n ? f(idx) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + (int)threadIdx.x + 256 * u;
      if (idx < n) dst[idx] = v[u];
    }
  }
}

// deterministic block sum of one float per thread (256 threads): xor-shuffle tree inside each warp,
// then the 8 warp totals in order.  Result valid in thread 0.  sbuf: 8 floats of shared memory.
__device__ __forceinline__ float block_sum256(float v, float* sbuf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sbuf[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) s += sbuf[w];
  return s;
}

// ---------------------------------------------------------------------------------------------
// Readout forward.  grid (8, 3), cluster (8, 1, 1): blockIdx.y = head (c, o, co), blockIdx.x = slice.
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __cluster_dims__(kRC, 1, 1) __launch_bounds__(256) k_readout_fwd(const Ctx c) {
  constexpr int H = 32 * VEC, HS = H / kRC, CT = H / 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cl = cg::this_cluster();
  const int B = clampB(c);
  const int h = blockIdx.y, j = blockIdx.x;
  const int K1 = (h == 2 && c.cat) ? 2 * H : H, KS = K1 / kRC;
  const int Bp = ceil_div(imax(c.Bm, 1), kChunk) * kChunk;
  const ReadoutSmem L = readout_smem(Bp, H, c.cat ? 2 * HS : HS, c.C, false);
  float* sU = reinterpret_cast<float*>(smem_raw + L.off_u);
  float* sW = reinterpret_cast<float*>(smem_raw + L.off_w);
  float* sP = reinterpret_cast<float*>(smem_raw + L.off_p);
  float* sH = reinterpret_cast<float*>(smem_raw + L.off_h);
  float* sLg = reinterpret_cast<float*>(smem_raw + L.off_lg);
  double* scratch = reinterpret_cast<double*>(smem_raw + L.off_misc);
  double* sum = scratch + 512;                        // [64]
  double* sq = sum + 64;                              // [64]
  float* sW2 = reinterpret_cast<float*>(smem_raw + L.off_w2);   // [C][HS] fc2 weight slice
  float* fc = reinterpret_cast<float*>(scratch + 512 + 256);
  float *sc1 = fc, *sh1 = fc + 32, *sc2 = fc + 64, *sh2 = fc + 96;          // [32] each
  float *g1 = fc + 128, *be1 = fc + 160, *g2 = fc + 192, *be2 = fc + 224;   // BatchNorm weights of the slices
  float *b1s = fc + 256, *rm1 = fc + 288, *rv1 = fc + 320, *rm2 = fc + 352, *rv2 = fc + 384, *b2s = fc + 416;
  const int t = threadIdx.x;
  const int k0 = j * KS, m0 = j * HS;
  const int bn1 = c.L + 3 + h, bn2 = c.L + 6 + h;
  const int C = c.C;

  // ---- before the dependency wait: everything that only reads parameters / BatchNorm buffers ----
  // fc1 weight slice, transposed copy [K1][H] made by k_param_prep: rows k0 .. k0 + KS
  PT_DECL
  stage_matrix_async(sW, c.wt_fc1(h) + (size_t)k0 * H, KS * H);
  if (t < KS) {
    g1[t] = c.params[c.bn_gamma[bn1] + k0 + t];
    be1[t] = c.params[c.bn_beta[bn1] + k0 + t];
    if (c.bn_buffers != nullptr && c.bn_rm[bn1] >= 0) {
      rm1[t] = c.bn_buffers[c.bn_rm[bn1] + k0 + t];
      rv1[t] = c.bn_buffers[c.bn_rv[bn1] + k0 + t];
    }
  } else if (t >= 32 && t < 32 + HS) {
    const int m = t - 32;
    g2[m] = c.params[c.bn_gamma[bn2] + m0 + m];
    be2[m] = c.params[c.bn_beta[bn2] + m0 + m];
    b1s[m] = c.params[c.po.fc1_b[h] + m0 + m];
    if (c.bn_buffers != nullptr && c.bn_rm[bn2] >= 0) {
      rm2[m] = c.bn_buffers[c.bn_rm[bn2] + m0 + m];
      rv2[m] = c.bn_buffers[c.bn_rv[bn2] + m0 + m];
    }
  } else if (t >= 64 && t < 64 + C) {
    b2s[t - 64] = c.params[c.po.fc2_b[h] + t - 64];
  }
  for (int i = t; i < C * HS; i += 256) sW2[i] = c.params[c.po.fc2_w[h] + (size_t)(i / HS) * H + m0 + (i % HS)];
  long long y0 = -1, y1 = -1;                          // labels of the rows CTA 0 scores (input data)
  if (j == 0 && c.with_loss && c.y != nullptr) {
    if (t < B) y0 = c.y[t];
    if (t + 256 < B) y1 = c.y[t + 256];
  }
  pdl_sync();
  PT_MARK();                                           // 0: dependency wait

  // ---- input slice + bn1 (all B rows are here: the statistics are local) ----
  load_batched(sU, Bp * KS, [&](int i) {
    const int b = i / KS, k = i - b * KS;
    return b < B ? head_input(c, h, b, k0 + k, H) : 0.f;
  });
  __syncthreads();
  PT_MARK();                                           // 1: input slice loaded
  if (c.train) {
    column_sums2(B, KS, scratch, sum, sq, [&](int b, int col, float& v0, float& v1) {
      const float v = sU[b * KS + col];
      v0 = v;
      v1 = v * v;
    });
    bn_local_finalize(c, bn1, k0, KS, B, sum, sq, g1, be1, rm1, rv1, sc1, sh1);
    if (j == 0 && t == 0 && c.nbt != nullptr) {
      c.nbt[bn1] += 1;
      c.nbt[bn2] += 1;
    }
  } else if (t < KS) {
    sc1[t] = c.bnf(bn1, BN_SCALE)[k0 + t];
    sh1[t] = c.bnf(bn1, BN_SHIFT)[k0 + t];
  }
  __syncthreads();
  for (int i = t; i < Bp * KS; i += 256) {
    const int b = i / KS, k = i - b * KS;
    sU[i] = b < B ? fmaf(sU[i], sc1[k], sh1[k]) : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  PT_MARK();                                           // 2: bn1 + normalise + W

  // ---- fc1: K split over the input slices; reduce-scatter of the partial products by hidden slice ----
  float* H1 = c.H1 + (size_t)h * c.Bm * H;
  const int tx = t & 15, ty = t >> 4;
  // thread (ty, tx): rows ty*8 .. +8 of the chunk; columns in NG groups of GW: group g starts at
  // g * (H / NG) + tx * GW, so the 16 lanes of a half-warp read one contiguous run per group
  constexpr int GW = CT >= 4 ? 4 : CT, NG = CT / GW;
  for (int r0 = 0; r0 < B; r0 += kChunk) {
    float acc[8][CT];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int q = 0; q < CT; ++q) acc[i][q] = 0.f;
#pragma unroll 4
    for (int k = 0; k < KS; ++k) {
      float a[8], w[CT];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = sU[(size_t)(r0 + ty * 8 + i) * KS + k];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float* wp = sW + (size_t)k * H + g * (H / NG) + tx * GW;
        if constexpr (GW == 4) {
          const float4 v = *reinterpret_cast<const float4*>(wp);
          w[g * 4] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < GW; ++e) w[g * GW + e] = wp[e];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < CT; ++q) acc[i][q] = fmaf(a[i], w[q], acc[i][q]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        float* pp = sP + (size_t)(ty * 8 + i) * H + g * (H / NG) + tx * GW;
        if constexpr (GW == 4) {
          *reinterpret_cast<float4*>(pp) = make_float4(acc[i][g * 4], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < GW; ++e) pp[e] = acc[i][g * GW + e];
        }
      }
    PT_MARK();                                         // 3: partial GEMM
    cl.sync();
    PT_MARK();                                         // 4: cluster sync
    {
      constexpr int F4 = HS / 4;                       // float4 per row of the slice
      const float* rp[kRC];
#pragma unroll
      for (int q = 0; q < kRC; ++q) rp[q] = cl.map_shared_rank(sP, q);
      constexpr int NI = (kChunk * F4 + 255) / 256;    // items per thread: every remote load in flight at once
      float4 vv[NI][kRC];
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int i = t + 256 * u;
        const int r = i / F4, m = (i - r * F4) * 4;
#pragma unroll
        for (int q = 0; q < kRC; ++q)
          vv[u][q] = i < kChunk * F4 ? *reinterpret_cast<const float4*>(rp[q] + (size_t)r * H + m0 + m)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < NI; ++u) {
        const int i = t + 256 * u;
        if (i >= kChunk * F4) break;
        const int r = i / F4, m = (i - r * F4) * 4;
        float4 s = vv[u][0];
#pragma unroll
        for (int q = 1; q < kRC; ++q) {
          s.x += vv[u][q].x; s.y += vv[u][q].y; s.z += vv[u][q].z; s.w += vv[u][q].w;
        }
        const int b = r0 + r;
        const float4 bb = *reinterpret_cast<const float4*>(b1s + m);
        float4 o = make_float4(fmaxf(s.x + bb.x, 0.f), fmaxf(s.y + bb.y, 0.f), fmaxf(s.z + bb.z, 0.f),
                               fmaxf(s.w + bb.w, 0.f));
        if (b >= B) o = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sH + (size_t)b * HS + m) = o;
        if (b < B) *reinterpret_cast<float4*>(H1 + (size_t)b * H + m0 + m) = o;
      }
    }
    cl.sync();
  }

  PT_MARK();                                           // 5: reduce-scatter + sync
  // ---- bn2 on the hidden slice (local) ----
  if (c.train) {
    column_sums2(B, HS, scratch, sum, sq, [&](int b, int col, float& v0, float& v1) {
      const float v = sH[b * HS + col];
      v0 = v;
      v1 = v * v;
    });
    bn_local_finalize(c, bn2, m0, HS, B, sum, sq, g2, be2, rm2, rv2, sc2, sh2);
  } else if (t < HS) {
    sc2[t] = c.bnf(bn2, BN_SCALE)[m0 + t];
    sh2[t] = c.bnf(bn2, BN_SHIFT)[m0 + t];
  }
  __syncthreads();

  PT_MARK();                                           // 6: bn2
  // ---- fc2: partial logits of the slice, summed over the cluster by CTA 0 ----
  float* sL = sP;                                      // [B][C]
  for (int i = t; i < B * C; i += 256) {
    const int b = i / C, cls = i - b * C;
    float s = 0.f;
#pragma unroll
    for (int m = 0; m < HS; ++m) s = fmaf(fmaf(sH[(size_t)b * HS + m], sc2[m], sh2[m]), sW2[cls * HS + m], s);
    sL[i] = s;
  }
  cl.sync();
  PT_MARK();                                           // 7: partial logits + sync
  if (j == 0) {
    const float* rl[kRC];
#pragma unroll
    for (int q = 0; q < kRC; ++q) rl[q] = cl.map_shared_rank(sL, q);
    for (int i = t; i < B * C; i += 256) {
      float v[kRC];
#pragma unroll
      for (int q = 0; q < kRC; ++q) v[q] = rl[q][i];
      float s = v[0];
#pragma unroll
      for (int q = 1; q < kRC; ++q) s += v[q];
      sLg[i] = s + b2s[i % C];
    }
    __syncthreads();
    float loss_part = 0.f, correct_part = 0.f;
    for (int b = t; b < B; b += 256) {
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
      float slp = 0.f, picked = 0.f;
      const long long yb = b == t ? y0 : (b == t + 256 ? y1 : ((c.with_loss && c.y != nullptr) ? c.y[b] : -1));
      for (int cls = 0; cls < C; ++cls) {
        const float lp = sLg[b * C + cls] - m - lse;
        c.logp[((size_t)h * c.Bm + b) * C + cls] = (c.raw_o && h == 1) ? sLg[b * C + cls] : lp;   // (model.py:288-289: raw logits)
        slp += lp;
        if ((long long)cls == yb) picked = lp;
      }
      if (c.with_loss) {
        loss_part += h == 0 ? -logf((float)C) - slp / (float)C : -picked;   // KL(uniform || .) row / NLL row
        correct_part += (long long)am == yb ? 1.f : 0.f;
      }
    }
    if (c.with_loss) {
      float* red = reinterpret_cast<float*>(scratch);
      const float ls = block_sum256(loss_part, red);
      const float cs = block_sum256(correct_part, red + 8);
      if (t == 0) {
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
  }
  PT_MARK();                                           // 8: CTA 0 tail (logits, log-softmax, loss)
  cl.sync();                                           // remote partial logits stay alive until CTA 0 is done
  PT_MARK();                                           // 9: final cluster sync
  PT_DUMP(c, 64);
}

// ---------------------------------------------------------------------------------------------
// Readout backward.  Same grid / cluster.  CTA j owns hidden slice m0 .. m0+HS (fc2, bn2, fc1 bias,
// rows of the fc1 weight) and input slice k0 .. k0+KS (bn1, columns of d fc1 weight, d input).
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __cluster_dims__(kRC, 1, 1) __launch_bounds__(256) k_readout_bwd(const Ctx c) {
  constexpr int H = 32 * VEC, HS = H / kRC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cl = cg::this_cluster();
  const int B = clampB(c);
  const int h = blockIdx.y, j = blockIdx.x;
  const int K1 = (h == 2 && c.cat) ? 2 * H : H, KS = K1 / kRC;
  const int Bp = ceil_div(imax(c.Bm, 1), kChunk) * kChunk;
  const ReadoutSmem L = readout_smem(Bp, H, c.cat ? 2 * HS : HS, c.C, true);
  float* sUin = reinterpret_cast<float*>(smem_raw + L.off_u);     // [Bp][KS] raw input slice
  float* sW = reinterpret_cast<float*>(smem_raw + L.off_w);       // [HS][K1] rows m0.. of the fc1 weight
  float* sP = reinterpret_cast<float*>(smem_raw + L.off_p);       // d logits / partial products / all-gathered u
  float* sH = reinterpret_cast<float*>(smem_raw + L.off_h);       // [Bp][HS] h1 slice
  float* sX = reinterpret_cast<float*>(smem_raw + L.off_x);       // [Bp][HS] d h2 -> u slice
  float* sDU = reinterpret_cast<float*>(smem_raw + L.off_du);     // [Bp][KS] d(bn1 out) slice
  double* scratch = reinterpret_cast<double*>(smem_raw + L.off_misc);
  double* s1 = scratch + 512;                         // [64]
  double* s2 = s1 + 64;                               // [64]
  float* fc = reinterpret_cast<float*>(scratch + 512 + 256);
  float *f_sc2 = fc, *f_sh2 = fc + 32, *f_mean2 = fc + 64, *f_rstd2 = fc + 96;       // hidden slice constants
  float *f_sc1 = fc + 128, *f_sh1 = fc + 160, *f_mean1 = fc + 192, *f_rstd1 = fc + 224;   // input slice constants
  float *c1 = fc + 256, *c2 = fc + 288;
  float* sW2 = reinterpret_cast<float*>(smem_raw + L.off_w2);     // [C][HS] fc2 weight slice
  const int t = threadIdx.x;
  const int k0 = j * KS, m0 = j * HS;
  const int bn1 = c.L + 3 + h, bn2 = c.L + 6 + h;
  const int C = c.C;

  PT_DECL
  stage_matrix_async(sW, c.params + c.po.fc1_w[h] + (size_t)m0 * K1, HS * K1);
  for (int i = t; i < C * HS; i += 256) sW2[i] = c.params[c.po.fc2_w[h] + (size_t)(i / HS) * H + m0 + (i % HS)];
  pdl_sync();
  PT_MARK();                                           // 0: dependency wait

  // ---- d logits for all rows (every CTA: B * C values) ----
  float* sDl = sP;
  for (int b = t; b < B; b += 256) {
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
    if (!(c.raw_o && h == 1))                          // (raw logits: the caller's gradient already is d logits)
      for (int cls = 0; cls < C; ++cls) {
        const float lp = c.logp[((size_t)h * c.Bm + b) * C + cls];
        sDl[b * C + cls] -= expf(lp) * sd;
      }
  }
  // ---- hidden slice, input slice, their BatchNorm records ----
  const float* H1 = c.H1 + (size_t)h * c.Bm * H;
  load_batched(sH, Bp * HS, [&](int i) {
    const int b = i / HS, m = i - b * HS;
    return b < B ? H1[(size_t)b * H + m0 + m] : 0.f;
  });
  load_batched(sUin, Bp * KS, [&](int i) {
    const int b = i / KS, k = i - b * KS;
    return b < B ? head_input(c, h, b, k0 + k, H) : 0.f;
  });
  if (t < HS) {
    f_sc2[t] = c.bnf(bn2, BN_SCALE)[m0 + t];
    f_sh2[t] = c.bnf(bn2, BN_SHIFT)[m0 + t];
    f_mean2[t] = c.bnf(bn2, BN_MEAN)[m0 + t];
    f_rstd2[t] = c.bnf(bn2, BN_RSTD)[m0 + t];
  }
  if (t < KS) {
    f_sc1[t] = c.bnf(bn1, BN_SCALE)[k0 + t];
    f_sh1[t] = c.bnf(bn1, BN_SHIFT)[k0 + t];
    f_mean1[t] = c.bnf(bn1, BN_MEAN)[k0 + t];
    f_rstd1[t] = c.bnf(bn1, BN_RSTD)[k0 + t];
  }
  __syncthreads();
  PT_MARK();                                           // 1: d logits, slices, constants

  // ---- fc2 backward on the slice: d W2[:, slice], d b2, d h2 = dl W2 ----
  // four lanes share one (class, hidden column) dot product over the B rows, combined by a fixed
  // xor-shuffle tree (deterministic); the trip count is uniform so every lane reaches the shuffles
  for (int base = 0; base < C * HS; base += 64) {
    const int task = base + (t >> 2), part = t & 3;
    const bool live = task < C * HS;
    const int cls = live ? task / HS : 0, m = live ? task - cls * HS : 0;
    float s = 0.f;
    if (live) {
#pragma unroll 4
      for (int b = part; b < B; b += 4) s = fmaf(sDl[b * C + cls], fmaf(sH[b * HS + m], f_sc2[m], f_sh2[m]), s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (live && part == 0) c.grads[c.po.fc2_w[h] + (size_t)cls * H + m0 + m] = s;
  }
  if (j == 0) {                                        // d b2: one warp per class, lanes over the rows
    const int warp = t >> 5, lane = t & 31;
    for (int cls = warp; cls < C; cls += 8) {
      float s = 0.f;
      for (int b = lane; b < B; b += 32) s += sDl[b * C + cls];
      s = warp_sum(s);
      if (lane == 0) c.grads[c.po.fc2_b[h] + cls] = s;
    }
  }
  for (int i = t; i < Bp * HS; i += 256) {
    const int b = i / HS, m = i - b * HS;
    float s = 0.f;
    if (b < B)
      for (int cls = 0; cls < C; ++cls) s = fmaf(sDl[b * C + cls], sW2[cls * HS + m], s);
    sX[i] = s;
  }
  __syncthreads();
  PT_MARK();                                           // 2: fc2 backward
  // ---- bn2 backward (local) -> u = relu'(h1) * bn2'(d h2);  d gamma2, d beta2, d b1 ----
  column_sums2(B, HS, scratch, s1, s2, [&](int b, int col, float& v0, float& v1) {
    const float dy = sX[b * HS + col];
    v0 = dy;
    v1 = dy * ((sH[b * HS + col] - f_mean2[col]) * f_rstd2[col]);
  });
  if (t < HS) {
    const double inv = B > 0 ? 1.0 / B : 0.0;
    c1[t] = (float)(s1[t] * inv);
    c2[t] = (float)(s2[t] * inv);
    c.grads[c.bn_gamma[bn2] + m0 + t] = (float)s2[t];
    c.grads[c.bn_beta[bn2] + m0 + t] = (float)s1[t];
  }
  __syncthreads();
  for (int i = t; i < Bp * HS; i += 256) {
    const int b = i / HS, m = i - b * HS;
    float u = 0.f;
    if (b < B) {
      const float x = sH[i];
      const float xh = (x - f_mean2[m]) * f_rstd2[m];
      u = x > 0.f ? f_sc2[m] * (sX[i] - c1[m] - xh * c2[m]) : 0.f;
    }
    sX[i] = u;
  }
  __syncthreads();
  column_sums2(B, HS, scratch, s1, s2, [&](int b, int col, float& v0, float& v1) {
    v0 = sX[b * HS + col];
    v1 = 0.f;
  });
  if (t < HS) c.grads[c.po.fc1_b[h] + m0 + t] = (float)s1[t];
  cp_async_wait_all();
  __syncthreads();
  PT_MARK();                                           // 3: bn2 backward, u, d b1

  // ---- fc1 backward, d input: K split over the hidden slices, reduce-scatter by input slice ----
  const int rchunk = K1 > H ? kChunk / 2 : kChunk;     // rows per exchange chunk (sP holds rchunk x K1 floats)
  const int tx = t & 15, ty = t >> 4;
  const int rt = rchunk / 16, ctb = K1 / 16;           // rows / columns per thread
  for (int r0 = 0; r0 < B; r0 += rchunk) {
    if (ctb >= 4) {
      // column groups of 4: group g starts at g * (K1 / ng) + tx * 4 (conflict-free float4 per half-warp)
      const int ng = ctb / 4;
      for (int g = 0; g < ng; ++g) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
        const float* wp = sW + g * (K1 / ng) + tx * 4;
#pragma unroll 4
        for (int m = 0; m < HS; ++m) {
          float a[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = i < rt ? sX[(size_t)(r0 + ty * rt + i) * HS + m] : 0.f;
          const float4 w = *reinterpret_cast<const float4*>(wp + (size_t)m * K1);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(a[i], w.x, acc[i][0]);
            acc[i][1] = fmaf(a[i], w.y, acc[i][1]);
            acc[i][2] = fmaf(a[i], w.z, acc[i][2]);
            acc[i][3] = fmaf(a[i], w.w, acc[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i < rt)
            *reinterpret_cast<float4*>(sP + (size_t)(ty * rt + i) * K1 + g * (K1 / ng) + tx * 4) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    } else {
      float acc[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.f;
      for (int m = 0; m < HS; ++m) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = i < rt ? sX[(size_t)(r0 + ty * rt + i) * HS + m] : 0.f;
#pragma unroll
          for (int q = 0; q < 2; ++q)
            if (q < ctb) acc[i][q] = fmaf(a, sW[(size_t)m * K1 + tx * ctb + q], acc[i][q]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (i < rt && q < ctb) sP[(size_t)(ty * rt + i) * K1 + tx * ctb + q] = acc[i][q];
    }
    cl.sync();
    {
      const int F4 = KS / 4;
      const float* rp[kRC];
#pragma unroll
      for (int q = 0; q < kRC; ++q) rp[q] = cl.map_shared_rank(sP, q);
      for (int i = t; i < rchunk * F4; i += 256) {
        const int r = i / F4, k = (i - r * F4) * 4;
        float4 v[kRC];
#pragma unroll
        for (int q = 0; q < kRC; ++q) v[q] = *reinterpret_cast<const float4*>(rp[q] + (size_t)r * K1 + k0 + k);
        float4 sm = v[0];
#pragma unroll
        for (int q = 1; q < kRC; ++q) {
          sm.x += v[q].x; sm.y += v[q].y; sm.z += v[q].z; sm.w += v[q].w;
        }
        if (r0 + r >= B) sm = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sDU + (size_t)(r0 + r) * KS + k) = sm;
      }
    }
    cl.sync();
  }
  PT_MARK();                                           // 4: d input GEMM + exchange
  // ---- bn1 backward on the input slice (local): d gamma1, d beta1, d input ----
  column_sums2(B, KS, scratch, s1, s2, [&](int b, int col, float& v0, float& v1) {
    const float dy = sDU[b * KS + col];
    v0 = dy;
    v1 = dy * ((sUin[b * KS + col] - f_mean1[col]) * f_rstd1[col]);
  });
  if (t < KS) {
    const double inv = B > 0 ? 1.0 / B : 0.0;
    c1[t] = (float)(s1[t] * inv);
    c2[t] = (float)(s2[t] * inv);
    c.grads[c.bn_gamma[bn1] + k0 + t] = (float)s2[t];
    c.grads[c.bn_beta[bn1] + k0 + t] = (float)s1[t];
  }
  __syncthreads();
  for (int i = t; i < B * KS; i += 256) {
    const int b = i / KS, k = i - b * KS;
    const float xh = (sUin[i] - f_mean1[k]) * f_rstd1[k];
    c.du[((size_t)h * c.Bm + b) * 2 * H + k0 + k] = f_sc1[k] * (sDU[i] - c1[k] - xh * c2[k]);
  }
  PT_MARK();                                           // 5: bn1 backward + d input store
  // ---- d W1[:, input slice] = sum_b u[b][:] (x) y1[b][slice]: all-gather u by 128-row chunks ----
  {
    // thread owns input column ki and the NA = H * KS / 256 contiguous hidden rows kgrp * NA ..
    const int ki = t % KS, kgrp = t / KS;
    const int NA = (H * KS) / 256;                     // 8 (H=128), 16 (H=128, cat), 2 / 4 (H=64); 0 for H=32
    float acc[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) acc[a] = 0.f;
    float* sUall = sP;                                 // [kChunk][H]
    const int ng = 256 / KS;
    for (int r0 = 0; r0 < B; r0 += kChunk) {
      __syncthreads();
      {
        const float* rp[kRC];
#pragma unroll
        for (int q = 0; q < kRC; ++q) rp[q] = cl.map_shared_rank(sX, q);
        constexpr int F4 = H / 4;
        constexpr int NI = kChunk * F4 / 256;          // float4 per thread: all remote loads in flight at once
        float4 v[NI];
#pragma unroll
        for (int u = 0; u < NI; ++u) {
          const int i = t + 256 * u;
          const int r = i / F4, m = (i - r * F4) * 4;
          const int q = m / HS;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r0 + r < B) v[u] = *reinterpret_cast<const float4*>(rp[q] + (size_t)(r0 + r) * HS + (m - q * HS));
        }
#pragma unroll
        for (int u = 0; u < NI; ++u) {
          const int i = t + 256 * u;
          const int r = i / F4, m = (i - r * F4) * 4;
          *reinterpret_cast<float4*>(sUall + (size_t)r * H + m) = v[u];
        }
      }
      __syncthreads();
      const int rows = imin(kChunk, B - r0);
      if (NA >= 4) {
        const float* up = sUall + kgrp * NA;
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
          const float y1 = fmaf(sUin[(size_t)(r0 + r) * KS + ki], f_sc1[ki], f_sh1[ki]);
#pragma unroll
          for (int f = 0; f < 4; ++f)
            if (f * 4 < NA) {
              const float4 u = *reinterpret_cast<const float4*>(up + (size_t)r * H + f * 4);
              acc[f * 4] = fmaf(u.x, y1, acc[f * 4]);
              acc[f * 4 + 1] = fmaf(u.y, y1, acc[f * 4 + 1]);
              acc[f * 4 + 2] = fmaf(u.z, y1, acc[f * 4 + 2]);
              acc[f * 4 + 3] = fmaf(u.w, y1, acc[f * 4 + 3]);
            }
        }
      } else {
        for (int r = 0; r < rows; ++r) {
          const float y1 = fmaf(sUin[(size_t)(r0 + r) * KS + ki], f_sc1[ki], f_sh1[ki]);
#pragma unroll
          for (int a = 0; a < 16; ++a)
            if (kgrp + ng * a < H) acc[a] = fmaf(sUall[(size_t)r * H + kgrp + ng * a], y1, acc[a]);
        }
      }
    }
    if (NA >= 4) {
#pragma unroll
      for (int a = 0; a < 16; ++a)
        if (a < NA) c.grads[c.po.fc1_w[h] + (size_t)(kgrp * NA + a) * K1 + k0 + ki] = acc[a];
    } else {
#pragma unroll
      for (int a = 0; a < 16; ++a)
        if (kgrp + ng * a < H) c.grads[c.po.fc1_w[h] + (size_t)(kgrp + ng * a) * K1 + k0 + ki] = acc[a];
    }
  }
  PT_MARK();                                           // 6: d W1
  cl.sync();                                           // peers may still be reading this CTA's u slice
  PT_MARK();                                           // 7: final cluster sync
  PT_DUMP(c, 80);
}

template <typename K>
int set_smem_h(K kernel, size_t bytes) {
  if (bytes > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace

size_t readout_smem_bytes(int Bm, int H, int cat, int C, int backward) {
  const int Bp = ceil_div(imax(Bm, 1), kChunk) * kChunk;
  return readout_smem(Bp, H, (cat ? 2 : 1) * (H / kRC), C, backward != 0).total;
}

// Which readout kernels run: 3 = the short-chain kernels of the fused small-graph path (head_ro.cu: H = 128, B <= 128,
// "add", C <= 8; the default there), 0 = the fp32 FFMA cluster kernels below, 1 = the streaming tensor-core kernels
// (head_tc.cu, B <= 512), 2 = the resident-tile tensor-core kernels (head_tc2.cu: B <= 128, "add", C <= 8).
// Default: the tensor cores when bf16 operands are requested (cal_model_desc.readout_bf16) or forced
// (readout_tc, or CAL_READOUT=tc in the environment); otherwise the FFMA kernels -- at B = 128 the readout is
// a latency chain on three SMs, not a FLOP problem, and the 24-CTA cluster kernels measure faster
// (profiles/README.md).  CAL_READOUT=legacy forces them even for bf16 (A/B measurements).
static int readout_path(const Ctx& c) {
  static const int forced = [] {
    const char* e = getenv("CAL_READOUT");
    if (e == nullptr) return -1;
    if (strcmp(e, "legacy") == 0) return 0;
    if (strcmp(e, "tc") == 0) return 1;
    return -1;
  }();
  if (forced < 0 && readout_ro_supported(c)) return 3;  // fused small-graph path: the short-chain kernels of head_ro.cu
  const bool legacy_fits = readout_smem_bytes(c.Bm, c.H, c.cat, c.C, 1) <= 225 * 1024 &&
                           readout_smem_bytes(c.Bm, c.H, c.cat, c.C, 0) <= 225 * 1024;
  const bool want_tc = forced == 1 || (forced != 0 && (c.readout_bf16 || c.readout_tc)) || !legacy_fits;
  if (!want_tc) return 0;
  if (readout_tc2_supported(c)) return 2;
  if (readout_tc_supported(c)) return 1;
  return 0;
}

bool readout_runs_ro(const Ctx& c) { return readout_path(c) == 3; }
int readout_path_id(const Ctx& c) { return readout_path(c); }         // 0 = FFMA cluster kernels

int launch_heads_forward(const Ctx& c, int with_loss, cudaStream_t s) {
  (void)with_loss;
  const int path = readout_path(c);
  if (path == 3) return launch_readout_ro_forward(c, s);
  if (path != 0) {
    if (!c.fsg_on) {                                   // (the fused small-graph forward pools in its epilogue)
      CAL_DISPATCH_VEC(c.H, { launch_k(k_pool<VEC>, dim3(c.Bm), dim3(256), 0, s, c); });
      note_launches(1);
      CAL_CUDA_CHECK_LAUNCH();
    }
    return path == 2 ? launch_readout_tc2_forward(c, s) : launch_readout_tc_forward(c, s);
  }
  CAL_DISPATCH_VEC(c.H, {
    if (!c.fsg_on) launch_k(k_pool<VEC>, dim3(c.Bm), dim3(256), 0, s, c);
    const size_t smem = readout_smem_bytes(c.Bm, c.H, c.cat, c.C, 0);
    int rc = set_smem_h(k_readout_fwd<VEC>, smem);
    if (rc) return rc;
    launch_k(k_readout_fwd<VEC>, dim3(kRC, 3), dim3(256), smem, s, c);
  });
  note_launches(c.fsg_on ? 1 : 2);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_heads_backward(const Ctx& c, cudaStream_t s) {
  const int path = readout_path(c);
  if (path == 3) return launch_readout_ro_backward(c, s);
  if (path != 0) return path == 2 ? launch_readout_tc2_backward(c, s) : launch_readout_tc_backward(c, s);
  CAL_DISPATCH_VEC(c.H, {
    const size_t smem = readout_smem_bytes(c.Bm, c.H, c.cat, c.C, 1);
    int rc = set_smem_h(k_readout_bwd<VEC>, smem);
    if (rc) return rc;
    launch_k(k_readout_bwd<VEC>, dim3(kRC, 3), dim3(256), smem, s, c);
  });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
