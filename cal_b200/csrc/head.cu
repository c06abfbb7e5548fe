// head.cu -- global_add_pool, the random-intervention mix, the three readout MLPs, the loss, and
// their backward passes (model.py:115-164, train_causal.py:178-186).
//
//   pool : xc_g[b] = sum_{batch_n = b} zc[n], xo_g likewise (CTA per graph, fixed order, no atomics)
//   head1: u_c = xc_g, u_o = xo_g, u_co = xc_g[perm] (+ or ||) xo_g;  h1 = relu(fc1(bn1(u)))
//   head2: logp = log_softmax(fc2(bn2(h1)));  KL(uniform || c) batchmean, NLL(o), NLL(co)
// BatchNorm over the B graph rows: bn1 statistics are recomputed by every head1 CTA from the
// pooled rows (B x H values, L2 resident); bn2 statistics use the partial-sum / last-CTA scheme.
#include "internal.cuh"

namespace cal {

namespace {

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

// smem: sW [K1][H] | sA [R][K1] | sRed f64 [8][H] | sSc [K1] | sSh [K1] | sSum f64 [2][256]
template <int VEC>
__global__ void __launch_bounds__(256) k_head1_fwd(const Ctx c) {
  constexpr int H = 32 * VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int B = clampB(c);
  const int h = blockIdx.y, tile = blockIdx.x;
  const int row0 = tile * kTileRows;
  if (row0 >= B) return;
  const int K1 = (h == 2 && c.cat) ? 2 * H : H;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sA = sW + (size_t)K1 * H;
  double* sRed = reinterpret_cast<double*>(sA + (size_t)kTileRows * K1);
  float* sSc = reinterpret_cast<float*>(sRed + kRowWarps * H);
  float* sSh = sSc + K1;
  double* sSum = reinterpret_cast<double*>(sSh + K1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn1 = c.L + 3 + h;
  stage_matrix_async(sW, c.wt_fc1(h), K1 * H);
  pdl_sync();                                        // everything below may read the predecessor's output

  if (c.train) {
    // bn1 statistics over all B rows (every CTA of this head computes them; tile 0 publishes)
    const int rpar = 256 / K1 > 0 ? 256 / K1 : 1;
    const int col = threadIdx.x % K1, rs = threadIdx.x / K1;
    double s = 0.0, q = 0.0;
    if (rs < rpar)
      for (int b = rs; b < B; b += rpar) {
        double v = (double)head_input(c, h, b, col, H);
        s += v;
        q += v * v;
      }
    sSum[threadIdx.x] = s;
    sSum[256 + threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < K1) {
      double a = 0.0, bq = 0.0;
      for (int r = 0; r < rpar; ++r) {
        a += sSum[r * K1 + threadIdx.x];
        bq += sSum[256 + r * K1 + threadIdx.x];
      }
      const int k = threadIdx.x;
      double mean = a / B, var = bq / B - mean * mean;
      if (var < 0.0) var = 0.0;
      float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
      float g = c.params[c.bn_gamma[bn1] + k], be = c.params[c.bn_beta[bn1] + k];
      float sc = g * rstd, sh = be - (float)mean * sc;
      sSc[k] = sc;
      sSh[k] = sh;
      if (tile == 0) {
        c.bnf(bn1, BN_SCALE)[k] = sc;
        c.bnf(bn1, BN_SHIFT)[k] = sh;
        c.bnf(bn1, BN_MEAN)[k] = (float)mean;
        c.bnf(bn1, BN_RSTD)[k] = rstd;
        if (c.bn_buffers != nullptr) {
          double unb = B > 1 ? var * ((double)B / (double)(B - 1)) : var;
          float* rm = c.bn_buffers + c.bn_rm[bn1];
          float* rv = c.bn_buffers + c.bn_rv[bn1];
          rm[k] = (1.f - c.momentum) * rm[k] + c.momentum * (float)mean;
          rv[k] = (1.f - c.momentum) * rv[k] + c.momentum * (float)unb;
        }
      }
    }
    if (tile == 0 && threadIdx.x == 0 && c.nbt != nullptr) c.nbt[bn1] += 1;
  } else {
    for (int k = threadIdx.x; k < K1; k += blockDim.x) {
      sSc[k] = c.bnf(bn1, BN_SCALE)[k];
      sSh[k] = c.bnf(bn1, BN_SHIFT)[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kTileRows * K1; i += blockDim.x) {
    const int r = i / K1, k = i - r * K1;
    float v = 0.f;
    if (row0 + r < B) v = fmaf(head_input(c, h, row0 + r, k, H), sSc[k], sSh[k]);
    sA[i] = v;
  }
  cp_async_wait_all();
  __syncthreads();
  float acc[kRPW][VEC];
#pragma unroll
  for (int r = 0; r < kRPW; ++r)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
  tile_gemm<VEC, kRPW>(sA, K1, sW, H, K1, acc);
  const float* b1 = c.params + c.po.fc1_b[h];
  float* H1 = c.H1 + (size_t)h * c.Bm * H;
  double st[2][VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) st[0][k] = st[1][k] = 0.0;
#pragma unroll
  for (int r = 0; r < kRPW; ++r) {
    const int b = row0 + warp * kRPW + r;
    if (b < B) {
      RowVec<VEC> o;
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        o.v[k] = fmaxf(acc[r][k] + b1[lane * VEC + k], 0.f);
        st[0][k] += (double)o.v[k];
        st[1][k] += (double)o.v[k] * (double)o.v[k];
      }
      o.store(H1 + (size_t)b * H, lane);
    }
  }
  if (c.train) {
    const int T1 = ceil_div(B, kTileRows);
    block_partial_store_ex<VEC, 2>(st, sRed, c.statp, H, h * T1 + tile, 2, 0, H, 0);
    if (grid_last_block(&c.counters[CNT_HEAD1], 3 * T1))
      for (int hh = 0; hh < 3; ++hh)
        bn_finalize(c, c.L + 6 + hh, c.statp + (size_t)hh * T1 * 2 * H, T1, 2, 0, 1, B);
  }
}

// fc2 + log_softmax + loss: warp per graph row, blockIdx.y = head.
template <int VEC>
__global__ void __launch_bounds__(256) k_head2_fwd(const Ctx c) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  __shared__ float sLoss[kRowWarps][2];
  const int B = clampB(c);
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kHeadRowsPerCta + warp;
  const int C = c.C;
  const int nct = ceil_div(B, kHeadRowsPerCta);
  if (blockIdx.x >= nct) return;
  float loss_row = 0.f, correct = 0.f;
  if (b < B) {
    BnLane<VEC> bn;
    bn.load_fwd(c, c.L + 6 + h, lane);
    RowVec<VEC> x;
    x.load_coherent(c.H1 + ((size_t)h * c.Bm + b) * H, lane);
#pragma unroll
    for (int k = 0; k < VEC; ++k) x.v[k] = fmaf(x.v[k], bn.sc[k], bn.sh[k]);
    const float* W2 = c.params + c.po.fc2_w[h];      // [C][H]
    const float* b2 = c.params + c.po.fc2_b[h];
    float mine = -INFINITY;                           // lane `cls` keeps logit[cls]
    for (int cls = 0; cls < C; ++cls) {
      RowVec<VEC> w;
      w.load(W2 + (size_t)cls * H, lane);
      float v = warp_sum(dot_lane<VEC>(x, w)) + b2[cls];
      if (lane == cls) mine = v;
    }
    float m = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float e = lane < C ? expf(mine - m) : 0.f;
    float se = warp_sum(e);
    float lp = mine - m - logf(se);
    if (lane < C) c.logp[((size_t)h * c.Bm + b) * C + lane] = lp;
    if (c.with_loss && c.y != nullptr) {
      const long long yb = c.y[b];
      // argmax (first maximal index)
      int am = lane < C && mine == m ? lane : 1 << 30;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) am = min(am, __shfl_xor_sync(0xffffffffu, am, o));
      correct = (long long)am == yb ? 1.f : 0.f;
      if (h == 0) {
        float s = warp_sum(lane < C ? lp : 0.f);
        loss_row = -logf((float)C) - s / (float)C;   // sum_j (1/C) (log(1/C) - logp_j)
      } else {
        float pick = (yb >= 0 && yb < C && lane == (int)yb) ? lp : 0.f;
        loss_row = -warp_sum(pick);
      }
    }
  }
  if (c.with_loss) {
    if (lane == 0) {
      sLoss[warp][0] = loss_row;
      sLoss[warp][1] = correct;
    }
    __syncthreads();
    float* lossp = c.loss + 8;                        // [3][2][g_head2]
    if (threadIdx.x < 2) {
      float s = 0.f;
      for (int w = 0; w < kRowWarps; ++w) s += sLoss[w][threadIdx.x];
      lossp[((size_t)h * 2 + threadIdx.x) * c.g_head2 + blockIdx.x] = s;
    }
    if (grid_last_block(&c.counters[CNT_HEAD2], 3 * nct)) {
      if (threadIdx.x < 6) {
        float s = 0.f;
        for (int g = 0; g < nct; ++g) s += lossp[(size_t)threadIdx.x * c.g_head2 + g];
        const int hh = threadIdx.x >> 1;
        if (threadIdx.x & 1) c.loss[4 + hh] = s;      // correct counts c / o / co
        else c.loss[1 + hh] = B > 0 ? s / (float)B : 0.f;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        c.loss[0] = c.w_c * c.loss[1] + c.w_o * c.loss[2] + c.w_co * c.loss[3];
        c.loss[7] = 0.f;
      }
    }
  }
}

// fc2 / log_softmax / bn2 backward: warp per graph row, blockIdx.y = head.
template <int VEC>
__global__ void __launch_bounds__(256) k_head2_bwd(const Ctx c) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  __shared__ __align__(16) float sH2[kHeadRowsPerCta][H];
  __shared__ float sDl[kHeadRowsPerCta][32];
  __shared__ double sRed[kRowWarps * H];
  const int B = clampB(c);
  const int h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kHeadRowsPerCta + warp;
  const int C = c.C;
  const int nct = ceil_div(B, kHeadRowsPerCta);
  if (blockIdx.x >= nct) return;
  const int bn2 = c.L + 6 + h;
  double st[2][VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) st[0][k] = st[1][k] = 0.0;
  RowVec<VEC> h2;
  h2.zero();
  float dl = 0.f;
  if (b < B) {
    BnLane<VEC> bn;
    bn.load_bwd(c, bn2, lane);
    RowVec<VEC> x;
    x.load_coherent(c.H1 + ((size_t)h * c.Bm + b) * H, lane);
    const float lp = lane < C ? c.logp[((size_t)h * c.Bm + b) * C + lane] : 0.f;
    float dlp = 0.f;
    if (c.grad_logp != nullptr) {
      if (lane < C) dlp = c.grad_logp[((size_t)h * B + b) * C + lane];
    } else if (lane < C) {
      const long long yb = c.y != nullptr ? c.y[b] : -1;
      if (h == 0) dlp = -c.w_c / ((float)C * (float)B);
      else if ((long long)lane == yb) dlp = -(h == 1 ? c.w_o : c.w_co) / (float)B;
    }
    const float sd = warp_sum(dlp);
    dl = lane < C ? dlp - expf(lp) * sd : 0.f;        // d logits
    const float* W2 = c.params + c.po.fc2_w[h];
    RowVec<VEC> dh;
    dh.zero();
    for (int cls = 0; cls < C; ++cls) {
      const float dv = __shfl_sync(0xffffffffu, dl, cls);
      RowVec<VEC> w;
      w.load(W2 + (size_t)cls * H, lane);
#pragma unroll
      for (int k = 0; k < VEC; ++k) dh.v[k] = fmaf(dv, w.v[k], dh.v[k]);
    }
    dh.store(c.dh + ((size_t)h * c.Bm + b) * H, lane);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      h2.v[k] = fmaf(x.v[k], bn.sc[k], bn.sh[k]);
      st[0][k] += (double)dh.v[k];
      st[1][k] += (double)dh.v[k] * (double)bn.xhat(k, x.v[k]);
    }
  }
  h2.store(&sH2[warp][0], lane);
  sDl[warp][lane] = dl;
  __syncthreads();
  // fc2 weight / bias gradient partials of this CTA's 8 rows
  float* gp = c.gpart + c.gp_fc2[h] + (size_t)blockIdx.x * (C * H + C);
  for (int t = threadIdx.x; t < C * H; t += blockDim.x) {
    const int cls = t / H, k = t - cls * H;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kHeadRowsPerCta; ++r) s = fmaf(sDl[r][cls], sH2[r][k], s);
    gp[t] = s;
  }
  if (threadIdx.x < C) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kHeadRowsPerCta; ++r) s += sDl[r][threadIdx.x];
    gp[C * H + threadIdx.x] = s;
  }
  block_partial_store_ex<VEC, 2>(st, sRed, c.statp, H, h * nct + blockIdx.x, 2, 0, H, 0);
  if (grid_last_block(&c.counters[CNT_BHEAD2], 3 * nct))
    for (int hh = 0; hh < 3; ++hh)
      bn_bwd_finalize(c, c.L + 6 + hh, c.statp + (size_t)hh * nct * 2 * H, nct, 2, 0, 1, B);
}

// fc1 / bn1 backward on a 32-row tile, blockIdx.y = head.
// smem: sW [H][K1] (fc1 weight as stored) | sU [R][H] | sY [R][K1] | sRed f64 [8][H]
template <int VEC>
__global__ void __launch_bounds__(256) k_head1_bwd(const Ctx c) {
  constexpr int H = 32 * VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int B = clampB(c);
  const int h = blockIdx.y, tile = blockIdx.x;
  const int row0 = tile * kTileRows;
  if (row0 >= B) return;
  const int K1 = (h == 2 && c.cat) ? 2 * H : H;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sU = sW + (size_t)H * K1;
  float* sY = sU + kTileRows * H;
  double* sRed = reinterpret_cast<double*>(sY + (size_t)kTileRows * K1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn1 = c.L + 3 + h, bn2 = c.L + 6 + h;
  stage_matrix_async(sW, c.params + c.po.fc1_w[h], H * K1);
  pdl_sync();                                        // everything below may read the predecessor's output
  BnLane<VEC> b2;
  b2.load_bwd(c, bn2, lane);
  const float* H1 = c.H1 + (size_t)h * c.Bm * H;
  const float* dh = c.dh + (size_t)h * c.Bm * H;
  float dbias[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) dbias[k] = 0.f;
#pragma unroll
  for (int r = 0; r < kRPW; ++r) {
    const int lr = warp * kRPW + r, b = row0 + lr;
    RowVec<VEC> u;
    u.zero();
    if (b < B) {
      RowVec<VEC> x, g;
      x.load_coherent(H1 + (size_t)b * H, lane);
      g.load_coherent(dh + (size_t)b * H, lane);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        u.v[k] = x.v[k] > 0.f ? b2.dx(k, g.v[k], x.v[k]) : 0.f;
        dbias[k] += u.v[k];
      }
    }
    u.store(sU + lr * H, lane);
  }
  const float* sc1 = c.bnf(bn1, BN_SCALE);
  const float* sh1 = c.bnf(bn1, BN_SHIFT);
  for (int i = threadIdx.x; i < kTileRows * K1; i += blockDim.x) {
    const int r = i / K1, k = i - r * K1;
    float v = 0.f;
    if (row0 + r < B) v = fmaf(head_input(c, h, row0 + r, k, H), sc1[k], sh1[k]);
    sY[i] = v;
  }
  cp_async_wait_all();
  __syncthreads();
  float* gp = c.gpart + c.gp_fc1[h] + (size_t)tile * (H * 2 * H + H);
  const int T1 = ceil_div(B, kTileRows);
  const int nhalf = K1 / H;
  for (int half = 0; half < nhalf; ++half) {
    // du'[b][half*H + j] = sum_k dh1[b][k] W1[k][half*H + j]
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm<VEC, kRPW>(sU, H, sW + half * H, K1, H, acc);
    BnLane<VEC> b1;
    b1.load_bwd(c, bn1, lane, half * H);
    double st[2][VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) st[0][k] = st[1][k] = 0.0;
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int b = row0 + warp * kRPW + r;
      if (b < B) {
        RowVec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          o.v[k] = acc[r][k];
          const float uin = head_input(c, h, b, half * H + lane * VEC + k, H);
          st[0][k] += (double)acc[r][k];
          st[1][k] += (double)acc[r][k] * (double)b1.xhat(k, uin);
        }
        o.store(c.du + ((size_t)h * c.Bm + b) * 2 * H + half * H, lane);
      }
    }
    block_partial_store_ex<VEC, 2>(st, sRed, c.statp + (size_t)h * T1 * 4 * H, H, tile, 2, 0, K1, half * H);
    // dW1[k_out][half*H + j] = sum_b dh1[b][k_out] * y1[b][half*H + j]
    OuterAcc<H> dW;
    dW.zero();
    dW.accumulate(sU, H, sY + half * H, K1, kTileRows);
    dW.store(gp + half * H, K1);
  }
  block_colsum_store<VEC>(dbias, reinterpret_cast<float*>(sRed), gp + H * K1, H);
  if (grid_last_block(&c.counters[CNT_BHEAD1], 3 * T1))
    for (int hh = 0; hh < 3; ++hh)
      bn_bwd_finalize(c, c.L + 3 + hh, c.statp + (size_t)hh * T1 * 4 * H, T1, 2, 0, 1, B);
}

// Gradient w.r.t. the pooled embeddings: bn1 backward of the three readouts, the c <- co path
// routed through the inverse permutation (model.py:152-157).
template <int VEC>
__global__ void __launch_bounds__(256) k_dpool(const Ctx c) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  const int B = clampB(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kRowWarps + warp;
  if (b >= B) return;
  const float* gc = c.pooled;
  const float* go = c.pooled + (size_t)c.Bm * H;
  BnLane<VEC> bc, bo, bco, bco2;
  bc.load_bwd(c, c.L + 3, lane);
  bo.load_bwd(c, c.L + 4, lane);
  bco.load_bwd(c, c.L + 5, lane);
  if (c.cat) bco2.load_bwd(c, c.L + 5, lane, H);
  const int bi = c.invperm[b];       // co row that consumed xc_g[b]
  RowVec<VEC> xc, xo, dc, dO, dco_c, dco_o, xo_bi, xc_pb;
  xc.load_coherent(gc + (size_t)b * H, lane);
  xo.load_coherent(go + (size_t)b * H, lane);
  dc.load_coherent(c.du + ((size_t)0 * c.Bm + b) * 2 * H, lane);
  dO.load_coherent(c.du + ((size_t)1 * c.Bm + b) * 2 * H, lane);
  dco_c.load_coherent(c.du + ((size_t)2 * c.Bm + bi) * 2 * H, lane);
  xo_bi.load_coherent(go + (size_t)bi * H, lane);
  xc_pb.load_coherent(gc + (size_t)c.perm[b] * H, lane);
  dco_o.load_coherent(c.du + ((size_t)2 * c.Bm + b) * 2 * H + (c.cat ? H : 0), lane);
  RowVec<VEC> oc, oo;
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    float g_c = bc.dx(k, dc.v[k], xc.v[k]);
    float g_o = bo.dx(k, dO.v[k], xo.v[k]);
    if (c.cat) {
      g_c += bco.dx(k, dco_c.v[k], xc.v[k]);
      g_o += bco2.dx(k, dco_o.v[k], xo.v[k]);
    } else {
      g_c += bco.dx(k, dco_c.v[k], xc.v[k] + xo_bi.v[k]);
      g_o += bco.dx(k, dco_o.v[k], xc_pb.v[k] + xo.v[k]);
    }
    oc.v[k] = g_c;
    oo.v[k] = g_o;
  }
  oc.store(c.dpool + (size_t)b * H, lane);
  oo.store(c.dpool + ((size_t)c.Bm + b) * H, lane);
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

int launch_heads_forward(const Ctx& c, int with_loss, cudaStream_t s) {
  (void)with_loss;
  const int H = c.H, K1m = c.cat ? 2 * H : H;
  CAL_DISPATCH_VEC(c.H, {
    launch_k(k_pool<VEC>, dim3(c.Bm), dim3(256), 0, s, c);
    size_t smem = (size_t)K1m * H * 4 + (size_t)kTileRows * K1m * 4 + (size_t)kRowWarps * H * 8 + 2 * K1m * 4 + 512 * 8;
    int rc = set_smem_h(k_head1_fwd<VEC>, smem);
    if (rc) return rc;
    launch_k(k_head1_fwd<VEC>, dim3(c.t_head1, 3), dim3(256), smem, s, c);
    launch_k(k_head2_fwd<VEC>, dim3(c.g_head2, 3), dim3(256), 0, s, c);
  });
  note_launches(3);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_heads_backward(const Ctx& c, cudaStream_t s) {
  const int H = c.H, K1m = c.cat ? 2 * H : H;
  CAL_DISPATCH_VEC(c.H, {
    launch_k(k_head2_bwd<VEC>, dim3(c.g_head2, 3), dim3(256), 0, s, c);
    size_t smem = (size_t)H * K1m * 4 + (size_t)kTileRows * H * 4 + (size_t)kTileRows * K1m * 4 + (size_t)kRowWarps * H * 8;
    int rc = set_smem_h(k_head1_bwd<VEC>, smem);
    if (rc) return rc;
    launch_k(k_head1_bwd<VEC>, dim3(c.t_head1, 3), dim3(256), smem, s, c);
    launch_k(k_dpool<VEC>, dim3(ceil_div(c.Bm, kRowWarps)), dim3(256), 0, s, c);
  });
  note_launches(3);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
