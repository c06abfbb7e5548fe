// conv.cu -- the node-level kernels: input transform, the fused GCNConv layer (forward and
// backward), the two attention-masked convs, and their weight gradients.
//
// One launch per BatchNorm boundary (training-mode BatchNorm is the only coupling between the
// graphs of a batch): a persistent CTA walks 32-row tiles; for every tile it
//   1. stages the tile's CSR segment in shared memory with coalesced loads,
//   2. gathers + normalises the neighbour rows (warp per destination row, lanes over channels,
//      previous BatchNorm applied on load) into a shared-memory tile        [gcn_conv.py:92-97]
//   3. multiplies the tile by the weight matrix held in shared memory (fp32 FFMA; aggregation and
//      the linear map commute, so A(XW) is evaluated as (AX)W)               [gcn_conv.py:75]
//   4. bias + ReLU, stores the rows, and accumulates the next BatchNorm's statistics in fp64;
// the last CTA to finish turns the per-CTA partial sums into the affine the next kernel applies.
#include "internal.cuh"

namespace cal {

namespace {

struct Dims {
  int N, E, B;
};


__device__ __forceinline__ Dims load_dims(const Ctx& c) {
  Dims d;
  d.N = imin(imax(c.dims[0], 0), c.Nm);
  d.E = imin(imax(c.dims[1], 0), c.Em);
  d.B = imin(imax(c.dims[2], 0), c.Bm);
  return d;
}

// ---------------------------------------------------------------------------------------------
// x_1 = relu(bn_feat(x) @ W_feat)    (model.py:90-91; conv_feat is gfn=True: no bias, no
// propagation, gcn_conv.py:75-77) + statistics of bns_conv[0].
// smem: sW [Fp][H] | sA [R][Fp] | sRed f64 [8][H]
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_feat_fwd(const Ctx c) {
  constexpr int H = 32 * VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sTot[4 * H];
  const Dims d = load_dims(c);
  const int N = d.N, F = c.F, Fp = (F + 3) & ~3;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sA = sW + (size_t)Fp * H;
  double* sRed = reinterpret_cast<double*>(sA + (size_t)kTileRows * Fp);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* W = c.params + c.po.conv_feat_w;
  for (int i = threadIdx.x; i < Fp * H; i += blockDim.x) sW[i] = i < F * H ? W[i] : 0.f;
  pdl_sync();                                        // everything below may read the predecessor's output
  // bn_feat (model.py:90): the column totals come from k_prep_init (statp[0 .. 2F)); every CTA turns
  // them into the affine it applies, CTA 0 also publishes the record and the running statistics
  float* sc = reinterpret_cast<float*>(sRed + kRowWarps * H);       // [Fp]
  float* sh = sc + Fp;                                               // [Fp]
  if (c.train) {
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
      const double s = c.statp[k], q = c.statp[F + k];
      double mean = 0.0, var = 0.0;
      if (N > 0) {
        mean = s / N;
        var = q / N - mean * mean;
        if (var < 0.0) var = 0.0;
      }
      const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
      const float scale = c.params[c.bn_gamma[0] + k] * rstd;
      const float shift = c.params[c.bn_beta[0] + k] - (float)mean * scale;
      sc[k] = scale;
      sh[k] = shift;
      if (blockIdx.x == 0) {
        c.bnf(0, BN_SCALE)[k] = scale;
        c.bnf(0, BN_SHIFT)[k] = shift;
        c.bnf(0, BN_MEAN)[k] = (float)mean;
        c.bnf(0, BN_RSTD)[k] = rstd;
        if (c.bn_buffers != nullptr && c.bn_rm[0] >= 0) {
          const double unb = N > 1 ? var * ((double)N / (double)(N - 1)) : var;
          float* rm = c.bn_buffers + c.bn_rm[0];
          float* rv = c.bn_buffers + c.bn_rv[0];
          rm[k] = (1.f - c.momentum) * rm[k] + c.momentum * (float)mean;
          rv[k] = (1.f - c.momentum) * rv[k] + c.momentum * (float)unb;
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && c.nbt != nullptr) c.nbt[0] += 1;
  } else {
    for (int k = threadIdx.x; k < F; k += blockDim.x) {
      sc[k] = c.bnf(0, BN_SCALE)[k];
      sh[k] = c.bnf(0, BN_SHIFT)[k];
    }
  }
  __syncthreads();
  float* out = c.Xl(0);
  double acc_s[VEC], acc_q[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc_s[i] = acc_q[i] = 0.0;
  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    __syncthreads();
    for (int i = threadIdx.x; i < kTileRows * Fp; i += blockDim.x) {
      int r = i / Fp, k = i - r * Fp;
      float v = 0.f;
      if (row0 + r < N && k < F) v = fmaf(c.feat[(size_t)(row0 + r) * F + k], sc[k], sh[k]);
      sA[i] = v;
    }
    __syncthreads();
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[r][i] = 0.f;
    tile_gemm<VEC, kRPW>(sA, Fp, sW, H, Fp, acc);
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int i = row0 + warp * kRPW + r;
      if (i < N) {
        RowVec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          o.v[k] = fmaxf(acc[r][k], 0.f);
          acc_s[k] += (double)o.v[k];
          acc_q[k] += (double)o.v[k] * (double)o.v[k];
        }
        o.store(out + (size_t)i * H, lane);
      }
    }
  }
  if (c.train && c.model != CAL_MODEL_GIN) {          // CausalGIN: no BatchNorm consumes x_1 (model.py:237-238)
    double accs[2][VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      accs[0][k] = acc_s[k];
      accs[1][k] = acc_q[k];
    }
    if (bn_consumer_side(c)) {                         // layer 0 sums the group vectors and finalises bns_conv[0]
      block_totals<VEC, 2>(accs, sRed, sTot, H, 0, H, 0);
      grid_sum_groups(c, bn_site(1), sTot, 2 * H, gridDim.x, blockIdx.x);
    } else {
      const BnPre pre = bn_prefetch(c, 1);
      block_totals<VEC, 2>(accs, sRed, sTot, H, 0, H, 0);
      if (grid_sum(c, 0, sTot, 2 * H, gridDim.x, blockIdx.x)) bn_finalize_tot(c, 1, sTot, sTot + H, N, &pre);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused GCNConv layer, forward.
//   MODE 0: backbone layer l < L-1      x_{l+2} = relu(A bn_l(x_{l+1}) W_l + b_l)   (model.py:93-95)
//   MODE 1: last backbone layer; epilogue also evaluates node_att_mlp / edge_att_mlp projections
//           (model.py:97-111) and the statistics of bnc / bno on att * x
//   (the two masked convs, model.py:112-113, have their own kernel below: k_masked_fwd_both)
//   MODE 3: first half of a CausalGIN layer (model.py:187-193, PyG GINConv): h = (x_i + sum_j x_j) W1^T + b1,
//           i.e. the same gather with unit weights and no BatchNorm on load, no ReLU; the epilogue
//           accumulates the statistics of the layer's inner BatchNorm (BN id 1 + layer)
// smem: sW [H][H] | sA [R][H] | sRed f64 [8][H] | sPtr [R+4] | sSrc [EC] | sNrm [EC]
// ---------------------------------------------------------------------------------------------
// forward layers: sW [H][H] | sA [R][H] | sRed f64 [8][H] | sPtr [R+4] | sX [stage][H] | sXn | sXa | sXs [stage]
template <int VEC>
constexpr size_t convf_smem_bytes(int stage_rows) {
  constexpr int H = 32 * VEC;
  return (size_t)H * H * 4 + (size_t)kTileRows * (H + kPad) * 4 + (size_t)kRowWarps * H * 8 + (kTileRows + 4) * 4 +
         (size_t)stage_rows * (H * 4 + 12);
}

template <int VEC, int MODE>
__global__ void __launch_bounds__(256) k_conv_fwd(const Ctx c, const int layer) {
  constexpr int H = 32 * VEC, LDA = H + kPad;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sTot[4 * H];
  const Dims d = load_dims(c);
  const int N = d.N;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sA = sW + H * H;
  double* sRed = reinterpret_cast<double*>(sA + kTileRows * LDA);
  int* sPtr = reinterpret_cast<int*>(sRed + kRowWarps * H);
  constexpr int kStage = kStageFwd;
  float* sX = reinterpret_cast<float*>(sPtr + kTileRows + 4);   // [kStage][H] staged neighbour rows (16-byte aligned)
  float* sXn = sX + (size_t)kStage * H;                // [kStage] per-entry norm factor
  int* sXs = reinterpret_cast<int*>(sXn + kStage);     // [kStage] source node of the entry
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn_in = MODE == 3 ? kBnIdentity : 1 + layer;
  const float* W = MODE == 3 ? c.wt_conv(layer)      // torch Linear stores [out, in]: the transposed copy is [in, out]
                             : c.params + c.po.convs_w[layer];
  const float* bias = c.params + c.po.convs_b[layer];
  const float* xin = c.Xl(layer);
  float* xout = MODE == 3 ? c.gin_h(layer) : c.Xl(layer + 1);

  stage_matrix_async(sW, W, H * H);
  // backbone layers in training: the BatchNorm on the input is finalised HERE from the group sums the
  // previous kernel left behind (bn_from_groups); its parameters are fetched before the dependency wait
  __shared__ float s_aff[(MODE == 0 || MODE == 1) ? 2 * H : 1];
  const bool cs = (MODE == 0 || MODE == 1) && bn_consumer_side(c);
  BnPre pre_in = {1.f, 0.f, 0.f, 1.f};
  if (cs) pre_in = bn_prefetch(c, bn_in);
  pdl_sync();                                        // everything below may read the predecessor's output

  BnLane<VEC> bn;
  if (cs) {
    bn_from_groups(c, bn_site(bn_in), bn_in, c.g_tile, 2 * H, 0, N, pre_in, sTot, s_aff, s_aff + H, blockIdx.x == 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      bn.sc[i] = s_aff[lane * VEC + i];
      bn.sh[i] = s_aff[H + lane * VEC + i];
    }
  } else {
    bn.load_fwd(c, bn_in, lane);
  }
  float bv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) bv[i] = bias[lane * VEC + i];

  LayerEpilogue<VEC, MODE == 1> epi;
  epi.init(c, lane);

  PT_DECL
  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    const int nrows = imin(kTileRows, N - row0);
    __syncthreads();                                   // previous tile's readers of sA / sPtr are done
    PT_MARK();                                         // 0: prologue (params, BN record, W issue)
    if (threadIdx.x <= nrows) sPtr[threadIdx.x] = c.in_ptr[row0 + threadIdx.x];
    __syncthreads();
    PT_MARK();                                         // 1: CSR pointers
    // ---- gather: the neighbour rows of a batch of target rows are staged in shared memory with
    // one bulk copy per CSR entry (all in flight at once), then summed warp-per-row from SMEM ----
    for (int rb = 0; rb < nrows;) {
      int re = rb;
      while (re < nrows && sPtr[re + 1] - sPtr[rb] <= kStage) ++re;
      const bool direct = re == rb;                    // a single row with more in-edges than the stage holds
      if (direct) re = rb + 1;
      const int pb = sPtr[rb], cnt = sPtr[re] - pb;
      if (!direct) {
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
          const int src = c.in_src[pb + e];
          sXs[e] = src;
          sXn[e] = MODE == 3 ? 1.f : c.in_norm[pb + e];
        }
        __syncthreads();
        // one 16-byte cp.async per lane and row: a warp moves one neighbour row per instruction and
        // every row of the batch is in flight at once
        for (int e = warp; e < cnt; e += kRowWarps)
          if (lane < H / 4) cp_async16(sX + (size_t)e * H + lane * 4, xin + (size_t)sXs[e] * H + lane * 4);
        cp_async_wait_all();
        __syncthreads();
        for (int lr = rb + warp; lr < re; lr += kRowWarps) {
          const int e0 = sPtr[lr] - pb, e1 = sPtr[lr + 1] - pb;
          RowVec<VEC> a;
          a.zero();
          for (int e = e0; e < e1; ++e) {
            RowVec<VEC> v;
            v.load_coherent(sX + (size_t)e * H, lane);
            const float w = sXn[e];                                      // dis[row] * dis[col] (1 for GIN)
#pragma unroll
            for (int k = 0; k < VEC; ++k) a.v[k] = fmaf(w, fmaf(v.v[k], bn.sc[k], bn.sh[k]), a.v[k]);
          }
          a.store(sA + lr * LDA, lane);
        }
        __syncthreads();                               // the stage is rewritten by the next batch
      } else {
        if (warp == 0) {                               // hub row: gather straight from global memory
          const int lr = rb;
          RowVec<VEC> a;
          a.zero();
          for (int p = sPtr[lr]; p < sPtr[lr + 1]; ++p) {
            const int src = c.in_src[p];
            RowVec<VEC> v;
            v.load_coherent(xin + (size_t)src * H, lane);
            const float w = MODE == 3 ? 1.f : c.in_norm[p];
#pragma unroll
            for (int k = 0; k < VEC; ++k) a.v[k] = fmaf(w, fmaf(v.v[k], bn.sc[k], bn.sh[k]), a.v[k]);
          }
          a.store(sA + lr * LDA, lane);
        }
      }
      rb = re;
    }
    if (warp * kRPW + kRPW > nrows) {                  // zero the rows of a ragged last tile
      for (int r = 0; r < kRPW; ++r) {
        const int lr = warp * kRPW + r;
        if (lr >= nrows) {
          RowVec<VEC> z;
          z.zero();
          z.store(sA + lr * LDA, lane);
        }
      }
    }
    PT_MARK();                                         // 3: gather (this warp)
    cp_async_wait_all();
    __syncthreads();
    PT_MARK();                                         // 4: wait for W + all warps
    // ---- tile GEMM ----
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm_fast<VEC, kRPW>(sA, LDA, sW, H, H, sX, acc);           // (the row stage is free: partial tiles of the K-split)
    PT_MARK();                                         // 5: GEMM
    // ---- epilogue ----
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int i = row0 + warp * kRPW + r;
      if (i < N) {                                      // warp-uniform
        RowVec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = MODE == 3 ? acc[r][k] + bv[k] : fmaxf(acc[r][k] + bv[k], 0.f);
        o.store(xout + (size_t)i * H, lane);
        epi.row(c, i, o.v, lane);
      }
    }
  }
  cp_async_wait_all();
  PT_MARK();                                           // 6: epilogue stores
  epi.finish(c, MODE == 3 ? layer - 1 : layer, sRed, sTot, N);    // MODE 3: BN id 2 + (layer - 1) = 1 + layer
  PT_MARK();                                           // 7: totals + grid sum + finalize
  PT_DUMP(c, 16);
}

// ---------------------------------------------------------------------------------------------
// context_convs and objects_convs (model.py:112-113) in ONE pass over the graph: both branches
// aggregate the same neighbour rows x_{L+1}[src] -- they differ only in the node attention, the
// BatchNorm affine (bnc / bno), the attention-weighted norm and the weight matrix -- so a CTA
// stages every neighbour row once and feeds both accumulators, then runs the two tile GEMMs.
// (A per-branch variant needs 2 CTAs per SM, which leaves room for only 40 staged rows: 2-3 dependent
// staging rounds per tile, measured 13.9 us vs 12.0 us.  Here one round of up to 128 rows does.)
// smem: sW [2][H][H] | sA [2][R][H] | sPtr [R+4] | sX [stage][H] | sXn [stage][2] | sXa [stage][2] | sXs [stage]
// ---------------------------------------------------------------------------------------------
constexpr int kStageBoth = 128;
template <int VEC>
constexpr size_t masked_both_smem_bytes() {
  constexpr int H = 32 * VEC;
  return 2 * (size_t)H * H * 4 + 2 * (size_t)kTileRows * (H + kPad) * 4 + (kTileRows + 4) * 4 +
         (size_t)kStageBoth * (H * 4 + 20);
}

template <int VEC>
__global__ void __launch_bounds__(256) k_masked_fwd_both(const Ctx c) {
  constexpr int H = 32 * VEC, LDA = H + kPad;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Dims d = load_dims(c);
  const int N = d.N;
  float* sW = reinterpret_cast<float*>(smem_raw);                 // [2][H][H]
  float* sA = sW + 2 * H * H;                                     // [2][R][H]
  int* sPtr = reinterpret_cast<int*>(sA + 2 * kTileRows * LDA);
  float* sX = reinterpret_cast<float*>(sPtr + kTileRows + 4);     // [kStageBoth][H]
  float2* sXn = reinterpret_cast<float2*>(sX + (size_t)kStageBoth * H);   // dis_w[source] * edge_att, both branches
  float2* sXa = sXn + kStageBoth;                                 // node_att[source], both branches
  int* sXs = reinterpret_cast<int*>(sXa + kStageBoth);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xin = c.Xl(c.L);

  stage_matrix_async(sW, c.params + c.po.context_w, H * H);
  stage_matrix_async(sW + H * H, c.params + c.po.objects_w, H * H);
  __shared__ double s_scr[4 * H];
  __shared__ float s_aff[4 * H];
  const bool cs = bn_consumer_side(c);               // bnc / bno finalised here from the last layer's group sums
  BnPre pre0 = {1.f, 0.f, 0.f, 1.f}, pre1 = pre0;
  if (cs) {
    pre0 = bn_prefetch(c, c.L + 1);
    pre1 = bn_prefetch(c, c.L + 2, H);               // threads [H, 2H) finalise bno
  }
  pdl_sync();                                        // everything below may read the predecessor's output

  BnLane<VEC> bn0, bn1;
  if (cs) {
    const int site = bn_site(c.L + 1);
    bn_from_groups2(c, site, c.L + 1, 0, site, c.L + 2, 2 * H, c.g_tile, 4 * H, N, pre0, pre1, s_scr, s_aff, s_aff + H,
                    s_aff + 2 * H, s_aff + 3 * H, blockIdx.x == 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      bn0.sc[i] = s_aff[lane * VEC + i];
      bn0.sh[i] = s_aff[H + lane * VEC + i];
      bn1.sc[i] = s_aff[2 * H + lane * VEC + i];
      bn1.sh[i] = s_aff[3 * H + lane * VEC + i];
    }
  } else {
    bn0.load_fwd(c, c.L + 1, lane);
    bn1.load_fwd(c, c.L + 2, lane);
  }
  float bv0[VEC], bv1[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    bv0[i] = c.params[c.po.context_b + lane * VEC + i];
    bv1[i] = c.params[c.po.objects_b + lane * VEC + i];
  }
  float* agg0 = c.agg;
  float* agg1 = c.agg + (size_t)c.Nm * H;
  float* z0 = c.Z;
  float* z1 = c.Z + (size_t)c.Nm * H;
  const float2* edge_wn = reinterpret_cast<const float2*>(c.edge_wn);
  const float2* edge_na = reinterpret_cast<const float2*>(c.edge_na);
  const float2* disw = reinterpret_cast<const float2*>(c.disw);

  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    const int nrows = imin(kTileRows, N - row0);
    __syncthreads();                                   // previous tile's readers of sA / sPtr are done
    if (threadIdx.x <= nrows) sPtr[threadIdx.x] = c.in_ptr[row0 + threadIdx.x];
    __syncthreads();
    for (int rb = 0; rb < nrows;) {
      int re = rb;
      while (re < nrows && sPtr[re + 1] - sPtr[rb] <= kStageBoth) ++re;
      const bool direct = re == rb;                    // a single row with more in-edges than the stage holds
      if (direct) re = rb + 1;
      const int pb = sPtr[rb], cnt = sPtr[re] - pb;
      if (!direct) {
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
          sXs[e] = c.in_src[pb + e];
          sXn[e] = edge_wn[pb + e];
          sXa[e] = edge_na[pb + e];
        }
        __syncthreads();
        for (int e = warp; e < cnt; e += kRowWarps)
          if (lane < H / 4) cp_async16(sX + (size_t)e * H + lane * 4, xin + (size_t)sXs[e] * H + lane * 4);
        float2 di_pre[kRPW];                           // target-side factors: in flight with the row copies
#pragma unroll
        for (int k = 0; k < kRPW; ++k) {
          const int lr = rb + warp + k * kRowWarps;
          di_pre[k] = lr < re ? disw[row0 + lr] : make_float2(0.f, 0.f);
        }
        cp_async_wait_all();
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kRPW; ++k) {
          const int lr = rb + warp + k * kRowWarps;
          if (lr >= re) break;
          const int i = row0 + lr;
          const int e0 = sPtr[lr] - pb, e1 = sPtr[lr + 1] - pb;
          const float2 di = di_pre[k];
          RowVec<VEC> a0, a1;
          a0.zero();
          a1.zero();
          for (int e = e0; e < e1; ++e) {
            RowVec<VEC> v;
            v.load_coherent(sX + (size_t)e * H, lane);
            const float2 wn = sXn[e], am = sXa[e];
            const float w0 = wn.x * di.x, w1 = wn.y * di.y;              // dis[row] * w * dis[col]
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              a0.v[k] = fmaf(w0, fmaf(am.x * v.v[k], bn0.sc[k], bn0.sh[k]), a0.v[k]);
              a1.v[k] = fmaf(w1, fmaf(am.y * v.v[k], bn1.sc[k], bn1.sh[k]), a1.v[k]);
            }
          }
          a0.store(agg0 + (size_t)i * H, lane);
          a1.store(agg1 + (size_t)i * H, lane);
          a0.store(sA + lr * LDA, lane);
          a1.store(sA + (kTileRows + lr) * LDA, lane);
        }
        __syncthreads();                               // the stage is rewritten by the next batch
      } else {
        if (warp == 0) {                               // hub row: gather straight from global memory
          const int lr = rb, i = row0 + lr;
          const float2 di = disw[i];
          RowVec<VEC> a0, a1;
          a0.zero();
          a1.zero();
          for (int p = sPtr[lr]; p < sPtr[lr + 1]; ++p) {
            const int src = c.in_src[p];
            RowVec<VEC> v;
            v.load_coherent(xin + (size_t)src * H, lane);
            const float w0 = (c.disw[(size_t)src * 2] * c.watt[(size_t)p * 2]) * di.x;
            const float w1 = (c.disw[(size_t)src * 2 + 1] * c.watt[(size_t)p * 2 + 1]) * di.y;
            const float am0 = c.natt[(size_t)src * 2], am1 = c.natt[(size_t)src * 2 + 1];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              a0.v[k] = fmaf(w0, fmaf(am0 * v.v[k], bn0.sc[k], bn0.sh[k]), a0.v[k]);
              a1.v[k] = fmaf(w1, fmaf(am1 * v.v[k], bn1.sc[k], bn1.sh[k]), a1.v[k]);
            }
          }
          a0.store(agg0 + (size_t)i * H, lane);
          a1.store(agg1 + (size_t)i * H, lane);
          a0.store(sA + lr * LDA, lane);
          a1.store(sA + (kTileRows + lr) * LDA, lane);
        }
      }
      rb = re;
    }
    if (warp * kRPW + kRPW > nrows) {                  // zero the rows of a ragged last tile
      for (int r = 0; r < kRPW; ++r) {
        const int lr = warp * kRPW + r;
        if (lr >= nrows) {
          RowVec<VEC> z;
          z.zero();
          z.store(sA + lr * LDA, lane);
          z.store(sA + (kTileRows + lr) * LDA, lane);
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int br = 0; br < 2; ++br) {
      float acc[kRPW][VEC];
#pragma unroll
      for (int r = 0; r < kRPW; ++r)
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
      if (br == 1) __syncthreads();                    // the partial tiles of branch 0 have been read back
      tile_gemm_fast<VEC, kRPW>(sA + br * kTileRows * LDA, LDA, sW + br * H * H, H, H, sX, acc);
      float* zo = br ? z1 : z0;
#pragma unroll
      for (int r = 0; r < kRPW; ++r) {
        const int i = row0 + warp * kRPW + r;
        if (i < N) {                                    // warp-uniform
          RowVec<VEC> o;
#pragma unroll
          for (int k = 0; k < VEC; ++k) o.v[k] = fmaxf(acc[r][k] + (br ? bv1[k] : bv0[k]), 0.f);
          o.store(zo + (size_t)i * H, lane);
        }
      }
    }
  }
  cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------
// Fused GCNConv layer, backward (backbone layers).  With g_z = relu'(x_out) * bn_up'(D_up):
//   u_j   = sum_{e: row_e = j} norm_e * g_z[col_e]          (transpose aggregate, by-source CSR)
//   D_j   = u_j W^T                                          (gradient w.r.t. bn_l output)
//   dW   += bn_l(x_in)_j^T u_j,   db += g_z (own rows)
//   and the two BatchNorm-backward sums of bn_l (sum D, sum D * xhat).
// smem: sW [H][H] (W^T) | sU [R][H] | sY [R][H] | sRed f64 [8][H] | sPtr | sDst | sNrm
// ---------------------------------------------------------------------------------------------
// sW [H][H] | sU [R][H] | sY [R][H] | sRed f64 [8][H] | sPtr [R+4] | sXD [stage][H] | sXX [stage][H] | sXn | sXs [stage]
template <int VEC>
constexpr size_t convb_smem_bytes() {
  constexpr int H = 32 * VEC;
  return (size_t)H * H * 4 + 2 * (size_t)kTileRows * (H + kPad) * 4 + (size_t)kRowWarps * H * 8 + (kTileRows + 4) * 4 +
         (size_t)kStageBwd * (2 * H * 4 + 8);
}

template <int VEC, bool GIN>
__global__ void __launch_bounds__(256) k_conv_bwd(const Ctx c, const int layer) {
  constexpr int H = 32 * VEC, LDA = H + kPad;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sTot[2 * H];
  const Dims d = load_dims(c);
  const int N = d.N;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sU = sW + H * H;
  float* sY = sU + kTileRows * LDA;
  double* sRed = reinterpret_cast<double*>(sY + kTileRows * LDA);
  int* sPtr = reinterpret_cast<int*>(sRed + kRowWarps * H);
  float* sXD = reinterpret_cast<float*>(sPtr + kTileRows + 4);   // [kStageBwd][H] staged D_up rows
  float* sXX = sXD + (size_t)kStageBwd * H;                       // [kStageBwd][H] staged x_up rows
  float* sXn = sXX + (size_t)kStageBwd * H;                       // [kStageBwd] norm of the entry
  int* sXs = reinterpret_cast<int*>(sXn + kStageBwd);             // [kStageBwd] target node of the entry
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // GCN: D = U W^T needs the transposed copy; GIN: h = agg W1^T, so D = U W1 with W1 as stored ([out, in])
  stage_matrix_async(sW, GIN ? c.params + c.po.convs_w[layer] : c.wt_conv(layer), H * H);
  pdl_sync();                                        // everything below may read the predecessor's output

  // GIN (first half of the layer, model.py:187-193): upstream = the layer's inner BatchNorm applied to h;
  // the rows of `Dup` already carry the ReLU mask (k_gin_b_bwd); no BatchNorm below, unit edge weights
  const int bn_in = GIN ? kBnIdentity : 1 + layer;
  const int bn_up = GIN ? 1 + layer : (layer == c.L - 1 ? kBnIdentity : 2 + layer);
  const float* xin = c.Xl(layer);
  const float* xup = GIN ? c.gin_h(layer) : c.Xl(layer + 1);
  const float* Dup = GIN ? c.gin_dr() : c.D + (size_t)((layer + 1) & 1) * c.Nm * H;
  float* Dout = c.D + (size_t)(layer & 1) * c.Nm * H;
  BnLane<VEC> bi, bu;
  bi.load_bwd(c, bn_in, lane);
  bu.load_bwd(c, bn_up, lane);
  if (bn_up != kBnIdentity) {
    // the upstream BatchNorm's backward sums arrive as group vectors (k_conv_bwd of the layer above, or
    // k_gin_b_bwd): every CTA turns them into c1 / c2, CTA 0 publishes d gamma / d beta
    __shared__ float s_c[2 * H];
    bn_bwd_from_groups(c, bn_site(bn_up), bn_up, c.g_tile, 2 * H, 0, N, sTot, s_c, s_c + H, blockIdx.x == 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      bu.c1[i] = s_c[lane * VEC + i];
      bu.c2[i] = s_c[H + lane * VEC + i];
    }
  }

  OuterAcc<H> dW;
  dW.zero();
  float dbias[VEC];
  double st[2][VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    dbias[k] = 0.f;
    st[0][k] = st[1][k] = 0.0;
  }

  PT_DECL
  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    const int nrows = imin(kTileRows, N - row0);
    __syncthreads();
    PT_MARK();                                         // 0: prologue
    if (threadIdx.x <= nrows) sPtr[threadIdx.x] = c.out_ptr[row0 + threadIdx.x];
    __syncthreads();
    PT_MARK();                                         // 1: CSR pointers
    // ---- transpose gather: D_up and x_up rows of every out-edge of a batch of source rows are
    // staged with bulk copies; g_z = relu'(x_up) * bn_up'(D_up) is formed while summing ----
    for (int rb = 0; rb < nrows;) {
      int re = rb;
      while (re < nrows && sPtr[re + 1] - sPtr[rb] <= kStageBwd) ++re;
      const bool direct = re == rb;
      if (direct) re = rb + 1;
      const int qb = sPtr[rb], cnt = sPtr[re] - qb;
      if (!direct) {
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
          sXs[e] = c.out_dst[qb + e];
          sXn[e] = GIN ? 1.f : c.out_norm[qb + e];
        }
        __syncthreads();
        for (int e = warp; e < cnt; e += kRowWarps)
          if (lane < H / 4) {
            const size_t off = (size_t)sXs[e] * H + lane * 4;
            cp_async16(sXD + (size_t)e * H + lane * 4, Dup + off);
            cp_async16(sXX + (size_t)e * H + lane * 4, xup + off);
          }
        RowVec<VEC> xi_pre[kRPW];                      // the rows' own inputs: in flight with the row copies
#pragma unroll
        for (int k = 0; k < kRPW; ++k) {
          const int lr = rb + warp + k * kRowWarps;
          if (lr < re) xi_pre[k].load_coherent(xin + (size_t)(row0 + lr) * H, lane);
          else xi_pre[k].zero();
        }
        cp_async_wait_all();
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kRPW; ++k) {
          const int lr = rb + warp + k * kRowWarps;
          if (lr >= re) break;
          const int j = row0 + lr;
          const int e0 = sPtr[lr] - qb, e1 = sPtr[lr + 1] - qb;
          RowVec<VEC> u, y, xi;
          u.zero();
          xi = xi_pre[k];
          for (int e = e0; e < e1; ++e) {
            RowVec<VEC> gv, xv;
            gv.load_coherent(sXD + (size_t)e * H, lane);
            xv.load_coherent(sXX + (size_t)e * H, lane);
            const float w = sXn[e];
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float g = (GIN || xv.v[k] > 0.f) ? bu.dx(k, gv.v[k], xv.v[k]) : 0.f;
              u.v[k] = fmaf(w, g, u.v[k]);
              if (e == e1 - 1) dbias[k] += g;          // the appended self loop is the row's last out-entry
            }
          }
#pragma unroll
          for (int k = 0; k < VEC; ++k) y.v[k] = fmaf(xi.v[k], bi.sc[k], bi.sh[k]);
          u.store(sU + lr * LDA, lane);
          y.store(sY + lr * LDA, lane);
        }
        __syncthreads();
      } else {
        if (warp == 0) {                               // hub row: straight from global memory
          const int lr = rb, j = row0 + lr;
          RowVec<VEC> u, y, xi;
          u.zero();
          xi.load_coherent(xin + (size_t)j * H, lane);
          for (int q = sPtr[lr]; q < sPtr[lr + 1]; ++q) {
            const int dd = c.out_dst[q];
            const float w = GIN ? 1.f : c.out_norm[q];
            RowVec<VEC> gv, xv;
            gv.load_coherent(Dup + (size_t)dd * H, lane);
            xv.load_coherent(xup + (size_t)dd * H, lane);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float g = (GIN || xv.v[k] > 0.f) ? bu.dx(k, gv.v[k], xv.v[k]) : 0.f;
              u.v[k] = fmaf(w, g, u.v[k]);
              if (q == sPtr[lr + 1] - 1) dbias[k] += g;
            }
          }
#pragma unroll
          for (int k = 0; k < VEC; ++k) y.v[k] = fmaf(xi.v[k], bi.sc[k], bi.sh[k]);
          u.store(sU + lr * LDA, lane);
          y.store(sY + lr * LDA, lane);
        }
      }
      rb = re;
    }
    if (warp * kRPW + kRPW > nrows) {                  // zero the rows of a ragged last tile
      for (int r = 0; r < kRPW; ++r) {
        const int lr = warp * kRPW + r;
        if (lr >= nrows) {
          RowVec<VEC> z;
          z.zero();
          z.store(sU + lr * LDA, lane);
          z.store(sY + lr * LDA, lane);
        }
      }
    }
    PT_MARK();                                         // 2: staging + gather (this warp)
    cp_async_wait_all();
    __syncthreads();
    PT_MARK();                                         // 3: wait W + all warps
    RowVec<VEC> xe[kRPW];                              // inputs of the rows this warp finishes: loaded under the GEMM
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int j = row0 + warp * kRPW + r;
      if (j < N) xe[r].load_coherent(xin + (size_t)j * H, lane);
      else xe[r].zero();
    }
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm_fast<VEC, kRPW>(sU, LDA, sW, H, H, sXD, acc);          // (the row stage is free: partial tiles of the K-split)
    PT_MARK();                                         // 4: GEMM dX
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int j = row0 + warp * kRPW + r;
      if (j < N) {
        RowVec<VEC> o, xi;
        xi = xe[r];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          o.v[k] = acc[r][k];
          if (!GIN) {
            st[0][k] += (double)acc[r][k];
            st[1][k] += (double)acc[r][k] * (double)bi.xhat(k, xi.v[k]);
          }
        }
        o.store(Dout + (size_t)j * H, lane);
      }
    }
    PT_MARK();                                         // 5: dX stores
    if (GIN) dW.accumulate(sU, LDA, sY, LDA, kTileRows);      // torch Linear: d W1 [out, in]
    else dW.accumulate(sY, LDA, sU, LDA, kTileRows);
    PT_MARK();                                         // 6: dW outer products
  }
  cp_async_wait_all();
  float* gp = c.gpart + c.gp_conv[layer] + (size_t)blockIdx.x * (H * H + H);
  if (blockIdx.x < ntiles) dW.store(gp, H);
  block_colsum_store<VEC>(dbias, reinterpret_cast<float*>(sRed), blockIdx.x < ntiles ? gp + H * H : nullptr, H);
  PT_MARK();                                           // 7: dW / db partial stores
  if (!GIN) {                                          // group sums only: the kernel below finalises (bn_bwd_from_groups)
    block_totals<VEC, 2>(st, sRed, sTot, H, 0, H, 0);
    grid_sum_groups(c, bn_site(bn_in), sTot, 2 * H, gridDim.x, blockIdx.x);
  }
  PT_MARK();                                           // 8: totals + grid sum + finalize
  PT_DUMP(c, 32);
}

// ---------------------------------------------------------------------------------------------
// Masked convs backward, dense part (blockIdx.y = branch):
//   dz = dpool[graph(i)] * relu'(z_i);  dagg_i = dz_i W^T;  dW += agg_i^T dz_i;  db += dz_i
// (global_add_pool backward is the broadcast of the pooled gradient, model.py:115-116.)
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256, 2) k_masked_bwd_gemm(const Ctx c) {      // two branch CTAs per SM: <= 128 registers
  constexpr int H = 32 * VEC, LDA = H + kPad;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Dims d = load_dims(c);
  const int N = d.N;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sU = sW + H * H;
  float* sY = sU + kTileRows * LDA;
  double* sRed = reinterpret_cast<double*>(sY + kTileRows * LDA);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int branch = blockIdx.y;
  stage_matrix_async(sW, c.wt_conv(c.L + branch), H * H);
  pdl_sync();                                        // everything below may read the predecessor's output
  const float* Z = c.Z + (size_t)branch * c.Nm * H;
  const float* A = c.agg + (size_t)branch * c.Nm * H;
  float* dagg = c.dagg + (size_t)branch * c.Nm * H;
  OuterAcc<H> dW;
  dW.zero();
  float dbias[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) dbias[k] = 0.f;
  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int lr = warp * kRPW + r;
      const int i = row0 + lr;
      RowVec<VEC> u, y;
      u.zero();
      y.zero();
      if (i < N) {
        RowVec<VEC> z, g;
        z.load_coherent(Z + (size_t)i * H, lane);
        // gradient of the pooled embedding of graph b, straight from the d-input rows of the three readouts
        // (global_add_pool backward = broadcast, model.py:115-116; the c <- co path goes through the inverse
        // permutation, model.py:152-157): no separate k_dpool launch
        {
          const int b = c.node_graph[i];
          RowVec<VEC> g2;
          if (branch == 0) {
            g.load_coherent(c.du + ((size_t)0 * c.Bm + b) * 2 * H, lane);
            g2.load_coherent(c.du + ((size_t)2 * c.Bm + c.invperm[b]) * 2 * H, lane);
          } else {
            g.load_coherent(c.du + ((size_t)1 * c.Bm + b) * 2 * H, lane);
            g2.load_coherent(c.du + ((size_t)2 * c.Bm + b) * 2 * H + (c.cat ? H : 0), lane);
          }
#pragma unroll
          for (int k = 0; k < VEC; ++k) g.v[k] += g2.v[k];
        }
        y.load_coherent(A + (size_t)i * H, lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          u.v[k] = z.v[k] > 0.f ? g.v[k] : 0.f;
          dbias[k] += u.v[k];
        }
      }
      u.store(sU + lr * LDA, lane);
      y.store(sY + lr * LDA, lane);
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm<VEC, kRPW>(sU, LDA, sW, H, H, acc);
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int i = row0 + warp * kRPW + r;
      if (i < N) {
        RowVec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = acc[r][k];
        o.store(dagg + (size_t)i * H, lane);
      }
    }
    dW.accumulate(sY, LDA, sU, LDA, kTileRows);
  }
  cp_async_wait_all();
  float* gp = c.gpart + c.gp_conv[c.L + branch] + (size_t)blockIdx.x * (H * H + H);
  if (blockIdx.x < ntiles) dW.store(gp, H);
  block_colsum_store<VEC>(dbias, reinterpret_cast<float*>(sRed), blockIdx.x < ntiles ? gp + H * H : nullptr, H);
}

// ---------------------------------------------------------------------------------------------
// CausalGIN layer, second half (model.py:187-193: ... BatchNorm1d -> ReLU -> Linear -> ReLU), dense:
//   forward   x_{l+2} = relu(relu(bn(h)) W2^T + b2)           (h = first-half output, k_conv_fwd<MODE 3>)
//   backward  u = relu'(x_{l+2}) * D_up;  d W2 += u^T relu(bn(h));  d b2 += u;
//             d r = (u W2) * [bn(h) > 0]  -> gin_dr, + the inner BatchNorm's backward sums.
// No BatchNorm follows a GIN layer, so only the LAST layer's forward has a statistics epilogue
// (node / edge attention projections and bnc / bno, as in k_conv_fwd<MODE 1>).
// smem: sW [H][H] | sA [R][LDA] (| sY [R][LDA]) | sRed f64 [8][H]
// ---------------------------------------------------------------------------------------------
template <int VEC, bool LASTL>
__global__ void __launch_bounds__(256) k_gin_b_fwd(const Ctx c, const int layer) {
  constexpr int H = 32 * VEC, LDA = H + kPad;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sTot[LASTL ? 4 * H : 2 * H];
  __shared__ float s_aff[2 * H];
  const Dims d = load_dims(c);
  const int N = d.N;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sA = sW + H * H;
  double* sRed = reinterpret_cast<double*>(sA + kTileRows * LDA);
  float* sPart = reinterpret_cast<float*>(sRed + kRowWarps * H);     // [4][R][LDA] partial tiles of the K-split GEMM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_matrix_async(sW, c.wt_gin2(layer), H * H);
  const bool cs = bn_consumer_side(c);               // the layer's inner BatchNorm is finalised here
  BnPre pre_in = {1.f, 0.f, 0.f, 1.f};
  if (cs) pre_in = bn_prefetch(c, 1 + layer);
  pdl_sync();                                        // everything below may read the predecessor's output
  BnLane<VEC> bn;
  if (cs) {
    bn_from_groups(c, bn_site(1 + layer), 1 + layer, c.g_tile, 2 * H, 0, N, pre_in, sTot, s_aff, s_aff + H, blockIdx.x == 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      bn.sc[i] = s_aff[lane * VEC + i];
      bn.sh[i] = s_aff[H + lane * VEC + i];
    }
  } else {
    bn.load_fwd(c, 1 + layer, lane);
  }
  float bv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) bv[i] = c.params[c.po.gin_b2[layer] + lane * VEC + i];
  LayerEpilogue<VEC, LASTL> epi;
  if (LASTL) epi.init(c, lane);
  const float* hin = c.gin_h(layer);
  float* xout = c.Xl(layer + 1);
  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int lr = warp * kRPW + r, i = row0 + lr;
      RowVec<VEC> a;
      a.zero();
      if (i < N) {
        a.load_coherent(hin + (size_t)i * H, lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) a.v[k] = fmaxf(fmaf(a.v[k], bn.sc[k], bn.sh[k]), 0.f);
      }
      a.store(sA + lr * LDA, lane);
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm_fast<VEC, kRPW>(sA, LDA, sW, H, H, sPart, acc);
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int i = row0 + warp * kRPW + r;
      if (i < N) {                                      // warp-uniform
        RowVec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) o.v[k] = fmaxf(acc[r][k] + bv[k], 0.f);
        o.store(xout + (size_t)i * H, lane);
        if (LASTL) epi.row(c, i, o.v, lane);
      }
    }
  }
  cp_async_wait_all();
  if (LASTL) epi.finish(c, layer, sRed, sTot, N);
}

template <int VEC>
__global__ void __launch_bounds__(256) k_gin_b_bwd(const Ctx c, const int layer) {
  constexpr int H = 32 * VEC, LDA = H + kPad;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sTot[2 * H];
  const Dims d = load_dims(c);
  const int N = d.N;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sU = sW + H * H;
  float* sY = sU + kTileRows * LDA;
  double* sRed = reinterpret_cast<double*>(sY + kTileRows * LDA);
  float* sPart = reinterpret_cast<float*>(sRed + kRowWarps * H);     // [4][R][LDA] partial tiles of the K-split GEMM
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  stage_matrix_async(sW, c.params + c.po.gin_w2[layer], H * H);     // [out, in] as stored: d r = u W2
  pdl_sync();                                        // everything below may read the predecessor's output
  const int bn_id = 1 + layer;
  BnLane<VEC> bn;
  bn.load_bwd(c, bn_id, lane);                       // scale / shift / mean / rstd of the forward; c1 / c2 are produced here
  const float* hin = c.gin_h(layer);
  const float* xup = c.Xl(layer + 1);
  const float* Dup = c.D + (size_t)((layer + 1) & 1) * c.Nm * H;
  float* DR = c.gin_dr();
  OuterAcc<H> dW;
  dW.zero();
  float dbias[VEC];
  double st[2][VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    dbias[k] = 0.f;
    st[0][k] = st[1][k] = 0.0;
  }
  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    __syncthreads();
    RowVec<VEC> hr[kRPW];                              // h rows of this warp: needed again after the GEMM
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int lr = warp * kRPW + r, i = row0 + lr;
      RowVec<VEC> u, y;
      u.zero();
      y.zero();
      hr[r].zero();
      if (i < N) {
        RowVec<VEC> xo, g;
        xo.load_coherent(xup + (size_t)i * H, lane);
        g.load_coherent(Dup + (size_t)i * H, lane);
        hr[r].load_coherent(hin + (size_t)i * H, lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          u.v[k] = xo.v[k] > 0.f ? g.v[k] : 0.f;
          y.v[k] = fmaxf(fmaf(hr[r].v[k], bn.sc[k], bn.sh[k]), 0.f);
          dbias[k] += u.v[k];
        }
      }
      u.store(sU + lr * LDA, lane);
      y.store(sY + lr * LDA, lane);
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm_fast<VEC, kRPW>(sU, LDA, sW, H, H, sPart, acc);
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int i = row0 + warp * kRPW + r;
      if (i < N) {
        RowVec<VEC> o;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float yb = fmaf(hr[r].v[k], bn.sc[k], bn.sh[k]);
          const float m = yb > 0.f ? acc[r][k] : 0.f;
          o.v[k] = m;
          st[0][k] += (double)m;
          st[1][k] += (double)m * (double)bn.xhat(k, hr[r].v[k]);
        }
        o.store(DR + (size_t)i * H, lane);
      }
    }
    dW.accumulate(sU, LDA, sY, LDA, kTileRows);        // torch Linear: d W2 [out, in]
  }
  cp_async_wait_all();
  float* gp = c.gpart + c.gp_gin2[layer] + (size_t)blockIdx.x * (H * H + H);
  if (blockIdx.x < ntiles) dW.store(gp, H);
  block_colsum_store<VEC>(dbias, reinterpret_cast<float*>(sRed), blockIdx.x < ntiles ? gp + H * H : nullptr, H);
  block_totals<VEC, 2>(st, sRed, sTot, H, 0, H, 0);
  grid_sum_groups(c, bn_site(bn_id), sTot, 2 * H, gridDim.x, blockIdx.x);      // k_conv_bwd<GIN> finalises
}

// ---------------------------------------------------------------------------------------------
// Input transform backward: with g = relu'(x_1) * bn_1'(D_0):
//   M = xhat_0^T g  [F, H]  and  cs = colsum(g); the reduce kernel turns them into
//   d W_feat = gamma_0 * M + beta_0 (x) cs,  d gamma_0[f] = sum_j W[f][j] M[f][j],
//   d beta_0[f] = sum_j W[f][j] cs[j]          (bn_feat is the first op: no gradient beyond it).
// grid (G, ceil(F / 64)); smem: sG [R][H] | sX [R][64]
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_feat_bwd(const Ctx c) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  constexpr int FW = kFeatChunk / kRowWarps;       // feature columns per warp
  __shared__ __align__(16) float sG[kTileRows * H];
  __shared__ float sX[kTileRows * kFeatChunk];
  __shared__ float sRed[kRowWarps * H];
  const Dims d = load_dims(c);
  const int N = d.N, F = c.F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = blockIdx.y * kFeatChunk;
  const float* x1 = c.Xl(0);
  const float* D0 = c.D;
  const float* mean0 = c.bnf(0, BN_MEAN);
  const float* rstd0 = c.bnf(0, BN_RSTD);
  BnLane<VEC> b1;
  b1.load_bwd(c, c.model == CAL_MODEL_GIN ? kBnIdentity : 1, lane);     // CausalGIN has no BatchNorm on x_1 (model.py:237-238)
  if (c.model == CAL_MODEL_GCN) {
    // bns_conv[0]'s backward sums arrive as group vectors from k_conv_bwd(layer 0); CTA (0, 0) publishes
    __shared__ double s_scr[2 * H];
    __shared__ float s_c[2 * H];
    bn_bwd_from_groups(c, bn_site(1), 1, c.g_tile, 2 * H, 0, imin(imax(c.dims[0], 0), c.Nm), s_scr, s_c, s_c + H,
                       blockIdx.x == 0 && blockIdx.y == 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      b1.c1[i] = s_c[lane * VEC + i];
      b1.c2[i] = s_c[H + lane * VEC + i];
    }
  }
  float M[FW][VEC], cs[VEC];
#pragma unroll
  for (int a = 0; a < FW; ++a)
#pragma unroll
    for (int k = 0; k < VEC; ++k) M[a][k] = 0.f;
#pragma unroll
  for (int k = 0; k < VEC; ++k) cs[k] = 0.f;
  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int lr = warp * kRPW + r;
      const int i = row0 + lr;
      RowVec<VEC> g;
      g.zero();
      if (i < N) {
        RowVec<VEC> dv, xv;
        dv.load_coherent(D0 + (size_t)i * H, lane);
        xv.load_coherent(x1 + (size_t)i * H, lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          g.v[k] = xv.v[k] > 0.f ? b1.dx(k, dv.v[k], xv.v[k]) : 0.f;
          cs[k] += g.v[k];
        }
      }
      g.store(sG + lr * H, lane);
    }
    for (int i = threadIdx.x; i < kTileRows * kFeatChunk; i += blockDim.x) {
      int r = i / kFeatChunk, f = f0 + (i - r * kFeatChunk);
      float v = 0.f;
      if (row0 + r < N && f < F) v = (c.feat[(size_t)(row0 + r) * F + f] - mean0[f]) * rstd0[f];
      sX[i] = v;
    }
    __syncthreads();
    for (int r = 0; r < kTileRows; ++r) {
      RowVec<VEC> g;
      g.load_coherent(sG + r * H, lane);
#pragma unroll
      for (int a = 0; a < FW; ++a) {
        float x = sX[r * kFeatChunk + warp * FW + a];
#pragma unroll
        for (int k = 0; k < VEC; ++k) M[a][k] = fmaf(x, g.v[k], M[a][k]);
      }
    }
  }
  float* gp = c.gpart + c.gp_feat + (size_t)blockIdx.x * ((size_t)F * H + H);
#pragma unroll
  for (int a = 0; a < FW; ++a) {
    const int f = f0 + warp * FW + a;
    if (f < F) {
      RowVec<VEC> o;
#pragma unroll
      for (int k = 0; k < VEC; ++k) o.v[k] = M[a][k];
      o.store(gp + (size_t)f * H, lane);
    }
  }
  if (blockIdx.y == 0) block_colsum_store<VEC>(cs, sRed, gp + (size_t)F * H, H);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 32 * 1024) {       // static shared memory counts against the 48 KB default too
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace

int launch_feat_forward(const Ctx& c, cudaStream_t s) {
  note_launches(1);
  const int Fp = (c.F + 3) & ~3;
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = (size_t)Fp * c.H * 4 + (size_t)kTileRows * Fp * 4 + (size_t)kRowWarps * c.H * 8 + 2 * (size_t)Fp * 4;
    int rc = set_smem(k_feat_fwd<VEC>, smem);
    if (rc) return rc;
    launch_k(k_feat_fwd<VEC>, dim3(c.g_tile), dim3(256), smem, s, c);
  });
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_conv_forward(const Ctx& c, int layer, cudaStream_t s) {
  const bool last = layer == c.L - 1;
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = convf_smem_bytes<VEC>(kStageFwd);
    if (last) {
      int rc = set_smem(k_conv_fwd<VEC, 1>, smem);
      if (rc) return rc;
      launch_k(k_conv_fwd<VEC, 1>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
    } else {
      int rc = set_smem(k_conv_fwd<VEC, 0>, smem);
      if (rc) return rc;
      launch_k(k_conv_fwd<VEC, 0>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
    }
  });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_masked_forward(const Ctx& c, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = masked_both_smem_bytes<VEC>();
    int rc = set_smem(k_masked_fwd_both<VEC>, smem);
    if (rc) return rc;
    launch_k(k_masked_fwd_both<VEC>, dim3(c.g_tile), dim3(256), smem, s, c);
  });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_conv_backward(const Ctx& c, int layer, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = convb_smem_bytes<VEC>();
    int rc = set_smem(k_conv_bwd<VEC, false>, smem);
    if (rc) return rc;
    launch_k(k_conv_bwd<VEC, false>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
  });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_gin_forward(const Ctx& c, int layer, cudaStream_t s) {
  const bool last = layer == c.L - 1;
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = convf_smem_bytes<VEC>(kStageFwd);
    int rc = set_smem(k_conv_fwd<VEC, 3>, smem);
    if (rc) return rc;
    launch_k(k_conv_fwd<VEC, 3>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
    smem = (size_t)c.H * c.H * 4 + 5 * (size_t)kTileRows * (c.H + kPad) * 4 + (size_t)kRowWarps * c.H * 8;
    if (last) {
      rc = set_smem(k_gin_b_fwd<VEC, true>, smem);
      if (rc) return rc;
      launch_k(k_gin_b_fwd<VEC, true>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
    } else {
      rc = set_smem(k_gin_b_fwd<VEC, false>, smem);
      if (rc) return rc;
      launch_k(k_gin_b_fwd<VEC, false>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
    }
  });
  note_launches(2);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_gin_backward(const Ctx& c, int layer, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = (size_t)c.H * c.H * 4 + 6 * (size_t)kTileRows * (c.H + kPad) * 4 + (size_t)kRowWarps * c.H * 8;
    int rc = set_smem(k_gin_b_bwd<VEC>, smem);
    if (rc) return rc;
    launch_k(k_gin_b_bwd<VEC>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
    smem = convb_smem_bytes<VEC>();
    rc = set_smem(k_conv_bwd<VEC, true>, smem);
    if (rc) return rc;
    launch_k(k_conv_bwd<VEC, true>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
  });
  note_launches(2);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_masked_bwd_gemm(const Ctx& c, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = (size_t)c.H * c.H * 4 + 2 * (size_t)kTileRows * (c.H + kPad) * 4 + (size_t)kRowWarps * c.H * 8;
    int rc = set_smem(k_masked_bwd_gemm<VEC>, smem);
    if (rc) return rc;
    launch_k(k_masked_bwd_gemm<VEC>, dim3(c.g_tile, 2), dim3(256), smem, s, c);
  });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_feat_backward(const Ctx& c, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, {
    launch_k(k_feat_bwd<VEC>, dim3(c.g_tile, ceil_div(c.F, kFeatChunk)), dim3(256), 0, s, c);
  });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
