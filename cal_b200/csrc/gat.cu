// gat.cu -- GATConv backbone layers of CausalGAT (model.py:340, 388-390): PyG 1.x
// GATConv(hidden, hidden / heads, heads, dropout) with concat=True, negative_slope 0.2.
//
//   x'    = bn_l(x) W                                      [N, heads * Ch]
//   e_p   = leaky_relu_0.2(<x'_i, a_i> + <x'_j, a_j>)      p = in-edge (j -> i) incl. the appended loop
//   alpha = exp(e - max_i) / (sum_i exp(e - max_i) + 1e-16)   per target i and head
//   out_i = concat_h sum_p alpha_p keep_p x'_j + bias      keep = attention dropout mask / (1 - p)
//
// Forward: k_gat_lin (tile GEMM + the two per-node attention scalars per head) and k_gat_agg
// (warp per target row: max / sum / normalise / aggregate in registers, no [E', heads, Ch]
// temporaries, no atomics).  Backward: k_gat_bwd_edge (warp per target: softmax + leaky-relu
// backward per in-edge) and k_gat_bwd_node (tile kernel by source row: transpose aggregate,
// attention-vector terms, dX = dX' W^T, dW, BatchNorm-backward sums), mirroring k_conv_bwd.
#include "internal.cuh"

namespace cal {

namespace {

struct GatBufs {
  float *xp, *asrc, *adst, *alpha;   // per layer: x' [Nm][H], a_src/a_dst [Nm][heads], alpha [EP][heads]
  float *dz, *dadst;                 // shared: d(pre-leaky-relu logit) [EP][heads], d a_dst [Nm][heads]
};

__host__ __device__ inline size_t up4(size_t x) { return (x + 3) & ~(size_t)3; }

__host__ __device__ inline GatBufs gat_bufs(const Ctx& c, int layer) {
  const size_t nh = up4((size_t)c.Nm * c.heads), eh = up4((size_t)c.EP * c.heads), xh = (size_t)c.Nm * c.H;
  const size_t per_layer = xh + 2 * nh + eh;
  GatBufs b;
  float* base = c.gat + (size_t)layer * per_layer;
  b.xp = base;
  b.asrc = base + xh;
  b.adst = b.asrc + nh;
  b.alpha = b.adst + nh;
  float* sh = c.gat + (size_t)c.L * per_layer;
  b.dz = sh;
  b.dadst = sh + eh;
  return b;
}

__device__ __forceinline__ int clampN(const Ctx& c) { return imin(imax(c.dims[0], 0), c.Nm); }

// sum over the lanes that own one head (a contiguous group of 32 / heads lanes)
__device__ __forceinline__ float head_sum(float v, int lanes_per_head) {
  for (int o = lanes_per_head >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// attention vector entries of a lane's channels: att is [1, heads, 2 * Ch] (a_i | a_j per head)
template <int VEC>
__device__ __forceinline__ void load_att(const Ctx& c, int layer, int lane, float (&ai)[VEC], float (&aj)[VEC]) {
  const int Ch = c.H / c.heads;
  const float* att = c.params + c.po.convs_att[layer];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int ch = lane * VEC + i, h = ch / Ch, cc = ch - h * Ch;
    ai[i] = att[(size_t)h * 2 * Ch + cc];
    aj[i] = att[(size_t)h * 2 * Ch + Ch + cc];
  }
}

// ---------------------------------------------------------------------------------------------
// x' = bn_l(x_l) W_l and the per-node attention scalars a_dst = <x', a_i>, a_src = <x', a_j>.
// smem: sW [H][H] | sA [R][H]
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_gat_lin(const Ctx c, const int layer) {
  constexpr int H = 32 * VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sA = sW + H * H;
  const int N = clampN(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatBufs g = gat_bufs(c, layer);
  stage_matrix_async(sW, c.params + c.po.convs_w[layer], H * H);
  pdl_sync();                                        // everything below may read the predecessor's output
  BnLane<VEC> bn;
  bn.load_fwd(c, 1 + layer, lane);
  float ai[VEC], aj[VEC];
  load_att<VEC>(c, layer, lane, ai, aj);
  const int heads = c.heads, lph = 32 / heads, head = lane / lph;
  const float* xin = c.Xl(layer);
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
        a.load_coherent(xin + (size_t)i * H, lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) a.v[k] = fmaf(a.v[k], bn.sc[k], bn.sh[k]);
      }
      a.store(sA + lr * H, lane);
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm<VEC, kRPW>(sA, H, sW, H, H, acc);
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int i = row0 + warp * kRPW + r;
      if (i < N) {
        RowVec<VEC> o;
        float sd = 0.f, ss = 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          o.v[k] = acc[r][k];
          sd = fmaf(acc[r][k], ai[k], sd);
          ss = fmaf(acc[r][k], aj[k], ss);
        }
        o.store(g.xp + (size_t)i * H, lane);
        sd = head_sum(sd, lph);
        ss = head_sum(ss, lph);
        if (lane % lph == 0) {
          g.adst[(size_t)i * heads + head] = sd;
          g.asrc[(size_t)i * heads + head] = ss;
        }
      }
    }
  }
  cp_async_wait_all();
}

// dropout keep factor of in-edge with key `key` (edge_index column, or E + node for the loop)
__device__ __forceinline__ float keep_of(const Ctx& c, int layer, int EN, int key, int head) {
  if (c.gat_keep == nullptr || !c.train) return 1.f;
  return c.gat_keep[((size_t)layer * EN + key) * c.heads + head];
}

// ---------------------------------------------------------------------------------------------
// Per-target softmax over the in-edges + weighted aggregate + bias + ReLU (+ the layer epilogue).
// Warp per target row; a lane works on the head that owns its channels.
// ---------------------------------------------------------------------------------------------
template <int VEC, bool LASTL>
__global__ void __launch_bounds__(256) k_gat_agg(const Ctx c, const int layer) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  __shared__ double sRed[kRowWarps * H];
  __shared__ double sTot[4 * H];
  const int N = clampN(c);
  const int EN = imin(imax(c.dims[1], 0), c.Em) + N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatBufs g = gat_bufs(c, layer);
  const int heads = c.heads, lph = 32 / heads, head = lane / lph;
  const float* bias = c.params + c.po.convs_b[layer];
  float bv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) bv[i] = bias[lane * VEC + i];
  LayerEpilogue<VEC, LASTL> epi;
  epi.init(c, lane);
  float* xout = c.Xl(layer + 1);
  for (int i = blockIdx.x * kRowWarps + warp; i < N; i += gridDim.x * kRowWarps) {
    const int p0 = c.in_ptr[i], p1 = c.in_ptr[i + 1];
    const float ad = g.adst[(size_t)i * heads + head];
    float m = -INFINITY;
    for (int p = p0; p < p1; ++p) {
      const float z = ad + g.asrc[(size_t)c.in_src[p] * heads + head];
      m = fmaxf(m, z > 0.f ? z : 0.2f * z);
    }
    float s = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float z = ad + g.asrc[(size_t)c.in_src[p] * heads + head];
      s += expf((z > 0.f ? z : 0.2f * z) - m);
    }
    const float inv = 1.0f / (s + 1e-16f);
    RowVec<VEC> a;
    a.zero();
    for (int p = p0; p < p1; ++p) {
      const int src = c.in_src[p];
      const float z = ad + g.asrc[(size_t)src * heads + head];
      const float al = expf((z > 0.f ? z : 0.2f * z) - m) * inv;
      if (lane % lph == 0) g.alpha[(size_t)p * heads + head] = al;
      const float w = al * keep_of(c, layer, EN, c.in_key[p], head);
      RowVec<VEC> v;
      v.load_coherent(g.xp + (size_t)src * H, lane);
#pragma unroll
      for (int k = 0; k < VEC; ++k) a.v[k] = fmaf(w, v.v[k], a.v[k]);
    }
    RowVec<VEC> o;
#pragma unroll
    for (int k = 0; k < VEC; ++k) o.v[k] = fmaxf(a.v[k] + bv[k], 0.f);
    o.store(xout + (size_t)i * H, lane);
    epi.row(c, i, o.v, lane);
  }
  epi.finish(c, layer, sRed, sTot, N);
}

// ---------------------------------------------------------------------------------------------
// Backward, per target row i (warp per row).  g_i = relu'(x_{l+2,i}) * bn_up'(D_up,i) is the
// gradient w.r.t. the layer output; for every in-edge p = (j -> i) and head h:
//   d alpha~_p = <g_i[h], x'_j[h]>,  d alpha_p = d alpha~_p * keep_p
//   d e_p = alpha_p (d alpha_p - sum_q alpha_q d alpha_q)        (softmax with the +1e-16 denominator)
//   d z_p = d e_p * (z_p > 0 ? 1 : 0.2)                           (leaky relu)
// writes dz[p][h] and d a_dst[i][h] = sum_p d z_p.
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_gat_bwd_edge(const Ctx c, const int layer) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  const int N = clampN(c);
  const int EN = imin(imax(c.dims[1], 0), c.Em) + N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatBufs g = gat_bufs(c, layer);
  const int heads = c.heads, lph = 32 / heads, head = lane / lph;
  const int bn_up = layer == c.L - 1 ? kBnIdentity : 2 + layer;
  BnLane<VEC> bu;
  bu.load_bwd(c, bn_up, lane);
  const float* xup = c.Xl(layer + 1);
  const float* Dup = c.D + (size_t)((layer + 1) & 1) * c.Nm * H;
  for (int i = blockIdx.x * kRowWarps + warp; i < N; i += gridDim.x * kRowWarps) {
    RowVec<VEC> go, xo;
    go.load_coherent(Dup + (size_t)i * H, lane);
    xo.load_coherent(xup + (size_t)i * H, lane);
    float gz[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) gz[k] = xo.v[k] > 0.f ? bu.dx(k, go.v[k], xo.v[k]) : 0.f;
    const int p0 = c.in_ptr[i], p1 = c.in_ptr[i + 1];
    const float ad = g.adst[(size_t)i * heads + head];
    float dot = 0.f;                               // sum_q alpha_q d alpha_q of this lane's head
    for (int p = p0; p < p1; ++p) {
      const int src = c.in_src[p];
      RowVec<VEC> v;
      v.load_coherent(g.xp + (size_t)src * H, lane);
      float da = 0.f;
#pragma unroll
      for (int k = 0; k < VEC; ++k) da = fmaf(gz[k], v.v[k], da);
      da = head_sum(da, lph) * keep_of(c, layer, EN, c.in_key[p], head);
      dot = fmaf(g.alpha[(size_t)p * heads + head], da, dot);
      if (lane % lph == 0) g.dz[(size_t)p * heads + head] = da;     // parked until the second pass
    }
    __syncwarp();
    float dsum = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float da = g.dz[(size_t)p * heads + head];
      const float al = g.alpha[(size_t)p * heads + head];
      const float z = ad + g.asrc[(size_t)c.in_src[p] * heads + head];
      const float dz = al * (da - dot) * (z > 0.f ? 1.f : 0.2f);
      dsum += dz;
      __syncwarp();
      if (lane % lph == 0) g.dz[(size_t)p * heads + head] = dz;
    }
    if (lane % lph == 0) g.dadst[(size_t)i * heads + head] = dsum;
  }
}

// ---------------------------------------------------------------------------------------------
// Backward, per source row j (32-row tiles, the GAT counterpart of k_conv_bwd):
//   dX'_j = sum_{q: j -> i} alpha_q keep_q g_i  +  d a_dst[j] a_i  +  (sum_q d z_q) a_j
//   D_j   = dX'_j W^T,   dW += bn_l(x_j)^T dX'_j,   d att += (d a_dst[j] x'_j | d a_src[j] x'_j),
//   db   += g_j, and the BatchNorm-backward sums of bn_l.
// smem: sW [H][H] (W^T) | sU [R][H] | sY [R][H] | sRed f64 [8][H] | sPtr
// ---------------------------------------------------------------------------------------------
template <int VEC>
constexpr size_t gatb_smem_bytes() {
  constexpr int H = 32 * VEC;
  return (size_t)H * H * 4 + 2 * (size_t)kTileRows * H * 4 + (size_t)kRowWarps * H * 8 + (kTileRows + 4) * 4;
}

template <int VEC>
__global__ void __launch_bounds__(256) k_gat_bwd_node(const Ctx c, const int layer) {
  constexpr int H = 32 * VEC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double sTot[2 * H];
  const int N = clampN(c);
  const int EN = imin(imax(c.dims[1], 0), c.Em) + N;
  float* sW = reinterpret_cast<float*>(smem_raw);
  float* sU = sW + H * H;
  float* sY = sU + kTileRows * H;
  double* sRed = reinterpret_cast<double*>(sY + kTileRows * H);
  int* sPtr = reinterpret_cast<int*>(sRed + kRowWarps * H);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GatBufs g = gat_bufs(c, layer);
  const int heads = c.heads, lph = 32 / heads, head = lane / lph;

  stage_matrix_async(sW, c.wt_conv(layer), H * H);
  pdl_sync();                                        // everything below may read the predecessor's output

  const int bn_in = 1 + layer;
  const int bn_up = layer == c.L - 1 ? kBnIdentity : 2 + layer;
  const float* xin = c.Xl(layer);
  const float* xup = c.Xl(layer + 1);
  const float* Dup = c.D + (size_t)((layer + 1) & 1) * c.Nm * H;
  float* Dout = c.D + (size_t)(layer & 1) * c.Nm * H;
  BnLane<VEC> bi, bu;
  bi.load_bwd(c, bn_in, lane);
  bu.load_bwd(c, bn_up, lane);
  float ai[VEC], aj[VEC];
  load_att<VEC>(c, layer, lane, ai, aj);

  OuterAcc<H> dW;
  dW.zero();
  float dbias[VEC], dai[VEC], daj[VEC];
  double st[2][VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    dbias[k] = dai[k] = daj[k] = 0.f;
    st[0][k] = st[1][k] = 0.0;
  }

  const int ntiles = ceil_div(N, kTileRows);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = tile * kTileRows;
    const int nrows = imin(kTileRows, N - row0);
    __syncthreads();
    if (threadIdx.x <= nrows) sPtr[threadIdx.x] = c.out_ptr[row0 + threadIdx.x];
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < kRPW; ++r) {
      const int lr = warp * kRPW + r;
      const int j = row0 + lr;
      RowVec<VEC> u, y;
      u.zero();
      y.zero();
      if (lr < nrows) {
        const int q0 = sPtr[lr], q1 = sPtr[lr + 1];
        float dasrc = 0.f;
        for (int q = q0; q < q1; ++q) {
          const int dd = c.out_dst[q], pos = c.out_pos[q];
          const float w = g.alpha[(size_t)pos * heads + head] * keep_of(c, layer, EN, c.out_key[q], head);
          dasrc += g.dz[(size_t)pos * heads + head];
          RowVec<VEC> gv, xv;
          gv.load_coherent(Dup + (size_t)dd * H, lane);
          xv.load_coherent(xup + (size_t)dd * H, lane);
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float gz = xv.v[k] > 0.f ? bu.dx(k, gv.v[k], xv.v[k]) : 0.f;
            u.v[k] = fmaf(w, gz, u.v[k]);
          }
        }
        const float dadst = g.dadst[(size_t)j * heads + head];
        RowVec<VEC> xi, go, xo, xp;
        xi.load_coherent(xin + (size_t)j * H, lane);
        go.load_coherent(Dup + (size_t)j * H, lane);
        xo.load_coherent(xup + (size_t)j * H, lane);
        xp.load_coherent(g.xp + (size_t)j * H, lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          u.v[k] = fmaf(dadst, ai[k], u.v[k]);
          u.v[k] = fmaf(dasrc, aj[k], u.v[k]);
          dai[k] = fmaf(dadst, xp.v[k], dai[k]);
          daj[k] = fmaf(dasrc, xp.v[k], daj[k]);
          y.v[k] = fmaf(xi.v[k], bi.sc[k], bi.sh[k]);
          dbias[k] += xo.v[k] > 0.f ? bu.dx(k, go.v[k], xo.v[k]) : 0.f;
        }
      }
      u.store(sU + lr * H, lane);
      y.store(sY + lr * H, lane);
    }
    cp_async_wait_all();
    __syncthreads();
    float acc[kRPW][VEC];
#pragma unroll
    for (int r = 0; r < kRPW; ++r)
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[r][k] = 0.f;
    tile_gemm<VEC, kRPW>(sU, H, sW, H, H, acc);
#pragma unroll
    for (int r = 0; r < kRPW; ++r) {
      const int j = row0 + warp * kRPW + r;
      if (j < N) {
        RowVec<VEC> o, xi;
        xi.load_coherent(xin + (size_t)j * H, lane);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          o.v[k] = acc[r][k];
          st[0][k] += (double)acc[r][k];
          st[1][k] += (double)acc[r][k] * (double)bi.xhat(k, xi.v[k]);
        }
        o.store(Dout + (size_t)j * H, lane);
      }
    }
    dW.accumulate(sY, H, sU, H, kTileRows);
  }
  cp_async_wait_all();
  const bool wrote = (int)blockIdx.x < ntiles;
  float* gp = c.gpart + c.gp_conv[layer] + (size_t)blockIdx.x * (H * H + H);
  if (wrote) dW.store(gp, H);
  block_colsum_store<VEC>(dbias, reinterpret_cast<float*>(sRed), wrote ? gp + H * H : nullptr, H);
  // attention-vector gradient partials in the att layout [heads][a_i (Ch) | a_j (Ch)]
  {
    float* sbuf = reinterpret_cast<float*>(sRed);
    float* ga = c.gpart + c.gp_gat[layer] + (size_t)blockIdx.x * 2 * H;
    const int Ch = H / heads;
    for (int part = 0; part < 2; ++part) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < VEC; ++k) sbuf[warp * H + lane * VEC + k] = part == 0 ? dai[k] : daj[k];
      __syncthreads();
      for (int ch = threadIdx.x; ch < H; ch += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kRowWarps; ++w) s += sbuf[w * H + ch];
        const int h = ch / Ch, cc = ch - h * Ch;
        if (wrote) ga[(size_t)h * 2 * Ch + part * Ch + cc] = s;
      }
    }
  }
  block_totals<VEC, 2>(st, sRed, sTot, H, 0, H, 0);
  if (grid_sum(c, 0, sTot, 2 * H, gridDim.x, blockIdx.x)) bn_bwd_finalize_tot(c, bn_in, sTot, sTot + H, N);
}

template <typename K>
int set_smem_g(K kernel, size_t bytes) {
  if (bytes > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

}  // namespace

size_t gat_workspace_floats(int Nm, int EP, int H, int L, int heads) {
  const size_t nh = up4((size_t)Nm * heads), eh = up4((size_t)EP * heads), xh = (size_t)Nm * H;
  return (size_t)L * (xh + 2 * nh + eh) + eh + nh;
}

int launch_gat_forward(const Ctx& c, int layer, cudaStream_t s) {
  const bool last = layer == c.L - 1;
  CAL_DISPATCH_VEC(c.H, {
    size_t smem = (size_t)c.H * c.H * 4 + (size_t)kTileRows * c.H * 4;
    int rc = set_smem_g(k_gat_lin<VEC>, smem);
    if (rc) return rc;
    launch_k(k_gat_lin<VEC>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
    if (last) launch_k(k_gat_agg<VEC, true>, dim3(c.g_row), dim3(256), 0, s, c, layer);
    else launch_k(k_gat_agg<VEC, false>, dim3(c.g_row), dim3(256), 0, s, c, layer);
  });
  note_launches(2);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_gat_backward(const Ctx& c, int layer, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, {
    launch_k(k_gat_bwd_edge<VEC>, dim3(c.g_row), dim3(256), 0, s, c, layer);
    size_t smem = gatb_smem_bytes<VEC>();
    int rc = set_smem_g(k_gat_bwd_node<VEC>, smem);
    if (rc) return rc;
    launch_k(k_gat_bwd_node<VEC>, dim3(c.g_tile), dim3(256), smem, s, c, layer);
  });
  note_launches(2);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
