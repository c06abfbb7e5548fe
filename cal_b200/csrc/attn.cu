// attn.cu -- the causal / shortcut attention masks around the two masked convs.
//
// Forward (model.py:97-104): edge_att = softmax(edge_att_mlp([x_row || x_col])) is evaluated as
// softmax(p[row] + q[col] + b) from the per-node projections p = x W_e[:, :H]^T and
// q = x W_e[:, H:]^T written by the last backbone layer's epilogue, so the [E, 2H] edge
// representation is never materialised.  The same pass accumulates the attention-weighted degree
// by source row and deg^-1/2 for both branches (gcn_conv.py:59-70 with edge_weight).
//
// Backward: gradient of the weighted normalisation norm_e = dis[row] * w_e * dis[col] w.r.t. w
// (through both endpoints' degrees), the two 2-way softmaxes, the projections, and the masked
// BatchNorms.
#include "internal.cuh"

namespace cal {

namespace {

__device__ __forceinline__ int clampN(const Ctx& c) { return imin(imax(c.dims[0], 0), c.Nm); }

// Warp per source node, lanes over its out-edges: the dependent chain is out_ptr -> (out_dst, out_pos)
// -> pq[dst], three L2 round trips per node regardless of its degree, and every lane keeps its
// edge's weights in registers for the per-entry factors (no reload).
__global__ void __launch_bounds__(256) k_edge_att(const Ctx c) {
  pdl_sync();
  const int N = clampN(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float be0 = c.params[c.po.edge_att_b], be1 = c.params[c.po.edge_att_b + 1];
  for (int n = blockIdx.x * kRowWarps + warp; n < N; n += gridDim.x * kRowWarps) {
    const float4 pn = *reinterpret_cast<const float4*>(c.pq + (size_t)n * 4);
    const float2 an = *reinterpret_cast<const float2*>(c.natt + (size_t)n * 2);
    const int q0 = c.out_ptr[n], q1 = c.out_ptr[n + 1] - 1;       // last slot = appended loop
    float deg0 = 0.f, deg1 = 0.f;
    if (q1 - q0 + 1 <= 32) {
      // the common case: the whole row (loop included) in one pass, weights stay in registers
      const int q = q0 + lane;
      const bool live = q <= q1, loop = q == q1;
      int pos = 0;
      float w0 = 0.f, w1 = 0.f;
      if (live) {
        pos = c.out_pos[q];
        w0 = w1 = loop ? 1.f : 0.5f;
        if (!loop && !c.no_eatt) {
          const int d = c.out_dst[q];
          const float4 pd = *reinterpret_cast<const float4*>(c.pq + (size_t)d * 4);
          const float t0 = pn.x + pd.z + be0, t1 = pn.y + pd.w + be1;
          const float m = fmaxf(t0, t1);
          const float e0 = expf(t0 - m), e1 = expf(t1 - m);
          const float inv = 1.0f / (e0 + e1);
          w0 = e0 * inv;
          w1 = e1 * inv;
        }
      }
      deg0 = warp_sum(w0);
      deg1 = warp_sum(w1);
      const float2 dn = make_float2(1.0f / sqrtf(deg0), 1.0f / sqrtf(deg1));
      if (lane == 0) *reinterpret_cast<float2*>(c.disw + (size_t)n * 2) = dn;
      if (live) {
        *reinterpret_cast<float2*>(c.watt + (size_t)pos * 2) = make_float2(w0, w1);
        // per-entry factors of the masked convs' gather (so that it needs no per-source lookups)
        *reinterpret_cast<float2*>(c.edge_wn + (size_t)pos * 2) = make_float2(dn.x * w0, dn.y * w1);
        *reinterpret_cast<float2*>(c.edge_na + (size_t)pos * 2) = an;
      }
      continue;
    }
    // hub row: two strided passes
    for (int q = q0 + lane; q <= q1; q += 32) {
      const int pos = c.out_pos[q];
      float w0 = 1.f, w1 = 1.f;
      if (q < q1) {
        w0 = w1 = 0.5f;
        if (!c.no_eatt) {
          const int d = c.out_dst[q];
          const float4 pd = *reinterpret_cast<const float4*>(c.pq + (size_t)d * 4);
          const float t0 = pn.x + pd.z + be0, t1 = pn.y + pd.w + be1;
          const float m = fmaxf(t0, t1);
          const float e0 = expf(t0 - m), e1 = expf(t1 - m);
          const float inv = 1.0f / (e0 + e1);
          w0 = e0 * inv;
          w1 = e1 * inv;
        }
      }
      *reinterpret_cast<float2*>(c.watt + (size_t)pos * 2) = make_float2(w0, w1);
      deg0 += w0;
      deg1 += w1;
    }
    deg0 = warp_sum(deg0);
    deg1 = warp_sum(deg1);
    const float2 dn = make_float2(1.0f / sqrtf(deg0), 1.0f / sqrtf(deg1));
    if (lane == 0) *reinterpret_cast<float2*>(c.disw + (size_t)n * 2) = dn;
    for (int q = q0 + lane; q <= q1; q += 32) {
      const int pos = c.out_pos[q];
      const float2 w = *reinterpret_cast<const float2*>(c.watt + (size_t)pos * 2);   // written by this lane above
      *reinterpret_cast<float2*>(c.edge_wn + (size_t)pos * 2) = make_float2(dn.x * w.x, dn.y * w.y);
      *reinterpret_cast<float2*>(c.edge_na + (size_t)pos * 2) = an;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Masked convs backward, sparse part (warp per source row j, blockIdx.y = branch k):
//   dy_j   = sum_{e: row_e = j} norm_e * dagg[col_e]           (gradient w.r.t. bn_k output)
//   dnorm_e = <dagg[col_e], y_j>,  y_j = bn_k(att_k[j] * x_j)
//   + BatchNorm-backward sums of bnc / bno.
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_masked_bwd_gather(const Ctx c) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  __shared__ double sRed[kRowWarps * H];
  __shared__ double sTot[2 * H];
  const int N = clampN(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.y;
  const int bn_id = c.L + 1 + k;
  const float* X = c.Xl(c.L);
  const float* dagg = c.dagg + (size_t)k * c.Nm * H;
  float* dym = c.dym + (size_t)k * c.Nm * H;
  BnLane<VEC> bn;
  bn.load_bwd(c, bn_id, lane);
  double st[2][VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) st[0][i] = st[1][i] = 0.0;
  for (int j = blockIdx.x * kRowWarps + warp; j < N; j += gridDim.x * kRowWarps) {
    const float aj = c.natt[(size_t)j * 2 + k];
    const float dj = c.disw[(size_t)j * 2 + k];
    RowVec<VEC> x, y, xh, dy;
    x.load_coherent(X + (size_t)j * H, lane);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float xm = aj * x.v[i];
      y.v[i] = fmaf(xm, bn.sc[i], bn.sh[i]);
      xh.v[i] = bn.xhat(i, xm);
    }
    dy.zero();
    const int q0 = c.out_ptr[j], q1 = c.out_ptr[j + 1];
    for (int q = q0; q < q1; q += 2) {
      int dd[2], pos[2];
      float w[2];
      RowVec<VEC> g[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const bool ok = q + t < q1;
        dd[t] = ok ? c.out_dst[q + t] : j;
        pos[t] = ok ? c.out_pos[q + t] : -1;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        g[t].load_coherent(dagg + (size_t)dd[t] * H, lane);
        w[t] = pos[t] >= 0 ? (dj * c.watt[(size_t)pos[t] * 2 + k]) * c.disw[(size_t)dd[t] * 2 + k] : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          dy.v[i] = fmaf(w[t], g[t].v[i], dy.v[i]);
          dot = fmaf(g[t].v[i], y.v[i], dot);
        }
        dot = warp_sum(dot);
        if (lane == 0 && pos[t] >= 0) c.dnrm[(size_t)pos[t] * 2 + k] = dot;
      }
    }
    dy.store(dym + (size_t)j * H, lane);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      st[0][i] += (double)dy.v[i];
      st[1][i] += (double)dy.v[i] * (double)xh.v[i];
    }
  }
  block_totals<VEC, 2>(st, sRed, sTot, H, 0, H, 0);
  grid_sum_groups(c, k, sTot, 2 * H, gridDim.x, blockIdx.x);       // k_att_bwd finalises bnc / bno (site = branch)
}

// ---------------------------------------------------------------------------------------------
// Weighted-norm backward (thread per node): d deg^-1/2, d deg, then for every out-edge of the
// node d w_e (both branches), the edge softmax backward dt_e, and dp[n] = sum_e dt_e.
// ---------------------------------------------------------------------------------------------
// Warp per node, lanes over its out- and in-edges (three dependent L2 round trips per node).
__global__ void __launch_bounds__(256) k_norm_bwd(const Ctx c) {
  pdl_sync();
  const int N = clampN(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = blockIdx.x * kRowWarps + warp; n < N; n += gridDim.x * kRowWarps) {
    const int q0 = c.out_ptr[n], q1 = c.out_ptr[n + 1] - 1;
    const int p0 = c.in_ptr[n], p1 = c.in_ptr[n + 1] - 1;        // p1 = the appended loop
    if (c.no_eatt) {
      for (int q = q0 + lane; q <= q1; q += 32)
        *reinterpret_cast<float2*>(c.dt + (size_t)c.out_pos[q] * 2) = make_float2(0.f, 0.f);
      if (lane == 0) *reinterpret_cast<float2*>(c.dp + (size_t)n * 2) = make_float2(0.f, 0.f);
      continue;
    }
    const float2 dn = *reinterpret_cast<const float2*>(c.disw + (size_t)n * 2);
    // this lane's first out-edge stays in registers for the second half (rows longer than 32 reload)
    int pos_r = -1;
    float2 g_r = make_float2(0.f, 0.f), w_r = g_r, dsd_r = g_r;
    float dd0 = 0.f, dd1 = 0.f;                                   // d dis[n]
    for (int q = q0 + lane; q < q1; q += 32) {
      const int pos = c.out_pos[q], d = c.out_dst[q];
      const float2 g = *reinterpret_cast<const float2*>(c.dnrm + (size_t)pos * 2);
      const float2 w = *reinterpret_cast<const float2*>(c.watt + (size_t)pos * 2);
      const float2 dsd = *reinterpret_cast<const float2*>(c.disw + (size_t)d * 2);
      if (q == q0 + lane) {
        pos_r = pos; g_r = g; w_r = w; dsd_r = dsd;
      }
      dd0 = fmaf(g.x * w.x, dsd.x, dd0);
      dd1 = fmaf(g.y * w.y, dsd.y, dd1);
    }
    for (int p = p0 + lane; p < p1; p += 32) {
      const int s = c.in_src[p];
      const float2 g = *reinterpret_cast<const float2*>(c.dnrm + (size_t)p * 2);
      const float2 w = *reinterpret_cast<const float2*>(c.watt + (size_t)p * 2);
      const float2 dss = *reinterpret_cast<const float2*>(c.disw + (size_t)s * 2);
      dd0 = fmaf(g.x * w.x, dss.x, dd0);
      dd1 = fmaf(g.y * w.y, dss.y, dd1);
    }
    if (lane == 0) {
      const float2 g = *reinterpret_cast<const float2*>(c.dnrm + (size_t)p1 * 2);
      dd0 = fmaf(2.f * dn.x, g.x, dd0);
      dd1 = fmaf(2.f * dn.y, g.y, dd1);
    }
    dd0 = warp_sum(dd0);
    dd1 = warp_sum(dd1);
    const float ddeg0 = -0.5f * dn.x * dn.x * dn.x * dd0;         // d deg = -1/2 deg^-3/2 d dis
    const float ddeg1 = -0.5f * dn.y * dn.y * dn.y * dd1;
    float dp0 = 0.f, dp1 = 0.f;
    for (int q = q0 + lane; q < q1; q += 32) {
      int pos = pos_r;
      float2 g = g_r, w = w_r, dsd = dsd_r;
      if (q != q0 + lane) {
        pos = c.out_pos[q];
        const int d = c.out_dst[q];
        g = *reinterpret_cast<const float2*>(c.dnrm + (size_t)pos * 2);
        w = *reinterpret_cast<const float2*>(c.watt + (size_t)pos * 2);
        dsd = *reinterpret_cast<const float2*>(c.disw + (size_t)d * 2);
      }
      const float dw0 = fmaf(g.x * dn.x, dsd.x, ddeg0);
      const float dw1 = fmaf(g.y * dn.y, dsd.y, ddeg1);
      const float dot = w.x * dw0 + w.y * dw1;
      const float dt0 = w.x * (dw0 - dot), dt1 = w.y * (dw1 - dot);
      *reinterpret_cast<float2*>(c.dt + (size_t)pos * 2) = make_float2(dt0, dt1);
      dp0 += dt0;
      dp1 += dt1;
    }
    dp0 = warp_sum(dp0);
    dp1 = warp_sum(dp1);
    if (lane == 0) {
      *reinterpret_cast<float2*>(c.dt + (size_t)c.out_pos[q1] * 2) = make_float2(0.f, 0.f);
      *reinterpret_cast<float2*>(c.dp + (size_t)n * 2) = make_float2(dp0, dp1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Attention backward (warp per node row): gathers dq[n] = sum_{e: col_e = n} dt_e, applies the
// bnc / bno backward, the node softmax backward, and forms the gradient w.r.t. x_{L+1}:
//   dX = a0 g_c + a1 g_o + ds W_n + dp W_e[:, :H] + dq W_e[:, H:]
// plus per-CTA partials of the node_att_mlp / edge_att_mlp gradients.
// ---------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256, 2) k_att_bwd(const Ctx c) {
  pdl_sync();
  constexpr int H = 32 * VEC;
  __shared__ float sRed[kRowWarps * H];
  __shared__ float sScal[kRowWarps][4];
  const int N = clampN(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* X = c.Xl(c.L);
  const float* dyc = c.dym;
  const float* dyo = c.dym + (size_t)c.Nm * H;
  float* Dout = c.D + (size_t)(c.L & 1) * c.Nm * H;
  BnLane<VEC> bc, bo;
  bc.load_bwd(c, c.L + 1, lane);
  bo.load_bwd(c, c.L + 2, lane);
  {
    // backward sums of bnc / bno: group vectors left by k_masked_bwd_gather (sites 0 / 1, grid g_row)
    __shared__ double s_scr[4 * H];
    __shared__ float s_c[4 * H];
    bn_bwd_from_groups2(c, 0, c.L + 1, 1, c.L + 2, c.g_row, 2 * H, N, s_scr, s_c, s_c + H, s_c + 2 * H, s_c + 3 * H,
                        blockIdx.x == 0);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      bc.c1[i] = s_c[lane * VEC + i];
      bc.c2[i] = s_c[H + lane * VEC + i];
      bo.c1[i] = s_c[2 * H + lane * VEC + i];
      bo.c2[i] = s_c[3 * H + lane * VEC + i];
    }
  }
  float wn0[VEC], wn1[VEC], wp0[VEC], wp1[VEC], wq0[VEC], wq1[VEC];
  {
    const float* Wn = c.params + c.po.node_att_w;
    const float* We = c.params + c.po.edge_att_w;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      int k = lane * VEC + i;
      wn0[i] = Wn[k];
      wn1[i] = Wn[H + k];
      wp0[i] = We[k];
      wp1[i] = We[2 * H + k];
      wq0[i] = We[H + k];
      wq1[i] = We[3 * H + k];
    }
  }
  float g_wn0[VEC], g_wn1[VEC], g_wp0[VEC], g_wp1[VEC], g_wq0[VEC], g_wq1[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) g_wn0[i] = g_wn1[i] = g_wp0[i] = g_wp1[i] = g_wq0[i] = g_wq1[i] = 0.f;
  float g_bn0 = 0.f, g_bn1 = 0.f, g_be0 = 0.f, g_be1 = 0.f;

  for (int n = blockIdx.x * kRowWarps + warp; n < N; n += gridDim.x * kRowWarps) {
    // dq: lanes stride over the in-edges (appended loop excluded: its dt is zero anyway)
    float dq0 = 0.f, dq1 = 0.f;
    for (int p = c.in_ptr[n] + lane; p < c.in_ptr[n + 1]; p += 32) {
      const float2 t = *reinterpret_cast<const float2*>(c.dt + (size_t)p * 2);
      dq0 += t.x;
      dq1 += t.y;
    }
    dq0 = warp_sum(dq0);
    dq1 = warp_sum(dq1);
    const float2 dpn = *reinterpret_cast<const float2*>(c.dp + (size_t)n * 2);
    const float2 a = *reinterpret_cast<const float2*>(c.natt + (size_t)n * 2);
    RowVec<VEC> x, gc, go;
    x.load_coherent(X + (size_t)n * H, lane);
    gc.load_coherent(dyc + (size_t)n * H, lane);
    go.load_coherent(dyo + (size_t)n * H, lane);
    float da0 = 0.f, da1 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      gc.v[i] = bc.dx(i, gc.v[i], a.x * x.v[i]);
      go.v[i] = bo.dx(i, go.v[i], a.y * x.v[i]);
      da0 = fmaf(gc.v[i], x.v[i], da0);
      da1 = fmaf(go.v[i], x.v[i], da1);
    }
    da0 = warp_sum(da0);
    da1 = warp_sum(da1);
    float ds0 = 0.f, ds1 = 0.f;
    if (!c.no_natt) {
      const float dot = a.x * da0 + a.y * da1;
      ds0 = a.x * (da0 - dot);
      ds1 = a.y * (da1 - dot);
    }
    RowVec<VEC> o;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float v = a.x * gc.v[i] + a.y * go.v[i];
      v = fmaf(ds0, wn0[i], v);
      v = fmaf(ds1, wn1[i], v);
      v = fmaf(dpn.x, wp0[i], v);
      v = fmaf(dpn.y, wp1[i], v);
      v = fmaf(dq0, wq0[i], v);
      v = fmaf(dq1, wq1[i], v);
      o.v[i] = v;
      g_wn0[i] = fmaf(ds0, x.v[i], g_wn0[i]);
      g_wn1[i] = fmaf(ds1, x.v[i], g_wn1[i]);
      g_wp0[i] = fmaf(dpn.x, x.v[i], g_wp0[i]);
      g_wp1[i] = fmaf(dpn.y, x.v[i], g_wp1[i]);
      g_wq0[i] = fmaf(dq0, x.v[i], g_wq0[i]);
      g_wq1[i] = fmaf(dq1, x.v[i], g_wq1[i]);
    }
    o.store(Dout + (size_t)n * H, lane);
    g_bn0 += ds0;
    g_bn1 += ds1;
    g_be0 += dpn.x;
    g_be1 += dpn.y;
  }
  if (blockIdx.x * kRowWarps < N) {
    float* gp = c.gpart + c.gp_att + (size_t)blockIdx.x * (8 * H + 4);
    // layout: node_att_w [2][H] | edge_att_w [2][2H] | node_att_b [2] | edge_att_b [2]
    block_colsum_store<VEC>(g_wn0, sRed, gp, H);
    block_colsum_store<VEC>(g_wn1, sRed, gp + H, H);
    block_colsum_store<VEC>(g_wp0, sRed, gp + 2 * H, H);
    block_colsum_store<VEC>(g_wq0, sRed, gp + 3 * H, H);
    block_colsum_store<VEC>(g_wp1, sRed, gp + 4 * H, H);
    block_colsum_store<VEC>(g_wq1, sRed, gp + 5 * H, H);
    __syncthreads();
    if (lane == 0) {
      sScal[warp][0] = g_bn0;
      sScal[warp][1] = g_bn1;
      sScal[warp][2] = g_be0;
      sScal[warp][3] = g_be1;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      float s = 0.f;
      for (int w = 0; w < kRowWarps; ++w) s += sScal[w][threadIdx.x];
      gp[6 * H + threadIdx.x] = s;
    }
  }
}

}  // namespace

int launch_edge_att(const Ctx& c, cudaStream_t s) {
  launch_k(k_edge_att, dim3(imax(1, imin(ceil_div(c.Nm, kRowWarps), 8 * kSMs))), dim3(256), 0, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_masked_bwd_gather(const Ctx& c, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, { launch_k(k_masked_bwd_gather<VEC>, dim3(c.g_row, 2), dim3(256), 0, s, c); });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_norm_backward(const Ctx& c, cudaStream_t s) {
  launch_k(k_norm_bwd, dim3(imax(1, imin(ceil_div(c.Nm, kRowWarps), 8 * kSMs))), dim3(256), 0, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_att_backward(const Ctx& c, cudaStream_t s) {
  CAL_DISPATCH_VEC(c.H, { launch_k(k_att_bwd<VEC>, dim3(c.g_row), dim3(256), 0, s, c); });
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
