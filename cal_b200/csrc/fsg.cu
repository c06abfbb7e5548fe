// fsg.cu -- the fused small-graph forward pass of CausalGCN (model.py:85-116): ONE persistent kernel from
// the input transform to the pooled graph embeddings.
//
// A CTA owns a block of whole graphs (<= kFsgRows nodes, <= kFsgEntries CSR entries; fsg.cuh) for the whole
// pass.  Its node features stay in shared memory from layer to layer, so message passing (gcn_conv.py:92-97)
// never leaves the SM; the node transforms x W (gcn_conv.py:75) run on the tensor cores (tcgen05.mma, 3xTF32,
// accumulators in TMEM) with the weight matrix as the M operand and the block's node rows on the N dimension
// (cost proportional to the number of rows, granularity 8); the weight operand of the next layer arrives by
// one cp.async.bulk of a pre-split image while the current layer's epilogue runs.  Training-mode BatchNorm is
// the only coupling between CTAs: per BatchNorm boundary one deterministic in-kernel all-reduce of the
// per-channel (sum, sum of squares) -- groups of 8 CTAs, fixed order, no float atomics -- whose latency is
// overlapped with the RAW aggregation of the next layer:  sum_e norm_e bn(x)[src_e] = sc * (sum_e norm_e x[src_e])
// + sh * (sum_e norm_e), so only an affine per row remains once the statistics arrive.
//
// The kernel writes the same workspace regions as the tiled kernels it replaces (conv.cu k_feat_fwd /
// k_conv_fwd / k_masked_fwd_both, attn.cu k_edge_att, head.cu k_pool), so the backward pass and the tests
// are agnostic of which forward ran.
#include "fsg_dev.cuh"

namespace cal {
namespace {

// ---- shared memory ----
struct FsgSmem {
  size_t a_hi, a_lo, b_hi, b_lo, x, in_ptr, out_ptr, in_src, in_nrm, out_dst, out_pos, w, wn, gptr, rowf, aff, part, tot, total;
};
__host__ __device__ inline FsgSmem fsg_smem() {
  FsgSmem s;
  size_t o = 0;
  s.a_hi = o;    o += 65536;
  s.a_lo = o;    o += 65536;
  s.b_hi = o;    o += kBPart;
  s.b_lo = o;    o += kBPart;
  s.x = o;       o += (size_t)kFsgRows * FH * 4;
  s.in_ptr = o;  o += 48 * 4;
  s.out_ptr = o; o += 48 * 4;
  s.in_src = o;  o += kFsgEntries * 4;
  s.in_nrm = o;  o += kFsgEntries * 4;
  s.out_dst = o; o += kFsgEntries * 4;
  s.out_pos = o; o += kFsgEntries * 4;
  s.w = o;       o += kFsgEntries * 8;                  // edge attention by in-CSR position (both branches)
  s.wn = o;      o += kFsgEntries * 8;                  // dis_w[source] * edge attention
  s.gptr = o;    o += 48 * 4;                           // first local row of every graph of the block
  s.rowf = o;    o += (size_t)kFsgRows * 12 * 4;        // per row: s (1), s_c / s_o (2), node att (2), p / q (4), dis_w (2), pad
  s.aff = o;     o += 4 * FH * 4;                       // BatchNorm scale / shift (two sets)
  s.part = o;    o += 512 * 8;
  s.tot = o;     o += 512 * 8;
  s.total = o;
  return s;
}

// What a thread needs to finalise channel k of BatchNorm `id`, fetched BEFORE the all-reduce wait: the indexed
// constant-bank loads (c.bn_*[id]) and the global loads of gamma / beta / the running statistics otherwise sit on
// the critical path right behind the wait (ncu: 6 % of the backward kernel's samples on one indexed LDC).
struct FsgBnPre {
  float g, b, rm, rv;
  long long rm_off, rv_off;
};
__device__ __forceinline__ FsgBnPre fsg_bn_prefetch(const Ctx& c, int id, int t0) {
  FsgBnPre p = {1.f, 0.f, 0.f, 1.f, -1, -1};
  const int k = (int)threadIdx.x - t0;
  if (k >= 0 && k < FH) {
    p.g = c.params[c.bn_gamma[id] + k];
    p.b = c.params[c.bn_beta[id] + k];
    if (blockIdx.x == 0 && c.bn_buffers != nullptr && c.bn_rm[id] >= 0) {
      p.rm_off = c.bn_rm[id] + k;
      p.rv_off = c.bn_rv[id] + k;
      p.rm = c.bn_buffers[p.rm_off];
      p.rv = c.bn_buffers[p.rv_off];
    }
  }
  return p;
}
// training-mode BatchNorm `id` (H channels) from the grid totals tot[0..H) (sum) and tot[H..2H) (sum of squares):
// threads [t0, t0 + H) write the affine into shared memory; CTA 0 publishes the record and the running statistics
__device__ __forceinline__ void fsg_bn_finalize(const Ctx& c, int id, int count, const FsgBnPre& pre, const double* tot, float* s_sc,
                                                float* s_sh, int t0) {
  const int k = (int)threadIdx.x - t0;
  if (k < 0 || k >= FH) return;
  const double s = tot[k], q = tot[FH + k];
  double mean = 0.0, var = 0.0;
  if (count > 0) {
    const double inv = 1.0 / (double)count;
    mean = s * inv;
    var = q * inv - mean * mean;
    if (var < 0.0) var = 0.0;
  }
  const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
  const float sc = pre.g * rstd;
  const float sh = pre.b - (float)mean * sc;
  s_sc[k] = sc;
  s_sh[k] = sh;
  if (blockIdx.x == 0) {
    c.bnf(id, BN_SCALE)[k] = sc;
    c.bnf(id, BN_SHIFT)[k] = sh;
    c.bnf(id, BN_MEAN)[k] = (float)mean;
    c.bnf(id, BN_RSTD)[k] = rstd;
    if (pre.rm_off >= 0) {
      const double unb = count > 1 ? var * ((double)count / (double)(count - 1)) : var;
      c.bn_buffers[pre.rm_off] = (1.f - c.momentum) * pre.rm + c.momentum * (float)mean;
      c.bn_buffers[pre.rv_off] = (1.f - c.momentum) * pre.rv + c.momentum * (float)unb;
    }
    if (k == 0 && c.nbt != nullptr) atomicAdd(reinterpret_cast<unsigned long long*>(c.nbt + id), 1ull);   // (RED: no load round trip)
  }
}

// ---------------------------------------------------------------------------------------------
// The forward kernel.  grid = min(max_graphs, 148), 256 threads, 1 CTA per SM.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FT, 1) k_fsg_forward(const Ctx c) {
  if (blockIdx.x == 0) CAL_TL(c.status, 11);                         // (timeline builds: kernel entry)
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_w, bar_mma;
  __shared__ uint32_t tmem_slot;
  const FsgSmem S = fsg_smem();
  unsigned char* sAh = smem + S.a_hi;
  unsigned char* sAl = smem + S.a_lo;
  unsigned char* sBh = smem + S.b_hi;
  unsigned char* sBl = smem + S.b_lo;
  float* sX = reinterpret_cast<float*>(smem + S.x);                  // [rows][128] current node features
  int* sInPtr = reinterpret_cast<int*>(smem + S.in_ptr);
  int* sOutPtr = reinterpret_cast<int*>(smem + S.out_ptr);
  int* sInSrc = reinterpret_cast<int*>(smem + S.in_src);
  float* sInNrm = reinterpret_cast<float*>(smem + S.in_nrm);
  int* sOutDst = reinterpret_cast<int*>(smem + S.out_dst);
  int* sOutPos = reinterpret_cast<int*>(smem + S.out_pos);
  float2* sW = reinterpret_cast<float2*>(smem + S.w);
  float2* sWn = reinterpret_cast<float2*>(smem + S.wn);
  int* sGptr = reinterpret_cast<int*>(smem + S.gptr);
  float* sRow = reinterpret_cast<float*>(smem + S.rowf);             // [rows][12]
  float* sAff = reinterpret_cast<float*>(smem + S.aff);              // sc0 | sh0 | sc1 | sh1
  double* sPart = reinterpret_cast<double*>(smem + S.part);
  double* sTot = reinterpret_cast<double*>(smem + S.tot);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const FsgWs ws = fsg_ws(c);
  const int L = c.L, F = c.F;
  const int Fp8 = (F + 7) & ~7;

  // ---- before the dependency wait: TMEM, barriers, attention projection weights (parameters) ----
  if (warp == 0) umma::tmem_alloc(&tmem_slot, kTmemCols);
  if (t == 0) {
    umma::mbar_init(&bar_w, 1);
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_fence_init();
  }
  float wn0[4], wn1[4], wp0[4], wp1[4], wq0[4], wq1[4];               // lane's 4 channels of node_att_mlp / edge_att_mlp
  {
    const float* Wn = c.params + c.po.node_att_w;                      // [2][H]
    const float* We = c.params + c.po.edge_att_w;                      // [2][2H]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = lane * 4 + i;
      wn0[i] = Wn[k];
      wn1[i] = Wn[FH + k];
      wp0[i] = We[k];
      wp1[i] = We[2 * FH + k];
      wq0[i] = We[FH + k];
      wq1[i] = We[3 * FH + k];
    }
  }
  const float bnat0 = c.params[c.po.node_att_b], bnat1 = c.params[c.po.node_att_b + 1];
  const float beat0 = c.params[c.po.edge_att_b], beat1 = c.params[c.po.edge_att_b + 1];
  // ---- still before the dependency wait: the immediate predecessor (k_fsg_prep) only writes the weight images; the
  // structure (cal_prep), the input features and their column totals were complete before it could start.
  // A block is ONE graph. ----
  const int N = imin(imax(c.dims[0], 0), c.Nm);
  int g0 = 0, g1 = 0, n0 = 0, n1 = 0, ie0 = 0, ie1 = 0, oe0 = 0;
  bool own = (int)blockIdx.x < imin(imax(c.dims[2], 0), c.Bm);
  if (own) {
    g0 = blockIdx.x; g1 = g0 + 1;
    n0 = c.graph_ptr[g0]; n1 = c.graph_ptr[g1];
    if (n1 < n0 || n1 - n0 > kFsgRows || n0 < 0 || n1 > c.Nm) own = false;
  }
  if (own) {
    ie0 = c.in_ptr[n0]; ie1 = c.in_ptr[n1]; oe0 = c.out_ptr[n0];
    if (ie1 < ie0 || ie1 - ie0 > kFsgEntries) own = false;
  }
  if (!own) g0 = g1 = n0 = n1 = ie0 = ie1 = oe0 = 0;
  const int Nc = n1 - n0, Ec = ie1 - ie0;
  const int npad = imax(8, (Nc + 7) & ~7);                             // MMA N
  uint32_t par_w = 0, par_m = 0;                                      // mbarrier phases
  if (own) {
    // the block's two CSR segments, local ids
    for (int i = t; i <= Nc; i += FT) {
      sInPtr[i] = c.in_ptr[n0 + i] - ie0;
      sOutPtr[i] = c.out_ptr[n0 + i] - oe0;
    }
    for (int e = t; e < Ec; e += FT) {
      sInSrc[e] = c.in_src[ie0 + e] - n0;
      sInNrm[e] = c.in_norm[ie0 + e];
      sOutDst[e] = c.out_dst[oe0 + e] - n0;
      sOutPos[e] = c.out_pos[oe0 + e] - ie0;
    }
    for (int g = t; g <= g1 - g0; g += FT) sGptr[g] = c.graph_ptr[g0 + g] - n0;
    // bn_feat (model.py:90): the column totals come from cal_prep (statp[0 .. 2F))
    float* sc = sAff;
    float* sh = sAff + FH;
    if (c.train) {
      for (int k = t; k < F; k += FT) {
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
      if (blockIdx.x == 0 && t == 0 && c.nbt != nullptr) atomicAdd(reinterpret_cast<unsigned long long*>(c.nbt + 0), 1ull);   // (RED: no load round trip in front of the barrier)
    } else {
      for (int k = t; k < F; k += FT) {
        sc[k] = c.bnf(0, BN_SCALE)[k];
        sh[k] = c.bnf(0, BN_SHIFT)[k];
      }
    }
    __syncthreads();
    // node operand of the input transform: bn_feat(x) rows, split
    const int nkc = Fp8 / 4;
    for (int i = warp; i < npad; i += 8) {
      if (lane < nkc) {
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < Nc) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = lane * 4 + e;
            if (k < F) v[e] = fmaf(c.feat[(size_t)(n0 + i) * F + k], sc[k], sh[k]);
          }
        }
        put_b(sBh, sBl, i, lane, make_float4(v[0], v[1], v[2], v[3]));
      }
    }
    umma::fence_async_smem();
  }
  umma::fence_before_sync();
  FSG_TDECL
  pdl_sync();                                                         // the weight images
  if (blockIdx.x == 0) CAL_TL(c.status, 2);
  CAL_TLC(c, 0, 0);
  FSG_T(0);                                                           // 0: dependency wait (set-up overlapped)
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int fx_set = fsg_epoch_begin(ws, 0, 0, L + 1);               // (every CTA of the grid, active or not)
  // One block per graph: G live blocks.  A live block whose graph does not fit the limits (the caller promised
  // cal_caps.small_graphs) reports it through the status word and only keeps the all-reduces of the others complete.
  const int G = imin(imax(c.dims[2], 0), c.Bm);
  // (opaque to the compiler on purpose: with `active` known before the wait it merges the set-up above into the body
  // below and the kernel runs 8 % slower -- 47.6 against 43.9 us, measured)
  int act_ = own ? 1 : 0;
  asm volatile("" : "+r"(act_));
  const bool active = act_ != 0;
  const bool unfit = !own && (int)blockIdx.x < G;
  if (t == 0 && (int)blockIdx.x < G) {                                // this block's record for the backward kernel
    int4* info = reinterpret_cast<int4*>(ws.info + (size_t)blockIdx.x * 8);
    info[0] = make_int4(g0, g1, n0, n1);
    info[1] = make_int4(ie0, ie1, oe0, own ? 1 : 0);
  }
  if (active) {
    // weight operand of the input transform
    if (t == 0) {
      const uint32_t bytes = (uint32_t)(Fp8 / 4) * kALbo;
      umma::mbar_expect_tx(&bar_w, 2 * bytes);
      umma::bulk_g2s(sAh, fsg_img_feat(ws, L), bytes, &bar_w);
      umma::bulk_g2s(sAl, fsg_img_feat(ws, L) + kFsgImgPart, bytes, &bar_w);
    }
    FSG_T(1);                                                         // 1: (set-up now happens before the wait)

    // ================= input transform: x_1 = relu(bn_feat(x) W_feat)  (gfn: no bias, no propagation) =================
    {
      umma::mbar_wait(&bar_w, par_w);
      par_w ^= 1u;
      if (t == 0) {
        umma::fence_after_sync();
        issue_3xtf32(sAh, sAl, sBh, sBl, tmem, tmem + 128u, Fp8 / 8, npad);
        umma::commit(&bar_mma);
      }
      umma::mbar_wait(&bar_mma, par_m);
      par_m ^= 1u;
      umma::fence_after_sync();
      FSG_T(2);                                                       // 2: input transform (MMA)
    }
  }

  // ================= layers + masked convs =================
  // phase p = 0 .. L-1: statistics of bns_conv[p] (input of layer p); phase L: statistics of bnc / bno
  float bias_next = 0.f;                                              // bias of the product whose epilogue comes next (fetched ahead of its MMA wait)
  for (int l = 0; l <= L; ++l) {
    const bool masked = l == L;
    if (active) {
      // -- weight operand of this phase (the previous MMAs have completed: the buffer is free) --
      if (t == 0) {
        const float* img = fsg_img_fwd(ws, l);
        umma::mbar_expect_tx(&bar_w, 2u * 65536u);
        umma::bulk_g2s(sAh, img, 65536u, &bar_w);
        umma::bulk_g2s(sAl, img + kFsgImgPart, 65536u, &bar_w);
      }
      // -- epilogue of the previous product: thread = output channel, the two warp sets split the row groups --
      {
        const int ch = (warp & 3) * 32 + lane;
        const float bias = bias_next;
        double s = 0.0, q = 0.0;
        for (int g8 = (warp >> 2) * 8; g8 < npad; g8 += 16) {
          float vm[8], vc[8];
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, g8), vm);
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + g8), vc);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = g8 + e;
            if (i < Nc) {
              const float x = fmaxf((vm[e] + vc[e]) + bias, 0.f);
              sX[i * FH + ch] = x;
              s += (double)x;
              q += (double)x * (double)x;
            }
          }
        }
        if (warp >= 4) {
          sTot[ch] = s;
          sTot[FH + ch] = q;
        }
        umma::fence_before_sync();
        __syncthreads();
        if (warp < 4) {
          sPart[ch] = s + sTot[ch];
          sPart[FH + ch] = q + sTot[FH + ch];
        }
      }
      umma::fence_before_sync();
      __syncthreads();
      umma::fence_after_sync();
      FSG_T(3);                                                       // 3: epilogues (TMEM -> x, statistics)
    }
    if (!masked) {
      // ---------------- backbone layer l: x_{l+2} = relu(A bn_l(x_{l+1}) W_l + b_l)  (model.py:93-95) ----------------
      if (active) {
        if (c.train) fsg_publish_fx(ws, fx_set, l, G, sPart, 2 * FH);
        const FsgBnPre pre = c.train ? fsg_bn_prefetch(c, 1 + l, 0) : FsgBnPre();
        FSG_T(4);                                                     // 4: publish
        // x_{l+1} rows to the workspace (the backward pass reads them): coalesced, after the publish so that the
        // all-reduce's fence does not wait for these stores
        {
          float4* xg = reinterpret_cast<float4*>(c.Xl(l) + (size_t)n0 * FH);
          const float4* xs = reinterpret_cast<const float4*>(sX);
          for (int i = t; i < Nc * (FH / 4); i += FT) xg[i] = xs[i];
        }
        // raw aggregate while the statistics travel: warp per target row, lanes over 4 channels
        for (int i = warp; i < Nc; i += 8) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          float sn = 0.f;
          for (int e = sInPtr[i]; e < sInPtr[i + 1]; ++e) {
            const float w = sInNrm[e];
            const float4 v = *reinterpret_cast<const float4*>(sX + sInSrc[e] * FH + lane * 4);
            acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            sn += w;
          }
          *reinterpret_cast<float4*>(sBh + b_off(i, lane)) = acc;
          if (lane == 0) sRow[i * 12] = sn;
        }
        FSG_T(5);                                                     // 5: raw aggregation
        float* sc = sAff;
        float* sh = sAff + FH;
        if (c.train) {
          fsg_wait_total_fx(ws, fx_set, l, G, 2 * FH, sTot);
          fsg_bn_finalize(c, 1 + l, N, pre, sTot, sc, sh, 0);
        } else if (t < FH) {
          sc[t] = c.bnf(1 + l, BN_SCALE)[t];
          sh[t] = c.bnf(1 + l, BN_SHIFT)[t];
        }
        __syncthreads();
        FSG_T(6);                                                     // 6: statistics wait + finalize
        const float4 sc4 = *reinterpret_cast<const float4*>(sc + lane * 4), sh4 = *reinterpret_cast<const float4*>(sh + lane * 4);
        for (int i = warp; i < npad; i += 8) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < Nc) {
            const float4 r = *reinterpret_cast<const float4*>(sBh + b_off(i, lane));
            const float sn = sRow[i * 12];
            a = make_float4(fmaf(sc4.x, r.x, sh4.x * sn), fmaf(sc4.y, r.y, sh4.y * sn), fmaf(sc4.z, r.z, sh4.z * sn),
                            fmaf(sc4.w, r.w, sh4.w * sn));
          }
          put_b(sBh, sBl, i, lane, a);
        }
        umma::fence_async_smem();
        __syncthreads();
        FSG_T(7);                                                     // 7: affine + split into the operand
        umma::mbar_wait(&bar_w, par_w);
        par_w ^= 1u;
        FSG_T(8);                                                     // 8: weight image wait
        if (t == 0) {
          umma::fence_after_sync();
          issue_3xtf32(sAh, sAl, sBh, sBl, tmem, tmem + 128u, FH / 8, npad);
          umma::commit(&bar_mma);
        }
        bias_next = c.params[c.po.convs_b[l] + (warp & 3) * 32 + lane];
        umma::mbar_wait(&bar_mma, par_m);
        par_m ^= 1u;
        umma::fence_after_sync();
        FSG_T(9);                                                     // 9: MMA
      }
      continue;
    }
    // ---------------- attention masks + the two masked convs + pooling (model.py:97-116) ----------------
    if (!active) break;
    // node attention softmax(node_att_mlp(x)), the per-node halves p, q of edge_att_mlp([x_row || x_col])
    for (int i = warp; i < Nc; i += 8) {
      const float4 v = *reinterpret_cast<const float4*>(sX + i * FH + lane * 4);
      const float o[4] = {v.x, v.y, v.z, v.w};
      float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s0 = fmaf(o[k], wn0[k], s0);
        s1 = fmaf(o[k], wn1[k], s1);
        p0 = fmaf(o[k], wp0[k], p0);
        p1 = fmaf(o[k], wp1[k], p1);
        q0 = fmaf(o[k], wq0[k], q0);
        q1 = fmaf(o[k], wq1[k], q1);
      }
#pragma unroll
      for (int o_ = 16; o_ > 0; o_ >>= 1) {                            // six warp sums at once: independent shuffles pipeline
        s0 += __shfl_xor_sync(0xffffffffu, s0, o_);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o_);
        p0 += __shfl_xor_sync(0xffffffffu, p0, o_);
        p1 += __shfl_xor_sync(0xffffffffu, p1, o_);
        q0 += __shfl_xor_sync(0xffffffffu, q0, o_);
        q1 += __shfl_xor_sync(0xffffffffu, q1, o_);
      }
      s0 += bnat0;
      s1 += bnat1;
      float a0 = 0.5f, a1 = 0.5f;
      if (!c.no_natt) {
        const float mx = fmaxf(s0, s1);
        const float e0 = expf(s0 - mx), e1 = expf(s1 - mx);
        const float inv = 1.0f / (e0 + e1);
        a0 = e0 * inv;
        a1 = e1 * inv;
      }
      if (lane == 0) {
        float* r = sRow + i * 12;
        r[3] = a0; r[4] = a1; r[5] = p0; r[6] = p1; r[7] = q0; r[8] = q1;
        *reinterpret_cast<float2*>(c.natt + (size_t)(n0 + i) * 2) = make_float2(a0, a1);
        *reinterpret_cast<float4*>(c.pq + (size_t)(n0 + i) * 4) = make_float4(p0, p1, q0, q1);
      }
    }
    __syncthreads();
    // statistics of bnc / bno on att * x: thread = channel, the two warp sets split the rows
    if (c.train) {
      const int ch = t & 127, half = t >> 7;
      double a = 0.0, b = 0.0, d = 0.0, e = 0.0;
      for (int i = half; i < Nc; i += 2) {
        const float x = sX[i * FH + ch];
        const float vc = sRow[i * 12 + 3] * x, vo = sRow[i * 12 + 4] * x;
        a += (double)vc; b += (double)vc * (double)vc;
        d += (double)vo; e += (double)vo * (double)vo;
      }
      if (half) {
        sTot[ch] = a; sTot[FH + ch] = b; sTot[2 * FH + ch] = d; sTot[3 * FH + ch] = e;
      }
      __syncthreads();
      if (!half) {
        sPart[ch] = a + sTot[ch];
        sPart[FH + ch] = b + sTot[FH + ch];
        sPart[2 * FH + ch] = d + sTot[2 * FH + ch];
        sPart[3 * FH + ch] = e + sTot[3 * FH + ch];
      }
    }
    __syncthreads();
    FSG_T(10);                                                        // 10: node attention + bnc / bno statistics
    if (c.train) fsg_publish_fx(ws, fx_set, L, G, sPart, 4 * FH, 0);   // (the launch's final site)
    const FsgBnPre pre_m = c.train ? fsg_bn_prefetch(c, t < FH ? L + 1 : L + 2, t < FH ? 0 : FH) : FsgBnPre();
    {
      float4* xg = reinterpret_cast<float4*>(c.Xl(L) + (size_t)n0 * FH);      // x_{L+1} rows to the workspace
      const float4* xs = reinterpret_cast<const float4*>(sX);
      for (int i = t; i < Nc * (FH / 4); i += FT) xg[i] = xs[i];
    }
    FSG_T(4);
    // edge attention softmax(p[row] + q[col] + b) per edge (never materialises [E, 2H]), the attention-weighted
    // degree by source row and dis_w = deg^-1/2 (gcn_conv.py:59-70 with edge_weight): warp per source node
    for (int n = warp; n < Nc; n += 8) {
      const float* rn = sRow + n * 12;
      const float pn0 = rn[5], pn1 = rn[6];
      const int q0 = sOutPtr[n], q1 = sOutPtr[n + 1] - 1;             // last slot = the appended self loop
      float deg0 = 0.f, deg1 = 0.f;
      for (int qb = q0; qb <= q1; qb += 32) {
        const int qq = qb + lane;
        float w0 = 0.f, w1 = 0.f;
        if (qq <= q1) {
          w0 = w1 = qq == q1 ? 1.f : 0.5f;
          if (qq != q1 && !c.no_eatt) {
            const float* rd = sRow + sOutDst[qq] * 12;
            const float t0 = pn0 + rd[7] + beat0, t1 = pn1 + rd[8] + beat1;
            const float mx = fmaxf(t0, t1);
            const float e0 = expf(t0 - mx), e1 = expf(t1 - mx);
            const float inv = 1.0f / (e0 + e1);
            w0 = e0 * inv;
            w1 = e1 * inv;
          }
          sW[sOutPos[qq]] = make_float2(w0, w1);
        }
        deg0 += warp_sum(w0);
        deg1 += warp_sum(w1);
      }
      const float2 dn = make_float2(1.0f / sqrtf(deg0), 1.0f / sqrtf(deg1));
      if (lane == 0) {
        sRow[n * 12 + 9] = dn.x;
        sRow[n * 12 + 10] = dn.y;
        *reinterpret_cast<float2*>(c.disw + (size_t)(n0 + n) * 2) = dn;
      }
      __syncwarp();
      const float2 an = make_float2(rn[3], rn[4]);
      for (int qq = q0 + lane; qq <= q1; qq += 32) {
        const int pos = sOutPos[qq];
        const float2 w = sW[pos];                                      // written by this lane above
        const float2 wn = make_float2(dn.x * w.x, dn.y * w.y);
        sWn[pos] = wn;
        *reinterpret_cast<float2*>(c.watt + (size_t)(ie0 + pos) * 2) = w;
        *reinterpret_cast<float2*>(c.edge_wn + (size_t)(ie0 + pos) * 2) = wn;
        *reinterpret_cast<float2*>(c.edge_na + (size_t)(ie0 + pos) * 2) = an;
      }
    }
    __syncthreads();
    FSG_T(11);                                                        // 11: edge attention
    float* sc0 = sAff;
    float* sh0 = sAff + FH;
    float* sc1 = sAff + 2 * FH;
    float* sh1 = sAff + 3 * FH;
    for (int br = 0; br < 2; ++br) {
      // raw aggregate of branch br: sum_e (wn_e * dis_w[i]) * (att[src] * x[src]) and the weight sum
      for (int i = warp; i < Nc; i += 8) {
        const float di = sRow[i * 12 + 9 + br];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float sn = 0.f;
        for (int e = sInPtr[i]; e < sInPtr[i + 1]; ++e) {
          const int src = sInSrc[e];
          const float2 wn = sWn[e];
          const float w = (br ? wn.y : wn.x) * di;
          const float am = sRow[src * 12 + 3 + br];
          const float4 v = *reinterpret_cast<const float4*>(sX + src * FH + lane * 4);
          acc.x = fmaf(w, am * v.x, acc.x); acc.y = fmaf(w, am * v.y, acc.y); acc.z = fmaf(w, am * v.z, acc.z); acc.w = fmaf(w, am * v.w, acc.w);
          sn += w;
        }
        *reinterpret_cast<float4*>(sBh + b_off(i, lane)) = acc;
        if (lane == 0) sRow[i * 12 + 1 + br] = sn;
      }
      FSG_T(5);
      if (br == 0) {
        if (c.train) {
          fsg_wait_total_fx(ws, fx_set, L, G, 4 * FH, sTot);
          CAL_TLC(c, 0, 1);
          fsg_bn_finalize(c, L + 1, N, pre_m, sTot, sc0, sh0, 0);
          fsg_bn_finalize(c, L + 2, N, pre_m, sTot + 2 * FH, sc1, sh1, FH);
        } else if (t < FH) {
          sc0[t] = c.bnf(L + 1, BN_SCALE)[t];
          sh0[t] = c.bnf(L + 1, BN_SHIFT)[t];
          sc1[t] = c.bnf(L + 2, BN_SCALE)[t];
          sh1[t] = c.bnf(L + 2, BN_SHIFT)[t];
        }
      }
      __syncthreads();
      FSG_T(6);
      {
        const float* sc = br ? sc1 : sc0;
        const float* sh = br ? sh1 : sh0;
        const float4 sc4 = *reinterpret_cast<const float4*>(sc + lane * 4), sh4 = *reinterpret_cast<const float4*>(sh + lane * 4);
        float* aggg = c.agg + (size_t)br * c.Nm * FH + (size_t)n0 * FH;
        for (int i = warp; i < npad; i += 8) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < Nc) {
            const float4 r = *reinterpret_cast<const float4*>(sBh + b_off(i, lane));
            const float sn = sRow[i * 12 + 1 + br];
            a = make_float4(fmaf(sc4.x, r.x, sh4.x * sn), fmaf(sc4.y, r.y, sh4.y * sn), fmaf(sc4.z, r.z, sh4.z * sn),
                            fmaf(sc4.w, r.w, sh4.w * sn));
            *reinterpret_cast<float4*>(aggg + (size_t)i * FH + lane * 4) = a;    // saved for the weight gradient
          }
          put_b(sBh, sBl, i, lane, a);
        }
      }
      umma::fence_async_smem();
      __syncthreads();
      FSG_T(7);
      umma::mbar_wait(&bar_w, par_w);
      par_w ^= 1u;
      FSG_T(8);
      if (t == 0) {
        umma::fence_after_sync();
        issue_3xtf32(sAh, sAl, sBh, sBl, tmem + (br ? 64u : 0u), tmem + 128u + (br ? 64u : 0u), FH / 8, npad);
        umma::commit(&bar_mma);
      }
      const float bias_br = c.params[(br ? c.po.objects_b : c.po.context_b) + (warp & 3) * 32 + lane];
      umma::mbar_wait(&bar_mma, par_m);
      par_m ^= 1u;
      umma::fence_after_sync();
      FSG_T(9);
      if (br == 0 && t == 0) {                                         // objects_convs weights while branch 0 is finished
        const float* img = fsg_img_fwd(ws, L + 1);
        umma::mbar_expect_tx(&bar_w, 2u * 65536u);
        umma::bulk_g2s(sAh, img, 65536u, &bar_w);
        umma::bulk_g2s(sAl, img + kFsgImgPart, 65536u, &bar_w);
      }
      // epilogue of the branch: z = relu(. + b), saved; global_add_pool over the rows of the block's graph
      // (model.py:115-116; a block is one graph).  thread = channel, the two warp sets split the row groups.
      {
        const int ch = (warp & 3) * 32 + lane;
        const float bias = bias_br;
        float* zg = c.Z + (size_t)br * c.Nm * FH + (size_t)n0 * FH;
        const uint32_t cm = br ? 64u : 0u;
        float pool = 0.f;
        for (int g8 = (warp >> 2) * 8; g8 < npad; g8 += 16) {
          float vm[8], vc[8];
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, cm + g8), vm);
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + cm + g8), vc);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = g8 + e;
            if (i < Nc) {
              const float z = fmaxf((vm[e] + vc[e]) + bias, 0.f);
              zg[(size_t)i * FH + ch] = z;
              pool += z;
            }
          }
        }
        float* sRed = reinterpret_cast<float*>(sPart);
        if (warp >= 4) sRed[br * FH + ch] = pool;
        umma::fence_before_sync();
        __syncthreads();
        if (warp < 4) c.pooled[((size_t)br * c.Bm + g0) * FH + ch] = pool + sRed[br * FH + ch];
      }
      umma::fence_before_sync();
      __syncthreads();
      umma::fence_after_sync();
      FSG_T(12);                                                      // 12: masked epilogue + pooling
    }
  }
  if (unfit && t == 0) {                                              // (skipped everything above)
    raise_status(c.status, kStCapacity);
    if (c.train) fsg_unfit_arrive(ws, fx_set, G, 0, L, 0);
  }
  FSG_TDUMP(c, 48);
  if (blockIdx.x == 0) CAL_TL(c.status, 3);
  CAL_TLC(c, 0, 2);

  // ---- teardown: TMEM (the all-reduce state needs none: fsg_epoch_begin) ----
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
  CAL_TLC(c, 0, 3);
}

// ---------------------------------------------------------------------------------------------
// Per-step preparation of the path: the pre-split (hi | lo) weight-operand images in the canonical K-major layout
// the MMA reads; every block builds 2048 elements of one image.  (A training loop can have the optimizer kernel
// write the image words instead: cal_image_sink / CAL_F_FSG_READY.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fsg_prep(const Ctx c) {
  // The blocks read only parameters: the optimizer step that wrote them is at least two launches back and complete by
  // the time this kernel can be scheduled (cal_prep sits between them; a caller whose cal_prep ran elsewhere passes
  // CAL_F_NO_OVERLAP), so they do their work BEFORE the dependency wait -- the images are built while the
  // predecessor (the structure kernel of cal_prep) still runs.  They still wait before they release the dependent
  // launch: every kernel of the chain relies on "my predecessor has passed its wait, so everything before it is
  // complete" (the fused forward reads the structure ahead of its own wait).
  const FsgWs ws = fsg_ws(c);
  const int t = threadIdx.x;
  const int L = c.L;
  {
    const int img = blockIdx.x >> 3, sub = blockIdx.x & 7;              // 8 blocks of 2048 elements per image
    float* dst;
    const float* W;
    int mode;                                                           // 0: A[m][k] = W[k][m] (forward), 1: A[m][k] = W[m][k] (backward), 2: feat
    int K = FH;
    if (img < 2 * (L + 2)) {
      const int j = img >> 1;
      mode = img & 1;
      W = c.params + (j < L ? c.po.convs_w[j] : (j == L ? c.po.context_w : c.po.objects_w));
      dst = mode ? fsg_img_bwd(ws, j) : fsg_img_fwd(ws, j);
    } else if (img == 2 * (L + 2)) {
      mode = 2;
      W = c.params + c.po.conv_feat_w;                                  // [F][H]
      dst = fsg_img_feat(ws, L);
      K = (c.F + 7) & ~7;
    } else {                                                            // fc1 of readout h ("add": [H][H] as stored = [out][in])
      const int q = img - 2 * (L + 2) - 1, h = q >> 1;
      mode = (q & 1) ? 0 : 1;                                           // forward image: A[m][k] = W[m][k]; backward: A[m][k] = W[k][m]
      W = c.params + c.po.fc1_w[h];
      dst = (q & 1) ? fsg_img_fc1_bwd(ws, L, h) : fsg_img_fc1_fwd(ws, L, h);
    }
    for (int e = sub * 2048 + t; e < (sub + 1) * 2048; e += 256) {
      const int kc = e >> 9, m = (e & 511) >> 2, kk = e & 3;            // float offset e = kc * 512 + m * 4 + kk
      const int k = kc * 4 + kk;
      if (k >= K) continue;
      float v;
      if (mode == 0) v = W[(size_t)k * FH + m];
      else if (mode == 1) v = W[(size_t)m * FH + k];
      else v = k < c.F ? W[(size_t)k * FH + m] : 0.f;
      float hi, lo;
      umma::split_tf32(v, hi, lo);
      dst[e] = hi;
      dst[kFsgImgPart + e] = lo;
    }
  }
  if (blockIdx.x == 0) CAL_TL(c.status, 1);
  pdl_sync();
}

}  // namespace

int fsg_grid(const Ctx& c) { return imax(1, imin(c.Bm, kSMs)); }
size_t fsg_region_bytes(int Bm, int L, int F) { return fsg_layout(Bm, L, F).total; }

// where the optimizer kernels write the image words of the parameters they update (cal_image_sink)
int fsg_fill_image_sink(const Ctx& c, cal_image_sink* sink) {
  const FsgWs ws = fsg_ws(c);
  const int L = c.L;
  int n = 0;
  auto add = [&](long long off, int rows, float* dst_t, float* dst_n) {
    sink->entry[n].offset = off;
    sink->entry[n].rows = rows;
    sink->entry[n].reserved = 0;
    sink->entry[n].dst_t = dst_t;
    sink->entry[n].dst_n = dst_n;
    ++n;
  };
  for (int j = 0; j < L + 2; ++j)                                     // forward image A[m][k] = W[k][m], backward A[m][k] = W[m][k]
    add(j < L ? c.po.convs_w[j] : (j == L ? c.po.context_w : c.po.objects_w), FH, fsg_img_fwd(ws, j), fsg_img_bwd(ws, j));
  add(c.po.conv_feat_w, c.F, fsg_img_feat(ws, L), nullptr);           // [F][H]; rows F .. of the image stay zero
  if (!c.cat)
    for (int h = 0; h < 3; ++h)                                       // fc1 [out][in]: forward A[m][k] = W[m][k], backward A[m][k] = W[k][m]
      add(c.po.fc1_w[h], FH, fsg_img_fc1_bwd(ws, L, h), fsg_img_fc1_fwd(ws, L, h));
  sink->count = n;
  return 0;
}

int launch_fsg_prep(const Ctx& c, cudaStream_t s, bool no_overlap) {
  const int n_img = (2 * (c.L + 2) + 1 + (c.cat ? 0 : 6)) * 8;       // (the fc1 images only exist for cat_or_add = "add")
  if (no_overlap) launch_k_plain(k_fsg_prep, dim3(n_img), dim3(256), 0, s, c);
  else launch_k(k_fsg_prep, dim3(n_img), dim3(256), 0, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_fsg_forward(const Ctx& c, cudaStream_t s, bool after_full_dependency) {
  const size_t smem = fsg_smem().total;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_fsg_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // CAL_F_FSG_READY: no k_fsg_prep ahead of this kernel -- its predecessor is then the optimizer kernel, whose output
  // (the attention projections) the set-up before the dependency wait reads: no programmatic overlap with it
  if (after_full_dependency) launch_k_plain(k_fsg_forward, dim3(fsg_grid(c)), dim3(FT), smem, s, c);
  else launch_k(k_fsg_forward, dim3(fsg_grid(c)), dim3(FT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
