// fsg_bwd.cu -- the fused small-graph BACKWARD pass of CausalGCN: ONE persistent kernel from the pooled-embedding
// gradients (left by the readout backward) down to the input transform, the mirror of fsg.cu.
//
// A CTA owns one graph (the block record its forward CTA left in the workspace).  Everything that couples nodes -- the transpose
// aggregates (gcn_conv.py:92-97 backward), the weighted-normalisation backward through both endpoints' degrees
// (gcn_conv.py:59-70), the edge / node attention softmax backward (model.py:97-111) -- is local to the CTA's
// shared memory; the gradient products  D = u W^T  (and  d agg = d z W^T  of the two masked convs) run on the
// tensor cores (tcgen05.mma, 3xTF32, accumulators in TMEM, pre-split weight images by cp.async.bulk); the weight
// gradients  d W = y^T u  are FFMA2 outer products that run while the BatchNorm all-reduce of the layer travels.
// Training-mode BatchNorm backward is the only coupling between CTAs: one deterministic in-kernel all-reduce of
// (sum dy, sum dy * xhat) per BatchNorm (bnc | bno together, then bns_conv[L-1 .. 0]): 1 + L per step.
//
// Replaces k_masked_bwd_gemm, k_masked_bwd_gather, k_norm_bwd, k_att_bwd, k_conv_bwd x L and k_feat_bwd
// (conv.cu / attn.cu) when the fused small-graph path is on; the per-block partial parameter gradients go to
// the CAL_WS_FSG region and are summed in block order by k_fsg_grad_reduce (deterministic).
#include "fsg_dev.cuh"

namespace cal {
namespace {

constexpr int kGBytes = kFsgRows * FH * 4;            // 20480: one row-major [rows][128] fp32 tile
constexpr uint32_t kTailOff = 65536u - (uint32_t)kGBytes;   // the gradient rows alias the last 10 K-chunks of the lo image
constexpr int kTailStep = (int)(kTailOff / (2u * kALbo));   // first K step (of 8) whose lo operand lies in the tail: 11
static_assert(kTailOff % (2u * kALbo) == 0, "the tail starts on a K-step boundary");
constexpr int kRowsPerWarp = kFsgRows / 8;            // 5

struct FsgBSmem {
  size_t a_hi, a_lo, g, b_hi, b_lo, x, in_ptr, out_ptr, in_src, out_dst, out_pos, out_row, out_nrm, w, dnrm, dt, rowf, vec, part, tot, pool, total;
};
__host__ __device__ inline FsgBSmem fsg_bsmem() {
  FsgBSmem s;
  size_t o = 0;
  s.a_hi = o;    o += 65536;
  s.a_lo = o;    o += 65536;
  s.g = s.a_lo + kTailOff;                              // [rows][128] gradient rows (layer loop), inside a_lo
  s.b_hi = o;    o += kBPart;
  s.b_lo = o;    o += kBPart;
  s.x = o;       o += (size_t)kGBytes;
  s.in_ptr = o;  o += 48 * 4;
  s.out_ptr = o; o += 48 * 4;
  s.in_src = o;  o += kFsgEntries * 4;
  s.out_dst = o; o += kFsgEntries * 4;
  s.out_pos = o; o += kFsgEntries * 4;
  s.out_row = o; o += kFsgEntries * 4;                  // source row of every out-CSR entry
  s.out_nrm = o; o += kFsgEntries * 4;
  s.w = o;       o += kFsgEntries * 8;                  // edge attention by in-CSR position (both branches)
  s.dnrm = o;    o += kFsgEntries * 8;                  // d norm by in-CSR position
  s.dt = o;      o += kFsgEntries * 8;                  // d (edge logits) by in-CSR position
  s.rowf = o;    o += (size_t)kFsgRows * 8 * 4;         // per row: node att (2), dis_w (2), dp (2), pad (2)
  s.vec = o;     o += 12 * FH * 4;                      // per-channel BatchNorm vectors (two sets of sc, sh, mean, rstd, c1, c2)
  s.part = o;    o += 512 * 8;
  s.tot = o;     o += 512 * 8;
  s.pool = o;    o += 2 * FH * 4;                       // pooled-embedding gradient of the block's graph, both branches
  s.total = o;
  return s;
}

// BatchNorm-backward totals tot[0..K) = sum dy, tot[K..2K) = sum dy * xhat  ->  c1 = mean(dy), c2 = mean(dy * xhat) in
// shared memory by threads [t0, t0 + K); CTA 0 publishes d gamma / d beta (and the record)
__device__ __forceinline__ void fsg_bn_bwd_finalize(const Ctx& c, int id, int count, long long gamma_off, long long beta_off,
                                                    const double* tot, float* s_c1, float* s_c2, int t0) {
  // (gamma_off / beta_off = c.bn_gamma[id] / c.bn_beta[id], read by the caller BEFORE the all-reduce wait: an indexed
  // constant-bank load right behind the wait was 6 % of this kernel's stall samples)
  const int k = (int)threadIdx.x - t0;
  if (k < 0 || k >= FH) return;
  const double inv = count > 0 ? 1.0 / count : 0.0;
  const double a = tot[k], b = tot[FH + k];
  const float c1 = (float)(a * inv), c2 = (float)(b * inv);
  s_c1[k] = c1;
  s_c2[k] = c2;
  if (blockIdx.x == 0) {
    c.bnf(id, BN_C1)[k] = c1;
    c.bnf(id, BN_C2)[k] = c2;
    c.grads[gamma_off + k] = (float)b;
    c.grads[beta_off + k] = (float)a;
  }
}

// ---- the weight gradient  dW[ka][kb] = sum_r P[r][ka] * Q[r][kb]  over the block's rows, on the tensor cores ----
// Both operands are needed "rows = K": transposed, K-major, in the layout of the weight images (chunk c of channel m at
// c * 2048 + m * 16, a chunk = 4 consecutive block rows).  They are built in the (idle) weight-image buffer:
// P^T hi | P^T lo | Q^T hi | Q^T lo, kOpT bytes each.  P comes from a row-major tile (optionally through the BatchNorm
// affine y = sc * x + sh), Q from the node operand of the gradient product (its hi / lo parts are copied as they are).
// Threads 0..127 build P^T (thread = channel), threads 128..255 build Q^T.
constexpr uint32_t kOpT = (uint32_t)(kFsgRows / 4) * kALbo;            // 20480
template <bool AFFINE>
__device__ __forceinline__ void build_dw_operands(unsigned char* wbuf, const float* sP, int ldp, const float* s_sc, const float* s_sh,
                                                  const unsigned char* bh, const unsigned char* bl, int rows, int npad) {
  const int t = threadIdx.x;
  if (t < FH) {
    const int ch = t;
    const float sc = AFFINE ? s_sc[ch] : 1.f, sh = AFFINE ? s_sh[ch] : 0.f;
    for (int kc = 0; kc < npad / 4; ++kc) {
      float v[4], h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = kc * 4 + e;
        v[e] = r < rows ? (AFFINE ? fmaf(sP[r * ldp + ch], sc, sh) : sP[r * ldp + ch]) : 0.f;
        umma::split_tf32(v[e], h[e], l[e]);
      }
      const uint32_t off = (uint32_t)kc * kALbo + (uint32_t)ch * 16u;
      *reinterpret_cast<float4*>(wbuf + off) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(wbuf + kOpT + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
  } else {
    const int ch = t - FH;
    for (int kc = 0; kc < npad / 4; ++kc) {
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t o = b_off(kc * 4 + e, ch >> 2) + (uint32_t)(ch & 3) * 4u;     // rows >= `rows` hold zeros
        h[e] = *reinterpret_cast<const float*>(bh + o);
        l[e] = *reinterpret_cast<const float*>(bl + o);
      }
      const uint32_t off = (uint32_t)kc * kALbo + (uint32_t)ch * 16u;
      *reinterpret_cast<float4*>(wbuf + 2 * kOpT + off) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(wbuf + 3 * kOpT + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
  }
}
// one thread: dW (128 TMEM columns at d) = P^T Q over npad rows (3xTF32, one accumulator: K <= 40)
__device__ __forceinline__ void issue_dw(const unsigned char* wbuf, uint32_t d, int npad) {
  const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, 128);
  const uint32_t base = umma::smem_addr(wbuf);
  uint64_t ph = umma::smem_desc(base, kALbo, kASbo), pl = umma::smem_desc(base + kOpT, kALbo, kASbo);
  uint64_t qh = umma::smem_desc(base + 2 * kOpT, kALbo, kASbo), ql = umma::smem_desc(base + 3 * kOpT, kALbo, kASbo);
  constexpr uint64_t dd = (2u * kALbo) >> 4;
  for (int s = 0; s < npad / 8; ++s) {
    umma::mma_tf32(d, pl, qh, idesc, s > 0);
    umma::mma_tf32(d, ph, ql, idesc, 1u);
    umma::mma_tf32(d, ph, qh, idesc, 1u);
    ph += dd; pl += dd; qh += dd; ql += dd;
  }
}
// all 8 warps: dW from TMEM -> dst [128][128].  A thread holds a ROW of the accumulator (lane = row), so a direct store
// would touch 32 different 128-byte lines per instruction; every warp transposes its 32 x 32 blocks through a private
// shared-memory scratch (stride 33: conflict-free both ways) and stores 4 rows x 128 contiguous bytes per instruction.
constexpr int kDrainScratch = 8 * 32 * 33;            // floats
__device__ __forceinline__ void drain_dw(uint32_t tmem_dw, float* dst, float* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (warp & 3) * 32;
  float* sw = scratch + warp * 32 * 33;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int col0 = (warp >> 2) * 64 + h * 32;
    float v[32];
    umma::ld32(umma::tmem_addr(tmem_dw, row0, col0), v);
#pragma unroll
    for (int cc = 0; cc < 32; ++cc) sw[lane * 33 + cc] = v[cc];
    __syncwarp();
    const int rr = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* q = sw + (rr + 4 * i) * 33 + c4;
      *reinterpret_cast<float4*>(dst + (size_t)(row0 + rr + 4 * i) * FH + col0 + c4) = make_float4(q[0], q[1], q[2], q[3]);
    }
    __syncwarp();
  }
}

// sum over the 8 warps of a lane-owned [4-channel] accumulator, fixed order: dst[k] (k < 128) by threads 0..127
__device__ __forceinline__ void colsum8(const float (&acc)[4], float* sRed /* [8][128] */, float* dst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  *reinterpret_cast<float4*>(sRed + warp * FH + lane * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  __syncthreads();
  if (threadIdx.x < FH) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sRed[w * FH + threadIdx.x];
    dst[threadIdx.x] = s;
  }
}

// global rows [n0, n0 + rows) of a [*, 128] fp32 matrix -> row-major shared tile (16-byte asynchronous copies)
__device__ __forceinline__ void rows_async(float* sdst, const float* gsrc, int rows) {
  for (int i = threadIdx.x; i < rows * (FH / 4); i += FT) cp_async16(sdst + i * 4, gsrc + (size_t)i * 4);
}

// ---------------------------------------------------------------------------------------------
// The backward kernel.  grid = min(max_graphs, 148), 256 threads, 1 CTA per SM.
// ---------------------------------------------------------------------------------------------
constexpr int kTmemColsB = 512;                       // D main 0 / 64, D correction 128 / 192, dW 256 .. 383
constexpr uint32_t kTmemDw = 256;
constexpr int kLdR = FH + 4;                          // row stride of the d agg / y tiles (conflict-free quarter-warp dots)

__global__ void __launch_bounds__(FT, 1) k_fsg_backward(const Ctx c) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_w, bar_w2, bar_mma, bar_dw;
  __shared__ uint32_t tmem_slot;
  const FsgBSmem S = fsg_bsmem();
  unsigned char* sAh = smem + S.a_hi;
  unsigned char* sAl = smem + S.a_lo;
  float* sG = reinterpret_cast<float*>(smem + S.g);                  // [rows][128]
  unsigned char* sBh = smem + S.b_hi;
  unsigned char* sBl = smem + S.b_lo;
  float* sX = reinterpret_cast<float*>(smem + S.x);                  // [rows][128]
  int* sInPtr = reinterpret_cast<int*>(smem + S.in_ptr);
  int* sOutPtr = reinterpret_cast<int*>(smem + S.out_ptr);
  int* sInSrc = reinterpret_cast<int*>(smem + S.in_src);
  int* sOutDst = reinterpret_cast<int*>(smem + S.out_dst);
  int* sOutPos = reinterpret_cast<int*>(smem + S.out_pos);
  int* sOutRow = reinterpret_cast<int*>(smem + S.out_row);
  float* sOutNrm = reinterpret_cast<float*>(smem + S.out_nrm);
  float2* sW = reinterpret_cast<float2*>(smem + S.w);
  float2* sDn = reinterpret_cast<float2*>(smem + S.dnrm);
  float2* sDt = reinterpret_cast<float2*>(smem + S.dt);
  float* sRow = reinterpret_cast<float*>(smem + S.rowf);             // [rows][8]
  float* sVec = reinterpret_cast<float*>(smem + S.vec);
  double* sPart = reinterpret_cast<double*>(smem + S.part);
  double* sTot = reinterpret_cast<double*>(smem + S.tot);
  float* sPool = reinterpret_cast<float*>(smem + S.pool);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const FsgWs ws = fsg_ws(c);
  const int L = c.L, F = c.F;

  // ---- before the dependency wait.  The immediate predecessor (the readout backward) only writes the gradient of
  // the pooled embeddings; everything else this kernel reads -- the block records, the CSRs, the forward pass's activations,
  // masks and BatchNorm records, the weight images -- was complete before the predecessor could start, so the whole
  // block-local set-up overlaps the predecessor's run. ----
  if (warp == 0) umma::tmem_alloc(&tmem_slot, kTmemColsB);
  if (t == 0) {
    umma::mbar_init(&bar_w, 1);
    umma::mbar_init(&bar_w2, 1);
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_init(&bar_dw, 1);
    umma::mbar_fence_init();
  }
  float wn0[4], wn1[4], wp0[4], wp1[4], wq0[4], wq1[4];
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
  // one block per graph: G live blocks, each with the record its forward block left (fitting the limits or not)
  const int N = imin(imax(c.dims[0], 0), c.Nm);
  const int G = imin(imax(c.dims[2], 0), c.Bm);
  int4 ia = make_int4(0, 0, 0, 0), ib = make_int4(0, 0, 0, 0);
  if ((int)blockIdx.x < G) {
    ia = *reinterpret_cast<const int4*>(ws.info + (size_t)blockIdx.x * 8);
    ib = *reinterpret_cast<const int4*>(ws.info + (size_t)blockIdx.x * 8 + 4);
  }
  const bool active = ib.w != 0;
  const bool unfit = (int)blockIdx.x < G && !active;
  uint32_t par_w = 0, par_w2 = 0, par_m = 0, par_d = 0;
  int g0 = 0, n0 = 0, Nc = 0, Ec = 0, ie0 = 0, oe0 = 0;
  if (active) {
    g0 = ia.x; n0 = ia.z; Nc = ia.w - ia.z;
    ie0 = ib.x; Ec = ib.y - ib.x; oe0 = ib.z;
    // ================= stage 0: block-local structure, masks, BatchNorm records =================
    if (t == 0) {                                                      // context_convs backward image (whole)
      umma::mbar_expect_tx(&bar_w, 2u * 65536u);
      umma::bulk_g2s(sAh, fsg_img_bwd(ws, L), 65536u, &bar_w);
      umma::bulk_g2s(sAl, fsg_img_bwd(ws, L) + kFsgImgPart, 65536u, &bar_w);
    }
    rows_async(sX, c.agg + (size_t)n0 * FH, Nc);                       // agg of the causal branch (weight gradient)
    for (int i = t; i <= Nc; i += FT) {
      sInPtr[i] = c.in_ptr[n0 + i] - ie0;
      sOutPtr[i] = c.out_ptr[n0 + i] - oe0;
    }
    for (int e = t; e < Ec; e += FT) {
      sInSrc[e] = c.in_src[ie0 + e] - n0;
      sOutDst[e] = c.out_dst[oe0 + e] - n0;
      sOutPos[e] = c.out_pos[oe0 + e] - ie0;
      sOutNrm[e] = c.out_norm[oe0 + e];
      sW[e] = *reinterpret_cast<const float2*>(c.watt + (size_t)(ie0 + e) * 2);
    }
    for (int i = t; i < Nc; i += FT) {
      const float2 a = *reinterpret_cast<const float2*>(c.natt + (size_t)(n0 + i) * 2);
      const float2 d = *reinterpret_cast<const float2*>(c.disw + (size_t)(n0 + i) * 2);
      float* r = sRow + i * 8;
      r[0] = a.x; r[1] = a.y; r[2] = d.x; r[3] = d.y; r[4] = 0.f; r[5] = 0.f;
      for (int q = c.out_ptr[n0 + i] - oe0, q1 = c.out_ptr[n0 + i + 1] - oe0; q < q1; ++q) sOutRow[q] = i;
    }
    {
      // records of bnc (threads 0..127) / bno (128..255): sc | sh | mean | rstd  (set br at sVec + br * 6 * FH)
      const int br = t >> 7, k = t & 127, id = L + 1 + br;
      float* v = sVec + br * 6 * FH;
      v[k] = c.bnf(id, BN_SCALE)[k];
      v[FH + k] = c.bnf(id, BN_SHIFT)[k];
      v[2 * FH + k] = c.bnf(id, BN_MEAN)[k];
      v[3 * FH + k] = c.bnf(id, BN_RSTD)[k];
    }
  }
  // the masked convs' outputs of this warp's rows (ReLU masks of stage 1): branch 0 now, branch 1 while branch 0 runs
  float4 zpre[kRowsPerWarp];
#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int i = warp + rr * 8;
    zpre[rr] = (active && i < Nc) ? __ldcg(reinterpret_cast<const float4*>(c.Z + (size_t)(n0 + i) * FH + lane * 4))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  umma::fence_before_sync();
  FSG_TDECL
  pdl_sync();
  // (after the wait: when this kernel is issued back to back its predecessor is its own previous launch, whose final
  // all-reduce site moves the epoch)
  const int fx_set = fsg_epoch_begin(ws, 1, 12, 13 + L);             // every CTA of the grid, active or not
  if (blockIdx.x == 0) CAL_TL(c.status, 8);
  CAL_TLC(c, 1, 0);
  FSG_T(0);                                                           // 0: dependency wait (set-up overlapped)
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (active) {
    const int npad = imax(8, (Nc + 7) & ~7);
    float* part = ws.part + (size_t)blockIdx.x * fsg_part_floats(L, F);
    {
      // global_add_pool backward = broadcast of the pooled gradient (model.py:115-116); the c <- co path goes
      // through the inverse permutation (model.py:152-157)
      const int br = t >> 7, k = t & 127;
      const size_t H2 = 2 * FH;
      float g;
      if (br == 0) g = c.du[((size_t)0 * c.Bm + g0) * H2 + k] + c.du[((size_t)2 * c.Bm + c.invperm[g0]) * H2 + k];
      else g = c.du[((size_t)1 * c.Bm + g0) * H2 + k] + c.du[((size_t)2 * c.Bm + g0) * H2 + (c.cat ? FH : 0) + k];
      sPool[br * FH + k] = g;
    }
    __syncthreads();
    FSG_T(1);                                                         // 1: pooled gradient

    // ================= stage 1: the two masked convs, dense part (model.py:112-113 backward) =================
    //   dz = dpool * relu'(z);  d agg = dz W^T (tensor cores);  dW += agg^T dz (tensor cores);  db += dz
    for (int br = 0; br < 2; ++br) {
      const float4 gp = *reinterpret_cast<const float4*>(sPool + br * FH + lane * 4);
      float dbias[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int i = warp + rr * 8;
        if (i >= npad) break;
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < Nc) {
          const float4 z = zpre[rr];
          u = make_float4(z.x > 0.f ? gp.x : 0.f, z.y > 0.f ? gp.y : 0.f, z.z > 0.f ? gp.z : 0.f, z.w > 0.f ? gp.w : 0.f);
          dbias[0] += u.x; dbias[1] += u.y; dbias[2] += u.z; dbias[3] += u.w;
        }
        put_b(sBh, sBl, i, lane, u);
      }
      if (br == 0) {                                                   // branch 1's masks: in flight under branch 0's products
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
          const int i = warp + rr * 8;
          if (i < Nc) zpre[rr] = __ldcg(reinterpret_cast<const float4*>(c.Z + (size_t)c.Nm * FH + (size_t)(n0 + i) * FH + lane * 4));
        }
      }
      umma::fence_async_smem();
      cp_async_wait_all();
      __syncthreads();
      if (t == 0) {
        umma::mbar_wait(&bar_w, par_w);
        umma::fence_after_sync();
        issue_3xtf32(sAh, sAl, sBh, sBl, tmem + (br ? 64u : 0u), tmem + 128u + (br ? 64u : 0u), FH / 8, npad);
        umma::commit(&bar_mma);
      }
      par_w ^= 1u;
      colsum8(dbias, reinterpret_cast<float*>(sPart), part + fsg_part_conv(L + br) + FH * FH);
      FSG_T(2);                                                       // 2: dz operand + issue
      umma::mbar_wait(&bar_mma, par_m);
      par_m ^= 1u;
      umma::fence_after_sync();
      FSG_T(4);                                                       // 4: MMA tail
      // the weight image is dead: the transposed operands of dW = agg^T dz go there
      build_dw_operands<false>(sAh, sX, FH, nullptr, nullptr, sBh, sBl, Nc, npad);
      umma::fence_async_smem();
      __syncthreads();
      if (t == 0) {
        umma::fence_after_sync();
        issue_dw(sAh, tmem + kTmemDw, npad);
        umma::commit(&bar_dw);
      }
      umma::mbar_wait(&bar_dw, par_d);
      par_d ^= 1u;
      umma::fence_after_sync();
      __syncthreads();                                                // sX / the operands / the image buffer are free
      if (br == 0) {
        rows_async(sX, c.agg + (size_t)c.Nm * FH + (size_t)n0 * FH, Nc);
        if (t == 0) {
          umma::mbar_expect_tx(&bar_w, 2u * 65536u);
          umma::bulk_g2s(sAh, fsg_img_bwd(ws, L + 1), 65536u, &bar_w);
          umma::bulk_g2s(sAl, fsg_img_bwd(ws, L + 1) + kFsgImgPart, 65536u, &bar_w);
        }
      } else {
        rows_async(sX, c.Xl(L) + (size_t)n0 * FH, Nc);                 // x_{L+1} rows for the sparse part
      }
      drain_dw(tmem + kTmemDw, part + fsg_part_conv(L + br), reinterpret_cast<float*>(sBh));
      umma::fence_before_sync();
      __syncthreads();                                                // (the scratch is the node-operand buffer)
      umma::fence_after_sync();
      FSG_T(3);                                                       // 3: weight gradient
    }
    // d agg of both branches, TMEM -> row-major tiles in the (free) operand buffers: thread = channel
    float* sR0 = reinterpret_cast<float*>(sBh);
    float* sR1 = reinterpret_cast<float*>(sBl);
    float* sY0 = reinterpret_cast<float*>(sAh);                        // y = bn_k(att_k x) tiles in the idle image buffer
    float* sY1 = sY0 + kFsgRows * kLdR;
    {
      const int ch = (warp & 3) * 32 + lane;
      for (int g8 = (warp >> 2) * 8; g8 < npad; g8 += 16) {
        float vm[8], vc[8], um[8], uc[8];
        umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, g8), vm);
        umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + g8), vc);
        umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, 64 + g8), um);
        umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, 192 + g8), uc);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int i = g8 + e;
          if (i < Nc) {
            sR0[i * kLdR + ch] = vm[e] + vc[e];
            sR1[i * kLdR + ch] = um[e] + uc[e];
          }
        }
      }
    }
    umma::fence_before_sync();
    cp_async_wait_all();
    __syncthreads();
    umma::fence_after_sync();
    FSG_T(5);                                                         // 5: d agg tiles

    // ================= stage 2: masked convs, sparse part =================
    //   dy_j = sum_{e: row_e = j} norm_e d agg[col_e]  (warp per source row j);
    //   d norm_e = <d agg[col_e], y_j>,  y_j = bn_k(att_k[j] x_j)  (quarter warp per (entry, branch))
    float4 dyk[kRowsPerWarp][2];                                      // dy of this warp's rows, kept across the all-reduce
    {
      double st[4][4];
#pragma unroll
      for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int i = 0; i < 4; ++i) st[v][i] = 0.0;
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int j = warp + rr * 8;
        dyk[rr][0] = dyk[rr][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= Nc) continue;
        const float4 x = *reinterpret_cast<const float4*>(sX + j * FH + lane * 4);
        const int q0 = sOutPtr[j], q1 = sOutPtr[j + 1];
        const float dj0 = sRow[j * 8 + 2], dj1 = sRow[j * 8 + 3];
        float4 dy0 = make_float4(0.f, 0.f, 0.f, 0.f), dy1 = dy0;
        for (int q = q0; q < q1; ++q) {
          const int dd = sOutDst[q];
          const float2 wa = sW[sOutPos[q]];
          const float w0 = (dj0 * wa.x) * sRow[dd * 8 + 2], w1 = (dj1 * wa.y) * sRow[dd * 8 + 3];
          const float4 ga = *reinterpret_cast<const float4*>(sR0 + dd * kLdR + lane * 4);
          const float4 gb = *reinterpret_cast<const float4*>(sR1 + dd * kLdR + lane * 4);
          dy0.x = fmaf(w0, ga.x, dy0.x); dy0.y = fmaf(w0, ga.y, dy0.y); dy0.z = fmaf(w0, ga.z, dy0.z); dy0.w = fmaf(w0, ga.w, dy0.w);
          dy1.x = fmaf(w1, gb.x, dy1.x); dy1.y = fmaf(w1, gb.y, dy1.y); dy1.z = fmaf(w1, gb.z, dy1.z); dy1.w = fmaf(w1, gb.w, dy1.w);
        }
        dyk[rr][0] = dy0;
        dyk[rr][1] = dy1;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float* v = sVec + k * 6 * FH;
          const float4 sc = *reinterpret_cast<const float4*>(v + lane * 4), sh = *reinterpret_cast<const float4*>(v + FH + lane * 4);
          const float4 mu = *reinterpret_cast<const float4*>(v + 2 * FH + lane * 4), rs = *reinterpret_cast<const float4*>(v + 3 * FH + lane * 4);
          const float aj = sRow[j * 8 + k];
          const float4 xm = make_float4(aj * x.x, aj * x.y, aj * x.z, aj * x.w);
          const float4 y = make_float4(fmaf(xm.x, sc.x, sh.x), fmaf(xm.y, sc.y, sh.y), fmaf(xm.z, sc.z, sh.z), fmaf(xm.w, sc.w, sh.w));
          const float4 xh = make_float4((xm.x - mu.x) * rs.x, (xm.y - mu.y) * rs.y, (xm.z - mu.z) * rs.z, (xm.w - mu.w) * rs.w);
          *reinterpret_cast<float4*>((k ? sY1 : sY0) + j * kLdR + lane * 4) = y;
          const float4 dy = k ? dy1 : dy0;
          st[2 * k][0] += (double)dy.x; st[2 * k][1] += (double)dy.y; st[2 * k][2] += (double)dy.z; st[2 * k][3] += (double)dy.w;
          st[2 * k + 1][0] += (double)dy.x * (double)xh.x; st[2 * k + 1][1] += (double)dy.y * (double)xh.y;
          st[2 * k + 1][2] += (double)dy.z * (double)xh.z; st[2 * k + 1][3] += (double)dy.w * (double)xh.w;
        }
      }
      __syncthreads();                                                // the y tiles are complete
      FSG_T(6);                                                       // 6: masked gather, rows (dy, y tiles)
      {
        const int sub = lane & 7, grp = t >> 3;                        // 32 quarter warps; lane `sub` owns 16 channels
        for (int it0 = 0; it0 < 2 * Ec; it0 += FT / 8) {               // warp-uniform trip count (full-mask shuffles)
          const int it = it0 + grp;
          const bool ok = it < 2 * Ec;
          const int q = ok ? it >> 1 : 0, k = it & 1;
          const float* g = (k ? sR1 : sR0) + sOutDst[q] * kLdR + sub * 16;
          const float* y = (k ? sY1 : sY0) + sOutRow[q] * kLdR + sub * 16;
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(g + 4 * i), b = *reinterpret_cast<const float4*>(y + 4 * i);
            dot = fmaf(a.x, b.x, dot); dot = fmaf(a.y, b.y, dot); dot = fmaf(a.z, b.z, dot); dot = fmaf(a.w, b.w, dot);
          }
          dot += __shfl_xor_sync(0xffffffffu, dot, 4);
          dot += __shfl_xor_sync(0xffffffffu, dot, 2);
          dot += __shfl_xor_sync(0xffffffffu, dot, 1);
          if (ok && sub == 0) {
            if (k) sDn[sOutPos[q]].y = dot;
            else sDn[sOutPos[q]].x = dot;
          }
        }
      }
      umma::fence_async_smem();                                       // (the y tiles sit where the next weight image lands)
      // block totals [bnc: sum dy | sum dy xhat | bno: sum dy | sum dy xhat]
      __syncthreads();                                                // the tiles are dead: the d agg buffer is scratch
      {
        double* sc8 = reinterpret_cast<double*>(sBh);                  // [4][8][128] doubles = 32 KB of the 45 KB operand area
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
          for (int i = 0; i < 4; ++i) sc8[(v * 8 + warp) * FH + lane * 4 + i] = st[v][i];
        __syncthreads();
        for (int i = t; i < 4 * FH; i += FT) {
          const int v = i >> 7, k = i & 127;
          double a = 0.0;
#pragma unroll
          for (int w = 0; w < 8; ++w) a += sc8[(v * 8 + w) * FH + k];
          sPart[i] = a;
        }
        __syncthreads();
      }
    }
    FSG_T(4);                                                         // 4 (+ MMA tail): masked gather, d-norm dots + block totals
    fsg_publish_fx(ws, fx_set, 12, G, sPart, 4 * FH);
    if (t == 0) {                                                      // image of the top backbone layer (all but the tail)
      umma::fence_async_smem();
      umma::mbar_expect_tx(&bar_w, 65536u + kTailOff);
      umma::bulk_g2s(sAh, fsg_img_bwd(ws, L - 1), 65536u, &bar_w);
      umma::bulk_g2s(sAl, fsg_img_bwd(ws, L - 1) + kFsgImgPart, kTailOff, &bar_w);
    }
    FSG_T(7);                                                         // 7: publish

    // ================= stage 3: weighted-norm backward (warp per node; overlaps the all-reduce) =================
    for (int n = warp; n < Nc; n += 8) {
      const int q0 = sOutPtr[n], q1 = sOutPtr[n + 1] - 1;              // q1 = the appended self loop
      const int p0 = sInPtr[n], p1 = sInPtr[n + 1] - 1;                // p1 = the appended self loop
      if (c.no_eatt) {
        for (int q = q0 + lane; q <= q1; q += 32) sDt[sOutPos[q]] = make_float2(0.f, 0.f);
        if (lane == 0) sRow[n * 8 + 4] = sRow[n * 8 + 5] = 0.f;
        continue;
      }
      const float2 dn = make_float2(sRow[n * 8 + 2], sRow[n * 8 + 3]);
      float dd0 = 0.f, dd1 = 0.f;                                     // d dis[n]
      for (int q = q0 + lane; q < q1; q += 32) {
        const int pos = sOutPos[q], d = sOutDst[q];
        const float2 g = sDn[pos], w = sW[pos];
        dd0 = fmaf(g.x * w.x, sRow[d * 8 + 2], dd0);
        dd1 = fmaf(g.y * w.y, sRow[d * 8 + 3], dd1);
      }
      for (int p = p0 + lane; p < p1; p += 32) {
        const int s = sInSrc[p];
        const float2 g = sDn[p], w = sW[p];
        dd0 = fmaf(g.x * w.x, sRow[s * 8 + 2], dd0);
        dd1 = fmaf(g.y * w.y, sRow[s * 8 + 3], dd1);
      }
      if (lane == 0) {
        const float2 g = sDn[p1];
        dd0 = fmaf(2.f * dn.x, g.x, dd0);
        dd1 = fmaf(2.f * dn.y, g.y, dd1);
      }
      dd0 = warp_sum(dd0);
      dd1 = warp_sum(dd1);
      const float ddeg0 = -0.5f * dn.x * dn.x * dn.x * dd0;           // d deg = -1/2 deg^-3/2 d dis
      const float ddeg1 = -0.5f * dn.y * dn.y * dn.y * dd1;
      float dp0 = 0.f, dp1 = 0.f;
      for (int q = q0 + lane; q < q1; q += 32) {
        const int pos = sOutPos[q], d = sOutDst[q];
        const float2 g = sDn[pos], w = sW[pos];
        const float dw0 = fmaf(g.x * dn.x, sRow[d * 8 + 2], ddeg0);
        const float dw1 = fmaf(g.y * dn.y, sRow[d * 8 + 3], ddeg1);
        const float dot = w.x * dw0 + w.y * dw1;
        const float dt0 = w.x * (dw0 - dot), dt1 = w.y * (dw1 - dot);
        sDt[pos] = make_float2(dt0, dt1);
        dp0 += dt0;
        dp1 += dt1;
      }
      dp0 = warp_sum(dp0);
      dp1 = warp_sum(dp1);
      if (lane == 0) {
        sDt[sOutPos[q1]] = make_float2(0.f, 0.f);
        sRow[n * 8 + 4] = dp0;
        sRow[n * 8 + 5] = dp1;
      }
    }
    __syncthreads();
    FSG_T(8);                                                         // 8: norm backward

    // ================= stage 4: attention backward -> gradient rows of the top backbone layer =================
    const long long go_c = c.bn_gamma[L + 1], bo_c = c.bn_beta[L + 1], go_o = c.bn_gamma[L + 2], bo_o = c.bn_beta[L + 2];
    fsg_wait_total_fx(ws, fx_set, 12, G, 4 * FH, sTot);
    CAL_TLC(c, 1, 4);
    fsg_bn_bwd_finalize(c, L + 1, N, go_c, bo_c, sTot, sVec + 4 * FH, sVec + 5 * FH, 0);
    fsg_bn_bwd_finalize(c, L + 2, N, go_o, bo_o, sTot + 2 * FH, sVec + 6 * FH + 4 * FH, sVec + 6 * FH + 5 * FH, FH);
    __syncthreads();
    FSG_T(9);                                                         // 9: all-reduce wait
    {
      float g_wn0[4], g_wn1[4], g_wp0[4], g_wp1[4], g_wq0[4], g_wq1[4], dbias[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) g_wn0[i] = g_wn1[i] = g_wp0[i] = g_wp1[i] = g_wq0[i] = g_wq1[i] = dbias[i] = 0.f;
      float g_bn0 = 0.f, g_bn1 = 0.f, g_be0 = 0.f, g_be1 = 0.f;
      float bsc[2][4], bmu[2][4], brs[2][4], bc1[2][4], bc2[2][4];
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* v = sVec + k * 6 * FH + lane * 4 + i;
          bsc[k][i] = v[0];
          bmu[k][i] = v[2 * FH];
          brs[k][i] = v[3 * FH];
          bc1[k][i] = v[4 * FH];
          bc2[k][i] = v[5 * FH];
        }
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int n = warp + rr * 8;
        if (n >= Nc) continue;
        float dq0 = 0.f, dq1 = 0.f;
        for (int p = sInPtr[n] + lane; p < sInPtr[n + 1]; p += 32) {
          const float2 d = sDt[p];
          dq0 += d.x;
          dq1 += d.y;
        }
        const float* r = sRow + n * 8;
        const float a0 = r[0], a1 = r[1], dpx = r[4], dpy = r[5];
        const float4 x4 = *reinterpret_cast<const float4*>(sX + n * FH + lane * 4);
        const float x[4] = {x4.x, x4.y, x4.z, x4.w};
        const float dc[4] = {dyk[rr][0].x, dyk[rr][0].y, dyk[rr][0].z, dyk[rr][0].w};
        const float dO[4] = {dyk[rr][1].x, dyk[rr][1].y, dyk[rr][1].z, dyk[rr][1].w};
        float gc[4], go[4];
        float da0 = 0.f, da1 = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xhc = (a0 * x[i] - bmu[0][i]) * brs[0][i];
          const float xho = (a1 * x[i] - bmu[1][i]) * brs[1][i];
          gc[i] = bsc[0][i] * (dc[i] - bc1[0][i] - xhc * bc2[0][i]);
          go[i] = bsc[1][i] * (dO[i] - bc1[1][i] - xho * bc2[1][i]);
          da0 = fmaf(gc[i], x[i], da0);
          da1 = fmaf(go[i], x[i], da1);
        }
        // four warp sums at once: the butterfly's independent shuffles pipeline
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          dq0 += __shfl_xor_sync(0xffffffffu, dq0, o);
          dq1 += __shfl_xor_sync(0xffffffffu, dq1, o);
          da0 += __shfl_xor_sync(0xffffffffu, da0, o);
          da1 += __shfl_xor_sync(0xffffffffu, da1, o);
        }
        float ds0 = 0.f, ds1 = 0.f;
        if (!c.no_natt) {
          const float dot = a0 * da0 + a1 * da1;
          ds0 = a0 * (da0 - dot);
          ds1 = a1 * (da1 - dot);
        }
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = a0 * gc[i] + a1 * go[i];
          v = fmaf(ds0, wn0[i], v);
          v = fmaf(ds1, wn1[i], v);
          v = fmaf(dpx, wp0[i], v);
          v = fmaf(dpy, wp1[i], v);
          v = fmaf(dq0, wq0[i], v);
          v = fmaf(dq1, wq1[i], v);
          // the top layer has no BatchNorm above it: gradient w.r.t. its pre-activation = relu'(x_{L+1}) * dX
          o[i] = x[i] > 0.f ? v : 0.f;
          dbias[i] += o[i];
          g_wn0[i] = fmaf(ds0, x[i], g_wn0[i]);
          g_wn1[i] = fmaf(ds1, x[i], g_wn1[i]);
          g_wp0[i] = fmaf(dpx, x[i], g_wp0[i]);
          g_wp1[i] = fmaf(dpy, x[i], g_wp1[i]);
          g_wq0[i] = fmaf(dq0, x[i], g_wq0[i]);
          g_wq1[i] = fmaf(dq1, x[i], g_wq1[i]);
        }
        *reinterpret_cast<float4*>(sG + n * FH + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
        g_bn0 += ds0;
        g_bn1 += ds1;
        g_be0 += dpx;
        g_be1 += dpy;
      }
      // per-block partials of the attention parameters and of the top layer's bias: one pass over [7][8 warps][128]
      float* sRed = reinterpret_cast<float*>(sBh);                    // 28 KB of the (idle) node-operand buffers
      __syncthreads();
      FSG_T(10);                                                      // 10: attention backward, rows
      {
        float* q = sRed + warp * FH + lane * 4;
        *reinterpret_cast<float4*>(q + 0 * 8 * FH) = make_float4(g_wn0[0], g_wn0[1], g_wn0[2], g_wn0[3]);
        *reinterpret_cast<float4*>(q + 1 * 8 * FH) = make_float4(g_wn1[0], g_wn1[1], g_wn1[2], g_wn1[3]);
        *reinterpret_cast<float4*>(q + 2 * 8 * FH) = make_float4(g_wp0[0], g_wp0[1], g_wp0[2], g_wp0[3]);
        *reinterpret_cast<float4*>(q + 3 * 8 * FH) = make_float4(g_wq0[0], g_wq0[1], g_wq0[2], g_wq0[3]);
        *reinterpret_cast<float4*>(q + 4 * 8 * FH) = make_float4(g_wp1[0], g_wp1[1], g_wp1[2], g_wp1[3]);
        *reinterpret_cast<float4*>(q + 5 * 8 * FH) = make_float4(g_wq1[0], g_wq1[1], g_wq1[2], g_wq1[3]);
        *reinterpret_cast<float4*>(q + 6 * 8 * FH) = make_float4(dbias[0], dbias[1], dbias[2], dbias[3]);
        if (lane == 0) {
          float* sc4 = sRed + 7 * 8 * FH + warp * 4;
          sc4[0] = g_bn0; sc4[1] = g_bn1; sc4[2] = g_be0; sc4[3] = g_be1;
        }
      }
      __syncthreads();
      float* pa = part + fsg_part_att(L);
      for (int i = t; i < 7 * FH; i += FT) {
        const int v = i >> 7, k = i & 127;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sRed[(v * 8 + w) * FH + k];
        if (v < 6) pa[v * FH + k] = s;
        else part[fsg_part_conv(L - 1) + FH * FH + k] = s;
      }
      if (t < 4) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += sRed[7 * 8 * FH + w * 4 + t];
        pa[6 * FH + t] = s;
      }
    }
    __syncthreads();
    FSG_T(1);                                                         // 1 (+ pooled gradient): attention backward, partial sums

    // ================= stage 5: backbone layers L-1 .. 0 =================
    //   sG: g = gradient w.r.t. the layer's pre-activation;  u_j = sum_{e: row_e = j} norm_e g[col_e];
    //   D = u W^T (gradient w.r.t. bn_l output, stays in TMEM);  dW += bn_l(x_in)^T u;  sums of bn_l backward.
    // Vector sets alternate: set (l & 1) holds bn_l = bns_conv[l] (id 1 + l): sc | sh | mean | rstd | c1 | c2.
    for (int l = L - 1; l >= 0; --l) {
      float* vin = sVec + (l & 1) * 6 * FH;
      rows_async(sX, c.Xl(l) + (size_t)n0 * FH, Nc);                   // x_in (x_up has been consumed)
      if (t < FH) {
        vin[t] = c.bnf(1 + l, BN_SCALE)[t];
        vin[FH + t] = c.bnf(1 + l, BN_SHIFT)[t];
        vin[2 * FH + t] = c.bnf(1 + l, BN_MEAN)[t];
        vin[3 * FH + t] = c.bnf(1 + l, BN_RSTD)[t];
      }
      for (int j = warp; j < npad; j += 8) {
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < Nc) {
          for (int q = sOutPtr[j]; q < sOutPtr[j + 1]; ++q) {
            const float w = sOutNrm[q];
            const float4 g = *reinterpret_cast<const float4*>(sG + sOutDst[q] * FH + lane * 4);
            u.x = fmaf(w, g.x, u.x); u.y = fmaf(w, g.y, u.y); u.z = fmaf(w, g.z, u.z); u.w = fmaf(w, g.w, u.w);
          }
        }
        put_b(sBh, sBl, j, lane, u);
      }
      umma::fence_async_smem();
      __syncthreads();                                                // the gradient rows are dead: the image tail may land
      FSG_T(11);                                                      // 11: transpose aggregate
      if (t == 0) {
        umma::mbar_expect_tx(&bar_w2, (uint32_t)kGBytes);
        umma::bulk_g2s(sAl + kTailOff, fsg_img_bwd(ws, l) + kFsgImgPart + kTailOff / 4, (uint32_t)kGBytes, &bar_w2);
        umma::mbar_wait(&bar_w, par_w);
        umma::fence_after_sync();
        const uint32_t idesc = umma::instr_desc(umma::kFmtTF32, 128, npad);
        constexpr uint64_t da = (2u * kALbo) >> 4, db = (2u * kBLbo) >> 4;      // descriptor step: start-address field only
        uint64_t ah = umma::smem_desc(umma::smem_addr(sAh), kALbo, kASbo), al = umma::smem_desc(umma::smem_addr(sAl), kALbo, kASbo);
        uint64_t bh = umma::smem_desc(umma::smem_addr(sBh), kBLbo, kBSbo), bl = umma::smem_desc(umma::smem_addr(sBl), kBLbo, kBSbo);
        const uint64_t bh0 = bh;
        // the two products of the hi image first, the product of the lo image last: its tail is still landing
#pragma unroll 4
        for (int s = 0; s < FH / 8; ++s) {
          umma::mma_tf32(tmem, ah, bh, idesc, s > 0);
          umma::mma_tf32(tmem + 128u, ah, bl, idesc, s > 0);
          ah += da; bh += db; bl += db;
        }
        bh = bh0;
#pragma unroll 4
        for (int s = 0; s < FH / 8; ++s) {
          if (s == kTailStep) umma::mbar_wait(&bar_w2, par_w2);
          umma::mma_tf32(tmem + 128u, al, bh, idesc, 1u);
          al += da; bh += db;
        }
        umma::commit(&bar_mma);
      }
      par_w ^= 1u;
      par_w2 ^= 1u;
      cp_async_wait_all();                                            // x_in rows (in flight since the top of the iteration)
      umma::mbar_wait(&bar_mma, par_m);
      par_m ^= 1u;
      umma::fence_after_sync();
      __syncthreads();
      FSG_T(12);                                                      // 12: MMA
      // sums of bn_l backward: thread = channel, the two warp sets split the row groups
      {
        const int ch = (warp & 3) * 32 + lane;
        const float mu = vin[2 * FH + ch], rs = vin[3 * FH + ch];
        double s = 0.0, q = 0.0;
        for (int g8 = (warp >> 2) * 8; g8 < npad; g8 += 16) {
          float vm[8], vc[8];
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, g8), vm);
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + g8), vc);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = g8 + e;
            if (i < Nc) {
              const float d = vm[e] + vc[e];
              const float xh = (sX[i * FH + ch] - mu) * rs;
              s += (double)d;
              q += (double)d * (double)xh;
            }
          }
        }
        if (warp >= 4) {
          sTot[ch] = s;
          sTot[FH + ch] = q;
        }
        __syncthreads();
        if (warp < 4) {
          sPart[ch] = s + sTot[ch];
          sPart[FH + ch] = q + sTot[FH + ch];
        }
        __syncthreads();
      }
      FSG_T(13);                                                      // 13: statistics epilogue
      fsg_publish_fx(ws, fx_set, 13 + (L - 1 - l), G, sPart, 2 * FH, l == 0 ? 1 : -1);   // (layer 0: the launch's final site)
      FSG_T(7);
      // dW = bn_l(x_in)^T u on the tensor cores while the all-reduce travels (operands in the idle image buffer)
      build_dw_operands<true>(sAh, sX, FH, vin, vin + FH, sBh, sBl, Nc, npad);
      umma::fence_async_smem();
      __syncthreads();
      if (t == 0) {
        umma::fence_after_sync();
        issue_dw(sAh, tmem + kTmemDw, npad);
        umma::commit(&bar_dw);
        if (l > 0) {                                                  // next image (all but the tail) as soon as the operands are dead
          umma::mbar_wait(&bar_dw, par_d);
          umma::mbar_expect_tx(&bar_w, 65536u + kTailOff);
          umma::bulk_g2s(sAh, fsg_img_bwd(ws, l - 1), 65536u, &bar_w);
          umma::bulk_g2s(sAl, fsg_img_bwd(ws, l - 1) + kFsgImgPart, kTailOff, &bar_w);
        }
      }
      FSG_T(3);
      const long long go_l = c.bn_gamma[1 + l], bo_l = c.bn_beta[1 + l];
      fsg_wait_total_fx(ws, fx_set, 13 + (L - 1 - l), G, 2 * FH, sTot);
      if (l == 0) CAL_TLC(c, 1, 1);
      fsg_bn_bwd_finalize(c, 1 + l, N, go_l, bo_l, sTot, vin + 4 * FH, vin + 5 * FH, 0);
      umma::mbar_wait(&bar_dw, par_d);
      par_d ^= 1u;
      umma::fence_after_sync();
      __syncthreads();
      FSG_T(9);
      // gradient w.r.t. the pre-activation of the layer below (the input transform for l == 0):
      //   g = relu'(x_in) * bn_l'(D)  -> sG (thread = channel), and its column sums = that layer's bias gradient
      {
        const int ch = (warp & 3) * 32 + lane;
        const float sc = vin[ch], mu = vin[2 * FH + ch], rs = vin[3 * FH + ch], c1 = vin[4 * FH + ch], c2 = vin[5 * FH + ch];
        float db = 0.f;
        for (int g8 = (warp >> 2) * 8; g8 < npad; g8 += 16) {
          float vm[8], vc[8];
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, g8), vm);
          umma::ld8(umma::tmem_addr(tmem, (warp & 3) * 32, 128 + g8), vc);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = g8 + e;
            if (i < Nc) {
              const float x = sX[i * FH + ch];
              const float xh = (x - mu) * rs;
              const float g = x > 0.f ? sc * ((vm[e] + vc[e]) - c1 - xh * c2) : 0.f;
              sG[i * FH + ch] = g;
              db += g;
            }
          }
        }
        float* sRed = reinterpret_cast<float*>(sPart);
        if (warp >= 4) sRed[ch] = db;
        drain_dw(tmem + kTmemDw, part + fsg_part_conv(l), reinterpret_cast<float*>(sBh));
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (warp < 4) {
          const float tot = db + sRed[ch];
          if (l > 0) part[fsg_part_conv(l - 1) + FH * FH + ch] = tot;
          else part[fsg_part_feat(L) + (size_t)F * FH + ch] = tot;     // column sums of g_1 (k_feat_bwd's cs)
        }
      }
      FSG_T(14);                                                      // 14: BatchNorm backward into the rows below + dW drain
    }

    // ================= stage 6: input transform backward: M = xhat_0^T g_1  [F, H] =================
    {
      __syncthreads();
      float* sF = sX;                                                  // [rows][F] xhat_0
      const float* mean0 = c.bnf(0, BN_MEAN);
      const float* rstd0 = c.bnf(0, BN_RSTD);
      for (int i = t; i < Nc * F; i += FT) {
        const int r = i / F, f = i - r * F;
        sF[i] = (c.feat[(size_t)(n0 + r) * F + f] - mean0[f]) * rstd0[f];
      }
      __syncthreads();
      const int m = t & 127;
      for (int f = t >> 7; f < F; f += 2) {
        float acc = 0.f;
        for (int j = 0; j < Nc; ++j) acc = fmaf(sF[j * F + f], sG[j * FH + m], acc);
        part[fsg_part_feat(L) + (size_t)f * FH + m] = acc;
      }
    }
    FSG_T(15);                                                        // 15: input transform backward
  }
  if (unfit && t == 0) fsg_unfit_arrive(ws, fx_set, G, 12, 12 + L, 1);   // (reported by the forward kernel)
  FSG_TDUMP(c, 64);
  if (blockIdx.x == 0) CAL_TL(c.status, 9);
  CAL_TLC(c, 1, 2);

  // ---- teardown: TMEM (the all-reduce state needs none: fsg_epoch_begin) ----
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kTmemColsB);
  CAL_TLC(c, 1, 3);
}

// ---------------------------------------------------------------------------------------------
// Block-order sum of the per-block partial gradients into the flat gradient buffer.
//   blocks [0, n_generic): 256 consecutive elements of the partial vector x 4 block slices (1024 threads);
//   blocks [n_generic, n_generic + F): the input transform, one feature row each:
//     d W_feat = gamma_0 * M + beta_0 (x) cs,  d gamma_0[f] = sum_j W[f][j] M[f][j],  d beta_0[f] = sum_j W[f][j] cs[j]
// ---------------------------------------------------------------------------------------------
constexpr int kRedMax = 2 * (CAL_MAX_LAYERS + 2) + 4;
struct FsgRedTable {
  int count;
  int V;                      // floats per block slot
  int generic_end;            // elements [0, generic_end) of a slot belong to the generic entries
  long long dst[kRedMax];
  int src[kRedMax], n[kRedMax];
};

__global__ void __launch_bounds__(1024) k_fsg_grad_reduce(const Ctx c, const FsgRedTable tb, const int n_generic) {
  pdl_sync();
  if (blockIdx.x == 0) CAL_TL(c.status, 10);
  __shared__ float s_a[4][256], s_b[4][256];
  const FsgWs ws = fsg_ws(c);
  const int np = imin(imax(c.dims[2], 0), c.Bm);                    // one partial per live block
  const int tx = threadIdx.x & 255, ty = threadIdx.x >> 8;
  if ((int)blockIdx.x < n_generic) {
    const int i = blockIdx.x * 256 + tx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (i < tb.generic_end) {
      const float* p = ws.part + i;
      int g = ty;
#pragma unroll 2
      for (; g + 12 < np; g += 16) {
        s0 += __ldcg(p + (size_t)g * tb.V);
        s1 += __ldcg(p + (size_t)(g + 4) * tb.V);
        s2 += __ldcg(p + (size_t)(g + 8) * tb.V);
        s3 += __ldcg(p + (size_t)(g + 12) * tb.V);
      }
      for (; g < np; g += 4) s0 += __ldcg(p + (size_t)g * tb.V);
    }
    s_a[ty][tx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ty == 0 && i < tb.generic_end) {
      const float sum = (s_a[0][tx] + s_a[1][tx]) + (s_a[2][tx] + s_a[3][tx]);
      for (int e = 0; e < tb.count; ++e)
        if (i >= tb.src[e] && i < tb.src[e] + tb.n[e]) {
          c.grads[tb.dst[e] + (i - tb.src[e])] = sum;
          break;
        }
    }
    return;
  }
  const int F = c.F, L = c.L;
  const int f = (int)blockIdx.x - n_generic;
  if (f >= F) return;
  const int j = tx & 127, sl = ty * 2 + (tx >> 7);                     // column, block slice (8 slices)
  const float* pm = ws.part + fsg_part_feat(L) + (size_t)f * FH + j;
  const float* pc = ws.part + fsg_part_feat(L) + (size_t)F * FH + j;
  float m0 = 0.f, m1 = 0.f, c0 = 0.f, c1 = 0.f;
  int g = sl;
  for (; g + 8 < np; g += 16) {
    m0 += __ldcg(pm + (size_t)g * tb.V);
    m1 += __ldcg(pm + (size_t)(g + 8) * tb.V);
    c0 += __ldcg(pc + (size_t)g * tb.V);
    c1 += __ldcg(pc + (size_t)(g + 8) * tb.V);
  }
  for (; g < np; g += 8) {
    m0 += __ldcg(pm + (size_t)g * tb.V);
    c0 += __ldcg(pc + (size_t)g * tb.V);
  }
  s_a[ty][tx] = m0 + m1;
  s_b[ty][tx] = c0 + c1;
  __syncthreads();
  __shared__ float s_g[4], s_d[4];
  float dg = 0.f, db = 0.f;
  if (threadIdx.x < FH) {
    float m = 0.f, cs = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      m += s_a[q][j] + s_a[q][FH + j];
      cs += s_b[q][j] + s_b[q][FH + j];
    }
    const float g0 = c.params[c.po.bn_feat_w + f], b0 = c.params[c.po.bn_feat_b + f];
    c.grads[c.po.conv_feat_w + (size_t)f * FH + j] = g0 * m + b0 * cs;
    const float w = c.params[c.po.conv_feat_w + (size_t)f * FH + j];
    dg = w * m;
    db = w * cs;
    if (f == 0 && c.po.conv_feat_b >= 0) c.grads[c.po.conv_feat_b + j] = 0.f;   // gfn=True: the bias never gets a gradient
    dg = warp_sum(dg);
    db = warp_sum(db);
    if ((threadIdx.x & 31) == 0) {
      s_g[threadIdx.x >> 5] = dg;
      s_d[threadIdx.x >> 5] = db;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    c.grads[c.po.bn_feat_w + f] = (s_g[0] + s_g[1]) + (s_g[2] + s_g[3]);
    c.grads[c.po.bn_feat_b + f] = (s_d[0] + s_d[1]) + (s_d[2] + s_d[3]);
  }
}

}  // namespace

int launch_fsg_backward(const Ctx& c, cudaStream_t s) {
  const size_t smem = fsg_bsmem().total;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_fsg_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  launch_k(k_fsg_backward, dim3(imax(1, imin(c.Bm, kSMs))), dim3(FT), smem, s, c);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

int launch_fsg_grad_reduce(const Ctx& c, cudaStream_t s) {
  FsgRedTable tb;
  const int L = c.L;
  tb.count = 0;
  tb.V = (int)fsg_part_floats(L, c.F);
  tb.generic_end = (int)fsg_part_feat(L);
  auto add = [&](long long dst, size_t src, int n) {
    if (dst < 0 || tb.count >= kRedMax) return;
    tb.dst[tb.count] = dst;
    tb.src[tb.count] = (int)src;
    tb.n[tb.count] = n;
    ++tb.count;
  };
  for (int l = 0; l < L; ++l) {
    add(c.po.convs_w[l], fsg_part_conv(l), FH * FH);
    add(c.po.convs_b[l], fsg_part_conv(l) + FH * FH, FH);
  }
  add(c.po.context_w, fsg_part_conv(L), FH * FH);
  add(c.po.context_b, fsg_part_conv(L) + FH * FH, FH);
  add(c.po.objects_w, fsg_part_conv(L + 1), FH * FH);
  add(c.po.objects_b, fsg_part_conv(L + 1) + FH * FH, FH);
  const size_t pa = fsg_part_att(L);
  add(c.po.node_att_w, pa, 2 * FH);
  add(c.po.edge_att_w, pa + 2 * FH, 4 * FH);
  add(c.po.node_att_b, pa + 6 * FH, 2);
  add(c.po.edge_att_b, pa + 6 * FH + 2, 2);
  const int n_generic = ceil_div(tb.generic_end, 256);
  launch_k(k_fsg_grad_reduce, dim3(n_generic + c.F), dim3(1024), 0, s, c, tb, n_generic);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
