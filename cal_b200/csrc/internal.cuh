// internal.cuh -- workspace layout, the per-call context handed to every kernel, and the
// launcher prototypes shared by the translation units of libcal_b200.so.
#pragma once
#include "common.cuh"

namespace cal {

constexpr int kTileRows = 24;                  // destination rows per GEMM tile (8 warps x 4 rows)
constexpr int kRPW = kTileRows / kRowWarps;    // rows per warp inside a tile
constexpr int kStageFwd = 128;                 // neighbour rows staged in shared memory per row batch (forward layers)
constexpr int kStageBwd = 112;                 // ... backward layers (two rows per entry)
constexpr int kNumBN = CAL_MAX_BN + 1;         // + the identity record used by the top layer's backward
constexpr int kBnIdentity = CAL_MAX_BN;
constexpr int kHeadRowsPerCta = 8;             // head2 kernels: one warp per graph row
constexpr int kFeatChunk = 64;                 // feature columns per CTA slice in the feat backward
constexpr int kGsGroup = 8;                    // CTAs per first-level group of the hierarchical grid sum
constexpr int kGsSites = 5;                    // grid-sum scratch sites: 0-2 in-kernel sums, 3-4 consumer-side BatchNorm hand-over (alternating)
constexpr int kGsCounters = 64;                // counters per site: [0] top level, [1 + group] first level

// BN record fields inside CAL_WS_BN: [id][field][KMAX]
enum { BN_SCALE = 0, BN_SHIFT, BN_MEAN, BN_RSTD, BN_C1, BN_C2, BN_FIELDS };

__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

struct Layout {
  size_t off[CAL_WS_REGION_COUNT];
  size_t size[CAL_WS_REGION_COUNT];
  size_t total;
  int kmax;        // floats per BN field
  size_t statp_legacy, gs_stride;   // doubles
  int gs_n;
  int g_tile;      // persistent grid of the row-tile (GEMM) kernels
  int g_row;       // grid of the warp-per-row kernels
  int t_head1;     // row tiles of the readout fc1 kernels (per head)
  int g_head2;     // CTAs of the readout fc2 kernels
  int g_feat;      // grid.x of feat kernels
  int n_fchunk;    // grid.y of the feat backward
  // GPART sub-offsets (floats)
  size_t gp_conv[CAL_MAX_LAYERS + 2];   // [G][H*H + H]
  size_t gp_att;                        // [G][8H + 4]
  size_t gp_feat;                       // [G][F*H + H]
  size_t gp_fc1[3];                     // [T1][H*2H + H]
  size_t gp_fc2[3];                     // [G2][C*H + C]
  size_t gp_gat[CAL_MAX_LAYERS];        // [G][2H] attention-vector partials (GAT)
  size_t gp_gin2[CAL_MAX_LAYERS];       // [G][H*H + H] second Linear of a GIN layer
};

// Everything a kernel needs, passed by value.
struct Ctx {
  // model
  int model, F, H, C, L, heads, cat, no_natt, no_eatt, train, readout_bf16, readout_tc;
  float eps, momentum, w_c, w_o, w_co, gat_p;
  // capacities and plan
  int Nm, Em, Bm, EP, kmax, g_tile, g_row, t_head1, g_head2;
  // batch
  const int* dims;
  const float* feat;
  const long long* ei_row;
  const long long* ei_col;
  const long long* batch;
  const long long* y;
  const int* perm_in;
  const float* gat_keep;
  // parameters
  const float* params;
  float* grads;
  float* bn_buffers;
  long long* nbt;
  cal_param_offsets po;
  long long bn_gamma[kNumBN], bn_beta[kNumBN], bn_rm[kNumBN], bn_rv[kNumBN];
  int bn_K[kNumBN];
  // workspace regions
  int* status;
  unsigned int* counters;
  int *in_ptr, *in_src, *in_key, *out_ptr, *out_dst, *out_pos, *out_key, *cnt_in, *cnt_out;
  int *graph_ptr, *node_graph, *perm, *invperm;
  float *in_norm, *dis, *X, *natt, *pq, *watt, *disw, *agg, *Z, *pooled, *H1, *logp, *loss, *bn;
  float *out_norm, *edge_wn, *edge_na;
  int with_loss;
  double* statp;
  double* gsum;             // hierarchical grid-sum scratch: [kGsSites][(Gmax + Gmax/8 + 1) * gs_n] doubles
  unsigned int* gs_cnt;     // [kGsSites][kGsCounters]
  int gs_n;                 // doubles per CTA vector slot (4 * kmax)
  size_t gs_stride;         // doubles per site
  float *WT, *gat, *dlogit, *dh, *du, *dagg, *dym, *dnrm, *dt, *dp, *D, *gpart;
  size_t gp_conv[CAL_MAX_LAYERS + 2], gp_att, gp_feat, gp_fc1[3], gp_fc2[3], gp_gat[CAL_MAX_LAYERS];
  size_t gp_gin2[CAL_MAX_LAYERS];
  const float* grad_logp;   // external dL/dlogp (nullptr = fused loss)
  unsigned char* fsg;       // CAL_WS_FSG region (fsg.cuh)
  int* egp;                 // [Bm+1] first edge column per graph | [Bm] self loops per graph | arrival counter (grouped_edges)
  int grouped;              // cal_caps.grouped_edges
  int raw_o;                // CAL_F_RAW_LOGITS_O: head 1 exchanges raw logits with the caller
  int fsg_on;               // the fused small-graph forward replaces feat .. masked_convs (and the pooling)
  int fsg_bwd_on;           // ... and the fused small-graph backward replaces masked_gemm_bwd .. feat_bwd

  __host__ __device__ float* bnf(int id, int field) const { return bn + ((size_t)id * BN_FIELDS + field) * kmax; }
  __host__ __device__ float* Xl(int l) const { return X + (size_t)l * Nm * H; }     // l = 0..L  (x_{l+1})
  __host__ __device__ float* wt_conv(int l) const { return WT + (size_t)l * H * H; }   // l = 0..L+1
  __host__ __device__ float* wt_fc1(int h) const { return WT + (size_t)(L + 2) * H * H + (size_t)h * 2 * H * H; }
  // CausalGIN: transposed nn.3 weights; h (nn.0 output) of every layer; masked gradient w.r.t. relu(bn(h))
  __host__ __device__ float* wt_gin2(int l) const { return WT + (size_t)(L + 2) * H * H + (size_t)3 * 2 * H * H + (size_t)l * H * H; }
  __host__ __device__ float* gin_h(int l) const { return gat + (size_t)l * Nm * H; }
  __host__ __device__ float* gin_dr() const { return gat + (size_t)L * Nm * H; }
};

// counter slots
enum {
  CNT_FEATSTAT = 0, CNT_FEAT, CNT_CONV0, CNT_MASKED = CNT_CONV0 + CAL_MAX_LAYERS, CNT_HEAD1, CNT_HEAD2,
  CNT_BHEAD2, CNT_BHEAD1, CNT_BGATHER, CNT_BCONV0, CNT_BFEAT = CNT_BCONV0 + CAL_MAX_LAYERS, CNT_GAT0,
  CNT_BGAT0 = CNT_GAT0 + CAL_MAX_LAYERS, CNT_END = CNT_BGAT0 + 2 * CAL_MAX_LAYERS
};
static_assert(CNT_END <= 64, "counter region too small");

void note_launches(int n);     // bumps the process-wide counter behind cal_launch_count()
int compute_layout(const cal_model_desc* m, const cal_caps* caps, Layout* lay);
int validate_model(const cal_model_desc* m);
size_t gat_workspace_floats(int Nm, int EP, int H, int L, int heads);
size_t readout_smem_bytes(int Bm, int H, int cat, int C, int backward);

// ---- launchers (each returns 0 or a cudaError_t) ----
int launch_prep(const Ctx& c, cudaStream_t s);
int launch_param_prep(const Ctx& c, cudaStream_t s);
int launch_feat_forward(const Ctx& c, cudaStream_t s);
int launch_conv_forward(const Ctx& c, int layer, cudaStream_t s);          // layer 0..L-1
int launch_edge_att(const Ctx& c, cudaStream_t s);
int launch_masked_forward(const Ctx& c, cudaStream_t s);
int launch_heads_forward(const Ctx& c, int with_loss, cudaStream_t s);
int launch_heads_backward(const Ctx& c, cudaStream_t s);
bool readout_tc_supported(const Ctx& c);                                   // tensor-core readout (head_tc.cu)
int launch_readout_tc_forward(const Ctx& c, cudaStream_t s);
int launch_readout_tc_backward(const Ctx& c, cudaStream_t s);
bool readout_ro_supported(const Ctx& c);                                   // short-chain readout of the fused small-graph path (head_ro.cu)
int readout_path_id(const Ctx& c);                                          // 0 = the FFMA cluster kernels of head.cu
bool readout_runs_ro(const Ctx& c);                                        // ... and they are the ones that will run (no override)
int launch_readout_ro_forward(const Ctx& c, cudaStream_t s);
int launch_readout_ro_backward(const Ctx& c, cudaStream_t s);
int launch_fsg_prep(const Ctx& c, cudaStream_t s, bool no_overlap = false);                         // fused small-graph path (fsg.cu)
int launch_fsg_forward(const Ctx& c, cudaStream_t s, bool after_full_dependency = false);
int fsg_fill_image_sink(const Ctx& c, cal_image_sink* sink);
size_t fsg_region_bytes(int Bm, int L, int F);
int launch_fsg_backward(const Ctx& c, cudaStream_t s);                      // masked convs .. input transform, one kernel
int launch_fsg_grad_reduce(const Ctx& c, cudaStream_t s);
bool readout_tc2_supported(const Ctx& c);                                  // resident-tile variant (head_tc2.cu)
int launch_readout_tc2_forward(const Ctx& c, cudaStream_t s);
int launch_readout_tc2_backward(const Ctx& c, cudaStream_t s);
int launch_masked_bwd_gemm(const Ctx& c, cudaStream_t s);
int launch_masked_bwd_gather(const Ctx& c, cudaStream_t s);
int launch_norm_backward(const Ctx& c, cudaStream_t s);
int launch_att_backward(const Ctx& c, cudaStream_t s);
int launch_conv_backward(const Ctx& c, int layer, cudaStream_t s);
int launch_feat_backward(const Ctx& c, cudaStream_t s);
int launch_grad_reduce(const Ctx& c, cudaStream_t s);
int launch_gat_forward(const Ctx& c, int layer, cudaStream_t s);
int launch_gat_backward(const Ctx& c, int layer, cudaStream_t s);
int launch_gin_forward(const Ctx& c, int layer, cudaStream_t s);
int launch_gin_backward(const Ctx& c, int layer, cudaStream_t s);

// ---- device helpers shared by the kernels ----

// ---------------------------------------------------------------------------------------------
// Hierarchical, deterministic cross-CTA sum of per-CTA fp64 vectors.
// Every participating CTA (rank in [0, G)) holds its own totals in shared memory sTot[0..n).
// Level 1: CTAs are grouped by 8; the last CTA of a group to arrive sums the group's vectors in
// rank order.  Level 2: the last group leader to arrive sums the group vectors in group order.
// Returns true in exactly ONE CTA, whose sTot then holds the grid totals.  Latency = two L2
// round trips of at most 8 / ceil(G/8) independent loads per thread, instead of one CTA walking
// G partials.  Counters reset themselves (CUDA-graph replay safe).  All threads must call.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool grid_sum(const Ctx& c, int site, double* sTot, int n, int G, int rank) {
  __shared__ int s_flag;
  double* l0 = c.gsum + (size_t)site * c.gs_stride;
  double* l1 = l0 + (size_t)kMaxStatBlocks * c.gs_n;
  unsigned int* cnt = c.gs_cnt + site * kGsCounters;
  const int t = threadIdx.x, T = blockDim.x;
  if (G <= 1) return true;
  const int grp = rank / kGsGroup, ngrp = (G + kGsGroup - 1) / kGsGroup;
  const int gsize = imin(kGsGroup, G - grp * kGsGroup);
  for (int i = t; i < n; i += T) l0[(size_t)rank * n + i] = sTot[i];
  // publish / observe like a cooperative-groups grid barrier: the CTA barrier orders the CTA's stores
  // before thread 0's device-scope fence (cumulative), the fence after the atomic orders the other
  // CTAs' stores before the loads below -- ONE thread fences instead of all 256
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned int old = atomicAdd(&cnt[1 + grp], 1u);
    s_flag = (old == (unsigned int)gsize - 1u);
    if (s_flag) {
      cnt[1 + grp] = 0u;
      __threadfence();
    }
  }
  __syncthreads();
  if (!s_flag) return false;
  for (int i = t; i < n; i += T) {
    double v[kGsGroup];
#pragma unroll
    for (int m = 0; m < kGsGroup; ++m)
      v[m] = m < gsize ? __ldcg(&l0[(size_t)(grp * kGsGroup + m) * n + i]) : 0.0;
    double sum = v[0];
#pragma unroll
    for (int m = 1; m < kGsGroup; ++m) sum += v[m];
    if (ngrp == 1) sTot[i] = sum;
    else l1[(size_t)grp * n + i] = sum;
  }
  if (ngrp == 1) {
    __syncthreads();
    return true;
  }
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned int old = atomicAdd(&cnt[0], 1u);
    s_flag = (old == (unsigned int)ngrp - 1u);
    if (s_flag) {
      cnt[0] = 0u;
      __threadfence();
    }
  }
  __syncthreads();
  if (!s_flag) return false;
  for (int i = t; i < n; i += T) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int g = 0;
    for (; g + 4 <= ngrp; g += 4) {
      s0 += __ldcg(&l1[(size_t)(g + 0) * n + i]);
      s1 += __ldcg(&l1[(size_t)(g + 1) * n + i]);
      s2 += __ldcg(&l1[(size_t)(g + 2) * n + i]);
      s3 += __ldcg(&l1[(size_t)(g + 3) * n + i]);
    }
    for (; g < ngrp; ++g) s0 += __ldcg(&l1[(size_t)g * n + i]);
    sTot[i] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  return true;
}

// The finishing CTA's parameter / running-statistic loads, issued BEFORE the grid sum by every CTA
// (L2 hits, 4 registers) so that they are not one more dependent round trip on the serial tail.
struct BnPre {
  float g, b, rm, rv;
};
__device__ __forceinline__ BnPre bn_prefetch(const Ctx& c, int id, int t0 = 0) {     // channel threadIdx.x - t0
  BnPre p = {1.f, 0.f, 0.f, 1.f};
  const int k = (int)threadIdx.x - t0;
  if (k >= 0 && k < c.bn_K[id]) {
    p.g = c.params[c.bn_gamma[id] + k];
    p.b = c.params[c.bn_beta[id] + k];
    if (c.bn_buffers != nullptr && c.bn_rm[id] >= 0) {
      p.rm = c.bn_buffers[c.bn_rm[id] + k];
      p.rv = c.bn_buffers[c.bn_rv[id] + k];
    }
  }
  return p;
}

// ---------------------------------------------------------------------------------------------
// Consumer-side BatchNorm hand-over (forward chain of CausalGCN / CausalGIN).  The producing kernel
// stops after the FIRST level of the tree: the last CTA of every group of 8 leaves the group's sums
// in l1[group]; nobody waits for the whole grid and nobody finalises.  The consuming kernel starts
// after the producer has completed (programmatic dependency), so every one of its CTAs sums the
// <= 37 group vectors itself -- independent loads, one L2 round trip, the same order as grid_sum's
// second level (bit-identical totals) -- and turns them into the affine it applies; CTA 0 also
// publishes the record (the backward pass reads it) and updates the running statistics.
// This removes the second atomic phase and the serial finalisation from the producer's tail.
// Sites 3 / 4 alternate along the chain (site = 3 + (first BN id & 1)) so that a kernel never
// writes the scratch its own late CTAs may still be reading.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int bn_site(int first_id) { return 3 + (first_id & 1); }

__device__ __forceinline__ void grid_sum_groups(const Ctx& c, int site, const double* sTot, int n, int G, int rank) {
  __shared__ int s_gflag;
  double* l0 = c.gsum + (size_t)site * c.gs_stride;
  double* l1 = l0 + (size_t)kMaxStatBlocks * c.gs_n;
  unsigned int* cnt = c.gs_cnt + site * kGsCounters;
  const int t = threadIdx.x, T = blockDim.x;
  const int grp = rank / kGsGroup;
  const int gsize = imin(kGsGroup, G - grp * kGsGroup);
  if (gsize == 1) {                                    // a group of one: its sums are the group sums
    for (int i = t; i < n; i += T) l1[(size_t)grp * n + i] = sTot[i];
    return;
  }
  for (int i = t; i < n; i += T) l0[(size_t)rank * n + i] = sTot[i];
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned int old = atomicAdd(&cnt[1 + grp], 1u);
    s_gflag = (old == (unsigned int)gsize - 1u);
    if (s_gflag) {
      cnt[1 + grp] = 0u;
      __threadfence();
    }
  }
  __syncthreads();
  if (!s_gflag) return;
  for (int i = t; i < n; i += T) {
    double v[kGsGroup];
#pragma unroll
    for (int m = 0; m < kGsGroup; ++m)
      v[m] = m < gsize ? __ldcg(&l0[(size_t)(grp * kGsGroup + m) * n + i]) : 0.0;
    double sum = v[0];
#pragma unroll
    for (int m = 1; m < kGsGroup; ++m) sum += v[m];
    l1[(size_t)grp * n + i] = sum;
  }
}

// Sum of the group vectors, value v of [voff, voff + nvals), second-level order of grid_sum.
__device__ __forceinline__ double gs_group_total(const double* p, int ngrp, int n_prod) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int g = 0;
  for (; g + 4 <= ngrp; g += 4) {
    s0 += __ldcg(p + (size_t)(g + 0) * n_prod);
    s1 += __ldcg(p + (size_t)(g + 1) * n_prod);
    s2 += __ldcg(p + (size_t)(g + 2) * n_prod);
    s3 += __ldcg(p + (size_t)(g + 3) * n_prod);
  }
  for (; g < ngrp; ++g) s0 += __ldcg(p + (size_t)g * n_prod);
  return ngrp == 1 ? s0 : (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ const double* gs_l1(const Ctx& c, int site) {
  return c.gsum + (size_t)site * c.gs_stride + (size_t)kMaxStatBlocks * c.gs_n;
}
// two vectors at once (both sets of loads in flight together): A -> scrA[0 .. nvals), B -> scrB[0 .. nvals)
__device__ __forceinline__ void gs_sum_pair(const Ctx& c, int siteA, int voffA, int siteB, int voffB, int G, int n_prod,
                                            int nvals, double* scrA, double* scrB) {
  const int ngrp = (G + kGsGroup - 1) / kGsGroup;
  const double* la = gs_l1(c, siteA) + voffA;
  const double* lb = gs_l1(c, siteB) + voffB;
  for (int v = threadIdx.x; v < nvals; v += blockDim.x) {
    const double* pa = la + v;
    const double* pb = lb + v;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    int g = 0;
    for (; g + 4 <= ngrp; g += 4) {
      const double x0 = __ldcg(pa + (size_t)(g + 0) * n_prod), x1 = __ldcg(pa + (size_t)(g + 1) * n_prod);
      const double x2 = __ldcg(pa + (size_t)(g + 2) * n_prod), x3 = __ldcg(pa + (size_t)(g + 3) * n_prod);
      const double y0 = __ldcg(pb + (size_t)(g + 0) * n_prod), y1 = __ldcg(pb + (size_t)(g + 1) * n_prod);
      const double y2 = __ldcg(pb + (size_t)(g + 2) * n_prod), y3 = __ldcg(pb + (size_t)(g + 3) * n_prod);
      a0 += x0; a1 += x1; a2 += x2; a3 += x3;
      b0 += y0; b1 += y1; b2 += y2; b3 += y3;
    }
    for (; g < ngrp; ++g) {
      a0 += __ldcg(pa + (size_t)g * n_prod);
      b0 += __ldcg(pb + (size_t)g * n_prod);
    }
    scrA[v] = ngrp == 1 ? a0 : (a0 + a1) + (a2 + a3);
    scrB[v] = ngrp == 1 ? b0 : (b0 + b1) + (b2 + b3);
  }
}

// forward finalisation of BatchNorm `id` from its totals tot[0 .. 2K) by threads [t0, t0 + K)
__device__ __forceinline__ void bn_fwd_finalize_smem(const Ctx& c, int id, int count, const BnPre& pre, const double* tot,
                                                     float* s_sc, float* s_sh, bool publish, int t0) {
  const int K = c.bn_K[id];
  const int t = (int)threadIdx.x - t0;
  if (t >= 0 && t < K) {
    const double s = tot[t], q = tot[K + t];
    double mean = 0.0, var = 0.0;
    if (count > 0) {
      mean = s / count;
      var = q / count - mean * mean;
      if (var < 0.0) var = 0.0;
    }
    const float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
    const float sc = pre.g * rstd;
    const float sh = pre.b - (float)mean * sc;
    s_sc[t] = sc;
    s_sh[t] = sh;
    if (publish) {
      c.bnf(id, BN_SCALE)[t] = sc;
      c.bnf(id, BN_SHIFT)[t] = sh;
      c.bnf(id, BN_MEAN)[t] = (float)mean;
      c.bnf(id, BN_RSTD)[t] = rstd;
      if (c.bn_buffers != nullptr && c.bn_rm[id] >= 0) {
        const double unb = count > 1 ? var * ((double)count / (double)(count - 1)) : var;
        c.bn_buffers[c.bn_rm[id] + t] = (1.f - c.momentum) * pre.rm + c.momentum * (float)mean;
        c.bn_buffers[c.bn_rv[id] + t] = (1.f - c.momentum) * pre.rv + c.momentum * (float)unb;
      }
      if (t == 0 && c.nbt != nullptr) c.nbt[id] += 1;
    }
  }
}
__device__ __forceinline__ void bn_bwd_finalize_smem(const Ctx& c, int id, int count, const double* tot, float* s_c1,
                                                     float* s_c2, bool publish, int t0) {
  const int K = c.bn_K[id];
  const int t = (int)threadIdx.x - t0;
  if (t >= 0 && t < K) {
    const double inv = count > 0 ? 1.0 / count : 0.0;
    const double a = tot[t], b = tot[K + t];
    const float c1 = (float)(a * inv), c2 = (float)(b * inv);
    s_c1[t] = c1;
    s_c2[t] = c2;
    if (publish) {
      c.bnf(id, BN_C1)[t] = c1;
      c.bnf(id, BN_C2)[t] = c2;
      c.grads[c.bn_gamma[id] + t] = (float)b;
      c.grads[c.bn_beta[id] + t] = (float)a;
    }
  }
}

// G / n_prod: grid size and vector length of the producer; the sums of BatchNorm `id` start at voff
// ([sum (K) | sum of squares (K)]).  scratch: >= 2K doubles of shared memory; s_sc / s_sh: K floats.
// All threads call; K <= blockDim.x.  `pre`: THIS thread's prefetch, i.e. of channel threadIdx.x.
__device__ __forceinline__ void bn_from_groups(const Ctx& c, int site, int id, int G, int n_prod, int voff, int count,
                                               const BnPre& pre, double* scratch, float* s_sc, float* s_sh,
                                               bool publish) {
  const int K = c.bn_K[id];
  const int ngrp = (G + kGsGroup - 1) / kGsGroup;
  const double* l1 = gs_l1(c, site) + voff;
  for (int v = threadIdx.x; v < 2 * K; v += blockDim.x) scratch[v] = gs_group_total(l1 + v, ngrp, n_prod);
  __syncthreads();
  bn_fwd_finalize_smem(c, id, count, pre, scratch, s_sc, s_sh, publish, 0);
  __syncthreads();
}
// Two BatchNorms whose sums sit in the same producer (bnc | bno): one pass, one pair of barriers.
// scratch: >= 4K doubles; threads [0, K) finalise A, [K, 2K) finalise B (2K <= blockDim.x);
// `preA` must be the prefetch of channel threadIdx.x of A, `preB` of channel threadIdx.x - K of B.
__device__ __forceinline__ void bn_from_groups2(const Ctx& c, int siteA, int idA, int voffA, int siteB, int idB,
                                                int voffB, int G, int n_prod, int count, const BnPre& preA,
                                                const BnPre& preB, double* scratch, float* s_scA, float* s_shA,
                                                float* s_scB, float* s_shB, bool publish) {
  const int K = c.bn_K[idA];
  gs_sum_pair(c, siteA, voffA, siteB, voffB, G, n_prod, 2 * K, scratch, scratch + 2 * K);
  __syncthreads();
  bn_fwd_finalize_smem(c, idA, count, preA, scratch, s_scA, s_shA, publish, 0);
  bn_fwd_finalize_smem(c, idB, count, preB, scratch + 2 * K, s_scB, s_shB, publish, K);
  __syncthreads();
}

// The same hand-over for the BatchNorm BACKWARD sums (sum dy | sum dy * xhat): c1 = mean(dy),
// c2 = mean(dy * xhat) into shared memory for this CTA; CTA `publish` writes d gamma / d beta and the record.
__device__ __forceinline__ void bn_bwd_from_groups(const Ctx& c, int site, int id, int G, int n_prod, int voff,
                                                   int count, double* scratch, float* s_c1, float* s_c2,
                                                   bool publish) {
  const int K = c.bn_K[id];
  const int ngrp = (G + kGsGroup - 1) / kGsGroup;
  const double* l1 = gs_l1(c, site) + voff;
  for (int v = threadIdx.x; v < 2 * K; v += blockDim.x) scratch[v] = gs_group_total(l1 + v, ngrp, n_prod);
  __syncthreads();
  bn_bwd_finalize_smem(c, id, count, scratch, s_c1, s_c2, publish, 0);
  __syncthreads();
}
__device__ __forceinline__ void bn_bwd_from_groups2(const Ctx& c, int siteA, int idA, int siteB, int idB, int G,
                                                    int n_prod, int count, double* scratch, float* s_c1A, float* s_c2A,
                                                    float* s_c1B, float* s_c2B, bool publish) {
  const int K = c.bn_K[idA];
  gs_sum_pair(c, siteA, 0, siteB, 0, G, n_prod, 2 * K, scratch, scratch + 2 * K);
  __syncthreads();
  bn_bwd_finalize_smem(c, idA, count, scratch, s_c1A, s_c2A, publish, 0);
  bn_bwd_finalize_smem(c, idB, count, scratch + 2 * K, s_c1B, s_c2B, publish, K);
  __syncthreads();
}

// true: this model's forward chain hands BatchNorm statistics over consumer-side (the GATConv
// kernels still read the finalised record)
__device__ __forceinline__ bool bn_consumer_side(const Ctx& c) { return c.train && c.model != CAL_MODEL_GAT; }

// Cross-warp reduction of per-lane fp64 accumulators (lane owns VEC channels, 8 warps) into the
// CTA totals sTot[(v0 + v) * K + koff + k].  sbuf holds kRowWarps * H doubles.  Fixed order.
template <int VEC, int NV>
__device__ __forceinline__ void block_totals(double (&acc)[NV][VEC], double* sbuf, double* sTot, int H, int v0,
                                             int K, int koff) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VEC; ++i) sbuf[warp * H + lane * VEC + i] = acc[v][i];
    __syncthreads();
    for (int k = threadIdx.x; k < H; k += blockDim.x) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) s += sbuf[w * H + k];
      sTot[(size_t)(v0 + v) * K + koff + k] = s;
    }
  }
  __syncthreads();
}

// Training-mode BatchNorm statistics from grid totals (sum[k], sumsq[k] in shared memory): the
// affine the next kernel applies, the saved mean / rstd, and the running-statistics update
// (biased variance to normalise, unbiased for running_var, momentum; torch BatchNorm1d).
__device__ __forceinline__ void bn_finalize_tot(const Ctx& c, int id, const double* sum, const double* sumsq,
                                                int count, const BnPre* pre = nullptr) {
  const int K = c.bn_K[id];
  const bool use_pre = pre != nullptr && K <= (int)blockDim.x;      // then k == threadIdx.x below
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const double s = sum[k], q = sumsq[k];
    double mean = 0.0, var = 0.0;
    if (count > 0) {
      mean = s / count;
      var = q / count - mean * mean;
      if (var < 0.0) var = 0.0;
    }
    float rstd = (float)(1.0 / sqrt(var + (double)c.eps));
    float g = use_pre ? pre->g : c.params[c.bn_gamma[id] + k], b = use_pre ? pre->b : c.params[c.bn_beta[id] + k];
    float sc = g * rstd;
    c.bnf(id, BN_SCALE)[k] = sc;
    c.bnf(id, BN_SHIFT)[k] = b - (float)mean * sc;
    c.bnf(id, BN_MEAN)[k] = (float)mean;
    c.bnf(id, BN_RSTD)[k] = rstd;
    if (c.bn_buffers != nullptr && c.bn_rm[id] >= 0) {
      double unb = count > 1 ? var * ((double)count / (double)(count - 1)) : var;
      float* rm = c.bn_buffers + c.bn_rm[id];
      float* rv = c.bn_buffers + c.bn_rv[id];
      rm[k] = (1.f - c.momentum) * (use_pre ? pre->rm : rm[k]) + c.momentum * (float)mean;
      rv[k] = (1.f - c.momentum) * (use_pre ? pre->rv : rv[k]) + c.momentum * (float)unb;
    }
  }
  if (threadIdx.x == 0 && c.nbt != nullptr) c.nbt[id] += 1;
}

// BatchNorm backward from grid totals: c1 = mean(dy), c2 = mean(dy * xhat); d gamma = sum dy*xhat,
// d beta = sum dy.
__device__ __forceinline__ void bn_bwd_finalize_tot(const Ctx& c, int id, const double* sum_dy,
                                                    const double* sum_dyx, int count) {
  const int K = c.bn_K[id];
  const double inv = count > 0 ? 1.0 / count : 0.0;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const double s1 = sum_dy[k], s2 = sum_dyx[k];
    c.bnf(id, BN_C1)[k] = (float)(s1 * inv);
    c.bnf(id, BN_C2)[k] = (float)(s2 * inv);
    c.grads[c.bn_gamma[id] + k] = (float)s2;
    c.grads[c.bn_beta[id] + k] = (float)s1;
  }
}

// Per-lane BatchNorm constants of the lane's VEC channels.
template <int VEC>
struct BnLane {
  float sc[VEC], sh[VEC], mean[VEC], rstd[VEC], c1[VEC], c2[VEC];
  __device__ __forceinline__ void load_fwd(const Ctx& c, int id, int lane, int koff = 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      sc[i] = c.bnf(id, BN_SCALE)[koff + lane * VEC + i];
      sh[i] = c.bnf(id, BN_SHIFT)[koff + lane * VEC + i];
    }
  }
  __device__ __forceinline__ void load_bwd(const Ctx& c, int id, int lane, int koff = 0) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      int k = koff + lane * VEC + i;
      sc[i] = c.bnf(id, BN_SCALE)[k];
      sh[i] = c.bnf(id, BN_SHIFT)[k];
      mean[i] = c.bnf(id, BN_MEAN)[k];
      rstd[i] = c.bnf(id, BN_RSTD)[k];
      c1[i] = c.bnf(id, BN_C1)[k];
      c2[i] = c.bnf(id, BN_C2)[k];
    }
  }
  // gradient w.r.t. the BatchNorm input given dy (gradient w.r.t. its output) and the input x
  __device__ __forceinline__ float dx(int i, float dy, float x) const {
    float xh = (x - mean[i]) * rstd[i];
    return sc[i] * (dy - c1[i] - xh * c2[i]);
  }
  __device__ __forceinline__ float xhat(int i, float x) const { return (x - mean[i]) * rstd[i]; }
};

// Epilogue shared by the backbone layer kernels (GCNConv and GATConv): given a finished output row
// o = relu(conv(x) + b) it accumulates the statistics of the BatchNorm that consumes the row
// (fp64 sum / sum of squares); for the LAST backbone layer (LASTL) it instead evaluates the
// node attention softmax(node_att_mlp(o)) (model.py:109), the per-node halves p, q of
// edge_att_mlp([x_row || x_col]) (model.py:97-102), and the statistics of bnc / bno on att * o.
template <int VEC, bool LASTL>
struct LayerEpilogue {
  static constexpr int NV = LASTL ? 4 : 2;
  double st[NV][VEC];
  float wn0[VEC], wn1[VEC], wp0[VEC], wp1[VEC], wq0[VEC], wq1[VEC];
  float bn0, bn1;
  __device__ __forceinline__ void init(const Ctx& c, int lane) {
    constexpr int H = 32 * VEC;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < VEC; ++i) st[v][i] = 0.0;
    bn0 = bn1 = 0.f;
    if (LASTL) {
      const float* Wn = c.params + c.po.node_att_w;    // [2][H]
      const float* We = c.params + c.po.edge_att_w;    // [2][2H]
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
      bn0 = c.params[c.po.node_att_b];
      bn1 = c.params[c.po.node_att_b + 1];
    }
  }
  // all 32 lanes of the warp call this for node i (warp-uniform)
  __device__ __forceinline__ void row(const Ctx& c, int i, const float (&o)[VEC], int lane) {
    if (!LASTL) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        st[0][k] += (double)o[k];
        st[1][k] += (double)o[k] * (double)o[k];
      }
      return;
    }
    float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      s0 = fmaf(o[k], wn0[k], s0);
      s1 = fmaf(o[k], wn1[k], s1);
      p0 = fmaf(o[k], wp0[k], p0);
      p1 = fmaf(o[k], wp1[k], p1);
      q0 = fmaf(o[k], wq0[k], q0);
      q1 = fmaf(o[k], wq1[k], q1);
    }
    s0 = warp_sum(s0) + bn0;
    s1 = warp_sum(s1) + bn1;
    p0 = warp_sum(p0);
    p1 = warp_sum(p1);
    q0 = warp_sum(q0);
    q1 = warp_sum(q1);
    float a0 = 0.5f, a1 = 0.5f;
    if (!c.no_natt) {                             // softmax over the two logits (model.py:109)
      float m = fmaxf(s0, s1);
      float e0 = expf(s0 - m), e1 = expf(s1 - m);
      float inv = 1.0f / (e0 + e1);
      a0 = e0 * inv;
      a1 = e1 * inv;
    }
    if (lane == 0) {
      *reinterpret_cast<float2*>(c.natt + (size_t)i * 2) = make_float2(a0, a1);
      *reinterpret_cast<float4*>(c.pq + (size_t)i * 4) = make_float4(p0, p1, q0, q1);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      double vc = (double)(a0 * o[k]), vo = (double)(a1 * o[k]);
      st[0][k] += vc;
      st[1][k] += vc * vc;
      st[2][k] += vo;
      st[3][k] += vo * vo;
    }
  }
  // CTA totals -> hierarchical grid sum -> BatchNorm finalisation by the one CTA that ends up
  // with the grid totals; all threads of the CTA call this
  __device__ __forceinline__ void finish(const Ctx& c, int layer, double* sRed, double* sTot, int N) {
    constexpr int H = 32 * VEC;
    if (!c.train) return;
    if (bn_consumer_side(c)) {                         // the next kernel sums the group vectors and finalises
      block_totals<VEC, NV>(st, sRed, sTot, H, 0, H, 0);
      grid_sum_groups(c, bn_site(LASTL ? c.L + 1 : 2 + layer), sTot, NV * H, gridDim.x, blockIdx.x);
      return;
    }
    const BnPre p0 = bn_prefetch(c, LASTL ? c.L + 1 : 2 + layer);
    const BnPre p1 = LASTL ? bn_prefetch(c, c.L + 2) : p0;
    block_totals<VEC, NV>(st, sRed, sTot, H, 0, H, 0);
    if (grid_sum(c, 0, sTot, NV * H, gridDim.x, blockIdx.x)) {
      if (!LASTL) {
        bn_finalize_tot(c, 2 + layer, sTot, sTot + H, N, &p0);
      } else {
        bn_finalize_tot(c, c.L + 1, sTot, sTot + H, N, &p0);
        bn_finalize_tot(c, c.L + 2, sTot + 2 * H, sTot + 3 * H, N, &p1);
      }
    }
  }
};

// Optional per-phase cycle counters (CTA 0, thread 0) for latency analysis: -DCAL_PHASE_TIMING.
#ifdef CAL_PHASE_TIMING
#define PT_DECL long long pt_t0 = clock64(); int pt_i = 0; long long pt_v[16];
#define PT_MARK() do { long long t_ = clock64(); if (pt_i < 16) pt_v[pt_i++] = t_ - pt_t0; pt_t0 = t_; } while (0)
#define PT_DUMP(c, base) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) for (int q_ = 0; q_ < pt_i; ++q_) (c).status[(base) + q_] = (int)pt_v[q_]; } while (0)
#else
#define PT_DECL
#define PT_MARK()
#define PT_DUMP(c, base)
#endif

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Stage an [rows x cols] fp32 matrix (cols % 4 == 0, 16-byte aligned) into shared memory.
__device__ __forceinline__ void stage_matrix_async(float* sdst, const float* gsrc, int n_floats) {
  for (int i = threadIdx.x * 4; i < n_floats; i += blockDim.x * 4) cp_async16(sdst + i, gsrc + i);
}

// acc[r][c] += sum_k sA[(warp*RPW + r) * lda + k] * sW[k * ldw + lane*VEC + c]   (K % 4 == 0)
// The operands of the next 4-k step are loaded into registers while the current step's FMAs
// issue (the CTA runs 2 warps per scheduler, too few to hide shared-memory latency otherwise).
// Packed fp32 pairs (sm_100a FFMA2, PTX fma.rn.f32x2): two IEEE round-to-nearest FMAs per instruction, so
// results are bit-identical to scalar fmaf while the FMA-bound inner loops need half the issue
// slots.  ptxas folds the duplicated scalar into the instruction's broadcast operand (R.F32).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// c + (a, a) * b, element-wise
__device__ __forceinline__ f32x2 ffma2_bcast(float a, f32x2 b, f32x2 c) {
  f32x2 r;
  const f32x2 aa = pack2(a, a);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(aa), "l"(b), "l"(c));
  return r;
}

template <int VEC>
__device__ __forceinline__ void load_w4(const float* __restrict__ wp, int ldw, float (&w)[4][VEC]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const float* q = wp + (size_t)kk * ldw;
    if constexpr (VEC == 4) {
      float4 t = *reinterpret_cast<const float4*>(q);
      w[kk][0] = t.x; w[kk][1] = t.y; w[kk][2] = t.z; w[kk][3] = t.w;
    } else if constexpr (VEC == 2) {
      float2 t = *reinterpret_cast<const float2*>(q);
      w[kk][0] = t.x; w[kk][1] = t.y;
    } else {
      w[kk][0] = *q;
    }
  }
}

template <int VEC, int RPW>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ sA, int lda, const float* __restrict__ sW,
                                          int ldw, int K, float (&acc)[RPW][VEC]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* a0 = sA + (size_t)warp * RPW * lda;
  const float* w0 = sW + lane * VEC;
  float4 a[RPW], an[RPW];
  float w[4][VEC], wn[4][VEC];
  f32x2 acc2[RPW][VEC / 2 > 0 ? VEC / 2 : 1];           // the accumulators as FFMA2 pairs (VEC even)
  if constexpr (VEC % 2 == 0) {
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int cc = 0; cc < VEC / 2; ++cc) acc2[r][cc] = pack2(acc[r][2 * cc], acc[r][2 * cc + 1]);
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) a[r] = *reinterpret_cast<const float4*>(a0 + r * lda);
  load_w4<VEC>(w0, ldw, w);
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 4) {
    const int kn = k0 + 4 < K ? k0 + 4 : k0;           // last step reloads itself (harmless)
#pragma unroll
    for (int r = 0; r < RPW; ++r) an[r] = *reinterpret_cast<const float4*>(a0 + r * lda + kn);
    load_w4<VEC>(w0 + (size_t)kn * ldw, ldw, wn);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
        if constexpr (VEC % 2 == 0) {
#pragma unroll
          for (int cc = 0; cc < VEC / 2; ++cc)
            acc2[r][cc] = ffma2_bcast(av, pack2(w[kk][2 * cc], w[kk][2 * cc + 1]), acc2[r][cc]);
        } else {
#pragma unroll
          for (int cc = 0; cc < VEC; ++cc) acc[r][cc] = fmaf(av, w[kk][cc], acc[r][cc]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) a[r] = an[r];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int cc = 0; cc < VEC; ++cc) w[kk][cc] = wn[kk][cc];
  }
  if constexpr (VEC % 2 == 0) {
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
      for (int cc = 0; cc < VEC / 2; ++cc) unpack2(acc2[r][cc], acc[r][2 * cc], acc[r][2 * cc + 1]);
  }
}

// Row padding of the shared-memory activation tiles (floats): rows of consecutive tile rows start 4
// banks apart.  (A column-slab mapping of this product -- warp w owns columns [16w, 16w+16) of all 24
// rows, 7 instead of 19 shared-memory wavefronts per warp per 4-k step -- was measured on B200 and is
// NOT faster: the loop is bound by FFMA issue with 2 warps per scheduler, not by LDS; profiles/README.md.)
constexpr int kPad = 4;

// The same product, K-split (H = 128 only).  tile_gemm is bound by the shared-memory RETURN path: every
// LDS.128 delivers 512 bytes per warp to the register file at 128 B/clk whatever it broadcasts, and a
// 3 x 4 register tile needs 7 of them per 24 FFMA2 (measured: 24x128x128 MACs in ~6.1 k cycles with
// every operand mapping, FFMA or FFMA2, 6 or 12 chains -- profiles/README.md).  Here the 8 warps form
// 4 K-groups of 64 threads; inside a group thread (rb, cb) owns a 6 x 8 register tile (rows 6rb..6rb+5,
// columns 4cb..4cb+3 and 64+4cb..64+4cb+3) over the group's 32 k: 14 LDS.128 per 96 FFMA2, half the
// bytes per FMA, 24 independent accumulator pairs.  The 4 partial tiles meet in `sPart`
// ([4][kTileRows][ldp], ldp >= H + kPad; must not alias sA / sW) and are added in group order while the
// tile is handed back in the row-per-warp layout of the epilogues.  All threads call; one
// __syncthreads inside.  (Sums of 4 x 32 consecutive k: deterministic, not bit-identical to tile_gemm.)
__device__ __forceinline__ void tile_gemm_ksplit128(const float* __restrict__ sA, int lda, const float* __restrict__ sW,
                                                    int ldw, float* __restrict__ sPart, int ldp,
                                                    float (&acc)[kRPW][4]) {
  static_assert(kTileRows == 24 && kRowWarps == 8, "4 K-groups x (4 row blocks x 16 column blocks)");
  constexpr int H = 128, KG = H / 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kg = warp >> 1;
  const int t64 = ((warp & 1) << 5) | lane;
  const int rb = t64 >> 4, cb = t64 & 15;
  const float* a0 = sA + (size_t)(6 * rb) * lda + kg * KG;
  const float* w0 = sW + (size_t)(kg * KG) * ldw + 4 * cb;
  f32x2 c2[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c2[i][j] = pack2(0.f, 0.f);
#pragma unroll
  for (int s = 0; s < KG / 4; ++s) {
    float4 a[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (size_t)i * lda + 4 * s);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float* wr = w0 + (size_t)(4 * s + kk) * ldw;
      const float4 wa = *reinterpret_cast<const float4*>(wr);
      const float4 wb = *reinterpret_cast<const float4*>(wr + 64);
      const f32x2 p0 = pack2(wa.x, wa.y), p1 = pack2(wa.z, wa.w), p2 = pack2(wb.x, wb.y), p3 = pack2(wb.z, wb.w);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        c2[i][0] = ffma2_bcast(av, p0, c2[i][0]);
        c2[i][1] = ffma2_bcast(av, p1, c2[i][1]);
        c2[i][2] = ffma2_bcast(av, p2, c2[i][2]);
        c2[i][3] = ffma2_bcast(av, p3, c2[i][3]);
      }
    }
  }
  float* pt = sPart + (size_t)kg * kTileRows * ldp;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack2(c2[i][j], v[2 * j], v[2 * j + 1]);
    float* o = pt + (size_t)(6 * rb + i) * ldp + 4 * cb;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 64) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRPW; ++r) {
    const float* o = sPart + (size_t)(warp * kRPW + r) * ldp + lane * 4;
    float4 t = *reinterpret_cast<const float4*>(o);
#pragma unroll
    for (int g = 1; g < 4; ++g) {
      const float4 u = *reinterpret_cast<const float4*>(o + (size_t)g * kTileRows * ldp);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    acc[r][0] = t.x; acc[r][1] = t.y; acc[r][2] = t.z; acc[r][3] = t.w;
  }
}

// H = 128: K-split; otherwise the row mapping.  sPart: [4][kTileRows][H + kPad] floats (H = 128 only).
template <int VEC, int RPW>
__device__ __forceinline__ void tile_gemm_fast(const float* __restrict__ sA, int lda, const float* __restrict__ sW,
                                               int ldw, int K, float* __restrict__ sPart, float (&acc)[RPW][VEC]) {
  if constexpr (VEC == 4 && RPW == kRPW) {
    tile_gemm_ksplit128(sA, lda, sW, ldw, sPart, 32 * VEC + kPad, acc);
  } else {
    tile_gemm<VEC, RPW>(sA, lda, sW, ldw, K, acc);
  }
}

// Outer-product accumulation over a row tile: acc[a][b] += sum_r sP[r*ld + kidx(a)] * sQ[r*ld + jidx(b)].
// 256 threads cover an [H x H] result: thread (ty = tid / 16, tx = tid % 16) owns MT x MT entries with
// MT = H / 16, index set {half*(H/2) + t*(MT/2) + i}.
template <int H>
struct OuterAcc {
  static constexpr int MT = H / 16;
  static constexpr int HM = MT / 2 > 0 ? MT / 2 : 1;   // contiguous run
  static constexpr int NH = MT / HM;                   // number of runs (2, or 2 when MT == 2 -> HM = 1)
  static constexpr bool kPairs = MT % 2 == 0;          // accumulators live as FFMA2 pairs over the column index
  f32x2 acc2[MT][kPairs ? MT / 2 : 1];
  float acc1[kPairs ? 1 : MT][kPairs ? 1 : MT];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int a = 0; a < MT; ++a) {
      if constexpr (kPairs) {
#pragma unroll
        for (int b = 0; b < MT / 2; ++b) acc2[a][b] = pack2(0.f, 0.f);
      } else {
#pragma unroll
        for (int b = 0; b < MT; ++b) acc1[a][b] = 0.f;
      }
    }
  }
  __device__ __forceinline__ float get(int a, int b) const {
    if constexpr (kPairs) {
      float lo, hi;
      unpack2(acc2[a][b >> 1], lo, hi);
      return (b & 1) ? hi : lo;
    } else {
      return acc1[a][b];
    }
  }
  __device__ __forceinline__ static int idx(int t, int a) { return (a / HM) * (H / NH) + t * HM + (a % HM); }
  // contiguous run of HM floats starting at element idx(t, run * HM); 16-byte aligned when HM == 4
  __device__ __forceinline__ static void load_runs(const float* __restrict__ row, int t, float (&v)[MT]) {
#pragma unroll
    for (int run = 0; run < NH; ++run) {
      const float* p = row + idx(t, run * HM);
      if constexpr (HM == 4) {
        float4 q = *reinterpret_cast<const float4*>(p);
        v[run * HM] = q.x; v[run * HM + 1] = q.y; v[run * HM + 2] = q.z; v[run * HM + 3] = q.w;
      } else if constexpr (HM == 2) {
        float2 q = *reinterpret_cast<const float2*>(p);
        v[run * HM] = q.x; v[run * HM + 1] = q.y;
      } else {
        v[run * HM] = *p;
      }
    }
  }
  __device__ __forceinline__ void accumulate(const float* __restrict__ sP, int ldp, const float* __restrict__ sQ,
                                             int ldq, int rows) {
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float p[MT], q[MT], pn[MT], qn[MT];
    load_runs(sP, ty, p);
    load_runs(sQ, tx, q);
    for (int r = 0; r < rows; ++r) {
      const int rn = r + 1 < rows ? r + 1 : r;
      load_runs(sP + (size_t)rn * ldp, ty, pn);
      load_runs(sQ + (size_t)rn * ldq, tx, qn);
      if constexpr (kPairs) {
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
          for (int b = 0; b < MT / 2; ++b) acc2[a][b] = ffma2_bcast(p[a], pack2(q[2 * b], q[2 * b + 1]), acc2[a][b]);
      } else {
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
          for (int b = 0; b < MT; ++b) acc1[a][b] = fmaf(p[a], q[b], acc1[a][b]);
      }
#pragma unroll
      for (int a = 0; a < MT; ++a) {
        p[a] = pn[a];
        q[a] = qn[a];
      }
    }
  }
  __device__ __forceinline__ void store(float* __restrict__ dst, int ldd) const {   // dst [H][ldd], 16-byte aligned rows
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
      for (int run = 0; run < NH; ++run) {
        float* p = dst + (size_t)idx(ty, a) * ldd + idx(tx, run * HM);
        if constexpr (HM == 4) {
          *reinterpret_cast<float4*>(p) = make_float4(get(a, run * HM), get(a, run * HM + 1), get(a, run * HM + 2),
                                                      get(a, run * HM + 3));
        } else if constexpr (HM == 2) {
          *reinterpret_cast<float2*>(p) = make_float2(get(a, run * HM), get(a, run * HM + 1));
        } else {
          *p = get(a, run * HM);
        }
      }
  }
};

// Block-reduce per-thread float column accumulators (lane owns VEC channels, 8 warps) in a fixed
// order and write H sums to dst.  sbuf holds kRowWarps * H floats.
template <int VEC>
__device__ __forceinline__ void block_colsum_store(const float (&acc)[VEC], float* sbuf, float* dst, int H) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < VEC; ++i) sbuf[warp * H + lane * VEC + i] = acc[i];
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) s += sbuf[w * H + k];
    if (dst != nullptr) dst[k] = s;
  }
}



}  // namespace cal
