// prep.cu -- per-batch structure preparation and per-step parameter preparation.
//
// Replaces the integer work the reference repeats inside every GCNConv.norm call
// (gcn_conv.py:44-70: remove_self_loops, add_self_loops, degree by ROW, deg^-1/2, norm) and the
// implicit segmentation of global_add_pool (model.py:115-116).  Done once per batch:
//   int64 -> int32, self loops dropped, one loop appended per node (LAST in every row, as PyG's
//   cat puts them last), CSR by target ("in") and by source ("out"), both ordered by edge_index
//   column so that sequential accumulation follows the CPU scatter_add order; graph_ptr; perm.
#include "internal.cuh"

namespace cal {

// Node pass + edge count + input-feature statistics in one launch:
//   graph ids / graph_ptr / perm, status word, identity BatchNorm record;
//   in- / out-degree counts of the kept edges (the count arrays are zero on entry: the workspace
//   is zero-filled once by the caller and k_prep_link re-zeroes them after use);
//   column sums / sums of squares of the input features (bn_feat, model.py:90) -> hierarchical grid
//   sum -> totals in the workspace (statp[0 .. 2F)), turned into the BatchNorm affine by k_feat_fwd.
__global__ void __launch_bounds__(256) k_prep_init(const Ctx c) {
  pdl_sync();
  __shared__ double s_s[256], s_q[256];
  __shared__ double sTot[2 * 512];               // F <= 512 (validate_model)
  const int N = c.dims[0], E = c.dims[1], B = c.dims[2];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const bool bad_caps = N < 0 || E < 0 || B < 0 || N > c.Nm || E > c.Em || B > c.Bm;
  if (tid == 0) {
    c.status[0] = bad_caps ? kStCapacity : 0;
    c.status[1] = c.status[2] = c.status[3] = 0;
    if (bad_caps) atomicOr(&c.status[kStSticky], kStCapacity);
  }
  if (tid < 64 + kGsSites * kGsCounters) c.counters[tid] = 0u;
  if (bad_caps) return;
  for (int k = tid; k < c.kmax; k += nth) {     // identity BatchNorm record
    c.bnf(kBnIdentity, BN_SCALE)[k] = 1.f;
    c.bnf(kBnIdentity, BN_SHIFT)[k] = 0.f;
    c.bnf(kBnIdentity, BN_MEAN)[k] = 0.f;
    c.bnf(kBnIdentity, BN_RSTD)[k] = 0.f;
    c.bnf(kBnIdentity, BN_C1)[k] = 0.f;
    c.bnf(kBnIdentity, BN_C2)[k] = 0.f;
  }
  for (int n = tid; n < N; n += nth) {
    long long g = c.batch[n];
    long long gp = n > 0 ? c.batch[n - 1] : -1;
    if (g < 0 || g >= B || g < gp) {
      raise_status(c.status, kStBadBatch);
      g = g < 0 ? 0 : (g >= B ? B - 1 : g);
    }
    c.node_graph[n] = (int)g;
    if (gp < -1) gp = -1;
    if (gp >= B) gp = B - 1;
    for (long long b = gp + 1; b <= g; ++b) c.graph_ptr[b] = n;     // graphs that start at n (empty ones too)
    if (n == N - 1)
      for (long long b = g + 1; b <= B; ++b) c.graph_ptr[b] = N;
  }
  if (N == 0)
    for (int b = tid; b <= B; b += nth) c.graph_ptr[b] = 0;
  for (int b = tid; b < B; b += nth) {
    int p = c.perm_in != nullptr ? c.perm_in[b] : b;
    if (p < 0 || p >= B) {
      raise_status(c.status, kStBadBatch);
      p = b;
    }
    c.perm[b] = p;
    c.invperm[p] = b;
  }
  if (c.grouped) {
    // ---- cal_caps.grouped_edges: the first edge_index column of every graph (k_prep_graph builds the CSRs graph by
    // graph) and the number of self loops per graph (they are dropped, so they shift the rows of all later graphs) ----
    int* egp = c.egp;
    int* nself = c.egp + c.Bm + 1;
    if (E == 0)
      for (int b = tid; b <= B; b += nth) egp[b] = 0;
    for (int e = tid; e < E; e += nth) {
      const long long r = c.ei_row[e], rp = e > 0 ? c.ei_row[e - 1] : -1;
      const bool ok = r >= 0 && r < N;
      long long g = ok ? c.batch[r] : 0, gp = (rp >= 0 && rp < N) ? c.batch[rp] : (e > 0 ? 0 : -1);
      if (!ok) raise_status(c.status, kStBadNode);
      g = g < 0 ? 0 : (g >= B ? B - 1 : g);                 // (an invalid `batch` is reported by the node pass above)
      gp = gp < -1 ? -1 : (gp >= B ? B - 1 : gp);
      if (g < gp) raise_status(c.status, kStBadBatch);       // the promise is broken: columns not grouped by graph
      for (long long b = gp + 1; b <= g; ++b) egp[b] = e;    // graphs whose columns start at e (edgeless ones too)
      if (e == E - 1)
        for (long long b = g + 1; b <= B; ++b) egp[b] = E;
      if (ok && r == c.ei_col[e]) atomicAdd(&nself[(int)g], 1);
    }
  } else {
    // ---- degree counts (gcn_conv.py:56: self loops are dropped) ----
    for (int e = tid; e < E; e += nth) {
      long long r = c.ei_row[e], d = c.ei_col[e];
      if (r < 0 || r >= N || d < 0 || d >= N) {
        raise_status(c.status, kStBadNode);
        continue;
      }
      if (r != d) {
        atomicAdd(&c.cnt_in[(int)d], 1);
        atomicAdd(&c.cnt_out[(int)r], 1);
      }
    }
  }
  // ---- input-feature column statistics (fp64), this CTA's slice of the rows ----
  {
    const int F = c.F, t = threadIdx.x;
    const int Fw = imin(F, 256), rpar = 256 / Fw;
    const int rows_per = ceil_div(imax(N, 1), gridDim.x);
    const int r0 = imin(blockIdx.x * rows_per, N), r1 = imin(r0 + rows_per, N);
    for (int cb = 0; cb < F; cb += Fw) {
      const int col = cb + t % Fw, rs = t / Fw;
      double s = 0.0, q = 0.0;
      if (rs < rpar && col < F)
        for (int r = r0 + rs; r < r1; r += rpar) {
          double v = (double)c.feat[(size_t)r * F + col];
          s += v;
          q += v * v;
        }
      __syncthreads();
      s_s[t] = s;
      s_q[t] = q;
      __syncthreads();
      if (t < Fw && cb + t < F) {
        double a = 0.0, b = 0.0;
        for (int k = 0; k < rpar; ++k) {
          a += s_s[k * Fw + t];
          b += s_q[k * Fw + t];
        }
        sTot[cb + t] = a;
        sTot[F + cb + t] = b;
      }
    }
    __syncthreads();
    if (grid_sum(c, 0, sTot, 2 * F, gridDim.x, blockIdx.x))
      for (int k = t; k < 2 * F; k += blockDim.x) c.statp[k] = sTot[k];
  }
}

// Exclusive scan of (count + 1) over the nodes (the +1 is the appended self loop), one CTA.
__global__ void __launch_bounds__(1024) k_prep_scan(const Ctx c) {
  pdl_sync();
  if (c.status[0] & kStCapacity) return;
  const int N = c.dims[0];
  __shared__ int s_part[2][1024];
  const int t = threadIdx.x, T = blockDim.x;
  const int per = (N + T - 1) / T;
  const int lo = imin(t * per, N), hi = imin(lo + per, N);
  int si = 0, so = 0;
  for (int n = lo; n < hi; ++n) {
    si += c.cnt_in[n] + 1;
    so += c.cnt_out[n] + 1;
  }
  s_part[0][t] = si;
  s_part[1][t] = so;
  __syncthreads();
  // Hillis-Steele inclusive scan over the T partials
  for (int d = 1; d < T; d <<= 1) {
    int a = 0, b = 0;
    if (t >= d) {
      a = s_part[0][t - d];
      b = s_part[1][t - d];
    }
    __syncthreads();
    s_part[0][t] += a;
    s_part[1][t] += b;
    __syncthreads();
  }
  int bi = s_part[0][t] - si, bo = s_part[1][t] - so;
  for (int n = lo; n < hi; ++n) {
    c.in_ptr[n] = bi;
    c.out_ptr[n] = bo;
    bi += c.cnt_in[n] + 1;
    bo += c.cnt_out[n] + 1;
    c.cnt_in[n] = 0;     // becomes the fill cursor
    c.cnt_out[n] = 0;
  }
  if (t == T - 1) {
    c.in_ptr[N] = s_part[0][T - 1];
    c.out_ptr[N] = s_part[1][T - 1];
  }
}

__global__ void k_prep_fill(const Ctx c) {
  pdl_sync();
  if (c.status[0] & kStCapacity) return;
  const int N = c.dims[0], E = c.dims[1];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int e = tid; e < E; e += nth) {
    long long r = c.ei_row[e], d = c.ei_col[e];
    if (r < 0 || r >= N || d < 0 || d >= N || r == d) continue;
    int pi = c.in_ptr[(int)d] + atomicAdd(&c.cnt_in[(int)d], 1);
    c.in_key[pi] = e;
    int po = c.out_ptr[(int)r] + atomicAdd(&c.cnt_out[(int)r], 1);
    c.out_key[po] = e;
  }
  for (int n = tid; n < N; n += nth) {      // appended self loop: last slot of the row
    c.in_key[c.in_ptr[n + 1] - 1] = E + n;
    c.out_key[c.out_ptr[n + 1] - 1] = E + n;
  }
}

__device__ __forceinline__ void insertion_sort(int* a, int n) {
  for (int i = 1; i < n; ++i) {
    int v = a[i], j = i - 1;
    while (j >= 0 && a[j] > v) {
      a[j + 1] = a[j];
      --j;
    }
    a[j + 1] = v;
  }
}

// Order every row by edge_index column, resolve endpoints, unweighted degree (by source row,
// gcn_conv.py:66) and deg^-1/2.
__global__ void k_prep_sort(const Ctx c) {
  pdl_sync();
  if (c.status[0] & kStCapacity) return;
  const int N = c.dims[0], E = c.dims[1];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    int p0 = c.in_ptr[n], p1 = c.in_ptr[n + 1];
    insertion_sort(c.in_key + p0, p1 - p0);
    for (int p = p0; p < p1; ++p) {
      int k = c.in_key[p];
      c.in_src[p] = k < E ? (int)c.ei_row[k] : n;
    }
    int q0 = c.out_ptr[n], q1 = c.out_ptr[n + 1];
    insertion_sort(c.out_key + q0, q1 - q0);
    for (int q = q0; q < q1; ++q) {
      int k = c.out_key[q];
      c.out_dst[q] = k < E ? (int)c.ei_col[k] : n;
    }
    float deg = (float)(q1 - q0);            // sum of unit weights over edges with row == n (+ the loop)
    c.dis[n] = 1.0f / sqrtf(deg);
  }
}

// Link the two orderings (out position -> in position) and the unweighted norm.
__global__ void k_prep_link(const Ctx c) {
  pdl_sync();
  if (c.status[0] & kStCapacity) return;
  const int N = c.dims[0];
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    c.cnt_in[n] = 0;                            // ready for the next batch (k_prep_init counts into them)
    c.cnt_out[n] = 0;
    const float dn = c.dis[n];
    for (int p = c.in_ptr[n]; p < c.in_ptr[n + 1]; ++p) {
      int s = c.in_src[p], key = c.in_key[p];
      const float nrm = c.dis[s] * dn;       // dis[row] * 1 * dis[col]
      c.in_norm[p] = nrm;
      int lo = c.out_ptr[s], hi = c.out_ptr[s + 1] - 1;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (c.out_key[mid] < key) lo = mid + 1; else hi = mid;
      }
      c.out_pos[lo] = p;
      c.out_norm[lo] = nrm;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cal_caps.grouped_edges: one CTA per graph builds that graph's rows of both CSRs in shared memory (edges packed to
// 2 x u16 local ids, shared-memory atomics for the degree counts and fill cursors, a shuffle scan, rank-by-counting
// inside the short rows) and writes them at the graph's place in the global arrays: rows of graph g start at
//   (edge_index columns before g) - (self loops before g) + (nodes before g).
// Same results as the other two paths (the row order is fixed by the ranks, not by the atomics).  Replaces the five
// global passes for batches the single-kernel path cannot hold (cfg 5: 512 graphs x 200 nodes: 625 us -> see profiles/).
// A graph beyond kPgNodes / kPgCols or with an endpoint outside itself raises the status word and gets empty rows.
// ---------------------------------------------------------------------------------------------
constexpr int kPgT = 256;
constexpr int kPgNodes = 512;
constexpr int kPgCols = 4096;

__global__ void __launch_bounds__(kPgT) k_prep_graph(const Ctx c) {
  pdl_sync();
  __shared__ int s_cin[kPgNodes], s_cout[kPgNodes], s_iptr[kPgNodes + 1], s_optr[kPgNodes + 1];
  __shared__ float s_dis[kPgNodes];
  __shared__ unsigned int s_edge[kPgCols];                             // (row << 16) | col, local ids; 0xffffffff = dropped
  __shared__ unsigned short s_ikey[kPgCols + kPgNodes], s_okey[kPgCols + kPgNodes];
  __shared__ int s_wi[kPgT / 32], s_wo[kPgT / 32], s_red[kPgT / 32], s_bad, s_last;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, T = kPgT;
  const int N = c.dims[0], E = c.dims[1], B = c.dims[2];
  if (N < 0 || E < 0 || B < 0 || N > c.Nm || E > c.Em || B > c.Bm) return;     // (k_prep_init reported it and stopped too)
  const int* egp = c.egp;
  int* nself = c.egp + c.Bm + 1;
  int* done = c.egp + 2 * c.Bm + 1;
  if (blockIdx.x == 0 && t == 0 && B == 0) c.in_ptr[0] = c.out_ptr[0] = 0;
  for (int g = blockIdx.x; g < B; g += gridDim.x) {
    const int n0 = c.graph_ptr[g], n1 = c.graph_ptr[g + 1], e0 = egp[g], e1 = egp[g + 1];
    const int Nc = n1 - n0, Ec = e1 - e0;
    // self loops of the graphs before this one
    int sl = 0;
    for (int b = t; b < g; b += T) sl += nself[b];
    for (int o = 16; o > 0; o >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, o);
    if (t == 0) s_bad = 0;
    if (lane == 0) s_red[warp] = sl;
    __syncthreads();
    sl = 0;
    for (int w = 0; w < T / 32; ++w) sl += s_red[w];
    const int base = e0 - sl + n0;
    bool fits = Nc >= 0 && Nc <= kPgNodes && Ec >= 0 && Ec <= kPgCols && n0 >= 0 && n1 <= N && e0 >= 0 && e1 <= E;
    if (!fits) {
      if (t == 0) raise_status(c.status, kStCapacity);
      if (Nc < 0 || n0 < 0 || n1 > N) {                               // (an invalid `batch`: nothing sane to write)
        __syncthreads();
        continue;
      }
    }
    for (int i = t; i < imin(Nc, kPgNodes); i += T) s_cin[i] = s_cout[i] = 0;
    __syncthreads();
    // 0. this graph's edge_index columns -> shared memory + degree counts (gcn_conv.py:56: self loops are dropped)
    if (fits)
      for (int el = t; el < Ec; el += T) {
        const long long r = c.ei_row[e0 + el] - n0, d = c.ei_col[e0 + el] - n0;
        unsigned int pk = 0xffffffffu;
        if (r < 0 || r >= Nc || d < 0 || d >= Nc) {
          s_bad = 1;                                                  // an endpoint outside the graph
        } else if (r != d) {
          pk = ((unsigned int)r << 16) | (unsigned int)d;
          atomicAdd(&s_cin[(int)d], 1);
          atomicAdd(&s_cout[(int)r], 1);
        }
        s_edge[el] = pk;
      }
    __syncthreads();
    const bool empty_rows = !fits || s_bad != 0;                      // reported; the graph keeps only its place
    if (s_bad != 0 && t == 0) raise_status(c.status, kStBadNode);
    if (empty_rows) {
      // rows without entries at the graph's place (the arrays stay consistent: later graphs do not move)
      for (int i = t; i < Nc; i += T) {
        c.in_ptr[n0 + i] = base;
        c.out_ptr[n0 + i] = base;
        c.dis[n0 + i] = 1.f;
      }
      if (g == B - 1 && t == 0) c.in_ptr[N] = c.out_ptr[N] = base;
      __syncthreads();
      continue;
    }
    // 1. exclusive scan of (count + 1) in chunks of T nodes; in- and out-counts packed in one word (both < 2^16)
    {
      unsigned int carry = 0u;
      for (int c0 = 0; c0 < Nc; c0 += T) {
        const int i = c0 + t;
        const unsigned int v = i < Nc ? (unsigned int)(s_cin[i] + 1) | ((unsigned int)(s_cout[i] + 1) << 16) : 0u;
        unsigned int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int a = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += a;
        }
        if (lane == 31) s_wi[warp] = (int)x;
        __syncthreads();
        if (warp == 0) {
          unsigned int a = lane < T / 32 ? (unsigned int)s_wi[lane] : 0u;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned int a2 = __shfl_up_sync(0xffffffffu, a, o);
            if (lane >= o) a += a2;
          }
          if (lane < T / 32) s_wo[lane] = (int)a;
        }
        __syncthreads();
        if (i < Nc) {
          const unsigned int ex = carry + (warp > 0 ? (unsigned int)s_wo[warp - 1] : 0u) + x - v;
          s_iptr[i] = (int)(ex & 0xffffu);
          s_optr[i] = (int)(ex >> 16);
          s_dis[i] = 1.0f / sqrtf((float)(v >> 16));                  // deg^-1/2, degree by source row incl. the loop (gcn_conv.py:66)
          s_cin[i] = 0;                                               // becomes the fill cursor
          s_cout[i] = 0;
        }
        carry += (unsigned int)s_wo[T / 32 - 1];
        __syncthreads();
      }
      if (t == 0) {
        s_iptr[Nc] = (int)(carry & 0xffffu);
        s_optr[Nc] = (int)(carry >> 16);
      }
    }
    __syncthreads();
    // 2. fill in arrival order (arbitrary); key = local column index, the appended self loop (key Ec + node) takes
    //    the last slot of its row
    for (int el = t; el < Ec; el += T) {
      const unsigned int pk = s_edge[el];
      if (pk == 0xffffffffu) continue;
      const int r = (int)(pk >> 16), d = (int)(pk & 0xffffu);
      s_ikey[s_iptr[d] + atomicAdd(&s_cin[d], 1)] = (unsigned short)el;
      s_okey[s_optr[r] + atomicAdd(&s_cout[r], 1)] = (unsigned short)el;
    }
    for (int i = t; i < Nc; i += T) {
      s_ikey[s_iptr[i + 1] - 1] = (unsigned short)(Ec + i);
      s_okey[s_optr[i + 1] - 1] = (unsigned short)(Ec + i);
    }
    __syncthreads();
    // 3. write-out.  Rows are ordered by edge_index column WITHOUT sorting: the final slot of an entry is its row
    //    start + the number of smaller keys in its (short) row, counted entry-parallel.
    for (int i = t; i < Nc; i += T) {
      c.in_ptr[n0 + i] = base + s_iptr[i];
      c.out_ptr[n0 + i] = base + s_optr[i];
      c.dis[n0 + i] = s_dis[i];
    }
    if (g == B - 1 && t == 0) {
      c.in_ptr[N] = base + s_iptr[Nc];
      c.out_ptr[N] = base + s_optr[Nc];
    }
    const int tot = s_iptr[Nc];
    for (int u = t; u < tot; u += T) {
      const int key = (int)s_ikey[u];
      int n, src;
      if (key < Ec) {
        const unsigned int pk = s_edge[key];
        src = (int)(pk >> 16);
        n = (int)(pk & 0xffffu);
      } else {
        n = src = key - Ec;
      }
      const int a = s_iptr[n], b = s_iptr[n + 1];
      int rank = 0;
      for (int v = a; v < b; ++v) rank += (int)s_ikey[v] < key;
      const int p = base + a + rank;
      c.in_key[p] = key < Ec ? e0 + key : E + n0 + n;
      c.in_src[p] = n0 + src;
      c.in_norm[p] = s_dis[src] * s_dis[n];                           // dis[row] * 1 * dis[col]
    }
    for (int u = t; u < tot; u += T) {
      const int key = (int)s_okey[u];
      int n, dst;
      if (key < Ec) {
        const unsigned int pk = s_edge[key];
        n = (int)(pk >> 16);
        dst = (int)(pk & 0xffffu);
      } else {
        n = dst = key - Ec;
      }
      const int a = s_optr[n], b = s_optr[n + 1];
      int rank = 0;
      for (int v = a; v < b; ++v) rank += (int)s_okey[v] < key;
      const int a2 = s_iptr[dst], b2 = s_iptr[dst + 1];
      int rank2 = 0;
      for (int v = a2; v < b2; ++v) rank2 += (int)s_ikey[v] < key;
      const int q = base + a + rank;
      c.out_key[q] = key < Ec ? e0 + key : E + n0 + n;
      c.out_dst[q] = n0 + dst;
      c.out_pos[q] = base + a2 + rank2;                               // in-CSR position of the same edge
      c.out_norm[q] = s_dis[n] * s_dis[dst];
    }
    __syncthreads();
  }
  // the last CTA to finish clears the self-loop counters for the next batch (every CTA has read its prefixes by now)
  __syncthreads();
  if (t == 0) {
    __threadfence();
    s_last = atomicAdd(done, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    for (int b = t; b < c.Bm; b += T) nself[b] = 0;
    if (t == 0) *done = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Small batches (the whole CSR fits one SM's shared memory; cfg 1: N ~ 3 200, E' ~ 10 400): the
// five kernels above collapse into ONE launch of G + kPrepS CTAs.
//   CTAs [0, G): input-feature statistics (+ housekeeping).
//   CTAs [G, G + kPrepS), the structure CTAs: EACH builds both complete CSRs in its own shared
//   memory -- edges packed to 2 x u16, shared-memory atomics for the degree counts and fill
//   cursors, a shuffle scan, per-row insertion sort; redundant work, but it only costs shared-memory
//   cycles and needs no inter-CTA communication -- and then writes only ITS slice of the nodes
//   (and of their in- / out-rows) to global memory, so the store bandwidth of kPrepS SMs is used.
// Same results as the multi-kernel path (the row order is fixed by the sort, not by the atomics).
// smem: cnt_in [Nm] | cnt_out [Nm] | in_ptr [Nm+1] | out_ptr [Nm+1] | dis [Nm] | edges [Em] (4-byte words) |
//       in_key [EP] | out_key [EP] (u16: keys are edge_index columns or E + node, all < 65535)
// ---------------------------------------------------------------------------------------------
constexpr int kPrepT = 1024;
constexpr int kPrepS = 8;
constexpr size_t kPrepSmallMaxSmem = 225 * 1024;
constexpr size_t kPrepStatSmem = (2 * kPrepT + 2 * 512) * sizeof(double);
inline size_t prep_small_smem(int Nm, int Em) {
  const size_t b = 4 * ((size_t)5 * Nm + (size_t)(Em + Nm) + (size_t)Em + 8);      // keys are u16 (E' < 65535)
  return b > kPrepStatSmem ? b : kPrepStatSmem;
}

__global__ void __launch_bounds__(kPrepT) k_prep_small(const Ctx c) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_status, s_wi[32], s_wo[32];
  const int t = threadIdx.x, T = kPrepT;
  const int N = c.dims[0], E = c.dims[1], B = c.dims[2];
  const bool bad_caps = N < 0 || E < 0 || B < 0 || N > c.Nm || E > c.Em || B > c.Bm;
  const int G = gridDim.x - kPrepS;
  if ((int)blockIdx.x < G) {
    // ---- housekeeping + input-feature column statistics (fp64), this CTA's slice of the rows ----
    const int tid = blockIdx.x * T + t, nth = G * T;
    PT_DECL
    if (tid < 64) c.counters[tid] = 0u;
    if (bad_caps) return;
    for (int k = tid; k < c.kmax; k += nth) {     // identity BatchNorm record
      c.bnf(kBnIdentity, BN_SCALE)[k] = 1.f;
      c.bnf(kBnIdentity, BN_SHIFT)[k] = 0.f;
      c.bnf(kBnIdentity, BN_MEAN)[k] = 0.f;
      c.bnf(kBnIdentity, BN_RSTD)[k] = 0.f;
      c.bnf(kBnIdentity, BN_C1)[k] = 0.f;
      c.bnf(kBnIdentity, BN_C2)[k] = 0.f;
    }
    double* s_s = reinterpret_cast<double*>(smem_raw);     // [T]   (the statistics CTAs use the dynamic
    double* s_q = s_s + kPrepT;                            // [T]    region for their scratch)
    double* sTot = s_q + kPrepT;                           // [2 * 512]: F <= 512 (validate_model)
    const int F = c.F;
    const int Fw = imin(F, T), rpar = T / Fw;
    const int rows_per = ceil_div(imax(N, 1), G);
    const int r0 = imin(blockIdx.x * rows_per, N), r1 = imin(r0 + rows_per, N);
    for (int cb = 0; cb < F; cb += Fw) {
      const int col = cb + t % Fw, rs = t / Fw;
      double s = 0.0, q = 0.0;
      if (rs < rpar && col < F)
        for (int r = r0 + rs; r < r1; r += rpar) {
          double v = (double)c.feat[(size_t)r * F + col];
          s += v;
          q += v * v;
        }
      __syncthreads();
      s_s[t] = s;
      s_q[t] = q;
      __syncthreads();
      // fixed-order two-level sum of the rpar row parts: 8 slices, then the 8 slice sums
      const int sl = t / Fw, cc = t % Fw;
      const int per_sl = ceil_div(rpar, 8);
      double a = 0.0, b = 0.0;
      if (sl < 8)
        for (int k = sl * per_sl; k < imin(rpar, (sl + 1) * per_sl); ++k) {
          a += s_s[k * Fw + cc];
          b += s_q[k * Fw + cc];
        }
      __syncthreads();
      if (sl < 8) {
        s_s[t] = a;
        s_q[t] = b;
      }
      __syncthreads();
      if (t < Fw && cb + t < F) {
        a = b = 0.0;
        for (int k = 0; k < 8 && k * Fw + t < T; ++k) {
          a += s_s[k * Fw + t];
          b += s_q[k * Fw + t];
        }
        sTot[cb + t] = a;
        sTot[F + cb + t] = b;
      }
    }
    __syncthreads();
    PT_MARK();                                         // 0: wait + column sums
    const bool fin = grid_sum(c, 0, sTot, 2 * F, G, blockIdx.x);
    if (fin)
      for (int k = t; k < 2 * F; k += T) c.statp[k] = sTot[k];
    PT_MARK();                                         // 1: grid sum
#ifdef CAL_PHASE_TIMING
    if (t == 0 && (blockIdx.x == 0 || fin)) for (int q_ = 0; q_ < pt_i; ++q_) c.status[(fin ? 116 : 112) + q_] = (int)pt_v[q_];
#endif
    return;
  }
  // ---- a structure CTA ----
  const int sj = (int)blockIdx.x - G;            // slice index
  if (bad_caps) {
    if (sj == 0 && t == 0) {
      c.status[0] = kStCapacity;
      c.status[1] = c.status[2] = c.status[3] = 0;
      atomicOr(&c.status[kStSticky], kStCapacity);
    }
    return;
  }
  const int Nm = c.Nm, EP = c.EP;
  int* s_cin = reinterpret_cast<int*>(smem_raw);
  int* s_cout = s_cin + Nm;
  int* s_iptr = s_cout + Nm;
  int* s_optr = s_iptr + Nm + 1;
  float* s_dis = reinterpret_cast<float*>(s_optr + Nm + 1);
  unsigned int* s_edge = reinterpret_cast<unsigned int*>(s_dis + Nm);      // (row << 16) | col; 0xffffffff = dropped
  unsigned short* s_ikey = reinterpret_cast<unsigned short*>(s_edge + c.Em);
  unsigned short* s_okey = s_ikey + EP;
  const int lane = t & 31, warp = t >> 5;
  PT_DECL
  if (t == 0) s_status = 0;
  if (sj == 0) CAL_TL(c.status, 0);
  for (int n = t; n < N; n += T) s_cin[n] = s_cout[n] = 0;
  __syncthreads();
  PT_MARK();                                           // 0: dependency wait + zero
  // 0. edges -> shared memory (8 independent 8-byte load pairs in flight per thread) + degree counts
  //    (gcn_conv.py:56: self loops are dropped)
  constexpr int kEU = 8;
  for (int e0 = 0; e0 < E; e0 += kEU * T) {
    long long r[kEU], d[kEU];
#pragma unroll
    for (int u = 0; u < kEU; ++u) {
      const int e = e0 + u * T + t;
      r[u] = e < E ? c.ei_row[e] : 0;
      d[u] = e < E ? c.ei_col[e] : 0;
    }
#pragma unroll
    for (int u = 0; u < kEU; ++u) {
      const int e = e0 + u * T + t;
      if (e >= E) continue;
      unsigned int pk = 0xffffffffu;
      if (r[u] < 0 || r[u] >= N || d[u] < 0 || d[u] >= N) {
        atomicOr(&s_status, kStBadNode);
      } else if (r[u] != d[u]) {
        pk = ((unsigned int)r[u] << 16) | (unsigned int)d[u];
        atomicAdd(&s_cin[(int)d[u]], 1);
        atomicAdd(&s_cout[(int)r[u]], 1);
      }
      s_edge[e] = pk;
    }
  }
  PT_MARK();                                           // 1: edges + counts
  // graph ids / graph_ptr / perm: this CTA's slice of the nodes; slice 0 also checks all of `batch`
  const int nper = ceil_div(imax(N, 1), kPrepS);
  const int nlo = imin(sj * nper, N), nhi = imin(nlo + nper, N);
  {
    constexpr int kNU = 4;
    const int a0 = sj == 0 ? 0 : nlo, a1 = sj == 0 ? N : nhi;
    for (int b0 = a0; b0 < a1; b0 += kNU * T) {
      long long gq[kNU], gpq[kNU];
#pragma unroll
      for (int u = 0; u < kNU; ++u) {
        const int n = b0 + u * T + t;
        gq[u] = n < a1 ? c.batch[n] : 0;
        gpq[u] = (n < a1 && n > 0) ? c.batch[n - 1] : -1;
      }
#pragma unroll
      for (int u = 0; u < kNU; ++u) {
        const int n = b0 + u * T + t;
        if (n >= a1) continue;
        long long g = gq[u], gp = gpq[u];
        if (g < 0 || g >= B || g < gp) {
          atomicOr(&s_status, kStBadBatch);
          g = g < 0 ? 0 : (g >= B ? B - 1 : g);
        }
        if (n < nlo || n >= nhi) continue;
        c.node_graph[n] = (int)g;
        if (gp < -1) gp = -1;
        if (gp >= B) gp = B - 1;
        for (long long b = gp + 1; b <= g; ++b) c.graph_ptr[b] = n;     // graphs that start at n (empty ones too)
        if (n == N - 1)
          for (long long b = g + 1; b <= B; ++b) c.graph_ptr[b] = N;
      }
    }
  }
  if (sj == 0) {
    if (N == 0)
      for (int b = t; b <= B; b += T) c.graph_ptr[b] = 0;
    for (int b = t; b < B; b += T) {
      int p = c.perm_in != nullptr ? c.perm_in[b] : b;
      if (p < 0 || p >= B) {
        atomicOr(&s_status, kStBadBatch);
        p = b;
      }
      c.perm[b] = p;
      c.invperm[p] = b;
    }
  }
  __syncthreads();
  PT_MARK();                                           // 2: node pass
  // 1. exclusive scan of (count + 1) in chunks of T nodes (node = chunk * T + t: conflict-free); the
  //    in- and out-counts travel packed in one word (both totals < 2^16, checked by the launcher)
  {
    unsigned int carry = 0u;
    for (int n0 = 0; n0 < N; n0 += T) {
      const int n = n0 + t;
      const unsigned int v = n < N ? (unsigned int)(s_cin[n] + 1) | ((unsigned int)(s_cout[n] + 1) << 16) : 0u;
      unsigned int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int a = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += a;
      }
      if (lane == 31) s_wi[warp] = (int)x;
      __syncthreads();
      if (warp == 0) {
        unsigned int a = (unsigned int)s_wi[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int a2 = __shfl_up_sync(0xffffffffu, a, o);
          if (lane >= o) a += a2;
        }
        s_wo[lane] = (int)a;
      }
      __syncthreads();
      if (n < N) {
        const unsigned int ex = carry + (warp > 0 ? (unsigned int)s_wo[warp - 1] : 0u) + x - v;
        s_iptr[n] = (int)(ex & 0xffffu);
        s_optr[n] = (int)(ex >> 16);
        s_dis[n] = 1.0f / sqrtf((float)(v >> 16));     // deg^-1/2, degree by source row incl. the loop (gcn_conv.py:66)
        s_cin[n] = 0;     // becomes the fill cursor
        s_cout[n] = 0;
      }
      carry += (unsigned int)s_wo[31];
      __syncthreads();
    }
    if (t == 0) {
      s_iptr[N] = (int)(carry & 0xffffu);
      s_optr[N] = (int)(carry >> 16);
    }
  }
  __syncthreads();
  PT_MARK();                                           // 3: scan
  // 2. fill in arrival order (arbitrary); the appended self loop takes the last slot of its row
  for (int e = t; e < E; e += T) {
    const unsigned int pk = s_edge[e];
    if (pk == 0xffffffffu) continue;
    const int r = (int)(pk >> 16), d = (int)(pk & 0xffffu);
    s_ikey[s_iptr[d] + atomicAdd(&s_cin[d], 1)] = (unsigned short)e;
    s_okey[s_optr[r] + atomicAdd(&s_cout[r], 1)] = (unsigned short)e;
  }
  for (int n = t; n < N; n += T) {
    s_ikey[s_iptr[n + 1] - 1] = (unsigned short)(E + n);
    s_okey[s_optr[n + 1] - 1] = (unsigned short)(E + n);
  }
  __syncthreads();
  PT_MARK();                                           // 4: fill
  // 3. this CTA's slice -> global memory.  Rows are ordered by edge_index column WITHOUT sorting: the
  //    final slot of an entry is its row start + the number of smaller keys in its (short) row, counted
  //    entry-parallel, so a hub row costs its length per thread instead of its length squared in one.
  for (int n = nlo + t; n < nhi; n += T) {
    c.in_ptr[n] = s_iptr[n];
    c.out_ptr[n] = s_optr[n];
    c.dis[n] = s_dis[n];
  }
  if (sj == kPrepS - 1 && t == 0) {
    c.in_ptr[N] = s_iptr[N];
    c.out_ptr[N] = s_optr[N];
  }
  {
    const int pb = s_iptr[nlo], pe = s_iptr[nhi];
    for (int u = pb + t; u < pe; u += T) {
      const int key = (int)s_ikey[u];
      int n, src;
      if (key < E) {
        const unsigned int pk = s_edge[key];
        src = (int)(pk >> 16);
        n = (int)(pk & 0xffffu);
      } else {
        n = src = key - E;
      }
      const int a = s_iptr[n], b = s_iptr[n + 1];
      int rank = 0;
      for (int v = a; v < b; ++v) rank += (int)s_ikey[v] < key;
      const int p = a + rank;
      c.in_key[p] = key;
      c.in_src[p] = src;
      c.in_norm[p] = s_dis[src] * s_dis[n];           // dis[row] * 1 * dis[col]
    }
    const int qb = s_optr[nlo], qe = s_optr[nhi];
    for (int u = qb + t; u < qe; u += T) {
      const int key = (int)s_okey[u];
      int n, dst;
      if (key < E) {
        const unsigned int pk = s_edge[key];
        n = (int)(pk >> 16);
        dst = (int)(pk & 0xffffu);
      } else {
        n = dst = key - E;
      }
      const int a = s_optr[n], b = s_optr[n + 1];
      int rank = 0;
      for (int v = a; v < b; ++v) rank += (int)s_okey[v] < key;
      const int a2 = s_iptr[dst], b2 = s_iptr[dst + 1];
      int rank2 = 0;
      for (int v = a2; v < b2; ++v) rank2 += (int)s_ikey[v] < key;
      const int q = a + rank;
      c.out_key[q] = key;
      c.out_dst[q] = dst;
      c.out_pos[q] = a2 + rank2;                      // in-CSR position of the same edge
      c.out_norm[q] = s_dis[n] * s_dis[dst];
    }
  }
  if (sj == 0) {
    __syncthreads();
    PT_MARK();                                         // 5: slice write-out
    if (t == 0) {
      c.status[0] = s_status;
      c.status[1] = c.status[2] = c.status[3] = 0;
      if (s_status != 0) atomicOr(&c.status[kStSticky], s_status);
    }
    CAL_TL(c.status, 12);
#ifdef CAL_PHASE_TIMING
    if (t == 0) for (int q_ = 0; q_ < pt_i; ++q_) c.status[96 + q_] = (int)pt_v[q_];
#endif
  }
}

int launch_prep(const Ctx& c, cudaStream_t s) {
  const size_t small = prep_small_smem(c.Nm, c.Em);
  if (small <= kPrepSmallMaxSmem && c.EP < 65535) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(k_prep_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPrepSmallMaxSmem);
      if (e != cudaSuccess) return (int)e;
      attr_set = true;
    }
    const int G = imax(1, imin(ceil_div(c.Nm, 128), 64));
    launch_k(k_prep_small, dim3(G + kPrepS), dim3(kPrepT), small, s, c);
    note_launches(1);
    CAL_CUDA_CHECK_LAUNCH();
    return 0;
  }
  const int T = 256;
  if (c.grouped) {                                  // cal_caps.grouped_edges: two launches, one CTA per graph
    const int gi2 = imax(1, imin(ceil_div(imax(imax(c.Nm, c.Bm + 1), c.Em), T), kMaxStatBlocks));
    launch_k(k_prep_init, dim3(gi2), dim3(T), 0, s, c);
    launch_k(k_prep_graph, dim3(imax(1, imin(c.Bm, 4 * kSMs))), dim3(kPgT), 0, s, c);
    note_launches(2);
    CAL_CUDA_CHECK_LAUNCH();
    return 0;
  }
  int gi = imax(1, imin(ceil_div(imax(imax(c.Nm, c.Bm + 1), c.Em), T), kMaxStatBlocks));
  int ge = imax(1, imin(ceil_div(imax(c.Em, c.Nm), T), 4 * kSMs));
  launch_k(k_prep_init, dim3(gi), dim3(T), 0, s, c);
  launch_k(k_prep_scan, dim3(1), dim3(1024), 0, s, c);
  launch_k(k_prep_fill, dim3(ge), dim3(T), 0, s, c);
  launch_k(k_prep_sort, dim3(imax(1, imin(ceil_div(c.Nm, 128), 8 * kSMs))), dim3(128), 0, s, c);
  launch_k(k_prep_link, dim3(imax(1, imin(ceil_div(c.Nm, 128), 8 * kSMs))), dim3(128), 0, s, c);
  note_launches(5);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Per-step parameter preparation: transposed copies of the conv weights (backward GEMMs) and of
// the fc1 weights (torch Linear stores [out, in]; the forward GEMM wants [in, out]); in eval mode
// also the BatchNorm affine from the running statistics.
// ---------------------------------------------------------------------------------------------
__global__ void k_param_prep(const Ctx c, const int n_tiles) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int H = c.H, hb = H / 32;
  const int n_conv = c.L + 2;
  int m = blockIdx.x;
  if (m < n_tiles) {
    // decode (matrix, 32x32 tile): conv matrices first (hb*hb tiles each), then the three fc1 weights
    const float* src;
    float* dst;
    int rows, cols, t;    // src is [rows][cols]; dst is [cols][rows]
    if (m < n_conv * hb * hb) {
      const int mat = m / (hb * hb);
      t = m - mat * hb * hb;
      long long off = mat < c.L ? c.po.convs_w[mat] : (mat == c.L ? c.po.context_w : c.po.objects_w);
      src = c.params + off;
      dst = c.wt_conv(mat);
      rows = H; cols = H;
    } else if (m < n_conv * hb * hb + 2 * hb * hb + hb * ((c.cat ? 2 * H : H) / 32)) {
      m -= n_conv * hb * hb;
      int h = 0;
      for (; h < 3; ++h) {
        const int nt = hb * (((h == 2 && c.cat) ? 2 * H : H) / 32);
        if (m < nt) break;
        m -= nt;
      }
      t = m;
      src = c.params + c.po.fc1_w[h];
      dst = c.wt_fc1(h);
      rows = H; cols = (h == 2 && c.cat) ? 2 * H : H;
    } else {                                           // CausalGIN: the second Linear of every layer
      m -= n_conv * hb * hb + 2 * hb * hb + hb * ((c.cat ? 2 * H : H) / 32);
      const int l = m / (hb * hb);
      t = m - l * hb * hb;
      src = c.params + c.po.gin_w2[l];
      dst = c.wt_gin2(l);
      rows = H; cols = H;
    }
    const int ct = cols / 32;
    const int r0 = (t / ct) * 32, c0 = (t % ct) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
      tile[i][threadIdx.x] = src[(size_t)(r0 + i) * cols + c0 + threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
      dst[(size_t)(c0 + i) * rows + r0 + threadIdx.x] = tile[threadIdx.x][i];
  } else if (!c.train) {
    // eval: y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta
    const int nbn = c.L + 9;
    const int t = threadIdx.y * blockDim.x + threadIdx.x;
    for (int id = 0; id < nbn; ++id) {
      const int K = c.bn_K[id];
      for (int k = t; k < K; k += blockDim.x * blockDim.y) {
        float rm = c.bn_buffers[c.bn_rm[id] + k], rv = c.bn_buffers[c.bn_rv[id] + k];
        float rstd = 1.0f / sqrtf(rv + c.eps);
        float sc = c.params[c.bn_gamma[id] + k] * rstd;
        c.bnf(id, BN_SCALE)[k] = sc;
        c.bnf(id, BN_SHIFT)[k] = c.params[c.bn_beta[id] + k] - rm * sc;
        c.bnf(id, BN_MEAN)[k] = rm;
        c.bnf(id, BN_RSTD)[k] = rstd;
      }
    }
  }
}

int launch_param_prep(const Ctx& c, cudaStream_t s) {
  const int hb = c.H / 32;
  const int n_tiles = (c.L + 2) * hb * hb + 2 * hb * hb + hb * ((c.cat ? 2 * c.H : c.H) / 32) +
                      (c.model == CAL_MODEL_GIN ? c.L * hb * hb : 0);
  launch_k(k_param_prep, dim3(n_tiles + 1), dim3(32, 8), 0, s, c, n_tiles);
  note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

}  // namespace cal
