// collate.cu -- mini-batch collation on the device.
//
// Replaces the host-side collate of the reference's loader (torch_geometric DataLoader ->
// Batch.from_data_list, train_causal.py:13-15,171-176: node features concatenated, edge_index
// offset by the running node count, `batch` = graph id per node, `y` concatenated) by one kernel
// over a dataset that lives in HBM.  The ids of the graphs of the
// step are read from a device-resident epoch order at a device-resident cursor, so one captured
// CUDA graph  collate -> prep -> forward -> backward -> Adam  serves a whole epoch with no
// host -> device traffic per step.
//
// Every CTA recomputes the exclusive scan of the B node / edge counts in shared memory (B <= a few
// thousand: cheaper than a second launch), then copies the graphs it owns.  Integer outputs are
// bit-exact with the host collate; features are copied verbatim.
#include "internal.cuh"

namespace cal {

namespace {

constexpr int kColT = 256;

struct ColArgs {
  cal_graph_store st;
  const int* order;       // i32[n_order] graph ids of the epoch, device
  int n_order;
  int* pos;               // device i32[4]: [0] offset of this step's first id, [1] arrival counter, [2] metrics pending, [3] graphs of the last collated batch
  int B;                  // graphs per step
  const int* perm_pool;   // i32[steps][B] random_idx per step (model.py:147-152), or NULL
  int Nm, Em, Bm;
  int* out_dims;
  int* out_perm;
  float* out_feat;
  long long* out_ei;
  long long ei_stride;
  long long* out_batch;
  long long* out_y;
  const float* prev_loss; // f32[8] loss parts of the previous step (CAL_WS_LOSS) or NULL
  float* acc;             // f32[8] epoch accumulators or NULL
  int advance;
};

__global__ void __launch_bounds__(kColT) k_collate(const ColArgs a) {
  pdl_sync();
  extern __shared__ int s_off[];                 // node_off [B+1] | edge_off [B+1] | gid [B]
  __shared__ int s_wn[kColT / 32], s_we[kColT / 32];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int B = a.B;
  int* s_no = s_off;
  int* s_eo = s_off + B + 1;
  int* s_gid = s_eo + B + 1;
  const int base = a.pos != nullptr ? *reinterpret_cast<volatile int*>(a.pos) : 0;
  const int Bn = imax(0, imin(B, a.n_order - base));        // graphs of this step (last batch may be short)
  // ---- exclusive scan of the node / edge counts of the step's graphs (chunks of kColT) ----
  int carry_n = 0, carry_e = 0;
  for (int b0 = 0; b0 < B; b0 += kColT) {
    const int b = b0 + t;
    int g = -1, nn = 0, ne = 0;
    if (b < Bn) {
      g = a.order[base + b];
      if (g < 0 || g >= a.st.num_graphs) g = -1;
      if (g >= 0) {
        nn = a.st.node_ptr[g + 1] - a.st.node_ptr[g];
        ne = a.st.edge_ptr[g + 1] - a.st.edge_ptr[g];
      }
    }
    int xn = nn, xe = ne;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int p = __shfl_up_sync(0xffffffffu, xn, o), q = __shfl_up_sync(0xffffffffu, xe, o);
      if (lane >= o) {
        xn += p;
        xe += q;
      }
    }
    if (lane == 31) {
      s_wn[warp] = xn;
      s_we[warp] = xe;
    }
    __syncthreads();
    int wn = 0, we = 0, tn = 0, te = 0;
#pragma unroll
    for (int w = 0; w < kColT / 32; ++w) {
      if (w < warp) {
        wn += s_wn[w];
        we += s_we[w];
      }
      tn += s_wn[w];
      te += s_we[w];
    }
    if (b < B) {
      s_no[b] = carry_n + wn + xn - nn;
      s_eo[b] = carry_e + we + xe - ne;
      s_gid[b] = g;
    }
    carry_n += tn;
    carry_e += te;
    __syncthreads();
  }
  if (t == 0) {
    s_no[B] = carry_n;
    s_eo[B] = carry_e;
  }
  __syncthreads();
  const int N = s_no[B], E = s_eo[B];
  const bool fits = N <= a.Nm && E <= a.Em && Bn <= a.Bm;
  if (blockIdx.x == 0 && t == 0) {
    // metrics of the PREVIOUS step (its kernels precede this one in stream order) -> epoch sums
    if (a.acc != nullptr && a.prev_loss != nullptr && a.pos != nullptr && a.pos[2] != 0) {
      const float pb = (float)a.pos[3];                                 // graphs of the batch collated before this one
      for (int k = 0; k < 4; ++k) a.acc[k] += a.prev_loss[k] * pb;      // batchmean losses -> sums over graphs
      for (int k = 4; k < 7; ++k) a.acc[k] += a.prev_loss[k];           // correct counts
      a.acc[7] += pb;
    }
    a.out_dims[0] = N;
    a.out_dims[1] = E;
    a.out_dims[2] = Bn;
    a.out_dims[3] = a.perm_pool != nullptr ? 1 : 0;
  }
  if (fits) {
    const int F = a.st.num_features;
    const int step = B > 0 ? base / B : 0;
    for (int b = blockIdx.x; b < Bn; b += gridDim.x) {
      const int g = s_gid[b];
      if (t == 0) {
        a.out_y[b] = g >= 0 ? a.st.y[g] : 0;
        int p = a.perm_pool != nullptr ? a.perm_pool[(size_t)step * B + b] : b;
        if (p < 0 || p >= Bn) p = b;
        a.out_perm[b] = p;
      }
      if (g < 0) continue;
      const int n0 = a.st.node_ptr[g], e0 = a.st.edge_ptr[g];
      const int no = s_no[b], nn = s_no[b + 1] - no;
      const int eo = s_eo[b], ne = s_eo[b + 1] - eo;
      const float* fs = a.st.feat + (size_t)n0 * F;
      float* fd = a.out_feat + (size_t)no * F;
      for (int i = t; i < nn * F; i += kColT) fd[i] = fs[i];
      for (int i = t; i < nn; i += kColT) a.out_batch[no + i] = b;
      for (int i = t; i < ne; i += kColT) {
        a.out_ei[eo + i] = (long long)(a.st.edge_src[e0 + i] + no);
        a.out_ei[a.ei_stride + eo + i] = (long long)(a.st.edge_dst[e0 + i] + no);
      }
    }
  }
  // every CTA has read pos[0] before it arrives here, so the last one may advance the cursor
  if (a.pos != nullptr) {
    __syncthreads();
    if (t == 0) {
      __threadfence();
      if (atomicAdd(reinterpret_cast<unsigned int*>(a.pos + 1), 1u) == gridDim.x - 1u) {
        a.pos[1] = 0;
        a.pos[2] = Bn > 0 ? 1 : 0;
        a.pos[3] = Bn;                                     // (the next collate / the flush weigh this step's losses with it:
        //                                                    with two staging buffers out_dims of THIS call's buffer is older)
        if (a.advance) a.pos[0] = base + Bn;
      }
    }
  }
}

__global__ void k_collate_flush(const float* prev_loss, const int* dims, int* pos, float* acc) {
  pdl_sync();
  (void)dims;
  if (threadIdx.x == 0 && blockIdx.x == 0 && pos[2] != 0) {
    const float pb = (float)pos[3];
    for (int k = 0; k < 4; ++k) acc[k] += prev_loss[k] * pb;
    for (int k = 4; k < 7; ++k) acc[k] += prev_loss[k];
    acc[7] += pb;
    pos[2] = 0;
  }
}

}  // namespace
}  // namespace cal

extern "C" int cal_collate(const cal_graph_store* st, const int32_t* order, int32_t n_order, int32_t* pos,
                           int32_t graphs_per_step, const int32_t* perm_pool, const cal_caps* caps,
                           const cal_batch* out, int advance, const float* prev_loss, float* acc, void* stream) {
  if (!st || !order || !caps || !out) return CAL_ENULL;
  if (!st->node_ptr || !st->edge_ptr || !st->feat || !st->edge_src || !st->edge_dst || !st->y) return CAL_ENULL;
  if (!out->dims || !out->feat || !out->edge_index || !out->batch || !out->y || !out->perm) return CAL_ENULL;
  if (graphs_per_step <= 0 || graphs_per_step > caps->max_graphs || n_order < 0 || st->num_features <= 0 ||
      st->num_graphs < 0 || out->edge_stride < caps->max_edges)
    return CAL_EINVAL;
  if ((acc != nullptr) != (prev_loss != nullptr)) return CAL_EINVAL;
  if (acc != nullptr && pos == nullptr) return CAL_EINVAL;
  cal::ColArgs a;
  a.st = *st;
  a.order = order;
  a.n_order = n_order;
  a.pos = pos;
  a.B = graphs_per_step;
  a.perm_pool = perm_pool;
  a.Nm = caps->max_nodes;
  a.Em = caps->max_edges;
  a.Bm = caps->max_graphs;
  a.out_dims = const_cast<int*>(out->dims);
  a.out_perm = const_cast<int*>(out->perm);
  a.out_feat = const_cast<float*>(out->feat);
  a.out_ei = reinterpret_cast<long long*>(const_cast<int64_t*>(out->edge_index));
  a.ei_stride = out->edge_stride;
  a.out_batch = reinterpret_cast<long long*>(const_cast<int64_t*>(out->batch));
  a.out_y = reinterpret_cast<long long*>(const_cast<int64_t*>(out->y));
  a.prev_loss = prev_loss;
  a.acc = acc;
  a.advance = advance;
  const size_t smem = (size_t)(3 * graphs_per_step + 2) * sizeof(int);
  if (smem > 48 * 1024) return CAL_EINVAL;
  const int grid = graphs_per_step < 2 * cal::kSMs ? graphs_per_step : 2 * cal::kSMs;
  cal::launch_k(cal::k_collate, dim3(grid), dim3(cal::kColT), smem, (cudaStream_t)stream, a);
  cal::note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}

extern "C" int cal_collate_flush(const float* prev_loss, const int32_t* dims, int32_t* pos, float* acc, void* stream) {
  if (!prev_loss || !dims || !pos || !acc) return CAL_ENULL;
  cal::launch_k(cal::k_collate_flush, dim3(1), dim3(32), 0, (cudaStream_t)stream, prev_loss, dims, pos, acc);
  cal::note_launches(1);
  CAL_CUDA_CHECK_LAUNCH();
  return 0;
}
