// api.cu -- the C ABI of libcal_b200.so (include/cal_b200.h): argument checks, context
// construction and the launch sequences of the forward and backward passes.
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "internal.cuh"

namespace cal {

static std::atomic<unsigned long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

namespace {

__global__ void k_copy_logp(const Ctx c, float* __restrict__ out) {
  pdl_sync();
  const int B = imin(imax(c.dims[2], 0), c.Bm), C = c.C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * B * C; i += gridDim.x * blockDim.x) {
    const int h = i / (B * C), r = i - h * B * C;
    out[i] = c.logp[(size_t)h * c.Bm * C + r];
  }
}

// CAL_F_STAGES(lo, hi) -> [lo, hi] clipped to the pass; absent -> unchanged (whole pass)
void stage_range(int flags, int* lo, int* hi) {
  const int a = (flags >> 8) & 0xff, b = (flags >> 16) & 0xff;
  if (a == 0 && b == 0) return;
  if (a > 0) *lo = imax(*lo, a - 1);
  if (b > 0) *hi = imin(*hi, b - 1);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int build_ctx(const cal_model_desc* m, const cal_caps* caps, const cal_param_offsets* po, const cal_bn_offsets* bo,
              const cal_batch* b, void* workspace, size_t ws_bytes, Ctx* out) {
  Layout lay;
  int rc = compute_layout(m, caps, &lay);
  if (rc != CAL_OK) return rc;
  if (workspace == nullptr || b == nullptr) return CAL_ENULL;
  if (!aligned16(workspace)) return CAL_EALIGN;
  if (ws_bytes < lay.total) return CAL_ECAPACITY;
  if (b->dims == nullptr || b->feat == nullptr || b->batch == nullptr) return CAL_ENULL;
  if (caps->max_edges > 0 && b->edge_index == nullptr) return CAL_ENULL;
  if (!aligned16(b->feat)) return CAL_EALIGN;
  Ctx& c = *out;
  memset(&c, 0, sizeof(Ctx));
  c.model = m->model;
  c.F = m->num_features;
  c.H = m->hidden;
  c.C = m->num_classes;
  c.L = m->layers;
  c.heads = m->model == CAL_MODEL_GAT ? m->heads : 1;
  c.cat = m->cat != 0;
  c.no_natt = m->without_node_attention != 0;
  c.no_eatt = m->without_edge_attention != 0;
  c.eps = m->bn_eps;
  c.momentum = m->bn_momentum;
  c.w_c = m->w_c;
  c.w_o = m->w_o;
  c.w_co = m->w_co;
  c.gat_p = m->gat_dropout;
  c.readout_bf16 = m->readout_bf16 != 0;
  c.readout_tc = m->readout_tc != 0;
  c.Nm = caps->max_nodes;
  c.Em = caps->max_edges;
  c.Bm = caps->max_graphs;
  c.EP = c.Em + c.Nm;
  c.kmax = lay.kmax;
  c.g_tile = lay.g_tile;
  c.g_row = lay.g_row;
  c.t_head1 = lay.t_head1;
  c.g_head2 = lay.g_head2;
  c.dims = b->dims;
  c.feat = b->feat;
  c.ei_row = reinterpret_cast<const long long*>(b->edge_index);
  c.ei_col = b->edge_index ? reinterpret_cast<const long long*>(b->edge_index) + b->edge_stride : nullptr;
  c.batch = reinterpret_cast<const long long*>(b->batch);
  c.y = reinterpret_cast<const long long*>(b->y);
  c.perm_in = b->perm;
  c.gat_keep = b->gat_keep;
  if (po != nullptr) c.po = *po;
  const int L = c.L;
  for (int i = 0; i < kNumBN; ++i) {
    c.bn_gamma[i] = c.bn_beta[i] = c.bn_rm[i] = c.bn_rv[i] = -1;
    c.bn_K[i] = c.H;
  }
  if (po != nullptr) {
    c.bn_gamma[0] = po->bn_feat_w;
    c.bn_beta[0] = po->bn_feat_b;
    c.bn_K[0] = c.F;
    for (int l = 0; l < L; ++l) {
      c.bn_gamma[1 + l] = po->bns_conv_w[l];
      c.bn_beta[1 + l] = po->bns_conv_b[l];
    }
    c.bn_gamma[L + 1] = po->bnc_w;
    c.bn_beta[L + 1] = po->bnc_b;
    c.bn_gamma[L + 2] = po->bno_w;
    c.bn_beta[L + 2] = po->bno_b;
    for (int h = 0; h < 3; ++h) {
      c.bn_gamma[L + 3 + h] = po->fc1_bn_w[h];
      c.bn_beta[L + 3 + h] = po->fc1_bn_b[h];
      c.bn_gamma[L + 6 + h] = po->fc2_bn_w[h];
      c.bn_beta[L + 6 + h] = po->fc2_bn_b[h];
    }
    if (c.cat) c.bn_K[L + 5] = 2 * c.H;
  }
  c.bn_K[kBnIdentity] = c.H;
  if (bo != nullptr)
    for (int i = 0; i < L + 9; ++i) {
      c.bn_rm[i] = bo->running_mean[i];
      c.bn_rv[i] = bo->running_var[i];
    }
  unsigned char* w = static_cast<unsigned char*>(workspace);
#define REG(T, r) reinterpret_cast<T*>(w + lay.off[r])
  c.status = REG(int, CAL_WS_STATUS);
  c.counters = REG(unsigned int, CAL_WS_COUNTERS);
  c.in_ptr = REG(int, CAL_WS_IN_PTR);
  c.in_src = REG(int, CAL_WS_IN_SRC);
  c.in_key = REG(int, CAL_WS_IN_KEY);
  c.in_norm = REG(float, CAL_WS_IN_NORM);
  c.out_ptr = REG(int, CAL_WS_OUT_PTR);
  c.out_dst = REG(int, CAL_WS_OUT_DST);
  c.out_pos = REG(int, CAL_WS_OUT_POS);
  c.out_key = REG(int, CAL_WS_OUT_KEY);
  c.cnt_in = REG(int, CAL_WS_CNT_IN);
  c.cnt_out = REG(int, CAL_WS_CNT_OUT);
  c.graph_ptr = REG(int, CAL_WS_GRAPH_PTR);
  c.node_graph = REG(int, CAL_WS_NODE_GRAPH);
  c.perm = REG(int, CAL_WS_PERM);
  c.invperm = REG(int, CAL_WS_INVPERM);
  c.dis = REG(float, CAL_WS_DIS);
  c.X = REG(float, CAL_WS_X);
  c.natt = REG(float, CAL_WS_NODE_ATT);
  c.pq = REG(float, CAL_WS_PQ);
  c.watt = REG(float, CAL_WS_EDGE_ATT);
  c.disw = REG(float, CAL_WS_DISW);
  c.agg = REG(float, CAL_WS_AGG);
  c.Z = REG(float, CAL_WS_Z);
  c.pooled = REG(float, CAL_WS_POOLED);
  c.H1 = REG(float, CAL_WS_H1);
  c.logp = REG(float, CAL_WS_LOGP);
  c.loss = REG(float, CAL_WS_LOSS);
  c.bn = REG(float, CAL_WS_BN);
  c.statp = REG(double, CAL_WS_STATP);
  c.gsum = c.statp + lay.statp_legacy;
  c.gs_n = lay.gs_n;
  c.gs_stride = lay.gs_stride;
  c.gs_cnt = c.counters + 64;
  c.WT = REG(float, CAL_WS_WT);
  c.gat = REG(float, CAL_WS_GAT);
  c.dlogit = REG(float, CAL_WS_DLOGIT);
  c.dh = REG(float, CAL_WS_DH);
  c.du = REG(float, CAL_WS_DU);
  c.dagg = REG(float, CAL_WS_DAGG);
  c.dym = REG(float, CAL_WS_DYM);
  c.dnrm = REG(float, CAL_WS_DNRM);
  c.dt = REG(float, CAL_WS_DT);
  c.dp = REG(float, CAL_WS_DP);
  c.D = REG(float, CAL_WS_D);
  c.gpart = REG(float, CAL_WS_GPART);
  c.out_norm = REG(float, CAL_WS_OUT_NORM);
  c.edge_wn = REG(float, CAL_WS_EDGE_WN);
  c.edge_na = REG(float, CAL_WS_EDGE_NA);
  c.fsg = REG(unsigned char, CAL_WS_FSG);
  c.egp = REG(int, CAL_WS_EDGE_GPTR);
  c.grouped = caps->grouped_edges != 0;
  c.fsg_on = m->model == CAL_MODEL_GCN && m->hidden == 128 && m->num_features <= 128 && caps->max_graphs <= kSMs &&
             caps->small_graphs != 0;
  c.fsg_bwd_on = c.fsg_on && caps->small_graphs != 2;
#undef REG
  for (int l = 0; l < CAL_MAX_LAYERS + 2; ++l) c.gp_conv[l] = lay.gp_conv[l];
  c.gp_att = lay.gp_att;
  c.gp_feat = lay.gp_feat;
  for (int h = 0; h < 3; ++h) {
    c.gp_fc1[h] = lay.gp_fc1[h];
    c.gp_fc2[h] = lay.gp_fc2[h];
  }
  for (int l = 0; l < CAL_MAX_LAYERS; ++l) {
    c.gp_gat[l] = lay.gp_gat[l];
    c.gp_gin2[l] = lay.gp_gin2[l];
  }
  return CAL_OK;
}

int check_offsets(const cal_model_desc* m, const cal_param_offsets* po) {
  if (po == nullptr) return CAL_ENULL;
  if (po->total <= 0) return CAL_EINVAL;
  const int64_t req[] = {po->bn_feat_w, po->bn_feat_b, po->conv_feat_w, po->edge_att_w, po->edge_att_b,
                         po->node_att_w, po->node_att_b, po->bnc_w, po->bnc_b, po->bno_w, po->bno_b,
                         po->context_w, po->context_b, po->objects_w, po->objects_b};
  for (int64_t v : req)
    if (v < 0 || v >= po->total) return CAL_EINVAL;
  for (int l = 0; l < m->layers; ++l) {
    if (po->bns_conv_w[l] < 0 || po->bns_conv_b[l] < 0 || po->convs_w[l] < 0 || po->convs_b[l] < 0) return CAL_EINVAL;
    if (m->model == CAL_MODEL_GAT && po->convs_att[l] < 0) return CAL_EINVAL;
    if (m->model == CAL_MODEL_GIN && (po->gin_w2[l] < 0 || po->gin_b2[l] < 0 || po->gin_w2[l] % 4 != 0)) return CAL_EINVAL;
    if (po->convs_w[l] % 4 != 0) return CAL_EALIGN;       // staged with 16-byte cp.async
  }
  if (po->context_w % 4 != 0 || po->objects_w % 4 != 0) return CAL_EALIGN;
  for (int h = 0; h < 3; ++h) {
    if (po->fc1_bn_w[h] < 0 || po->fc1_bn_b[h] < 0 || po->fc1_w[h] < 0 || po->fc1_b[h] < 0 || po->fc2_bn_w[h] < 0 ||
        po->fc2_bn_b[h] < 0 || po->fc2_w[h] < 0 || po->fc2_b[h] < 0)
      return CAL_EINVAL;
    if (po->fc1_w[h] % 4 != 0 || po->fc2_w[h] % 4 != 0) return CAL_EALIGN;
  }
  return CAL_OK;
}

}  // namespace
}  // namespace cal

using namespace cal;

extern "C" {

int cal_abi_version(void) { return CAL_ABI_VERSION; }

uint64_t cal_launch_count(void) { return (uint64_t)g_launches.load(std::memory_order_relaxed); }

int cal_stage_count(const cal_model_desc* m, int pass) {
  if (validate_model(m) != CAL_OK) return CAL_EINVAL;
  if (pass == CAL_PASS_FORWARD) return m->layers + 6;
  if (pass == CAL_PASS_BACKWARD) return m->layers + 7;
  return CAL_EINVAL;
}

const char* cal_stage_name(const cal_model_desc* m, int pass, int stage) {
  static thread_local char buf[32];
  if (validate_model(m) != CAL_OK || stage < 0) return "";
  const int L = m->layers;
  if (pass == CAL_PASS_FORWARD) {
    if (stage == 0) return "param_prep";
    if (stage == 1) return "feat";
    if (stage < 2 + L) { snprintf(buf, sizeof buf, "layer_%d", stage - 2); return buf; }
    const char* tail[] = {"edge_att", "masked_convs", "readout", "copy_out"};
    return stage - 2 - L < 4 ? tail[stage - 2 - L] : "";
  }
  if (pass == CAL_PASS_BACKWARD) {
    const char* head[] = {"readout_bwd", "masked_gemm_bwd", "masked_gather_bwd", "norm_bwd", "att_bwd"};
    if (stage < 5) return head[stage];
    if (stage < 5 + L) { snprintf(buf, sizeof buf, "layer_%d_bwd", L - 1 - (stage - 5)); return buf; }
    if (stage == 5 + L) return "feat_bwd";
    if (stage == 6 + L) return "grad_reduce";
  }
  return "";
}

const char* cal_error_string(int code) {
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case CAL_OK: return "ok";
    case CAL_EINVAL: return "invalid argument or unsupported shape";
    case CAL_ENULL: return "required pointer is NULL";
    case CAL_EALIGN: return "pointer or offset not 16-byte aligned";
    case CAL_ECAPACITY: return "workspace too small for the requested capacities";
    case CAL_EUNSUPPORTED: return "unsupported model configuration";
    case CAL_ETIMEOUT: return "data-parallel exchange timed out waiting for a peer";
    default: return "unknown error";
  }
}

size_t cal_workspace_bytes(const cal_model_desc* m, const cal_caps* caps) {
  Layout lay;
  if (compute_layout(m, caps, &lay) != CAL_OK) return 0;
  return lay.total;
}

int cal_workspace_region(const cal_model_desc* m, const cal_caps* caps, int region, size_t* offset_bytes,
                         size_t* size_bytes) {
  Layout lay;
  int rc = compute_layout(m, caps, &lay);
  if (rc != CAL_OK) return rc;
  if (region < 0 || region >= CAL_WS_REGION_COUNT) return CAL_EINVAL;
  if (offset_bytes) *offset_bytes = lay.off[region];
  if (size_bytes) *size_bytes = lay.size[region];
  return CAL_OK;
}

int cal_image_sink_init(const cal_model_desc* m, const cal_caps* caps, const cal_param_offsets* po, void* workspace,
                        size_t ws_bytes, cal_image_sink* sink) {
  if (sink == nullptr) return CAL_ENULL;
  memset(sink, 0, sizeof(*sink));
  int rc = validate_model(m);
  if (rc != CAL_OK) return rc;
  rc = check_offsets(m, po);
  if (rc != CAL_OK) return rc;
  alignas(16) static const int64_t no_data[2] = {0, 0};             // (the batch pointers are only validated here, never read)
  cal_batch nb;
  memset(&nb, 0, sizeof(nb));
  nb.dims = reinterpret_cast<const int32_t*>(no_data);
  nb.feat = reinterpret_cast<const float*>(no_data);
  nb.edge_index = no_data;
  nb.batch = no_data;
  Ctx c;
  rc = build_ctx(m, caps, po, nullptr, &nb, workspace, ws_bytes, &c);
  if (rc != CAL_OK) return rc;
  c.train = 1;
  if (!(c.fsg_bwd_on && readout_runs_ro(c))) return CAL_OK;          // count = 0: the fused path is not taken
  fsg_fill_image_sink(c, sink);
  for (int q = 0; q < sink->count; ++q)
    if (sink->entry[q].offset % 4 != 0) {                             // (the optimizer kernels write four image words at a time)
      memset(sink, 0, sizeof(*sink));
      return CAL_OK;
    }
  return CAL_OK;
}

int cal_prep(const cal_model_desc* m, const cal_caps* caps, const cal_batch* b, void* workspace, size_t ws_bytes,
             void* stream) {
  Ctx c;
  int rc = build_ctx(m, caps, nullptr, nullptr, b, workspace, ws_bytes, &c);
  if (rc != CAL_OK) return rc;
  return launch_prep(c, (cudaStream_t)stream);
}

int cal_read_status(const cal_model_desc* m, const cal_caps* caps, void* workspace, void* stream) {
  Layout lay;
  int rc = compute_layout(m, caps, &lay);
  if (rc != CAL_OK) return rc;
  if (workspace == nullptr) return CAL_ENULL;
  // status[0]: the batch prepared last; status[kStSticky]: everything raised since the previous read (cleared here)
  int st[kStSticky + 1] = {0};
  unsigned char* base = static_cast<unsigned char*>(workspace) + lay.off[CAL_WS_STATUS];
  cudaError_t e = cudaMemcpyAsync(st, base, sizeof(st), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e == cudaSuccess && st[kStSticky] != 0) e = cudaMemsetAsync(base + 4 * kStSticky, 0, 4, (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  return st[0] | st[kStSticky];
}

int cal_causal_forward(const cal_model_desc* m, const cal_caps* caps, const cal_param_offsets* po,
                       const cal_bn_offsets* bo, const float* params, float* bn_buffers,
                       int64_t* bn_num_batches_tracked, const cal_batch* b, int flags, float* out_logp,
                       void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate_model(m);
  if (rc != CAL_OK) return rc;
  rc = check_offsets(m, po);
  if (rc != CAL_OK) return rc;
  if (params == nullptr) return CAL_ENULL;
  if (!aligned16(params)) return CAL_EALIGN;
  const int train = (flags & CAL_F_TRAIN) != 0;
  if (!train && (bo == nullptr || bn_buffers == nullptr)) return CAL_ENULL;   // eval needs running statistics
  if ((flags & CAL_F_LOSS) && b != nullptr && b->y == nullptr) return CAL_ENULL;
  Ctx c;
  rc = build_ctx(m, caps, po, bo, b, workspace, ws_bytes, &c);
  if (rc != CAL_OK) return rc;
  c.params = params;
  c.bn_buffers = bn_buffers;
  c.nbt = reinterpret_cast<long long*>(bn_num_batches_tracked);
  c.train = train;
  c.with_loss = (flags & CAL_F_LOSS) != 0;
  c.raw_o = (flags & CAL_F_RAW_LOGITS_O) != 0;
  if (c.raw_o && (readout_path_id(c) != 0 || c.with_loss)) return CAL_EUNSUPPORTED;   // FFMA readout kernels, loss by the caller
  cudaStream_t s = (cudaStream_t)stream;
  const int L = c.L;
  int lo = 0, hi = L + 5;
  stage_range(flags, &lo, &hi);
  // CAL_F_FSG_READY (include/cal_b200.h): honoured only where stage 0 is k_fsg_prep alone
  const bool ready = (flags & CAL_F_FSG_READY) != 0 && c.train && c.fsg_bwd_on && readout_runs_ro(c);
  for (int st = lo; st <= hi; ++st) {
    if (st == 0) {
      // (the transposed weight copies of k_param_prep serve the tiled backward and the FFMA readouts; the fused
      // small-graph path reads the operand images of k_fsg_prep instead -- in training mode it skips the launch)
      if (ready) continue;                               // CAL_F_FSG_READY: the operand images are in place
      if (!(c.train && c.fsg_bwd_on && readout_runs_ro(c))) rc = launch_param_prep(c, s);
      if (rc == 0 && c.fsg_on) rc = launch_fsg_prep(c, s, (flags & CAL_F_NO_OVERLAP) != 0);
    } else if (c.fsg_on && st >= 1 && st <= 3 + L) {
      // fused small-graph path: stage "feat" runs the whole forward up to the pooled embeddings (fsg.cu)
      if (st == 1) rc = launch_fsg_forward(c, s, ready);
    } else if (st == 1) rc = launch_feat_forward(c, s);
    else if (st < 2 + L)
      rc = c.model == CAL_MODEL_GAT ? launch_gat_forward(c, st - 2, s)
                                    : (c.model == CAL_MODEL_GIN ? launch_gin_forward(c, st - 2, s) : launch_conv_forward(c, st - 2, s));
    else if (st == 2 + L) rc = launch_edge_att(c, s);
    else if (st == 3 + L) rc = launch_masked_forward(c, s);
    else if (st == 4 + L) rc = launch_heads_forward(c, c.with_loss, s);
    else if (st == 5 + L && out_logp != nullptr) {
      launch_k(k_copy_logp, dim3(imax(1, imin(ceil_div(3 * c.Bm * c.C, 256), kSMs))), dim3(256), 0, s, c, out_logp);
      note_launches(1);
      CAL_CUDA_CHECK_LAUNCH();
    }
    if (rc != 0) return rc;
  }
  return CAL_OK;
}

int cal_causal_backward(const cal_model_desc* m, const cal_caps* caps, const cal_param_offsets* po,
                        const float* params, const cal_batch* b, const float* grad_logp, float* grads, int flags,
                        void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate_model(m);
  if (rc != CAL_OK) return rc;
  rc = check_offsets(m, po);
  if (rc != CAL_OK) return rc;
  if (params == nullptr || grads == nullptr) return CAL_ENULL;
  if (!aligned16(params) || !aligned16(grads)) return CAL_EALIGN;
  if (grad_logp == nullptr && (b == nullptr || b->y == nullptr)) return CAL_ENULL;
  Ctx c;
  rc = build_ctx(m, caps, po, nullptr, b, workspace, ws_bytes, &c);
  if (rc != CAL_OK) return rc;
  c.params = params;
  c.grads = grads;
  c.train = 1;
  c.grad_logp = grad_logp;
  c.raw_o = (flags & CAL_F_RAW_LOGITS_O) != 0;
  if (c.raw_o && (readout_path_id(c) != 0 || grad_logp == nullptr)) return CAL_EUNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  const int L = c.L;
  int lo = 0, hi = L + 6;
  stage_range(flags, &lo, &hi);
  for (int st = lo; st <= hi; ++st) {
    if (st == 0) rc = launch_heads_backward(c, s);
    else if (c.fsg_bwd_on && st >= 1 && st <= 5 + L) {
      // fused small-graph path: stage "masked_gemm_bwd" runs the whole backward down to the input transform (fsg_bwd.cu)
      if (st == 1) rc = launch_fsg_backward(c, s);
    } else if (c.fsg_bwd_on && st == 6 + L) rc = launch_fsg_grad_reduce(c, s);
    else if (st == 1) rc = launch_masked_bwd_gemm(c, s);
    else if (st == 2) rc = launch_masked_bwd_gather(c, s);
    else if (st == 3) rc = launch_norm_backward(c, s);
    else if (st == 4) rc = launch_att_backward(c, s);
    else if (st < 5 + L) {
      const int l = L - 1 - (st - 5);
      rc = c.model == CAL_MODEL_GAT ? launch_gat_backward(c, l, s)
                                    : (c.model == CAL_MODEL_GIN ? launch_gin_backward(c, l, s) : launch_conv_backward(c, l, s));
    } else if (st == 5 + L) rc = launch_feat_backward(c, s);
    else if (st == 6 + L) rc = launch_grad_reduce(c, s);
    if (rc != 0) return rc;
  }
  return CAL_OK;
}

}  // extern "C"
