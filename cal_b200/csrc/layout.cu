// layout.cu -- workspace layout and argument validation (host only).
#include "internal.cuh"

namespace cal {

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int validate_model(const cal_model_desc* m) {
  if (m == nullptr) return CAL_ENULL;
  if (m->model != CAL_MODEL_GCN && m->model != CAL_MODEL_GAT && m->model != CAL_MODEL_GIN) return CAL_EINVAL;
  if (m->hidden != 32 && m->hidden != 64 && m->hidden != 128) return CAL_EUNSUPPORTED;
  if (m->num_features < 1 || m->num_features > 512) return CAL_EUNSUPPORTED;
  if (m->num_classes < 2 || m->num_classes > 32) return CAL_EUNSUPPORTED;
  if (m->layers < 1 || m->layers > CAL_MAX_LAYERS) return CAL_EINVAL;
  if (m->model == CAL_MODEL_GAT) {
    if (m->heads != 1 && m->heads != 2 && m->heads != 4 && m->heads != 8) return CAL_EINVAL;
    if (!(m->gat_dropout >= 0.f && m->gat_dropout < 1.f)) return CAL_EINVAL;
  }
  return CAL_OK;
}

int compute_layout(const cal_model_desc* m, const cal_caps* caps, Layout* lay) {
  int rc = validate_model(m);
  if (rc != CAL_OK) return rc;
  if (caps == nullptr) return CAL_ENULL;
  if (caps->max_nodes < 1 || caps->max_edges < 0 || caps->max_graphs < 1) return CAL_EINVAL;
  // the FFMA readout kernels keep all graph rows of a column slice in one CTA's shared memory; beyond that
  // the streaming tensor-core kernels (head_tc.cu) serve up to 512 graphs
  if ((readout_smem_bytes(caps->max_graphs, m->hidden, m->cat != 0, m->num_classes, 1) > 225 * 1024 ||
       readout_smem_bytes(caps->max_graphs, m->hidden, m->cat != 0, m->num_classes, 0) > 225 * 1024) &&
      caps->max_graphs > 512)
    return CAL_EUNSUPPORTED;
  const size_t Nm = caps->max_nodes, Em = caps->max_edges, Bm = caps->max_graphs, EP = Em + Nm;
  const size_t H = m->hidden, F = m->num_features, C = m->num_classes, L = m->layers;
  const size_t Fp = align_up(F, 4);
  lay->kmax = (int)align_up(Fp > 2 * H ? Fp : 2 * H, 32);
  lay->g_tile = imax(1, imin(ceil_div((int)Nm, kTileRows), kSMs));
  lay->g_row = imax(1, imin(ceil_div((int)Nm, 2 * kRowWarps), 2 * kSMs));
  lay->t_head1 = ceil_div((int)Bm, kTileRows);
  lay->g_head2 = ceil_div((int)Bm, kHeadRowsPerCta);
  lay->g_feat = lay->g_tile;
  lay->n_fchunk = ceil_div((int)F, kFeatChunk);
  const size_t G = lay->g_tile;

  size_t sz[CAL_WS_REGION_COUNT] = {0};
  sz[CAL_WS_STATUS] = 128 * 4;       // [0] status bits; [16..] optional phase-timing slots (CAL_PHASE_TIMING builds)
  sz[CAL_WS_COUNTERS] = (64 + kGsSites * kGsCounters) * 4;
  sz[CAL_WS_IN_PTR] = (Nm + 1) * 4;
  sz[CAL_WS_IN_SRC] = EP * 4;
  sz[CAL_WS_IN_KEY] = EP * 4;
  sz[CAL_WS_IN_NORM] = EP * 4;
  sz[CAL_WS_OUT_PTR] = (Nm + 1) * 4;
  sz[CAL_WS_OUT_DST] = EP * 4;
  sz[CAL_WS_OUT_POS] = EP * 4;
  sz[CAL_WS_OUT_KEY] = EP * 4;
  sz[CAL_WS_CNT_IN] = Nm * 4;
  sz[CAL_WS_CNT_OUT] = Nm * 4;
  sz[CAL_WS_GRAPH_PTR] = (Bm + 1) * 4;
  sz[CAL_WS_NODE_GRAPH] = Nm * 4;
  sz[CAL_WS_PERM] = Bm * 4;
  sz[CAL_WS_INVPERM] = Bm * 4;
  sz[CAL_WS_DIS] = Nm * 4;
  sz[CAL_WS_X] = (L + 1) * Nm * H * 4;
  sz[CAL_WS_NODE_ATT] = Nm * 2 * 4;
  sz[CAL_WS_PQ] = Nm * 4 * 4;
  sz[CAL_WS_EDGE_ATT] = EP * 2 * 4;
  sz[CAL_WS_DISW] = Nm * 2 * 4;
  sz[CAL_WS_AGG] = 2 * Nm * H * 4;
  sz[CAL_WS_Z] = 2 * Nm * H * 4;
  sz[CAL_WS_POOLED] = 2 * Bm * H * 4;
  sz[CAL_WS_H1] = 3 * Bm * H * 4;
  sz[CAL_WS_LOGP] = 3 * Bm * C * 4;
  sz[CAL_WS_LOSS] = (8 + 6 * (size_t)lay->g_head2) * 4;
  sz[CAL_WS_BN] = (size_t)kNumBN * BN_FIELDS * lay->kmax * 4;
  // legacy single-level partials (readout kernels) + the hierarchical grid-sum scratch
  lay->statp_legacy = (size_t)imax(kMaxStatBlocks, 3 * lay->g_head2) * 4 * lay->kmax;
  lay->gs_n = 4 * lay->kmax;
  lay->gs_stride = (size_t)(kMaxStatBlocks + kMaxStatBlocks / kGsGroup + 1) * lay->gs_n;
  sz[CAL_WS_STATP] = (lay->statp_legacy + kGsSites * lay->gs_stride) * 8;
  sz[CAL_WS_WT] = ((L + 2) * H * H + 3 * 2 * H * H + (m->model == CAL_MODEL_GIN ? L * H * H : 0)) * 4;
  if (m->model == CAL_MODEL_GIN) sz[CAL_WS_GAT] = (L + 1) * Nm * H * 4;
  if (m->model == CAL_MODEL_GAT)
    sz[CAL_WS_GAT] = gat_workspace_floats((int)Nm, (int)EP, (int)H, (int)L, m->heads) * 4;
  sz[CAL_WS_DLOGIT] = 3 * Bm * C * 4;
  sz[CAL_WS_DH] = 3 * Bm * H * 4;
  sz[CAL_WS_DU] = 3 * Bm * 2 * H * 4;
  sz[CAL_WS_DAGG] = 2 * Nm * H * 4;
  sz[CAL_WS_DYM] = 2 * Nm * H * 4;
  sz[CAL_WS_DNRM] = EP * 2 * 4;
  sz[CAL_WS_DT] = EP * 2 * 4;
  sz[CAL_WS_DP] = Nm * 2 * 4;
  sz[CAL_WS_D] = 2 * Nm * H * 4;
  sz[CAL_WS_OUT_NORM] = EP * 4;
  sz[CAL_WS_EDGE_WN] = EP * 2 * 4;
  sz[CAL_WS_EDGE_NA] = EP * 2 * 4;
  if (m->model == CAL_MODEL_GCN && m->hidden == 128) sz[CAL_WS_FSG] = fsg_region_bytes((int)Bm, (int)L, (int)F);
  sz[CAL_WS_EDGE_GPTR] = (2 * Bm + 8) * 4;

  size_t gp = 0;
  for (size_t l = 0; l < L + 2; ++l) {
    lay->gp_conv[l] = gp;
    gp += G * (H * H + H);
  }
  lay->gp_att = gp;
  gp += (size_t)lay->g_row * (8 * H + 4);
  lay->gp_feat = gp;
  gp += G * (F * H + H);
  for (int h = 0; h < 3; ++h) {
    lay->gp_fc1[h] = gp;
    gp += (size_t)lay->t_head1 * (H * 2 * H + H);
  }
  for (int h = 0; h < 3; ++h) {
    lay->gp_fc2[h] = gp;
    gp += (size_t)lay->g_head2 * (C * H + C);
  }
  for (size_t l = 0; l < L; ++l) {
    lay->gp_gat[l] = gp;
    if (m->model == CAL_MODEL_GAT) gp += G * 2 * H;
  }
  for (size_t l = 0; l < L; ++l) {
    lay->gp_gin2[l] = gp;
    if (m->model == CAL_MODEL_GIN) gp += G * (H * H + H);
  }
  sz[CAL_WS_GPART] = gp * 4;

  size_t off = 0;
  for (int r = 0; r < CAL_WS_REGION_COUNT; ++r) {
    lay->off[r] = off;
    lay->size[r] = sz[r];
    off += align_up(sz[r], 256);
  }
  lay->total = off;
  return CAL_OK;
}

}  // namespace cal
