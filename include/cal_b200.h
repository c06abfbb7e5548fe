/*
 * cal_b200.h -- C ABI of the B200-native CAL hot path (libcal_b200.so).
 *
 * Drop-in boundary for the CausalGCN / CausalGAT forward + backward of
 * yongduosui/CAL.  Every entry point names the reference interface it
 * replaces (file:line relative to the upstream repo).  Plain pointers and
 * sizes only: no torch types, no C++ types, no exceptions.
 *
 * Conventions
 *  - All data pointers are DEVICE pointers, 16-byte aligned, contiguous
 *    row-major.  The library never allocates, frees or retains them.
 *  - `stream` is a cudaStream_t passed as void*.  Calls are asynchronous
 *    w.r.t. the host, never synchronise, and are CUDA-graph capturable.
 *  - Return value: 0 = ok, < 0 = CAL_E* argument error (checked on the host
 *    before any launch), > 0 = cudaError_t of a failed launch.
 *  - Batch sizes live in DEVICE memory (`dims`: int32[4] = {N nodes,
 *    E edge_index columns, B graphs, 0}) so that one captured CUDA graph
 *    serves every batch that fits the capacities in `cal_caps`.
 *  - The workspace must be zero-filled ONCE after allocation (cal_prep keeps its degree counters
 *    zero between calls instead of clearing them at the start of every call).
 *  - Data-dependent violations (node id out of range, `batch` not sorted)
 *    are reported through the int32 status word at workspace region
 *    CAL_WS_STATUS (0 = ok), not through the return value.
 */
#ifndef CAL_B200_H
#define CAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAL_ABI_VERSION 3
#define CAL_MAX_LAYERS 8
#define CAL_MAX_BN (1 + CAL_MAX_LAYERS + 2 + 6)

/* error codes */
#define CAL_OK 0
#define CAL_EINVAL (-1)      /* bad scalar argument / unsupported shape */
#define CAL_ENULL (-2)       /* required pointer is NULL */
#define CAL_EALIGN (-3)      /* pointer not 16-byte aligned */
#define CAL_ECAPACITY (-4)   /* workspace too small for the capacities */
#define CAL_EUNSUPPORTED (-5)
#define CAL_ETIMEOUT (-6)    /* a data-parallel peer did not deliver its gradients in time */

/* model kind */
#define CAL_MODEL_GCN 0      /* CausalGCN, model.py:12-164 */
#define CAL_MODEL_GAT 1      /* CausalGAT, model.py:315-450 */
#define CAL_MODEL_GIN 2      /* CausalGIN, model.py:166-313: layers = GINConv(Linear-BN-ReLU-Linear-ReLU), no bns_conv */

/* Model hyper-parameters: the `args` fields the reference models read
 * (model.py:24-37,65-77; opts.py:32-52) plus the loss weights of
 * train_causal.py:183 (opts.py:43-45). */
typedef struct {
  int32_t model;               /* CAL_MODEL_* */
  int32_t num_features;        /* F */
  int32_t hidden;              /* H: 32, 64 or 128 */
  int32_t num_classes;         /* C: 2..32 */
  int32_t layers;              /* L: 1..CAL_MAX_LAYERS */
  int32_t heads;               /* GAT heads (model.py:319): 1, 2, 4 or 8 */
  int32_t cat;                 /* args.cat_or_add == "cat" (model.py:65-75) */
  int32_t without_node_attention; /* model.py:106-107 */
  int32_t without_edge_attention; /* model.py:99-100 */
  float gat_dropout;           /* model.py:320,340 */
  float bn_eps;                /* 1e-5 */
  float bn_momentum;           /* 0.1 */
  float w_c, w_o, w_co;        /* loss weights args.c / args.o / args.co */
  int32_t readout_bf16;        /* 0: readout MLP GEMMs as 3xTF32 on the tensor cores (fp32 accuracy, the 1e-5 parity path);
                                  1: bf16 operands, fp32 accumulate (BASELINE.json configs[4] "bf16 MLP / fp32 aggregate") */
  int32_t readout_tc;          /* 0: auto -- the fp32 readout runs on the FFMA cluster kernels (measured faster at B = 128, where the
                                  readout is latency- not FLOP-bound: profiles/README.md), bf16 on the tensor cores;
                                  1: always the tensor-core kernels (tcgen05.mma, accumulators in TMEM) */
} cal_model_desc;

/* Capacities a workspace is sized for. */
typedef struct {
  int32_t max_nodes;
  int32_t max_edges;           /* edge_index columns before self-loop surgery */
  int32_t max_graphs;
  int32_t small_graphs;        /* != 0: the caller guarantees that every graph of every batch has <= 40 nodes and <= 320 edges after
                                  self-loop surgery (edge_index columns without self loops + one per node).  CausalGCN with
                                  hidden 128, <= 128 features and <= 148 graphs per batch then runs the fused small-graph
                                  forward (one persistent kernel, graph blocks resident in shared memory, tensor-core node
                                  transforms; csrc/fsg.cu) and the fused small-graph backward (csrc/fsg_bwd.cu).  2: the fused
                                  forward only (the tiled backward kernels run on what it saved; A/B tests).  A batch that
                                  breaks the promise sets CAL_ST_CAPACITY. */
  int32_t grouped_edges;       /* != 0: the caller guarantees that the edge_index columns of every batch are grouped by graph, in
                                  graph order, with both endpoints inside the graph (what Batch.from_data_list / PyG collate
                                  and cal_collate produce), and that no graph has more than 512 nodes or 4096 edge_index
                                  columns.  Batches too large for the single-kernel structure path are then prepared by
                                  one CTA per graph in shared memory (k_prep_graph) instead of the global counting sort.
                                  A batch that breaks the promise sets CAL_ST_BAD_BATCH / CAL_ST_CAPACITY. */
} cal_caps;

/* Offsets (in floats) of every parameter inside the flat parameter buffer;
 * the flat gradient buffer uses the same offsets.  Names follow the
 * reference state_dict (model.py:38-75, gcn_conv.py:30-35).  Index 0/1/2 of
 * the fc* arrays = the c / o / co readout (model.py:55-75). -1 = absent. */
typedef struct {
  int64_t bn_feat_w, bn_feat_b;
  int64_t conv_feat_w, conv_feat_b;
  int64_t bns_conv_w[CAL_MAX_LAYERS], bns_conv_b[CAL_MAX_LAYERS];
  int64_t convs_w[CAL_MAX_LAYERS], convs_b[CAL_MAX_LAYERS], convs_att[CAL_MAX_LAYERS];
  int64_t edge_att_w, edge_att_b, node_att_w, node_att_b;
  int64_t bnc_w, bnc_b, bno_w, bno_b;
  int64_t context_w, context_b, objects_w, objects_b;
  int64_t fc1_bn_w[3], fc1_bn_b[3], fc1_w[3], fc1_b[3];
  int64_t fc2_bn_w[3], fc2_bn_b[3], fc2_w[3], fc2_b[3];
  /* CausalGIN (model.py:187-193), layer i = convs.i.nn: convs_w/convs_b = nn.0 (Linear, [out, in] like
   * every torch Linear), bns_conv_w/bns_conv_b = nn.1 (the BatchNorm inside the layer), and: */
  int64_t gin_w2[CAL_MAX_LAYERS], gin_b2[CAL_MAX_LAYERS];   /* nn.3 (Linear) */
  int64_t total;               /* number of floats in the flat buffer */
} cal_param_offsets;

/* BatchNorm running statistics (torch BatchNorm1d buffers).  BN ids:
 * 0 = bn_feat, 1..L = bns_conv[i], L+1 = bnc, L+2 = bno,
 * L+3+h = fc1_bn_{c,o,co}, L+6+h = fc2_bn_{c,o,co}.  CausalGIN: 1..L = convs[i].nn[1]. */
typedef struct {
  int64_t running_mean[CAL_MAX_BN];   /* offsets in floats into bn_buffers */
  int64_t running_var[CAL_MAX_BN];
} cal_bn_offsets;

/* One mini-batch (PyG Batch fields read at model.py:87-89 and
 * train_causal.py:176), all device pointers. */
typedef struct {
  const int32_t* dims;         /* int32[4] = {N, E, B, 0} */
  const float* feat;           /* f32[N,F]  data.x / data.feat */
  const int64_t* edge_index;   /* i64[2,E]  row 0 = source, row 1 = target; rows are E apart */
  const int64_t* batch;        /* i64[N]    non-decreasing graph id */
  const int64_t* y;            /* i64[B]    labels (may be NULL for forward-only) */
  const int32_t* perm;         /* i32[B]    random_idx of model.py:152 (NULL = identity) */
  const float* gat_keep;       /* f32[L][E+N][heads] GATConv attention-dropout keep mask scaled by 1/(1-p), or NULL;
                                  row e < E = edge_index column e, row E+n = the self loop appended for node n */
  int64_t edge_stride;         /* elements between edge_index[0,0] and edge_index[1,0] */
} cal_batch;

/* Named workspace regions (for tests / debugging / saved activations).
 * EP = max_edges + max_nodes (edge slots after self-loop surgery).  "in-CSR
 * position" p = index into the by-target CSR; every per-edge float array is
 * stored in that order. */
enum cal_ws_region {
  CAL_WS_STATUS = 0,   /* i32[4]: [0] = status bits (CAL_ST_*) */
  CAL_WS_COUNTERS,     /* u32[64 + 3 * 64]: self-resetting grid arrival counters */
  CAL_WS_IN_PTR,       /* i32[maxN+1] CSR by target (edge_index[1]) incl. appended self loops */
  CAL_WS_IN_SRC,       /* i32[EP] source node of in-CSR position p */
  CAL_WS_IN_KEY,       /* i32[EP] edge_index column e (< E) or E + node for the appended loop */
  CAL_WS_IN_NORM,      /* f32[EP] unweighted GCN norm deg^-1/2[row] * deg^-1/2[col] */
  CAL_WS_OUT_PTR,      /* i32[maxN+1] CSR by source (edge_index[0]) */
  CAL_WS_OUT_DST,      /* i32[EP] */
  CAL_WS_OUT_POS,      /* i32[EP] in-CSR position of the same edge */
  CAL_WS_OUT_KEY,      /* i32[EP] */
  CAL_WS_CNT_IN, CAL_WS_CNT_OUT,   /* i32[maxN] scratch */
  CAL_WS_GRAPH_PTR,    /* i32[maxB+1] first node of every graph */
  CAL_WS_NODE_GRAPH,   /* i32[maxN] batch as int32 */
  CAL_WS_PERM,         /* i32[maxB] random_idx (identity when the batch gives none) */
  CAL_WS_INVPERM,      /* i32[maxB] */
  CAL_WS_DIS,          /* f32[maxN] unweighted deg^-1/2 */
  CAL_WS_X,            /* f32[L+1][maxN][H]: x_1 .. x_{L+1} (post-ReLU layer outputs) */
  CAL_WS_NODE_ATT,     /* f32[maxN][2] softmax(node_att_mlp(x)) */
  CAL_WS_PQ,           /* f32[maxN][4] edge_att_mlp split: x W_e[:, :H]^T (2), x W_e[:, H:]^T (2) */
  CAL_WS_EDGE_ATT,     /* f32[EP][2] edge_att by in-CSR position (appended loops hold 1,1) */
  CAL_WS_DISW,         /* f32[maxN][2] weighted deg^-1/2 of the causal / shortcut norm */
  CAL_WS_AGG,          /* f32[2][maxN][H] normalised aggregate entering context/objects weight */
  CAL_WS_Z,            /* f32[2][maxN][H] relu(context_convs), relu(objects_convs) */
  CAL_WS_POOLED,       /* f32[2][maxB][H] global_add_pool of the two */
  CAL_WS_H1,           /* f32[3][maxB][H] relu(fc1_*(bn(.))) of the c / o / co readouts */
  CAL_WS_LOGP,         /* f32[3][maxB][C] the three outputs (log-probabilities) */
  CAL_WS_LOSS,         /* f32[8]: loss, c_loss, o_loss, co_loss, correct_c, correct_o, correct_co, 0 */
  CAL_WS_BN,           /* f32[CAL_MAX_BN+1][6][KMAX]: scale, shift, mean, rstd, c1, c2 per BatchNorm */
  CAL_WS_STATP,        /* f64 scratch of the cross-CTA BatchNorm reductions (hierarchical grid sum) */
  CAL_WS_WT,           /* f32 transposed copies of conv / fc1 weights */
  CAL_WS_GAT,          /* f32 backbone scratch.  CausalGIN: h [L][maxN][H] (nn.0 output, before the inner BatchNorm), then d r [maxN][H];
                          GATConv: per layer x' [maxN][H], a_src / a_dst [maxN][heads], alpha [EP][heads]; then dz [EP][heads], d a_dst [maxN][heads] */
  CAL_WS_DLOGIT,       /* f32[3][maxB][C] */
  CAL_WS_DH,           /* f32[3][maxB][H] */
  CAL_WS_DU,           /* f32[3][maxB][2H] */
  CAL_WS_DAGG,         /* f32[2][maxN][H] */
  CAL_WS_DYM,          /* f32[2][maxN][H] */
  CAL_WS_DNRM,         /* f32[EP][2] */
  CAL_WS_DT,           /* f32[EP][2] */
  CAL_WS_DP,           /* f32[maxN][2] */
  CAL_WS_D,            /* f32[2][maxN][H] ping-pong gradient w.r.t. BatchNorm outputs */
  CAL_WS_GPART,        /* f32 per-CTA partial parameter gradients */
  CAL_WS_OUT_NORM,     /* f32[EP] unweighted norm by out-CSR position (= IN_NORM[OUT_POS]) */
  CAL_WS_EDGE_WN,      /* f32[EP][2] dis_w[source] * edge_att by in-CSR position (weighted norm without the target factor) */
  CAL_WS_EDGE_NA,      /* f32[EP][2] node_att[source] by in-CSR position */
  CAL_WS_FSG,          /* fused small-graph path: block records, all-reduce accumulators and counters, pre-split weight images, per-block partial gradients */
  CAL_WS_EDGE_GPTR,    /* i32[B+1] first edge_index column of every graph | i32[B] self loops per graph | arrival counter (grouped_edges) */
  CAL_WS_REGION_COUNT
};

/* status bits */
#define CAL_ST_BAD_NODE 1    /* edge endpoint outside [0, N) */
#define CAL_ST_BAD_BATCH 2   /* batch not sorted / outside [0, B) */
#define CAL_ST_CAPACITY 4    /* N / E / B exceed the workspace capacities */

/* flags for cal_causal_forward */
#define CAL_F_TRAIN 1        /* BatchNorm batch statistics + running-stat update; GAT dropout */
#define CAL_F_LOSS 2         /* also evaluate train_causal.py:178-186 (needs batch.y) */
#define CAL_F_RAW_LOGITS_O 16 /* cal_causal_forward: the objects head returns its RAW logits (before log_softmax) in place of the
                                log-probabilities -- CausalGIN's train_type="irm", model.py:288-289; cal_causal_backward: grad_logp of
                                that head is the gradient with respect to those raw logits (grad_logp must be given).  FFMA readout
                                kernels only (CAL_EUNSUPPORTED otherwise) */
#define CAL_F_NO_OVERLAP 8   /* cal_causal_forward: the kernel that precedes this call in the stream is not cal_prep's (e.g. the optimizer's:
                                cal_prep ran ahead, elsewhere) -- the first kernel must not overlap its tail */
#define CAL_F_FSG_READY 4    /* cal_causal_forward, training: structure and operand images are complete (see cal_image_sink) */
/* Run only stages [lo, hi] of the pass (see cal_stage_count / cal_stage_name); flags without a
 * range run the whole pass.  Used for per-operator tests and live per-kernel timing. */
#define CAL_F_STAGES(lo, hi) ((((lo) + 1) << 8) | (((hi) + 1) << 16))
#define CAL_PASS_FORWARD 0
#define CAL_PASS_BACKWARD 1

int cal_abi_version(void);
const char* cal_error_string(int code);

/* Number of kernels this library has launched in this process (monotonic; the difference
 * around a call is that call's launch count). */
uint64_t cal_launch_count(void);

/* The operators ("stages") a pass is made of, in execution order, and their names:
 * forward  = param_prep, feat, layer_0 .. layer_{L-1}, edge_att, masked_convs, readout, copy_out
 * backward = readout_bwd, masked_gemm_bwd, masked_gather_bwd, norm_bwd, att_bwd,
 *            layer_{L-1}_bwd .. layer_0_bwd, feat_bwd, grad_reduce */
int cal_stage_count(const cal_model_desc* m, int pass);
const char* cal_stage_name(const cal_model_desc* m, int pass, int stage);

/* Size in bytes of the workspace for a model at given capacities; the offset /
 * size of one named region inside it (returns CAL_EINVAL for unknown ids). */
size_t cal_workspace_bytes(const cal_model_desc* m, const cal_caps* caps);
int cal_workspace_region(const cal_model_desc* m, const cal_caps* caps, int region,
                         size_t* offset_bytes, size_t* size_bytes);

/* Structure preparation, once per batch (replaces the per-layer work of
 * GCNConv.norm, gcn_conv.py:44-70, and the implicit graph segmentation of
 * global_add_pool, model.py:115): int64 -> int32, self-loop removal, one
 * appended self loop per node, stable CSR by target and by source, the
 * unweighted symmetric normalisation, graph_ptr.  Zeroes the status word. */
int cal_prep(const cal_model_desc* m, const cal_caps* caps, const cal_batch* b,
             void* workspace, size_t ws_bytes, void* stream);

/* CausalGCN.forward / CausalGAT.forward (model.py:85-122 / 380-409).  Requires
 * cal_prep on the same workspace.  Outputs land in CAL_WS_LOGP (and are copied
 * to `out_logp` f32[3][B][C] if non-NULL).  With CAL_F_LOSS the loss parts and
 * correct counts land in CAL_WS_LOSS.  With CAL_F_TRAIN the activations needed
 * by cal_causal_backward stay in the workspace. */
int cal_causal_forward(const cal_model_desc* m, const cal_caps* caps, const cal_param_offsets* po,
                       const cal_bn_offsets* bo, const float* params, float* bn_buffers,
                       int64_t* bn_num_batches_tracked, const cal_batch* b, int flags,
                       float* out_logp, void* workspace, size_t ws_bytes, void* stream);

/* Backward of the above (autograd of loss.backward(), train_causal.py:187).
 * `grad_logp` f32[3][B][C] = dL/d(outputs); NULL means "use the loss evaluated
 * by the forward with CAL_F_LOSS".  Writes (not accumulates) every parameter
 * gradient into `grads` (flat, same offsets as params).  `flags`: 0 or CAL_F_STAGES(lo, hi). */
int cal_causal_backward(const cal_model_desc* m, const cal_caps* caps, const cal_param_offsets* po,
                        const float* params, const cal_batch* b, const float* grad_logp,
                        float* grads, int flags, void* workspace, size_t ws_bytes, void* stream);

/* torch.optim.Adam step on a flat buffer (train_causal.py:21,192): m, v moments;
 * `step` is a device int32[2], zero-initialised by the caller: step[0] = number of updates applied so
 * far (this call applies update step[0] + 1 and advances it), step[1] = scratch;
 * `lr_device` (f32[1], device) overrides `lr` when non-NULL so that a captured CUDA graph follows
 * the per-epoch learning-rate schedule (train_causal.py:22,29);
 * grad_scale multiplies the gradient first (1/world_size after all-reduce). */
int cal_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  int64_t n, int32_t* step, float lr, const float* lr_device, float beta1,
                  float beta2, float eps, float weight_decay, float grad_scale, void* stream);
int cal_adam_tick(int32_t* step, void* stream);   /* ++step[0] on the device (manual stepping) */

/* ---- fused small-graph path: operand images as a side product of the optimizer step ------------
 * The fused kernels (cal_caps.small_graphs, csrc/fsg.cu) multiply by hi / lo split, K-major IMAGES of the conv /
 * fc1 weights that a small kernel (k_fsg_prep) rebuilds from the parameters at the head of every forward pass.  A
 * training loop can take that kernel off the critical path: cal_image_sink_init describes where the images of
 * this model live in `workspace`, cal_adam_step_images / cal_dp_adam_step_images (= cal_adam_step /
 * cal_dp_adam_step + `sink`) write the image words of every parameter they update, and the NEXT
 * cal_causal_forward on that workspace may then be called with CAL_F_FSG_READY, which asserts
 *   (1) cal_prep of its batch completed on this workspace before any kernel of the call can start, and
 *   (2) the parameters were last modified by an optimizer step that carried this workspace's sink
 * and starts with the fused forward kernel itself (no k_fsg_prep, no k_param_prep).  sink->count == 0 when the
 * model / capacities do not take the fused path -- CAL_F_FSG_READY is then ignored. */
typedef struct {
  int64_t offset;              /* first float of a [rows][128] row-major matrix in the flat parameter buffer */
  float* dst_t;                /* image with A[m][k] = W[k][m] (or NULL) */
  float* dst_n;                /* image with A[m][k] = W[m][k] (or NULL) */
  int32_t rows, reserved;
} cal_image_entry;
typedef struct {
  int32_t count, reserved;
  cal_image_entry entry[16];
} cal_image_sink;
int cal_image_sink_init(const cal_model_desc* m, const cal_caps* caps, const cal_param_offsets* po,
                        void* workspace, size_t ws_bytes, cal_image_sink* sink);
int cal_adam_step_images(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                         int64_t n, int32_t* step, float lr, const float* lr_device, float beta1,
                         float beta2, float eps, float weight_decay, float grad_scale,
                         const cal_image_sink* sink, void* stream);

/* ---- mini-batch collation on the device ---------------------------------------------------------
 * Replaces the host-side collate of the reference's loader (torch_geometric DataLoader ->
 * Batch.from_data_list, train_causal.py:13-15,171-176): node features concatenated, edge_index
 * offset by the running node count, `batch` = graph id per node, `y` concatenated.  The dataset is
 * uploaded once (all graphs back to back, graph-local node ids in the edge lists). */
typedef struct {
  int32_t num_graphs, num_features;
  const int32_t* node_ptr;     /* i32[G+1] first node of every graph */
  const int32_t* edge_ptr;     /* i32[G+1] first edge_index column of every graph */
  const float* feat;           /* f32[node_ptr[G], F] */
  const int32_t* edge_src;     /* i32[edge_ptr[G]] edge_index[0] of every graph, graph-local ids */
  const int32_t* edge_dst;     /* i32[edge_ptr[G]] edge_index[1] */
  const int64_t* y;            /* i64[G] */
} cal_graph_store;

/* Collate the graphs order[pos[0] .. pos[0] + graphs_per_step) (fewer at the end of `order`) into the
 * buffers `out` points at (they are WRITTEN: dims, perm, feat, edge_index, batch, y; capacities from
 * `caps`; out->edge_stride >= caps->max_edges).  `order` i32[n_order] and `pos` i32[4] (zero-initialised
 * cursor, may be NULL = offset 0) live in device memory; with `advance` != 0 the cursor moves on by
 * the number of graphs taken, so replaying one captured CUDA graph walks through the epoch.
 * `perm_pool` i32[steps][graphs_per_step] (or NULL = identity): random_idx of model.py:147-152 for
 * every step of the epoch, drawn on the host like the reference does.  `acc` f32[8] (or NULL, then
 * `prev_loss` is NULL too): epoch sums -- before the new batch is written, the previous step's
 * CAL_WS_LOSS (`prev_loss`) is added as sum over graphs of loss / c / o / co, the three correct
 * counts and the number of graphs (train_causal.py:186-191); cal_collate_flush adds the last step.
 * pos[3] carries the graph count of the batch collated last (the weight of `prev_loss`), so consecutive calls may
 * write to different `out` buffers -- the next batch can be collated while the current one is still in use. */
int cal_collate(const cal_graph_store* store, const int32_t* order, int32_t n_order, int32_t* pos,
                int32_t graphs_per_step, const int32_t* perm_pool, const cal_caps* caps,
                const cal_batch* out, int advance, const float* prev_loss, float* acc, void* stream);
int cal_collate_flush(const float* prev_loss, const int32_t* dims, int32_t* pos, float* acc, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory ---------------------------------
 * The reference trains in ONE process (train_causal.py:171-192: loss.backward(); optimizer.step());
 * its data-parallel form is DDP's gradient all-reduce between the two.  Here that pair -- sum of the
 * flat gradient buffer over the ranks, then Adam on the mean -- is ONE kernel per rank that pushes
 * its gradient chunks into its peers' exchange regions over NVLink, waits per chunk, sums the copies
 * in rank order (bit-identical replicas) and applies the update (cal_b200/csrc/comm.cu).
 *
 * Set-up, once per process group (one process per GPU, all on one node):
 *   bytes = cal_dp_region_bytes(world, n);  cal_dp_alloc(device, bytes, &mine);
 *   cal_dp_export(mine, handle)  -> send the 64 handle bytes to every peer (any transport);
 *   cal_dp_import(device, peer_handle, &region[q]) for q != rank;  region[rank] = mine.
 * These are the only entry points that allocate (the region must be a whole cudaMalloc allocation
 * to be exportable) and that synchronise.  cal_dp_adam_step itself is asynchronous and CUDA-graph
 * capturable; every rank must issue the same sequence of cal_dp_adam_step calls.  `n` % 4 == 0. */
#define CAL_MAX_WORLD 16
#define CAL_DP_HANDLE_BYTES 64
typedef struct {
  int32_t world, rank;
  void* region[CAL_MAX_WORLD];   /* region[q] = rank q's exchange region as mapped in THIS process */
} cal_dp_comm;
size_t cal_dp_region_bytes(int32_t world, int64_t n);
int cal_dp_alloc(int32_t device, size_t bytes, void** region);
int cal_dp_free(void* region);
int cal_dp_export(void* region, unsigned char handle[CAL_DP_HANDLE_BYTES]);
int cal_dp_import(int32_t device, const unsigned char handle[CAL_DP_HANDLE_BYTES], void** mapped);
int cal_dp_unmap(void* mapped);
/* = cal_adam_step on the rank-ordered mean of the `world` gradient buffers (arguments as there). */
int cal_dp_adam_step(const cal_dp_comm* comm, float* params, const float* grads, float* exp_avg,
                     float* exp_avg_sq, int64_t n, int32_t* step, float lr, const float* lr_device,
                     float beta1, float beta2, float eps, float weight_decay, void* stream);
/* ... + the operand images of the fused small-graph path (cal_image_sink above; NULL = none) */
int cal_dp_adam_step_images(const cal_dp_comm* comm, float* params, const float* grads, float* exp_avg,
                            float* exp_avg_sq, int64_t n, int32_t* step, float lr, const float* lr_device,
                            float beta1, float beta2, float eps, float weight_decay,
                            const cal_image_sink* sink, void* stream);
/* A peer that does not deliver within CAL_DP_TIMEOUT_S seconds (environment, default 600; 0 = wait for ever)
 * makes the kernel record an error word and trap: the rank fails with a launch error instead of
 * applying an update from stale data.  Consecutive calls on one stream are safe (the exchange number is
 * read after the programmatic dependency wait).
 * cal_dp_read_error: 0, or CAL_ETIMEOUT if an exchange gave up waiting for a peer (synchronises the stream;
 * after a trap it returns the CUDA launch failure). */
int cal_dp_read_error(const cal_dp_comm* comm, void* stream);

/* Self-test of the tensor-core building block (tcgen05.mma with the accumulator in TMEM, the operand
 * layouts and precisions the readout kernels use): D[M,N] = A[M,K] * B[N,K]^T on one CTA.
 * kind 0 = 3xTF32 (fp32-accurate split product), 1 = TF32, 2 = BF16; M <= 128, N <= 256; `variant`
 * bit 0 selects the second of the two shared-memory core-matrix arrangements.  Device pointers. */
int cal_selftest_umma(int kind, int M, int N, int K, const float* A, const float* B, float* D, int variant,
                      void* stream);

/* Poll the status words (synchronises the stream): returns the CAL_ST_* bits of the batch prepared last OR-ed with
 * every bit raised since the previous call (a second word that only this call clears: a batch prepared ahead of its
 * step does not hide what the step before it reported), or a negative CAL_E* / positive cudaError_t. */
int cal_read_status(const cal_model_desc* m, const cal_caps* caps, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAL_B200_H */
