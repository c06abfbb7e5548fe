"""ctypes binding of ``libcal_b200.so`` (the C ABI declared in ``include/cal_b200.h``).

There is no CPU fallback: if the CUDA library has not been built the import of
this module raises, and so does every product entry point."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CAL_B200_LIB: an alternative build of the same library (instrumented -DCAL_PHASE_TIMING builds)
LIB_PATH = os.environ.get("CAL_B200_LIB") or os.path.join(HERE, "libcal_b200.so")

CAL_MAX_LAYERS = 8
CAL_MAX_BN = 1 + CAL_MAX_LAYERS + 2 + 6
CAL_MODEL_GCN, CAL_MODEL_GAT, CAL_MODEL_GIN = 0, 1, 2
CAL_F_TRAIN, CAL_F_LOSS, CAL_F_FSG_READY, CAL_F_NO_OVERLAP, CAL_F_RAW_LOGITS_O = 1, 2, 4, 8, 16
CAL_ST_BAD_NODE, CAL_ST_BAD_BATCH, CAL_ST_CAPACITY = 1, 2, 4

# enum cal_ws_region, in header order
WS_REGIONS = [
    "STATUS", "COUNTERS", "IN_PTR", "IN_SRC", "IN_KEY", "IN_NORM", "OUT_PTR", "OUT_DST", "OUT_POS",
    "OUT_KEY", "CNT_IN", "CNT_OUT", "GRAPH_PTR", "NODE_GRAPH", "PERM", "INVPERM", "DIS", "X",
    "NODE_ATT", "PQ", "EDGE_ATT", "DISW", "AGG", "Z", "POOLED", "H1", "LOGP", "LOSS", "BN", "STATP",
    "WT", "GAT", "DLOGIT", "DH", "DU", "DAGG", "DYM", "DNRM", "DT", "DP", "D", "GPART",
    "OUT_NORM", "EDGE_WN", "EDGE_NA", "FSG", "EDGE_GPTR",
]
WS = {n: i for i, n in enumerate(WS_REGIONS)}

EXPORTS = [
    "cal_abi_version", "cal_error_string", "cal_workspace_bytes", "cal_workspace_region", "cal_prep",
    "cal_causal_forward", "cal_causal_backward", "cal_adam_step", "cal_adam_tick", "cal_read_status",
    "cal_launch_count", "cal_stage_count", "cal_stage_name",
    "cal_dp_region_bytes", "cal_dp_alloc", "cal_dp_free", "cal_dp_export", "cal_dp_import", "cal_dp_unmap",
    "cal_dp_adam_step", "cal_dp_read_error", "cal_collate", "cal_collate_flush", "cal_selftest_umma",
    "cal_image_sink_init", "cal_adam_step_images", "cal_dp_adam_step_images",
]
CAL_MAX_WORLD, CAL_DP_HANDLE_BYTES = 16, 64
CAL_PASS_FORWARD, CAL_PASS_BACKWARD = 0, 1


def stages_flag(lo, hi):
    """CAL_F_STAGES(lo, hi) of include/cal_b200.h."""
    return ((lo + 1) << 8) | ((hi + 1) << 16)


class ModelDesc(C.Structure):
    _fields_ = [("model", C.c_int32), ("num_features", C.c_int32), ("hidden", C.c_int32),
                ("num_classes", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("cat", C.c_int32), ("without_node_attention", C.c_int32),
                ("without_edge_attention", C.c_int32), ("gat_dropout", C.c_float),
                ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
                ("w_c", C.c_float), ("w_o", C.c_float), ("w_co", C.c_float), ("readout_bf16", C.c_int32), ("readout_tc", C.c_int32)]


class Caps(C.Structure):
    _fields_ = [("max_nodes", C.c_int32), ("max_edges", C.c_int32), ("max_graphs", C.c_int32),
                ("small_graphs", C.c_int32), ("grouped_edges", C.c_int32)]


class ParamOffsets(C.Structure):
    _fields_ = [("bn_feat_w", C.c_int64), ("bn_feat_b", C.c_int64),
                ("conv_feat_w", C.c_int64), ("conv_feat_b", C.c_int64),
                ("bns_conv_w", C.c_int64 * CAL_MAX_LAYERS), ("bns_conv_b", C.c_int64 * CAL_MAX_LAYERS),
                ("convs_w", C.c_int64 * CAL_MAX_LAYERS), ("convs_b", C.c_int64 * CAL_MAX_LAYERS),
                ("convs_att", C.c_int64 * CAL_MAX_LAYERS),
                ("edge_att_w", C.c_int64), ("edge_att_b", C.c_int64),
                ("node_att_w", C.c_int64), ("node_att_b", C.c_int64),
                ("bnc_w", C.c_int64), ("bnc_b", C.c_int64), ("bno_w", C.c_int64), ("bno_b", C.c_int64),
                ("context_w", C.c_int64), ("context_b", C.c_int64),
                ("objects_w", C.c_int64), ("objects_b", C.c_int64),
                ("fc1_bn_w", C.c_int64 * 3), ("fc1_bn_b", C.c_int64 * 3),
                ("fc1_w", C.c_int64 * 3), ("fc1_b", C.c_int64 * 3),
                ("fc2_bn_w", C.c_int64 * 3), ("fc2_bn_b", C.c_int64 * 3),
                ("fc2_w", C.c_int64 * 3), ("fc2_b", C.c_int64 * 3),
                ("gin_w2", C.c_int64 * CAL_MAX_LAYERS), ("gin_b2", C.c_int64 * CAL_MAX_LAYERS),
                ("total", C.c_int64)]


class BnOffsets(C.Structure):
    _fields_ = [("running_mean", C.c_int64 * CAL_MAX_BN), ("running_var", C.c_int64 * CAL_MAX_BN)]


class Batch(C.Structure):
    _fields_ = [("dims", C.c_void_p), ("feat", C.c_void_p), ("edge_index", C.c_void_p),
                ("batch", C.c_void_p), ("y", C.c_void_p), ("perm", C.c_void_p),
                ("gat_keep", C.c_void_p), ("edge_stride", C.c_int64)]


class GraphStoreDesc(C.Structure):
    _fields_ = [("num_graphs", C.c_int32), ("num_features", C.c_int32), ("node_ptr", C.c_void_p),
                ("edge_ptr", C.c_void_p), ("feat", C.c_void_p), ("edge_src", C.c_void_p),
                ("edge_dst", C.c_void_p), ("y", C.c_void_p)]


class DpComm(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("region", C.c_void_p * CAL_MAX_WORLD)]


class ImageEntry(C.Structure):
    _fields_ = [("offset", C.c_int64), ("dst_t", C.c_void_p), ("dst_n", C.c_void_p), ("rows", C.c_int32),
                ("reserved", C.c_int32)]


class ImageSink(C.Structure):
    """cal_image_sink: where an optimizer step writes the fused small-graph path's operand images."""
    _fields_ = [("count", C.c_int32), ("reserved", C.c_int32), ("entry", ImageEntry * 16)]


class CalError(RuntimeError):
    pass


_lib = None


def load():
    """Load libcal_b200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CalError(
            "cal_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C cal_b200/csrc`; there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.cal_abi_version.restype = C.c_int
    lib.cal_error_string.restype = C.c_char_p
    lib.cal_error_string.argtypes = [C.c_int]
    lib.cal_workspace_bytes.restype = C.c_size_t
    lib.cal_workspace_bytes.argtypes = [P(ModelDesc), P(Caps)]
    lib.cal_workspace_region.restype = C.c_int
    lib.cal_workspace_region.argtypes = [P(ModelDesc), P(Caps), C.c_int, P(C.c_size_t), P(C.c_size_t)]
    lib.cal_prep.restype = C.c_int
    lib.cal_prep.argtypes = [P(ModelDesc), P(Caps), P(Batch), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cal_causal_forward.restype = C.c_int
    lib.cal_causal_forward.argtypes = [P(ModelDesc), P(Caps), P(ParamOffsets), P(BnOffsets), C.c_void_p,
                                       C.c_void_p, C.c_void_p, P(Batch), C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cal_causal_backward.restype = C.c_int
    lib.cal_causal_backward.argtypes = [P(ModelDesc), P(Caps), P(ParamOffsets), C.c_void_p, P(Batch),
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.cal_adam_step.restype = C.c_int
    lib.cal_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_float, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                  C.c_void_p]
    lib.cal_adam_step_images.restype = C.c_int
    lib.cal_adam_step_images.argtypes = lib.cal_adam_step.argtypes[:-1] + [P(ImageSink), C.c_void_p]
    lib.cal_image_sink_init.restype = C.c_int
    lib.cal_image_sink_init.argtypes = [P(ModelDesc), P(Caps), P(ParamOffsets), C.c_void_p, C.c_size_t, P(ImageSink)]
    lib.cal_launch_count.restype = C.c_uint64
    lib.cal_stage_count.restype = C.c_int
    lib.cal_stage_count.argtypes = [P(ModelDesc), C.c_int]
    lib.cal_stage_name.restype = C.c_char_p
    lib.cal_stage_name.argtypes = [P(ModelDesc), C.c_int, C.c_int]
    lib.cal_adam_tick.restype = C.c_int
    lib.cal_adam_tick.argtypes = [C.c_void_p, C.c_void_p]
    lib.cal_read_status.restype = C.c_int
    lib.cal_read_status.argtypes = [P(ModelDesc), P(Caps), C.c_void_p, C.c_void_p]
    lib.cal_collate.restype = C.c_int
    lib.cal_collate.argtypes = [P(GraphStoreDesc), C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                P(Caps), P(Batch), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cal_collate_flush.restype = C.c_int
    lib.cal_collate_flush.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cal_dp_region_bytes.restype = C.c_size_t
    lib.cal_dp_region_bytes.argtypes = [C.c_int32, C.c_int64]
    lib.cal_dp_alloc.restype = C.c_int
    lib.cal_dp_alloc.argtypes = [C.c_int32, C.c_size_t, P(C.c_void_p)]
    lib.cal_dp_free.restype = C.c_int
    lib.cal_dp_free.argtypes = [C.c_void_p]
    lib.cal_dp_export.restype = C.c_int
    lib.cal_dp_export.argtypes = [C.c_void_p, C.c_char_p]
    lib.cal_dp_import.restype = C.c_int
    lib.cal_dp_import.argtypes = [C.c_int32, C.c_char_p, P(C.c_void_p)]
    lib.cal_dp_unmap.restype = C.c_int
    lib.cal_dp_unmap.argtypes = [C.c_void_p]
    lib.cal_dp_adam_step.restype = C.c_int
    lib.cal_dp_adam_step.argtypes = [P(DpComm), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_void_p]
    lib.cal_dp_adam_step_images.restype = C.c_int
    lib.cal_dp_adam_step_images.argtypes = lib.cal_dp_adam_step.argtypes[:-1] + [P(ImageSink), C.c_void_p]
    lib.cal_dp_read_error.restype = C.c_int
    lib.cal_dp_read_error.argtypes = [P(DpComm), C.c_void_p]
    lib.cal_selftest_umma.restype = C.c_int
    lib.cal_selftest_umma.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p]
    if lib.cal_abi_version() != 3:
        raise CalError("cal_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        raise CalError("%s failed: %s (code %d)" % (what, load().cal_error_string(rc).decode(), rc))
