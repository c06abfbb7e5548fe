"""CausalGCN / CausalGAT with the reference's constructor and forward signature
(model.py:14-22,85 and model.py:316-320,380 upstream), executed by the sm_100a
kernels of ``libcal_b200.so`` through the C ABI in ``include/cal_b200.h``.

The modules own ordinary ``nn.Parameter``s / BatchNorm buffers under the
reference's ``state_dict`` names (model.py:38-75, gcn_conv.py:30-35), so a
reference checkpoint loads unchanged and the reference training loop
(train_causal.py:162-223: ``model(data)`` -> losses -> ``loss.backward()`` ->
``optimizer.step()``) runs unchanged.  Underneath, all parameters are views of
one flat device buffer (and all gradients of one flat gradient buffer), which is
what the kernels, the NCCL all-reduce and the fused Adam step consume.

There is no CPU path: ``forward`` on a CPU tensor or without the built library
raises."""
from __future__ import annotations

import ctypes as C
import math
import random

import torch
import torch.nn as nn
from torch.nn import BatchNorm1d, Linear, Parameter

from . import _lib
from ._lib import WS

__all__ = ["CausalGCN", "CausalGAT", "CausalGIN", "GCNConv", "GATConv", "GINConv", "Engine"]


def _glorot(t):
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))      # torch_geometric.nn.inits.glorot
    t.data.uniform_(-stdv, stdv)


class GCNConv(nn.Module):
    """Parameter holder with the layout and init of gcn_conv.py:12-42
    (``weight`` [in, out] glorot, ``bias`` [out] zeros).  The arithmetic lives in
    the fused CUDA layer kernels, not here."""

    def __init__(self, in_channels, out_channels, improved=False, cached=False, bias=True,
                 edge_norm=True, gfn=False):
        super().__init__()
        if improved or cached or not edge_norm or not bias:
            raise NotImplementedError("cal_b200 GCNConv: only the configuration CausalGCN/CausalGAT use")
        self.in_channels, self.out_channels, self.gfn = in_channels, out_channels, gfn
        self.weight = Parameter(torch.empty(in_channels, out_channels))
        self.bias = Parameter(torch.empty(out_channels))
        _glorot(self.weight)
        self.bias.data.fill_(0)

    def forward(self, *a, **k):
        raise RuntimeError("cal_b200.GCNConv is executed inside CausalGCN/CausalGAT.forward")


class GATConv(nn.Module):
    """Parameter holder for PyG-1.x ``GATConv(in, out, heads, dropout)`` (model.py:340):
    ``weight`` [in, heads*out], ``att`` [1, heads, 2*out], ``bias`` [heads*out]."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, bias=True):
        super().__init__()
        assert concat and bias and negative_slope == 0.2
        self.in_channels, self.out_channels, self.heads, self.dropout = in_channels, out_channels, heads, dropout
        self.weight = Parameter(torch.empty(in_channels, heads * out_channels))
        self.att = Parameter(torch.empty(1, heads, 2 * out_channels))
        self.bias = Parameter(torch.empty(heads * out_channels))
        _glorot(self.weight)
        _glorot(self.att)
        self.bias.data.fill_(0)

    def forward(self, *a, **k):
        raise RuntimeError("cal_b200.GATConv is executed inside CausalGAT.forward")


class GINConv(nn.Module):
    """Parameter holder for PyG-1.x ``GINConv(nn, eps=0, train_eps=False)`` (model.py:187-193): the
    wrapped ``nn`` (Linear -> BatchNorm1d -> ReLU -> Linear -> ReLU) and the ``eps`` buffer PyG keeps
    in the ``state_dict``.  Executed inside CausalGIN.forward by the fused CUDA layer kernels."""

    def __init__(self, nn_module, eps=0.0, train_eps=False):
        super().__init__()
        if train_eps or float(eps) != 0.0:
            raise NotImplementedError("cal_b200.GINConv: eps = 0, train_eps = False (what CausalGIN uses)")
        self.nn = nn_module
        self.register_buffer("eps", torch.tensor([float(eps)]))

    def forward(self, *a, **k):
        raise RuntimeError("cal_b200.GINConv is executed inside CausalGIN.forward")


# ----------------------------------------------------------------------------------------------
# Engine: flat buffers, workspace, C-ABI calls
# ----------------------------------------------------------------------------------------------

def _round_up(x, m):
    return (x + m - 1) // m * m


def flat_offsets(module):
    """Offsets (in floats) of every parameter of ``module`` inside the flat parameter / gradient
    buffer, in ``named_parameters()`` order, each tensor padded to 16 bytes.  -> (dict, total)."""
    offs, total = {}, 0
    for n, p in module.named_parameters():
        offs[n] = total
        total += _round_up(p.numel(), 4)
    return offs, total


class _Staged:
    """Device-side view of one mini-batch handed to the C ABI."""
    __slots__ = ("N", "E", "B", "cbatch", "keep", "gen", "raw_o")


class Engine:
    """Owns the flat parameter / gradient / BatchNorm buffers and the workspace of one model on
    one device and issues the C-ABI calls."""

    RING = 64

    def __init__(self, module, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CalError("cal_b200 runs on CUDA devices only (no CPU fallback)")
        self.module = module
        m = module
        L, H = len(m.convs), m.hidden
        self.L, self.H, self.C, self.F = L, H, m.num_classes, m.num_features
        self.is_gat = isinstance(m, CausalGAT)
        self.is_gin = isinstance(m, CausalGIN)
        d = _lib.ModelDesc()
        d.model = _lib.CAL_MODEL_GAT if self.is_gat else (_lib.CAL_MODEL_GIN if self.is_gin else _lib.CAL_MODEL_GCN)
        d.num_features, d.hidden, d.num_classes, d.layers = self.F, H, self.C, L
        d.heads = m.head if self.is_gat else 1
        d.cat = int(m.args.cat_or_add == "cat")
        d.without_node_attention = int(bool(m.without_node_attention))
        d.without_edge_attention = int(bool(m.without_edge_attention))
        d.gat_dropout = float(m.dropout) if self.is_gat else 0.0
        d.bn_eps, d.bn_momentum = 1e-5, 0.1
        d.w_c = float(getattr(m.args, "c", 0.5))
        d.w_o = float(getattr(m.args, "o", 1.0))
        d.w_co = float(getattr(m.args, "co", 0.5))
        # not a reference option: bf16 operands (fp32 accumulate) in the readout MLP GEMMs (BASELINE.json configs[4])
        d.readout_bf16 = int(bool(getattr(m.args, "readout_bf16", False)))
        d.readout_tc = int(bool(getattr(m.args, "readout_tc", False)))      # force the tensor-core readout kernels
        self.desc = d
        self._flatten()
        self.caps = None
        self.ws = None
        self.gen = 0
        self._ring = None
        self._ring_i = 0
        self.opt_state = None
        self._owner = None          # weakref to the Trainer whose captured CUDA graphs point into self.ws
        self.fsg_mode = "auto"      # "off": never take the fused small-graph kernels; "fwd": fused forward, tiled backward (A/B tests)

    # ---- flat parameter / buffer storage ----
    def _flatten(self):
        m, dev = self.module, self.device
        named = list(m.named_parameters())
        offs, total = flat_offsets(m)                     # every tensor stays 16-byte aligned
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        for n, p in named:
            v = flat[offs[n]:offs[n] + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            p.grad = None
        self.flat, self.flat_grad = flat, torch.zeros_like(flat)
        self.param_names, self.param_offs, self.total = [n for n, _ in named], offs, total
        self.params = [p for _, p in named]
        po = _lib.ParamOffsets()
        for f, _t in po._fields_:
            v = getattr(po, f)
            if isinstance(v, int):
                setattr(po, f, -1)
            else:
                for i in range(len(v)):
                    v[i] = -1
        g = offs.get
        po.bn_feat_w, po.bn_feat_b = g("bn_feat.weight"), g("bn_feat.bias")
        po.conv_feat_w, po.conv_feat_b = g("conv_feat.weight"), g("conv_feat.bias")
        for i in range(self.L):
            if self.is_gin:                                # convs.i.nn = Linear, BatchNorm1d, ReLU, Linear, ReLU
                po.bns_conv_w[i], po.bns_conv_b[i] = g("convs.%d.nn.1.weight" % i), g("convs.%d.nn.1.bias" % i)
                po.convs_w[i], po.convs_b[i] = g("convs.%d.nn.0.weight" % i), g("convs.%d.nn.0.bias" % i)
                po.gin_w2[i], po.gin_b2[i] = g("convs.%d.nn.3.weight" % i), g("convs.%d.nn.3.bias" % i)
                continue
            po.bns_conv_w[i], po.bns_conv_b[i] = g("bns_conv.%d.weight" % i), g("bns_conv.%d.bias" % i)
            po.convs_w[i], po.convs_b[i] = g("convs.%d.weight" % i), g("convs.%d.bias" % i)
            po.convs_att[i] = g("convs.%d.att" % i, -1)
        po.edge_att_w, po.edge_att_b = g("edge_att_mlp.weight"), g("edge_att_mlp.bias")
        po.node_att_w, po.node_att_b = g("node_att_mlp.weight"), g("node_att_mlp.bias")
        po.bnc_w, po.bnc_b, po.bno_w, po.bno_b = g("bnc.weight"), g("bnc.bias"), g("bno.weight"), g("bno.bias")
        po.context_w, po.context_b = g("context_convs.weight"), g("context_convs.bias")
        po.objects_w, po.objects_b = g("objects_convs.weight"), g("objects_convs.bias")
        for h, t in enumerate(("c", "o", "co")):
            po.fc1_bn_w[h], po.fc1_bn_b[h] = g("fc1_bn_%s.weight" % t), g("fc1_bn_%s.bias" % t)
            po.fc1_w[h], po.fc1_b[h] = g("fc1_%s.weight" % t), g("fc1_%s.bias" % t)
            po.fc2_bn_w[h], po.fc2_bn_b[h] = g("fc2_bn_%s.weight" % t), g("fc2_bn_%s.bias" % t)
            po.fc2_w[h], po.fc2_b[h] = g("fc2_%s.weight" % t), g("fc2_%s.bias" % t)
        po.total = total
        self.po = po
        # BatchNorm buffers, in BN-id order (include/cal_b200.h)
        inner = [conv.nn[1] for conv in m.convs] if self.is_gin else list(m.bns_conv)
        bns = [m.bn_feat] + inner + [m.bnc, m.bno, m.fc1_bn_c, m.fc1_bn_o, m.fc1_bn_co,
                                     m.fc2_bn_c, m.fc2_bn_o, m.fc2_bn_co]
        bo = _lib.BnOffsets()
        tot = 0
        for i, bn in enumerate(bns):
            bo.running_mean[i] = tot
            tot += _round_up(bn.num_features, 4)
            bo.running_var[i] = tot
            tot += _round_up(bn.num_features, 4)
        buf = torch.zeros(tot, dtype=torch.float32, device=dev)
        nbt = torch.zeros(len(bns), dtype=torch.int64, device=dev)
        for i, bn in enumerate(bns):
            k = bn.num_features
            rm = buf[bo.running_mean[i]:bo.running_mean[i] + k]
            rv = buf[bo.running_var[i]:bo.running_var[i] + k]
            rm.copy_(bn.running_mean)
            rv.copy_(bn.running_var)
            nbt[i] = int(bn.num_batches_tracked)
            bn._buffers["running_mean"], bn._buffers["running_var"] = rm, rv
            bn._buffers["num_batches_tracked"] = nbt[i]
        self.bo, self.bn_buf, self.nbt, self.bns = bo, buf, nbt, bns

    def grad_views(self):
        """Per-parameter views of the flat gradient buffer (reference parameter order)."""
        return [self.flat_grad[self.param_offs[n]:self.param_offs[n] + p.numel()].view(p.shape)
                for n, p in zip(self.param_names, self.params)]

    # ---- workspace ----
    def ensure_caps(self, N, E, B):
        c = self.caps
        if c is not None and N <= c.max_nodes and E <= c.max_edges and B <= c.max_graphs:
            return False
        owner = self._owner() if self._owner is not None else None
        if owner is not None and not owner._dead:
            # growing would free the workspace baked into the Trainer's captured CUDA graphs (their
            # replays would then read and write freed memory): the capacities are frozen
            raise _lib.CalError(
                "cal_b200: batch (N=%d, E=%d, B=%d) exceeds the capacities (N<=%d, E<=%d, B<=%d) frozen by the "
                "Trainer that owns this model's workspace -- size the Trainer's caps to cover the evaluation "
                "batches too (batch_caps(train_batches + eval_batches))" % (N, E, B, c.max_nodes, c.max_edges, c.max_graphs))
        grow = lambda need, cur, q: max(cur, _round_up(int(need * 1.25) + 1, q))
        return self.set_caps(grow(N, c.max_nodes if c else 0, 256), grow(E, c.max_edges if c else 0, 256),
                             grow(B, c.max_graphs if c else 0, 32) if c else _round_up(max(B, 1), 32),
                             small_graphs=bool(c.small_graphs) if c else False,
                             grouped_edges=bool(c.grouped_edges) if c else False)

    def set_caps(self, max_nodes, max_edges, max_graphs, small_graphs=False, grouped_edges=False):
        """(Re)allocate the workspace for explicit capacities.  ``small_graphs``: the caller guarantees
        <= 40 nodes and <= 320 CSR entries per graph (cal_caps.small_graphs: the fused small-graph forward);
        ``grouped_edges``: the caller guarantees edge_index columns grouped by graph, <= 512 nodes and <= 4096 columns
        per graph (cal_caps.grouped_edges: per-graph structure preparation).  A Trainer that owned the previous
        workspace is invalidated (its captured CUDA graphs point into freed memory): its next step raises."""
        owner = self._owner() if self._owner is not None else None
        if owner is not None:
            owner._invalidate()
        self._owner = None
        caps = _lib.Caps()
        caps.max_nodes, caps.max_edges, caps.max_graphs = int(max_nodes), int(max_edges), int(max_graphs)
        caps.small_graphs = self._fsg_level(small_graphs) if int(max_graphs) <= self.FSG_GRAPHS else 0
        caps.grouped_edges = 1 if grouped_edges else 0
        nbytes = self.lib.cal_workspace_bytes(C.byref(self.desc), C.byref(caps))
        if nbytes == 0:
            raise _lib.CalError("cal_b200: unsupported model configuration or capacities (hidden must be 32/64/128, "
                                "2 <= classes <= 32, features <= 512; max_graphs=%d may exceed what the readout "
                                "kernels hold in one SM's shared memory)" % caps.max_graphs)
        self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self.caps, self.ws_bytes = caps, nbytes
        self._regions = {}
        ring_elems = 4 + caps.max_graphs
        self._ring = torch.zeros(self.RING, ring_elems, dtype=torch.int32).pin_memory()
        self._meta_dev = torch.zeros(ring_elems, dtype=torch.int32, device=self.device)
        self.out_logp = torch.zeros(3 * caps.max_graphs * self.C, dtype=torch.float32, device=self.device)
        return True

    def region(self, name, dtype=torch.float32):
        """A named workspace region as a flat tensor view (tests / debugging)."""
        v = self._regions.get((name, dtype))
        if v is None:
            off, size = C.c_size_t(), C.c_size_t()
            _lib.check(self.lib.cal_workspace_region(C.byref(self.desc), C.byref(self.caps), WS[name],
                                                     C.byref(off), C.byref(size)), "cal_workspace_region")
            v = self._regions[(name, dtype)] = self.ws[off.value:off.value + size.value].view(dtype)
        return v

    # ---- batches ----
    def stage(self, data, perm=None, gat_keep=None):
        """Bind the tensors of a PyG-style batch (already on the device) for the C ABI."""
        x = data.x if getattr(data, "x", None) is not None else data.feat       # model.py:87
        if x.device != self.device:
            raise _lib.CalError("cal_b200: batch is on %s but the model is on %s (call data.to(device) first, "
                                "as train_causal.py:174 does)" % (x.device, self.device))
        x = x.contiguous().float()
        ei = data.edge_index.contiguous()
        bvec = data.batch.contiguous()
        if ei.dtype != torch.int64 or bvec.dtype != torch.int64:
            ei, bvec = ei.long(), bvec.long()
        y = getattr(data, "y", None)
        if y is not None:
            y = y.view(-1).contiguous().long()
        N, E = int(x.size(0)), int(ei.size(1))
        B = int(getattr(data, "num_graphs", 0) or 0)
        if B <= 0:
            B = int(y.numel()) if y is not None else int(bvec.max().item()) + 1
        if x.size(1) != self.F:
            raise _lib.CalError("cal_b200: batch has %d features, model expects %d" % (x.size(1), self.F))
        self.ensure_caps(N, E, B)
        self._auto_small_graphs(bvec, ei, N, B)
        slot = self._ring[self._ring_i]
        self._ring_i = (self._ring_i + 1) % self.RING
        slot[0], slot[1], slot[2], slot[3] = N, E, B, 0
        if perm is not None:
            slot[4:4 + B] = torch.as_tensor(perm, dtype=torch.int32)
        self._meta_dev.copy_(slot, non_blocking=True)
        cb = _lib.Batch()
        cb.dims = self._meta_dev.data_ptr()
        cb.feat = x.data_ptr()
        cb.edge_index = ei.data_ptr() if E > 0 else bvec.data_ptr()     # never read when E == 0, but not NULL
        cb.edge_stride = E
        cb.batch = bvec.data_ptr()
        cb.y = y.data_ptr() if y is not None else 0
        cb.perm = self._meta_dev.data_ptr() + 16 if perm is not None else 0
        cb.gat_keep = gat_keep.data_ptr() if gat_keep is not None else 0
        st = _Staged()
        st.raw_o = False
        st.N, st.E, st.B, st.cbatch = N, E, B, cb
        st.keep = (x, ei, bvec, y, gat_keep)
        return st

    FSG_ROWS, FSG_ENTRIES, FSG_GRAPHS = 40, 320, 148     # csrc/fsg.cuh

    def _auto_small_graphs(self, bvec, ei, N, B):
        """The module path (``model(data)``) decides per batch whether the fused small-graph forward applies
        (one device reduction + sync); a Trainer declares it once for its whole dataset instead."""
        owner = self._owner() if self._owner is not None else None
        if owner is not None and not owner._dead:
            return                                         # frozen by the Trainer's declaration
        small = False
        if (self.fsg_mode != "off" and not self.is_gat and not self.is_gin and self.H == 128 and self.F <= 128 and 0 < B <= self.FSG_GRAPHS
                and self.caps.max_graphs <= self.FSG_GRAPHS and N > 0):
            nodes = torch.bincount(bvec, minlength=B)
            ents = nodes.clone()
            if ei.numel() > 0:
                ents += torch.bincount(bvec[ei[0]], minlength=B)
            small = bool((nodes.max() <= self.FSG_ROWS) & (ents.max() <= self.FSG_ENTRIES))
        self.caps.small_graphs = self._fsg_level(small)
        # batches beyond the single-kernel structure path (csrc/prep.cu launch_prep): per-graph preparation when the
        # edge_index columns are grouped by graph (PyG collate) and every graph fits a CTA's shared memory
        c = self.caps
        beyond = (4 * (5 * c.max_nodes + (c.max_edges + c.max_nodes) + c.max_edges + 8) > 225 * 1024
                  or c.max_edges + c.max_nodes >= 65535)
        grouped = False
        if beyond and N > 0 and B > 0 and ei.numel() > 0:
            gs = bvec[ei[0]]
            grouped = bool((gs[1:] >= gs[:-1]).all() & (gs == bvec[ei[1]]).all()
                           & (torch.bincount(bvec, minlength=B).max() <= self.PG_NODES)
                           & (torch.bincount(gs, minlength=B).max() <= self.PG_COLS))
        c.grouped_edges = 1 if grouped else 0

    PG_NODES, PG_COLS = 512, 4096                        # csrc/prep.cu k_prep_graph

    def _fsg_level(self, small):
        """cal_caps.small_graphs: 0 = tiled kernels, 1 = fused small-graph forward and backward, 2 = fused forward only."""
        if not small or self.fsg_mode == "off":
            return 0
        if self.is_gat or self.is_gin or self.H != 128 or self.F > 128:
            return 0                                       # (the C side would ignore the promise: keep the flag truthful)
        return 2 if self.fsg_mode == "fwd" else 1

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def prep(self, st):
        _lib.check(self.lib.cal_prep(C.byref(self.desc), C.byref(self.caps), C.byref(st.cbatch),
                                     self.ws.data_ptr(), self.ws_bytes, self._stream()), "cal_prep")

    def forward(self, st, train, with_loss=False, copy_out=True, stages=None, raw_o=False):
        flags = (_lib.CAL_F_TRAIN if train else 0) | (_lib.CAL_F_LOSS if with_loss else 0)
        if raw_o:                                          # objects head: raw logits instead of log-probabilities
            flags |= _lib.CAL_F_RAW_LOGITS_O
        if stages is not None:
            flags |= _lib.stages_flag(*stages)
        self.gen += 1
        st.gen = self.gen
        _lib.check(self.lib.cal_causal_forward(
            C.byref(self.desc), C.byref(self.caps), C.byref(self.po), C.byref(self.bo), self.flat.data_ptr(),
            self.bn_buf.data_ptr(), self.nbt.data_ptr(), C.byref(st.cbatch), flags,
            self.out_logp.data_ptr() if copy_out else 0, self.ws.data_ptr(), self.ws_bytes, self._stream()),
            "cal_causal_forward")
        if copy_out:
            return self.out_logp[:3 * st.B * self.C].view(3, st.B, self.C)
        return None

    def backward(self, st, grad_logp=None, stages=None, raw_o=False):
        if st.gen != self.gen:
            raise _lib.CalError("cal_b200: backward through a stale forward (the workspace holds the "
                                "activations of the most recent forward only)")
        gp = 0
        if grad_logp is not None:
            grad_logp = grad_logp.contiguous().float()
            gp = grad_logp.data_ptr()
        _lib.check(self.lib.cal_causal_backward(
            C.byref(self.desc), C.byref(self.caps), C.byref(self.po), self.flat.data_ptr(), C.byref(st.cbatch),
            gp, self.flat_grad.data_ptr(),
            (_lib.stages_flag(*stages) if stages is not None else 0) | (_lib.CAL_F_RAW_LOGITS_O if raw_o else 0),
            self.ws.data_ptr(), self.ws_bytes, self._stream()),
            "cal_causal_backward")

    def status(self):
        rc = self.lib.cal_read_status(C.byref(self.desc), C.byref(self.caps), self.ws.data_ptr(), self._stream())
        if rc < 0 or rc > 7:
            _lib.check(rc, "cal_read_status")
        return rc

    def loss_parts(self):
        """f32[7] device view: loss, c_loss, o_loss, co_loss, correct_c, correct_o, correct_co."""
        return self.region("LOSS")[:7]

    def loss_parts_full(self):
        return self.region("LOSS")[:8]

    # ---- fused Adam on the flat buffers (train_causal.py:21,192) ----
    # ``images_version``: the parameter version signature at the last optimizer step that also wrote the fused path's
    # operand images (None: the images in the workspace may be stale).  Our kernels do not move torch's version
    # counters, any torch-side in-place modification of a parameter (optimizers, load_state_dict, p.mul_()) or of the
    # flat buffer does -- so "unchanged" means the images still match the parameters.  The one hole is the one autograd
    # has too: writes through ``p.data`` / ``p.detach()`` aliases are invisible; after such a write call
    # ``images_stale()`` (or Trainer.params_changed()).
    images_version = None

    def param_version(self):
        v = self.flat._version
        for p in self.params:
            v += p._version
        return v

    def images_fresh(self):
        return self.images_version is not None and self.images_version == self.param_version()

    def images_stale(self):
        self.images_version = None

    def image_sink(self):
        """cal_image_sink of the current workspace (count == 0 unless the fused small-graph path is taken)."""
        key = (self.ws.data_ptr(), int(self.caps.small_graphs))
        if getattr(self, "_sink_key", None) != key:
            sk = _lib.ImageSink()
            _lib.check(self.lib.cal_image_sink_init(C.byref(self.desc), C.byref(self.caps), C.byref(self.po), self.ws.data_ptr(),
                                                    self.ws_bytes, C.byref(sk)), "cal_image_sink_init")
            self._sink, self._sink_key = sk, key
        return self._sink

    def adam_step(self, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0, lr_device=None, sink=None):
        """``lr_device`` (f32[1] device tensor) overrides ``lr`` so captured graphs follow a schedule.
        ``sink``: cal_image_sink -- also write the fused path's operand images of the updated parameters."""
        if self.opt_state is None:
            self.opt_state = (torch.zeros_like(self.flat), torch.zeros_like(self.flat),
                              torch.zeros(2, dtype=torch.int32, device=self.device))
        m, v, step = self.opt_state
        s = self._stream()
        _lib.check(self.lib.cal_adam_step_images(self.flat.data_ptr(), self.flat_grad.data_ptr(), m.data_ptr(),
                                                 v.data_ptr(), self.total, step.data_ptr(), float(lr),
                                                 lr_device.data_ptr() if lr_device is not None else 0,
                                                 betas[0], betas[1], eps, weight_decay, grad_scale,
                                                 C.byref(sink) if sink is not None else None, s), "cal_adam_step")
        self.images_version = self.param_version() if sink is not None else None

    def stage_names(self, backward=False):
        p = _lib.CAL_PASS_BACKWARD if backward else _lib.CAL_PASS_FORWARD
        n = self.lib.cal_stage_count(C.byref(self.desc), p)
        return [self.lib.cal_stage_name(C.byref(self.desc), p, i).decode() for i in range(n)]


class _CausalFn(torch.autograd.Function):
    """Autograd bridge for the reference-style loop (loss computed by torch from the three
    log-probability outputs, then ``loss.backward()``, train_causal.py:178-187)."""

    @staticmethod
    def forward(ctx, eng, st, *params):
        raw_o = bool(getattr(st, "raw_o", False))           # objects head returns raw logits (CausalGIN "irm")
        out = eng.forward(st, train=True, raw_o=raw_o)
        ctx.eng, ctx.st, ctx.raw_o = eng, st, raw_o
        return out.clone()

    @staticmethod
    def backward(ctx, gout):
        eng = ctx.eng
        eng.backward(ctx.st, gout, raw_o=ctx.raw_o)
        g = eng.flat_grad.clone()
        outs = []
        for n, p in zip(eng.param_names, eng.params):
            if n == "conv_feat.bias":            # gfn=True: the bias is never used (gcn_conv.py:76-77)
                outs.append(None)
            else:
                o = eng.param_offs[n]
                outs.append(g[o:o + p.numel()].view(p.shape))
        return (None, None) + tuple(outs)


# ----------------------------------------------------------------------------------------------
# The modules
# ----------------------------------------------------------------------------------------------

class _CausalBase(nn.Module):
    def _build_tail(self, hidden, num_classes):
        """model.py:47-83 / 342-378, same construction order (=> same RNG draws)."""
        self.edge_att_mlp = nn.Linear(hidden * 2, 2)
        self.node_att_mlp = nn.Linear(hidden, 2)
        self.bnc = BatchNorm1d(hidden)
        self.bno = BatchNorm1d(hidden)
        self.context_convs = GCNConv(hidden, hidden)
        self.objects_convs = GCNConv(hidden, hidden)
        self.fc1_bn_c = BatchNorm1d(hidden)
        self.fc1_c = Linear(hidden, hidden)
        self.fc2_bn_c = BatchNorm1d(hidden)
        self.fc2_c = Linear(hidden, num_classes)
        self.fc1_bn_o = BatchNorm1d(hidden)
        self.fc1_o = Linear(hidden, hidden)
        self.fc2_bn_o = BatchNorm1d(hidden)
        self.fc2_o = Linear(hidden, num_classes)
        if self.args.cat_or_add == "cat":
            self.fc1_bn_co = BatchNorm1d(hidden * 2)
            self.fc1_co = Linear(hidden * 2, hidden)
        elif self.args.cat_or_add == "add":
            self.fc1_bn_co = BatchNorm1d(hidden)
            self.fc1_co = Linear(hidden, hidden)
        else:
            assert False                                  # model.py:77
        self.fc2_bn_co = BatchNorm1d(hidden)
        self.fc2_co = Linear(hidden, num_classes)
        for m in self.modules():                          # model.py:80-83
            if isinstance(m, BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0.0001)
        self._engine = None

    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float() replace the parameter storages: the flat views are rebuilt lazily
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        if self._engine is not None:       # num_batches_tracked scalars are copied in place; nothing to do
            pass
        return r

    @property
    def engine(self):
        if self._engine is None:
            p = next(self.parameters())
            self._engine = Engine(self, p.device)
        return self._engine

    def _shuffles(self, eval_random):
        raise NotImplementedError

    def _perm(self, num, eval_random):
        """random_idx of model.py:147-152 / 433-438: Python's RNG, one shuffle per forward."""
        if not self._shuffles(eval_random):
            return None
        l = [i for i in range(num)]
        random.shuffle(l)
        return l

    def _gat_keep(self, st_N, st_E):
        return None

    def forward(self, data, eval_random=True, perm=None, raw_o=False):
        """-> (xc_logis, xo_logis, xco_logis), three [B, C] log-probability tensors
        (model.py:85-122 / 380-409).  ``raw_o``: the objects head's RAW logits in place of xo_logis."""
        eng = self.engine
        x = data.x if getattr(data, "x", None) is not None else data.feat
        B = int(getattr(data, "num_graphs", 0) or 0) or int(data.y.numel())
        if perm is None:
            perm = self._perm(B, eval_random)
        keep = self._gat_keep(int(x.size(0)), int(data.edge_index.size(1))) if self.training else None
        st = eng.stage(data, perm=perm, gat_keep=keep)
        st.raw_o = bool(raw_o)
        eng.prep(st)
        if self.training and torch.is_grad_enabled():
            out = _CausalFn.apply(eng, st, *eng.params)
        else:
            out = eng.forward(st, train=self.training, raw_o=bool(raw_o)).clone()
        return out[0], out[1], out[2]


class CausalGCN(_CausalBase):
    """Drop-in for the reference ``CausalGCN`` (model.py:12-164)."""

    def __init__(self, num_features, num_classes, args, gfn=False, collapse=False, residual=False,
                 res_branch="BNConvReLU", global_pool="sum", dropout=0, edge_norm=True):
        super().__init__()
        if gfn or not edge_norm:
            raise NotImplementedError("cal_b200.CausalGCN: gfn=False, edge_norm=True only (the reference defaults)")
        hidden = args.hidden
        assert global_pool == "sum"                        # model.py:26
        self.args = args
        self.hidden, self.num_features = hidden, num_features
        self.global_pool = global_pool
        self.dropout = dropout
        self.with_random = args.with_random
        self.without_node_attention = args.without_node_attention
        self.without_edge_attention = args.without_edge_attention
        self.num_classes = num_classes
        self.fc_num = getattr(args, "fc_num", "222")
        self.bn_feat = BatchNorm1d(num_features)
        self.conv_feat = GCNConv(num_features, hidden, gfn=True)
        self.bns_conv = nn.ModuleList()
        self.convs = nn.ModuleList()
        for _ in range(args.layers):
            self.bns_conv.append(BatchNorm1d(hidden))
            self.convs.append(GCNConv(hidden, hidden))
        self._build_tail(hidden, num_classes)

    def _shuffles(self, eval_random):                      # model.py:149-151
        return bool(self.with_random and eval_random)


class CausalGIN(_CausalBase):
    """Drop-in for the reference ``CausalGIN`` (model.py:166-313): same constructor, same
    ``state_dict`` keys (``convs.i.nn.{0,1,3}.*``, ``convs.i.eps``), ``forward(data, eval_random=True,
    train_type="base")``."""

    def __init__(self, num_features, num_classes, args, gfn=False, edge_norm=True):
        super().__init__()
        if gfn or not edge_norm:
            raise NotImplementedError("cal_b200.CausalGIN: gfn=False, edge_norm=True only (the reference defaults)")
        hidden = args.hidden
        self.args = args
        self.hidden, self.num_features = hidden, num_features
        self.dropout = 0.0
        self.without_node_attention = False
        self.without_edge_attention = False
        self.num_classes = num_classes
        self.fc_num = getattr(args, "fc_num", "222")
        self.bn_feat = BatchNorm1d(num_features)
        self.conv_feat = GCNConv(num_features, hidden, gfn=True)
        self.bns_conv = nn.ModuleList()                    # stays empty (model.py:185)
        self.convs = nn.ModuleList()
        for _ in range(args.layers):
            self.convs.append(GINConv(nn.Sequential(Linear(hidden, hidden), BatchNorm1d(hidden), nn.ReLU(),
                                                    Linear(hidden, hidden), nn.ReLU())))
        self._build_tail(hidden, num_classes)

    def _shuffles(self, eval_random):                      # model.py:296-297
        return bool(eval_random)

    def forward(self, data, eval_random=True, train_type="base", perm=None):
        """model.py:234-270.  ``train_type="irm"``: the objects head also returns its raw logits -- the tuple
        ``(xc_logis, (xo, xo_logis), xco_logis)`` of model.py:281-292.  The kernels hand back the raw logits of that head;
        its log_softmax (and, in the backward pass, the sum of both gradients) is a [B, C] torch op."""
        if train_type != "irm":
            return super().forward(data, eval_random=eval_random, perm=perm)
        xc_logis, xo, xco_logis = super().forward(data, eval_random=eval_random, perm=perm, raw_o=True)
        return xc_logis, (xo, torch.log_softmax(xo, dim=-1)), xco_logis


class CausalGAT(_CausalBase):
    """Drop-in for the reference ``CausalGAT`` (model.py:315-450)."""

    def __init__(self, num_features, num_classes, args, head=4, dropout=0.2):
        super().__init__()
        hidden = args.hidden
        self.args = args
        self.hidden, self.num_features = hidden, num_features
        self.head = head
        self.dropout = dropout
        self.without_node_attention = False
        self.without_edge_attention = False
        self.num_classes = num_classes
        self.fc_num = getattr(args, "fc_num", "222")
        self.bn_feat = BatchNorm1d(num_features)
        self.conv_feat = GCNConv(num_features, hidden, gfn=True)
        self.bns_conv = nn.ModuleList()
        self.convs = nn.ModuleList()
        for _ in range(args.layers):
            self.bns_conv.append(BatchNorm1d(hidden))
            self.convs.append(GATConv(hidden, int(hidden / head), heads=head, dropout=dropout))
        self._build_tail(hidden, num_classes)
        self.dropout_mask = None          # optional injected keep-mask [L, E+N, heads] (tests)

    def _shuffles(self, eval_random):                      # model.py:435
        return bool(eval_random)

    def _gat_keep(self, N, E):
        """Scaled keep-mask of GATConv's attention dropout (F.dropout(alpha, p), PyG 1.x)."""
        p = float(self.dropout)
        if self.dropout_mask is not None:
            return (self.dropout_mask.to(self.engine.device).float() / (1.0 - p)).contiguous()
        if p <= 0.0:
            return None
        dev = next(self.parameters()).device
        keep = (torch.rand(len(self.convs), E + N, self.head, device=dev) >= p).float() / (1.0 - p)
        return keep
