"""The training step of the reference (train_causal.py:171-192: forward -> KL + 2 NLL loss ->
backward -> Adam) as one asynchronous, CUDA-graph-replayed sequence of C-ABI calls, with a
data-parallel gradient all-reduce between backward and the optimizer.

``Trainer.step`` never synchronises: loss parts and correct counts stay on the device
(``Trainer.metrics``) and are read once per epoch -- or per step through ``step_host``, which is the
end-to-end path (one pinned-host packed batch in, H2D copy, step, loss parts back to the host).

Batches travel as ONE packed buffer (``PackedLayout``): ``dims | perm | feat | edge_index | batch |
y`` at capacity-fixed offsets, so a single ``cudaMemcpyAsync`` moves a batch and a captured graph
whose kernels point into a staging buffer serves every batch.  Index tensors stay int64 at this
boundary because that is what the reference's DataLoader hands over (train_causal.py:174-176)."""
from __future__ import annotations

import ctypes as C
import random
import weakref

import numpy as np
import torch

from . import _lib

__all__ = ["PackedLayout", "Trainer", "batch_caps", "allreduce_flat_grads", "PeerExchange", "GraphStore",
           "epoch_order"]


def _up(x, m):
    return (x + m - 1) // m * m


FSG_ROWS, FSG_ENTRIES = 40, 320          # csrc/fsg.cuh: the fused small-graph path's per-graph limits
PG_NODES, PG_COLS = 512, 4096            # csrc/prep.cu k_prep_graph: per-graph limits of cal_caps.grouped_edges


def batch_caps(batches, slack=1.0):
    """Capacities covering ``batches``: (max nodes, max edge_index columns, max graphs, small_graphs, grouped_edges) --
    ``small_graphs`` is True when every graph has <= 40 nodes and <= 320 CSR entries (edges + one self loop
    per node), which lets CausalGCN run the fused small-graph forward (cal_caps.small_graphs);
    ``grouped_edges`` is True when the edge_index columns of every batch are grouped by graph in graph order with both
    endpoints inside the graph (what PyG's collate produces) and no graph exceeds 512 nodes / 4096 columns
    (cal_caps.grouped_edges: per-graph structure preparation for batches beyond the single-kernel path)."""
    n = max(int(b.batch.numel()) for b in batches)
    e = max(int(b.edge_index.size(1)) for b in batches)
    g = max(int(b.num_graphs) for b in batches)
    small, grouped = True, True
    for b in batches:
        bv, ei = b.batch.numpy(), b.edge_index.numpy()
        nodes = np.bincount(bv, minlength=int(b.num_graphs))
        gsrc = bv[ei[0]]
        cols = np.bincount(gsrc, minlength=int(b.num_graphs))
        if nodes.max(initial=0) > FSG_ROWS or (nodes + cols).max(initial=0) > FSG_ENTRIES:
            small = False
        if (nodes.max(initial=0) > PG_NODES or cols.max(initial=0) > PG_COLS or np.any(np.diff(gsrc) < 0)
                or np.any(gsrc != bv[ei[1]])):
            grouped = False
    return _up(int(n * slack), 32), _up(max(int(e * slack), 1), 32), _up(g, 8), small, grouped


def epoch_order(num_graphs, epoch, seed=0, rank=0, world_size=1, graphs_per_step=None):
    """This rank's graph order for one epoch (input of ``Trainer.begin_epoch``): the shuffle of the
    reference's ``DataLoader(train_dataset, batch_size, shuffle=True)`` (train_causal.py:13-15), sharded
    DistributedSampler-style -- every rank draws the SAME permutation from (seed, epoch) and takes
    ``perm[rank::world_size]``.  With ``graphs_per_step`` the shard is cut to a
    whole number of steps common to all ranks, so every rank issues the same number of gradient
    exchanges (a peer exchange / all-reduce with a missing rank would wait forever).  NB: this drops the
    partial tail batch of every epoch (like ``DataLoader(drop_last=True)``), unlike the reference's
    single-process loader, which trains on the short last batch too."""
    perm = np.random.RandomState((int(seed) * 1000003 + int(epoch)) % (2 ** 31 - 1)).permutation(int(num_graphs))
    shard = perm[int(rank)::int(world_size)]
    if graphs_per_step:
        per_rank = int(num_graphs) // int(world_size)                 # the shortest shard
        shard = shard[:per_rank // int(graphs_per_step) * int(graphs_per_step)]
    return shard.astype(np.int32)


def cosine_lr(epoch, base_lr, min_lr=0.0, epochs=100):
    """Learning rate of ``CosineAnnealingLR(optimizer, T_max=epochs, eta_min=min_lr)`` after ``epoch`` calls of
    ``lr_scheduler.step()`` (train_causal.py:22,29: one call per epoch, so epoch e of 1..epochs trains with
    ``cosine_lr(e - 1, ...)``) -- what a loop built on ``Trainer`` hands to ``Trainer.set_lr``."""
    import math
    return float(min_lr) + (float(base_lr) - float(min_lr)) * (1.0 + math.cos(math.pi * int(epoch) / int(epochs))) / 2.0


def allreduce_flat_grads(flat_grad, group=None):
    """The ONE collective of the data-parallel step: sum the flat gradient
    buffer over the ranks in place and return the scale (1 / world_size) the optimizer applies.
    NCCL on the GPUs; any torch.distributed backend works (the CPU tests use gloo)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class GraphStore:
    """A dataset resident in device memory for ``cal_collate`` (csrc/collate.cu): all graphs back
    to back -- ``node_ptr`` / ``edge_ptr`` i32[G+1], ``feat`` f32[sum n, F], graph-local
    ``edge_src`` / ``edge_dst`` i32[sum e], ``y`` i64[G].  Replaces the per-step host collate of
    the reference's DataLoader (train_causal.py:13-15,171-176)."""

    def __init__(self, graphs, device):
        graphs = list(graphs)
        self.device = torch.device(device)
        self.num_graphs = len(graphs)
        feat = lambda d: d.x if getattr(d, "x", None) is not None else d.feat
        n = np.array([int(feat(d).size(0)) for d in graphs], dtype=np.int64)
        e = np.array([int(d.edge_index.size(1)) for d in graphs], dtype=np.int64)
        if n.sum() >= 2 ** 31 or e.sum() >= 2 ** 31:
            raise _lib.CalError("cal_b200: GraphStore is limited to 2^31 nodes / edge columns")
        self.node_counts, self.edge_counts = n, e
        self.F = int(feat(graphs[0]).size(1)) if graphs else 0
        to = lambda t: t.to(self.device)
        self.node_ptr = to(torch.from_numpy(np.concatenate([[0], np.cumsum(n)]).astype(np.int32)))
        self.edge_ptr = to(torch.from_numpy(np.concatenate([[0], np.cumsum(e)]).astype(np.int32)))
        self.feat = to(torch.cat([feat(d).float() for d in graphs]).contiguous())
        ei = torch.cat([d.edge_index for d in graphs], dim=1)
        self.edge_src = to(ei[0].to(torch.int32).contiguous())
        self.edge_dst = to(ei[1].to(torch.int32).contiguous())
        self.y = to(torch.cat([d.y.view(-1)[:1] for d in graphs]).long().contiguous())
        d = _lib.GraphStoreDesc()
        d.num_graphs, d.num_features = self.num_graphs, self.F
        d.node_ptr, d.edge_ptr = self.node_ptr.data_ptr(), self.edge_ptr.data_ptr()
        d.feat, d.edge_src, d.edge_dst, d.y = (self.feat.data_ptr(), self.edge_src.data_ptr(),
                                               self.edge_dst.data_ptr(), self.y.data_ptr())
        self.desc = d

    @classmethod
    def from_flat(cls, flat, device):
        """From ``cal_b200.datasets.FlatGraphs`` (the vectorised generator / TU reader): the flat arrays go to the
        device as they are -- no per-graph Python objects on the way."""
        self = cls.__new__(cls)
        self.device = torch.device(device)
        self.num_graphs = len(flat)
        n = (flat.node_ptr[1:] - flat.node_ptr[:-1]).cpu().numpy().astype(np.int64)
        e = (flat.edge_ptr[1:] - flat.edge_ptr[:-1]).cpu().numpy().astype(np.int64)
        if n.sum() >= 2 ** 31 or e.sum() >= 2 ** 31:
            raise _lib.CalError("cal_b200: GraphStore is limited to 2^31 nodes / edge columns")
        self.node_counts, self.edge_counts = n, e
        self.F = int(flat.feat.size(1))
        to = lambda t, dt: t.to(device=self.device, dtype=dt).contiguous()
        self.node_ptr, self.edge_ptr = to(flat.node_ptr, torch.int32), to(flat.edge_ptr, torch.int32)
        self.feat = to(flat.feat, torch.float32)
        self.edge_src, self.edge_dst = to(flat.edge_index[0], torch.int32), to(flat.edge_index[1], torch.int32)
        self.y = to(flat.y, torch.long)
        d = _lib.GraphStoreDesc()
        d.num_graphs, d.num_features = self.num_graphs, self.F
        d.node_ptr, d.edge_ptr = self.node_ptr.data_ptr(), self.edge_ptr.data_ptr()
        d.feat, d.edge_src, d.edge_dst, d.y = (self.feat.data_ptr(), self.edge_src.data_ptr(),
                                               self.edge_dst.data_ptr(), self.y.data_ptr())
        self.desc = d
        return self

    def caps(self, graphs_per_step, order=None):
        """Capacities covering every step of ``order`` (default: the worst ``graphs_per_step``
        graphs of the dataset, which covers any order)."""
        B = int(graphs_per_step)
        if order is None:
            n = int(np.sort(self.node_counts)[::-1][:B].sum())
            e = int(np.sort(self.edge_counts)[::-1][:B].sum())
        else:
            o = np.asarray(order)
            pad = (-len(o)) % B
            nn = np.concatenate([self.node_counts[o], np.zeros(pad, dtype=np.int64)]).reshape(-1, B).sum(1)
            ee = np.concatenate([self.edge_counts[o], np.zeros(pad, dtype=np.int64)]).reshape(-1, B).sum(1)
            n, e = int(nn.max()), int(ee.max())
        small = bool(self.node_counts.max(initial=0) <= FSG_ROWS and
                     (self.node_counts + self.edge_counts).max(initial=0) <= FSG_ENTRIES)
        return _up(n, 32), _up(max(e, 1), 32), _up(B, 8), small


class PeerExchange:
    """NVLink peer-memory gradient exchange fused with Adam (``cal_dp_adam_step``, csrc/comm.cu).

    One exchange region per rank, allocated by the library, exported as a CUDA IPC handle and mapped
    by every peer; the 64-byte handles travel through ``torch.distributed.all_gather_object`` (plumbing
    only -- no collective runs per step).  Raises ``CalError`` when the peers cannot map each other
    (different nodes, IPC disabled); the Trainer then keeps the NCCL all-reduce."""

    def __init__(self, lib, device, n_floats, group=None):
        import torch.distributed as dist
        self.lib = lib
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.CAL_MAX_WORLD:
            raise _lib.CalError("cal_b200: peer exchange supports at most %d ranks" % _lib.CAL_MAX_WORLD)
        self.device = torch.device(device)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        nbytes = lib.cal_dp_region_bytes(self.world, int(n_floats))
        if nbytes == 0:
            raise _lib.CalError("cal_b200: cal_dp_region_bytes rejected world=%d n=%d" % (self.world, n_floats))
        mine = C.c_void_p()
        self._mine, self._mapped = None, []
        ok, handle = 1, b"\0" * _lib.CAL_DP_HANDLE_BYTES
        rc = lib.cal_dp_alloc(dev_index, nbytes, C.byref(mine))
        if rc == 0:
            self._mine = mine.value
            buf = C.create_string_buffer(_lib.CAL_DP_HANDLE_BYTES)
            rc = lib.cal_dp_export(mine, buf)
            handle = buf.raw
        ok = int(rc == 0)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (ok, dev_index, handle), group=group)
        comm = _lib.DpComm()
        comm.world, comm.rank = self.world, self.rank
        err = None if all(g[0] for g in gathered) else "a rank could not allocate / export its exchange region"
        if err is None:
            for q, (_ok, _dev, h) in enumerate(gathered):
                if q == self.rank:
                    comm.region[q] = self._mine
                    continue
                m = C.c_void_p()
                rc = lib.cal_dp_import(dev_index, h, C.byref(m))
                if rc != 0:
                    err = "cudaIpcOpenMemHandle of rank %d's region failed: %s" % (q, lib.cal_error_string(rc).decode())
                    break
                self._mapped.append(m.value)
                comm.region[q] = m.value
        flags = [None] * self.world
        dist.all_gather_object(flags, err is None, group=group)       # all ranks agree on the outcome
        if not all(flags):
            self.close()
            raise _lib.CalError("cal_b200: peer exchange unavailable (%s)" % (err or "a peer failed to map this rank"))
        self.comm = comm
        self.nbytes = nbytes

    def adam_step(self, eng, lr, betas, eps, weight_decay, lr_device=None, sink=None):
        if eng.opt_state is None:
            eng.opt_state = (torch.zeros_like(eng.flat), torch.zeros_like(eng.flat),
                             torch.zeros(2, dtype=torch.int32, device=eng.device))
        m, v, step = eng.opt_state
        _lib.check(self.lib.cal_dp_adam_step_images(C.byref(self.comm), eng.flat.data_ptr(), eng.flat_grad.data_ptr(),
                                                    m.data_ptr(), v.data_ptr(), eng.total, step.data_ptr(), float(lr),
                                                    lr_device.data_ptr() if lr_device is not None else 0,
                                                    betas[0], betas[1], eps, weight_decay,
                                                    C.byref(sink) if sink is not None else None, eng._stream()),
                   "cal_dp_adam_step")
        eng.images_version = eng.param_version() if sink is not None else None

    def check(self, eng):
        """Raise if an exchange timed out waiting for a peer (synchronises)."""
        _lib.check(self.lib.cal_dp_read_error(C.byref(self.comm), eng._stream()), "cal_dp_adam_step (peer exchange)")

    def close(self):
        for m in self._mapped:
            self.lib.cal_dp_unmap(m)
        self._mapped = []
        if self._mine:
            self.lib.cal_dp_free(self._mine)
            self._mine = None


class PackedLayout:
    """Byte offsets of the fields of a packed batch for given capacities."""

    def __init__(self, max_nodes, max_edges, max_graphs, num_features):
        self.Nm, self.Em, self.Bm, self.F = int(max_nodes), int(max_edges), int(max_graphs), int(num_features)
        o = 0
        self.off_dims = o
        o += 16
        self.off_perm = o
        o = _up(o + 4 * self.Bm, 16)
        self.off_feat = o
        o = _up(o + 4 * self.Nm * self.F, 16)
        self.off_ei = o
        o = _up(o + 8 * 2 * self.Em, 16)
        self.off_batch = o
        o = _up(o + 8 * self.Nm, 16)
        self.off_y = o
        o = _up(o + 8 * self.Bm, 16)
        self.nbytes = o

    def used_bytes(self, N, E, B):
        """Bytes of a batch that carry information (the h2d payload a dense packing would need)."""
        return 16 + 4 * B + 4 * N * self.F + 16 * E + 8 * N + 8 * B

    def pack(self, data, out, perm=None):
        """Write ``data`` (CPU tensors) into the uint8 tensor ``out`` (pinned or pageable host memory)."""
        a = out.numpy()
        x = data.x if getattr(data, "x", None) is not None else data.feat
        N, E, B = int(x.size(0)), int(data.edge_index.size(1)), int(data.num_graphs)
        if N > self.Nm or E > self.Em or B > self.Bm or int(x.size(1)) != self.F:
            raise _lib.CalError("cal_b200: batch (N=%d, E=%d, B=%d, F=%d) exceeds the packed layout "
                                "(N<=%d, E<=%d, B<=%d, F=%d)" % (N, E, B, x.size(1), self.Nm, self.Em, self.Bm, self.F))
        a[self.off_dims:self.off_dims + 16].view(np.int32)[:] = (N, E, B, 1 if perm is not None else 0)
        pv = a[self.off_perm:self.off_perm + 4 * self.Bm].view(np.int32)
        pv[:B] = np.arange(B, dtype=np.int32) if perm is None else np.asarray(perm, dtype=np.int32)
        a[self.off_feat:self.off_feat + 4 * N * self.F].view(np.float32)[:] = x.numpy().reshape(-1)
        ei = data.edge_index.numpy()
        ev = a[self.off_ei:self.off_ei + 16 * self.Em].view(np.int64)
        ev[:E] = ei[0]
        ev[self.Em:self.Em + E] = ei[1]
        a[self.off_batch:self.off_batch + 8 * N].view(np.int64)[:] = data.batch.numpy()
        a[self.off_y:self.off_y + 8 * B].view(np.int64)[:] = data.y.numpy().reshape(-1)
        return out

    def cbatch(self, base_ptr):
        """The C-ABI ``cal_batch`` whose pointers address a packed buffer at ``base_ptr``."""
        cb = _lib.Batch()
        cb.dims = base_ptr + self.off_dims
        cb.perm = base_ptr + self.off_perm
        cb.feat = base_ptr + self.off_feat
        cb.edge_index = base_ptr + self.off_ei
        cb.edge_stride = self.Em
        cb.batch = base_ptr + self.off_batch
        cb.y = base_ptr + self.off_y
        cb.gat_keep = 0
        return cb


class Trainer:
    """Flat-buffer trainer around one ``CausalGCN`` / ``CausalGAT`` on one device.

    ``process_group``: a ``torch.distributed`` group (or ``True`` for the default group); when its
    world size is > 1 the flat gradient buffer is all-reduced (sum) after backward and the Adam step
    scales by 1/world_size.  BatchNorm statistics stay per rank."""

    def __init__(self, model, caps, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 process_group=None, use_graph=True, with_random=None, max_graphs_cached=1024, collective="auto"):
        self.model = model
        self.eng = model.engine
        eng = self.eng
        self.device = eng.device
        self._dead = False
        eng.set_caps(*caps)              # invalidates a Trainer that owned the previous workspace
        eng._owner = weakref.ref(self)   # ... and freezes the capacities: Engine.ensure_caps raises instead of growing
        self.fused_small_graphs = bool(eng.caps.small_graphs)
        self.layout = PackedLayout(eng.caps.max_nodes, eng.caps.max_edges, eng.caps.max_graphs, eng.F)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
        self.use_graph = use_graph
        self.pg = None
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.pg = dist.group.WORLD if process_group is True else process_group
            self.world = dist.get_world_size(self.pg)
        # the gradient exchange: "peer" = NVLink peer-memory push fused with Adam (one captured graph per
        # step, no collective launch), "nccl" = ncclAllReduce between a compute and an update graph,
        # "auto" = peer when the ranks can map each other's memory, else nccl
        self.peer = None
        self.collective = "none"
        if self.world > 1:
            if collective not in ("auto", "peer", "nccl"):
                raise _lib.CalError("cal_b200: collective must be 'auto', 'peer' or 'nccl'")
            self.collective = "nccl"
            if collective in ("auto", "peer"):
                try:
                    self.peer = PeerExchange(eng.lib, self.device, eng.total, self.pg)
                    self.collective = "peer"
                except _lib.CalError:
                    if collective == "peer":
                        raise
        self._graphs = {}
        self._max_graphs = max_graphs_cached
        self.staging = torch.zeros(self.layout.nbytes, dtype=torch.uint8, device=self.device)
        self._host_loss = torch.zeros(8, dtype=torch.float32).pin_memory()
        self.launches_per_step = None
        # train_causal.py:177 calls model(data, eval_random=args.with_random): no shuffle when args.with_random is off
        self.with_random = (bool(getattr(model.args, "with_random", True)) and bool(model._shuffles(True))
                            if with_random is None else with_random)
        self._is_gat = eng.is_gat
        self._warm = False
        self._update_graph = None
        self._update_launches = 0

    def _invalidate(self):
        """The engine's workspace was reallocated (Engine.set_caps / a newer Trainer): the captured
        graphs of this Trainer point into freed memory and must never be replayed."""
        self._dead = True
        self._graphs = {}
        self._update_graph = None
        if getattr(self, "_epoch", None) is not None:
            self._epoch["graph"] = None

    def _check_alive(self):
        if self._dead:
            raise _lib.CalError("cal_b200: this Trainer's workspace was reallocated (Engine.set_caps or a newer "
                                "Trainer on the same model); create a new Trainer")

    # ---- batches ----
    def set_lr(self, lr):
        self.lr_dev.fill_(float(lr))

    def draw_perm(self, B):
        """random_idx of model.py:147-152: one Python-RNG shuffle per training forward."""
        if not self.with_random:
            return None
        l = list(range(B))
        random.shuffle(l)
        return l

    def pack(self, data, out=None, perm="draw"):
        if out is None:
            out = torch.empty(self.layout.nbytes, dtype=torch.uint8)
            out.zero_()
            if self.device.type == "cuda":
                out = out.pin_memory()
        if perm == "draw":
            perm = self.draw_perm(int(data.num_graphs))
        return self.layout.pack(data, out, perm)

    def upload(self, data, perm="draw"):
        """Pack a CPU batch and keep it resident on the device."""
        return self.pack(data, perm=perm).to(self.device, non_blocking=False)

    # ---- the step ----
    def params_changed(self):
        """Tell the trainer that parameters were written behind torch's version counters (through ``p.data`` or a
        detached alias): the next step rebuilds the fused path's operand images from the parameters."""
        self.eng.images_stale()

    def _sink(self):
        sk = self.eng.image_sink()
        return sk if sk.count > 0 else None

    def _issue(self, base_ptr, gat_keep=None, part="all", next_ptr=None, prepped=False, ready=False, loss_out=None,
               side_first=None):
        """Enqueue the step on the current stream.  part: "all", or "compute" (prep + forward + loss +
        backward) / "update" (gradient all-reduce + Adam) for the two-graph data-parallel replay.
        ``prepped``: cal_prep of this batch already ran (at the end of the previous step);  ``next_ptr``: run
        cal_prep of the NEXT batch on a forked branch next to this step's update (the structure work needs only the
        batch itself, so it hides under the gradient exchange / Adam instead of heading the next step);
        ``ready``: prepped AND the operand images of the fused small-graph path are current (the optimizer steps issued
        here write them, cal_image_sink) -- the forward pass then starts with the fused kernel itself;
        ``loss_out``: pinned f32[8] host tensor -- the loss parts / correct counts are copied into it on a branch forked
        right behind the forward pass (the copy engine works beside the backward pass);
        ``side_first``: callable issued on a branch forked right behind the forward pass; cal_prep(next) waits for it
        (the device-resident epoch collates the next batch there: it touches only the other staging buffer and the
        loss parts the forward pass has just written, so it runs beside the readout backward kernel)."""
        eng, lib = self.eng, self.eng.lib
        sink = self._sink()
        side2 = None
        if part in ("all", "compute"):
            cb = self.layout.cbatch(base_ptr)
            if gat_keep is not None:
                cb.gat_keep = gat_keep.data_ptr()
            s = eng._stream()
            d, caps = C.byref(eng.desc), C.byref(eng.caps)
            if not prepped:
                _lib.check(lib.cal_prep(d, caps, C.byref(cb), eng.ws.data_ptr(), eng.ws_bytes, s), "cal_prep")
            fflags = _lib.CAL_F_TRAIN | _lib.CAL_F_LOSS
            if prepped:                                        # the kernel ahead of this call is the optimizer's, not cal_prep's
                fflags |= _lib.CAL_F_NO_OVERLAP | (_lib.CAL_F_FSG_READY if ready else 0)
            _lib.check(lib.cal_causal_forward(d, caps, C.byref(eng.po), C.byref(eng.bo), eng.flat.data_ptr(),
                                              eng.bn_buf.data_ptr(), eng.nbt.data_ptr(), C.byref(cb),
                                              fflags, 0, eng.ws.data_ptr(), eng.ws_bytes, s),
                       "cal_causal_forward")
            if loss_out is not None:
                main = torch.cuda.current_stream(self.device)
                side2 = self._side_stream(1)
                side2.wait_stream(main)
                with torch.cuda.stream(side2):
                    loss_out.copy_(eng.loss_parts_full(), non_blocking=True)
            if side_first is not None and next_ptr is not None and part == "all":
                # (the same branch that later prepares the batch: one chain  collate(next) -> [backward done] -> prep(next))
                main = torch.cuda.current_stream(self.device)
                early = self._side_stream()
                early.wait_stream(main)
                with torch.cuda.stream(early):
                    side_first()
            fork = next_ptr is not None and part == "all"
            last = eng.L + 6                                   # backward stage "grad_reduce"

            def bwd(flags):
                _lib.check(lib.cal_causal_backward(d, caps, C.byref(eng.po), eng.flat.data_ptr(), C.byref(cb), 0,
                                                   eng.flat_grad.data_ptr(), flags, eng.ws.data_ptr(), eng.ws_bytes, s),
                           "cal_causal_backward")
            bwd(_lib.stages_flag(0, last - 1) if fork else 0)
            eng.gen += 1
        side = None
        if next_ptr is not None and part == "all":
            # fork: the next batch's structure preparation runs beside the block-order gradient reduction and the
            # update.  It writes only workspace regions (CSR, norms, graph_ptr, input-feature statistics, status) that
            # nothing after the last backward kernel reads: the reduction reads the partial gradients and the batch
            # dims, the update touches the flat buffers / the exchange region.
            main = torch.cuda.current_stream(self.device)
            side = self._side_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                cbn = self.layout.cbatch(next_ptr)
                if gat_keep is not None:
                    cbn.gat_keep = gat_keep.data_ptr()
                _lib.check(lib.cal_prep(C.byref(eng.desc), C.byref(eng.caps), C.byref(cbn), eng.ws.data_ptr(), eng.ws_bytes,
                                        side.cuda_stream), "cal_prep")
            bwd(_lib.stages_flag(last, last))
        if self.peer is not None:
            if part in ("all", "update"):
                self.peer.adam_step(eng, 0.0, self.betas, self.eps, self.weight_decay, lr_device=self.lr_dev, sink=sink)
            for sd in (side, side2):
                if sd is not None:
                    torch.cuda.current_stream(self.device).wait_stream(sd)
            return
        if part in ("all", "allreduce"):
            self._scale = allreduce_flat_grads(eng.flat_grad, self.pg) if self.world > 1 else 1.0
        if part in ("all", "update"):
            eng.adam_step(0.0, self.betas, self.eps, self.weight_decay, 1.0 / self.world, lr_device=self.lr_dev, sink=sink)
        for sd in (side, side2):
            if sd is not None:
                torch.cuda.current_stream(self.device).wait_stream(sd)

    def _side_stream(self, i=0):
        if not hasattr(self, "_side"):
            self._side = [torch.cuda.Stream(self.device) for _ in range(2)]
        return self._side[i]

    def _gat_keep_for(self, key):
        """Attention-dropout keep mask of CausalGAT (model.py:340 dropout=0.2), regenerated on the
        device before every step into a fixed buffer the captured graph points at."""
        if not self._is_gat or float(self.model.dropout) <= 0.0:
            return None
        if not hasattr(self, "_keep"):
            eng = self.eng
            self._keep = torch.ones(eng.L, eng.caps.max_edges + eng.caps.max_nodes, self.model.head,
                                    dtype=torch.float32, device=self.device)
        return self._keep

    def _refresh_keep(self):
        if getattr(self, "_keep", None) is not None:
            p = float(self.model.dropout)
            self._keep.bernoulli_(1.0 - p).mul_(1.0 / (1.0 - p))

    def step(self, packed_dev, next_packed=None, loss_out=None):
        """Enqueue one training step on a device-resident packed batch (asynchronous).

        ``next_packed``: the device-resident batch of the NEXT call (a loader that knows it one step ahead): its
        structure preparation (cal_prep) is then issued on a forked branch beside this step's update and the next
        ``step(next_packed, ...)`` skips it -- every batch is still prepared exactly once.  (Single captured graph
        per step only: not with the two-graph NCCL replay.)
        ``loss_out``: pinned f32[8] host tensor that receives this step's loss parts / correct counts (copied beside the
        backward pass; valid once the step has completed)."""
        self._check_alive()
        self.pipe_flush()                                # (no-op unless a step_host_async batch is pending)
        if packed_dev.device != self.device:
            raise _lib.CalError("cal_b200: Trainer.step needs a device-resident packed batch (use step_host)")
        keep = self._gat_keep_for(None)
        if keep is not None:
            self._refresh_keep()
        ahead_ok = self.world == 1 or self.peer is not None
        ptr = packed_dev.data_ptr()
        prepped = ahead_ok and getattr(self, "_prepped_ptr", None) == ptr
        nxt = next_packed.data_ptr() if (next_packed is not None and ahead_ok) else None
        self._prepped_ptr = nxt
        ready = prepped and self._sink() is not None and self.eng.images_fresh()
        if not self.use_graph:
            c0 = self.eng.lib.cal_launch_count()
            self._issue(ptr, keep, next_ptr=nxt, prepped=prepped, ready=ready, loss_out=loss_out)
            self.launches_per_step = int(self.eng.lib.cal_launch_count() - c0) + (
                1 if self.world > 1 and self.peer is None else 0)
            return
        key = ptr if (nxt is None and not prepped and loss_out is None) else (
            ptr, nxt, prepped, ready, loss_out.data_ptr() if loss_out is not None else None)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self._max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            if not self._warm:
                self._warmup(packed_dev, keep)
            count = self.eng.lib.cal_launch_count
            c0 = count()
            if self.world == 1 or self.peer is not None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    self._issue(ptr, keep, next_ptr=nxt, prepped=prepped, ready=ready, loss_out=loss_out)
                g = (g, None)
                self.launches_per_step = int(count() - c0)
            else:
                # data parallel: the NCCL all-reduce is issued eagerly between two captured graphs
                # (compute | update); the update graph is shared by every batch
                ga = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga, stream=self._capture_stream()):
                    self._issue(ptr, keep, part="compute")
                n_compute = int(count() - c0)
                if self._update_graph is None:
                    c1 = count()
                    self._update_graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self._update_graph, stream=self._capture_stream()):
                        self._issue(ptr, keep, part="update")
                    self._update_launches = int(count() - c1)
                g = (ga, self._update_graph)
                self.launches_per_step = n_compute + 1 + self._update_launches     # + the NCCL all-reduce kernel
            self._graphs[key] = (g, packed_dev, loss_out)
        else:
            g = g[0]
        g[0].replay()
        if g[1] is not None:
            self._issue(ptr, keep, part="allreduce")
            g[1].replay()
        if loss_out is not None and g[1] is not None:      # (two-graph NCCL replay: the copy is not part of a graph)
            loss_out.copy_(self.eng.loss_parts_full(), non_blocking=True)
        # (the replayed optimizer kernel wrote the operand images of the parameters it produced)
        self.eng.images_version = self.eng.param_version() if self._sink() is not None else None

    def step_many(self, batches, next_packed=None, loss_out=None):
        """``len(batches)`` consecutive training steps as ONE captured graph (asynchronous) -- the same kernels in the
        same order as ``step(b0, b1); step(b1, b2); ...; step(b_last, next_packed)`` (bit-identical parameters), but the
        steps inside the group are separated by a kernel-to-kernel dependency instead of the gap between two graph
        launches (measured: 7.7 us).  ``loss_out``: one pinned f32[8] tensor per step, or None.  Falls back to single
        steps where a step needs host work of its own (CausalGAT's per-step dropout mask, the two-graph NCCL replay,
        eager mode)."""
        batches = list(batches)
        K = len(batches)
        outs = list(loss_out) if loss_out is not None else [None] * K
        ahead_ok = self.world == 1 or self.peer is not None
        if K == 1 or not self.use_graph or not ahead_ok or self._gat_keep_for(None) is not None:
            for k, b in enumerate(batches):
                self.step(b, batches[k + 1] if k + 1 < K else next_packed, outs[k])
            return
        self._check_alive()
        self.pipe_flush()
        for b in batches:
            if b.device != self.device:
                raise _lib.CalError("cal_b200: Trainer.step_many needs device-resident packed batches")
        ptrs = tuple(b.data_ptr() for b in batches)
        prepped = getattr(self, "_prepped_ptr", None) == ptrs[0]
        nxt = next_packed.data_ptr() if next_packed is not None else None
        self._prepped_ptr = nxt
        sink = self._sink()
        ready = prepped and sink is not None and self.eng.images_fresh()
        key = (ptrs, nxt, prepped, ready, tuple(o.data_ptr() if o is not None else None for o in outs))
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self._max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            if not self._warm:
                self._warmup(batches[0], None)
            count = self.eng.lib.cal_launch_count
            c0 = count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream()):
                for k, p in enumerate(ptrs):
                    # inside the group every step finds its batch prepared and the images written by the step before it
                    self._issue(p, None, next_ptr=ptrs[k + 1] if k + 1 < K else nxt,
                                prepped=prepped if k == 0 else True, ready=ready if k == 0 else sink is not None,
                                loss_out=outs[k])
            self.launches_per_step = int(round((count() - c0) / K))
            self._graphs[key] = ((g, None), batches, outs)
        else:
            g = g[0][0]
        g.replay()
        self.eng.images_version = self.eng.param_version() if sink is not None else None

    def _capture_stream(self):
        if not hasattr(self, "_cap_stream"):
            self._cap_stream = torch.cuda.Stream(self.device)
        return self._cap_stream

    def _save_state(self):
        eng = self.eng
        return (eng.flat.clone(), eng.flat_grad.clone(), eng.bn_buf.clone(), eng.nbt.clone(),
                None if eng.opt_state is None else [t.clone() for t in eng.opt_state])

    def _warmup(self, packed_dev, keep):
        """Load every kernel (lazy module loading, cudaFuncSetAttribute) outside capture without
        touching the model state: run the pass, then restore parameters / statistics."""
        eng = self.eng
        self._prepped_ptr = None
        saved = self._save_state()
        self._issue(packed_dev.data_ptr(), keep)
        self._restore_state(saved)
        self._warm = True

    def _restore_state(self, saved):
        eng = self.eng
        torch.cuda.synchronize(self.device)
        eng.flat.copy_(saved[0])
        eng.flat_grad.copy_(saved[1])
        eng.bn_buf.copy_(saved[2])
        eng.nbt.copy_(saved[3])
        if saved[4] is None:
            for t in (eng.opt_state or ()):
                t.zero_()
        else:
            for t, s in zip(eng.opt_state, saved[4]):
                t.copy_(s)
        torch.cuda.synchronize(self.device)

    def profile_stages(self, packed_dev, reps=20):
        """Live per-operator timing: every stage of the step issued ``reps`` times back to back
        (as one CUDA-graph replay) between two CUDA events on the launching stream.  Model state is restored afterwards.
        -> list of (name, kernel launches per issue, average milliseconds per issue)."""
        eng, lib = self.eng, self.eng.lib
        self.pipe_flush()
        self._prepped_ptr = None
        saved = self._save_state()
        keep = self._gat_keep_for(None)
        # every buffer holds a consistent step.  LOCAL update: this is a single-rank measurement -- a gradient exchange
        # here would wait for peers that are not taking part (bench.py times the stages on rank 0 only)
        self._issue(packed_dev.data_ptr(), keep, part="compute")
        eng.adam_step(0.0, self.betas, self.eps, self.weight_decay, 1.0, lr_device=self.lr_dev)
        cb = self.layout.cbatch(packed_dev.data_ptr())
        if keep is not None:
            cb.gat_keep = keep.data_ptr()
        d, caps = C.byref(eng.desc), C.byref(eng.caps)
        ws, nb = eng.ws.data_ptr(), eng.ws_bytes
        # NB: the stream is looked up at issue time -- inside torch.cuda.graph it is the capture stream

        def fwd(i):
            _lib.check(lib.cal_causal_forward(d, caps, C.byref(eng.po), C.byref(eng.bo), eng.flat.data_ptr(),
                                              eng.bn_buf.data_ptr(), eng.nbt.data_ptr(), C.byref(cb),
                                              _lib.CAL_F_TRAIN | _lib.CAL_F_LOSS | _lib.stages_flag(i, i), 0, ws, nb,
                                              eng._stream()), "cal_causal_forward")

        def bwd(i):
            _lib.check(lib.cal_causal_backward(d, caps, C.byref(eng.po), eng.flat.data_ptr(), C.byref(cb), 0,
                                               eng.flat_grad.data_ptr(), _lib.stages_flag(i, i), ws, nb,
                                               eng._stream()), "cal_causal_backward")

        jobs = [("prep", lambda: _lib.check(lib.cal_prep(d, caps, C.byref(cb), ws, nb, eng._stream()), "cal_prep"))]
        jobs += [(n, (lambda i=i: fwd(i))) for i, n in enumerate(eng.stage_names()) if n != "copy_out"]
        jobs += [(n, (lambda i=i: bwd(i))) for i, n in enumerate(eng.stage_names(backward=True))]
        jobs += [("adam", lambda: eng.adam_step(0.0, self.betas, self.eps, self.weight_decay, 1.0, lr_device=self.lr_dev))]
        out = []
        for name, fn in jobs:
            c0 = lib.cal_launch_count()
            fn()
            nl = int(lib.cal_launch_count() - c0)
            # `reps` issues captured in one CUDA graph: host launch latency stays out of the timing
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream()):
                for _ in range(reps):
                    fn()
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(self.device)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize(self.device)
            out.append((name, nl, e0.elapsed_time(e1) / reps))
            del g
        self._restore_state(saved)
        return out

    def step_host(self, packed_host, sync=True):
        """End-to-end step: H2D copy of a (pinned) packed host batch, the step, and the loss parts /
        correct counts back on the host.  Returns the pinned f32[8] result when ``sync``."""
        self.staging.copy_(packed_host, non_blocking=True)
        self.step(self.staging, loss_out=self._host_loss)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
            return self._host_loss
        return None

    # ---- pipelined end-to-end path: H2D of batch i+1 overlaps the step of batch i ----
    PIPE_RING = 16              # staging buffers = result slots: batch i uses buffer / slot i % PIPE_RING

    def _pipe_init(self):
        dev = self.device
        n = self.PIPE_RING
        self._pipe = {
            "i": 0,
            "stage": [torch.zeros(self.layout.nbytes, dtype=torch.uint8, device=dev) for _ in range(n)],
            "copy": torch.cuda.Stream(dev),
            "ready": [torch.cuda.Event() for _ in range(n)],
            "ring": torch.zeros(self.PIPE_RING, 8, dtype=torch.float32).pin_memory(),
            "ring_ev": [torch.cuda.Event() for _ in range(self.PIPE_RING)],
            "pending": None,
        }

    def step_host_async(self, packed_host):
        """Enqueue one end-to-end step without synchronising the host.  The packed batch is uploaded on a copy
        stream into one of PIPE_RING staging buffers (overlapping earlier steps) and its loss parts / correct counts
        are copied into the buffer's pinned result slot beside its backward pass (a copy node of the step's captured
        graph).  Returns the slot index; ``pipe_result(slot)`` waits for it.

        The step of a batch is ISSUED one call late -- together with the upload of the batch that follows it -- so
        that the follower's structure preparation can ride beside this step's gradient reduction and update
        (``step(cur, next)``); ``pipe_result`` / ``pipe_flush`` (and every other entry point of the trainer) issue
        the step that is still pending.  Every batch is prepared and stepped exactly once, in call order."""
        if not hasattr(self, "_pipe"):
            self._pipe_init()
        p = self._pipe
        i = p["i"]
        b = r = i % self.PIPE_RING
        main = torch.cuda.current_stream(self.device)
        if i >= self.PIPE_RING:
            p["ring_ev"][r].synchronize()                # bound the host run-ahead to the ring depth (and: the
            #                                              step that last read this staging buffer / wrote this slot is done)
        cs = p["copy"]
        with torch.cuda.stream(cs):
            p["stage"][b].copy_(packed_host, non_blocking=True)
            p["ready"][b].record(cs)
        main.wait_event(p["ready"][b])
        pend = p["pending"]
        p["pending"] = None
        if pend is not None:
            self._pipe_issue(pend, b)
        p["pending"] = (b, r)
        p["i"] = i + 1
        return r

    def _pipe_issue(self, pend, nxt):
        p = self._pipe
        b, r = pend
        main = torch.cuda.current_stream(self.device)
        self.step(p["stage"][b], p["stage"][nxt] if nxt is not None else None, loss_out=p["ring"][r])
        p["ring_ev"][r].record(main)

    def pipe_flush(self):
        """Issue the step of the most recent ``step_host_async`` batch if it is still pending."""
        p = getattr(self, "_pipe", None)
        if p is not None and p["pending"] is not None:
            pend, p["pending"] = p["pending"], None
            self._pipe_issue(pend, None)

    def pipe_reset(self):
        """Issue the pending step, wait for everything and rewind the buffer cursor (the next ``step_host_async`` uses
        buffer / slot 0 again -- a repeated sequence of calls then replays the same captured graphs)."""
        self.pipe_flush()
        torch.cuda.current_stream(self.device).synchronize()
        if hasattr(self, "_pipe"):
            self._pipe["i"] = 0

    def pipe_result(self, slot):
        p = self._pipe
        if p["pending"] is not None and p["pending"][1] == slot:
            self.pipe_flush()
        p["ring_ev"][slot].synchronize()
        return p["ring"][slot]

    # ---- device-resident epochs: collate on the GPU, one captured graph for every step ----
    def begin_epoch(self, store, order, graphs_per_step, perms="draw"):
        """Upload the epoch's graph order (and the random-intervention permutations, drawn on the
        host with Python's RNG like model.py:147-152) and rewind the device cursor.  Afterwards
        ``step_epoch()`` runs one training step per call with no host -> device traffic."""
        if store.F != self.eng.F:
            raise _lib.CalError("cal_b200: store has %d features, model expects %d" % (store.F, self.eng.F))
        B = int(graphs_per_step)
        order = np.asarray(order, dtype=np.int32)
        n_steps = -(-len(order) // B)
        need = store.caps(B, order)
        caps = self.eng.caps
        if need[0] > caps.max_nodes or need[1] > caps.max_edges or B > caps.max_graphs:
            raise _lib.CalError("cal_b200: epoch needs capacities %s, trainer has (%d, %d, %d)"
                                % (need, caps.max_nodes, caps.max_edges, caps.max_graphs))
        ep = getattr(self, "_epoch", None)
        if ep is None or ep["B"] != B or ep["store"] is not store or ep["cap"] < len(order):
            cap = max(len(order), 1)
            ep = self._epoch = {
                "B": B, "store": store, "cap": cap,
                "order": torch.zeros(cap, dtype=torch.int32, device=self.device),
                "perm": torch.zeros(-(-cap // B) * B, dtype=torch.int32, device=self.device),
                "pos": torch.zeros(4, dtype=torch.int32, device=self.device),
                "acc": torch.zeros(8, dtype=torch.float32, device=self.device),
                "graph": None,
            }
        ep["n"], ep["steps"], ep["done"] = len(order), n_steps, 0
        ep["order"][:len(order)].copy_(torch.from_numpy(order), non_blocking=True)
        ep["with_perm"] = False
        if isinstance(perms, str):                  # "draw"
            perms = None
            if self.with_random:
                perms = np.zeros((n_steps, B), dtype=np.int32)
                for s in range(n_steps):
                    bn = min(B, len(order) - s * B)
                    perms[s, :bn] = self.draw_perm(bn)
        if perms is not None:
            pp = np.ascontiguousarray(np.asarray(perms, dtype=np.int32)).reshape(-1)
            ep["perm"][:pp.size].copy_(torch.from_numpy(pp), non_blocking=True)
            ep["with_perm"] = True
        ep["pos"].zero_()
        ep["acc"].zero_()
        ep["next_collated"], ep["slot"] = False, 0         # (look-ahead state of _step_epoch_ahead)
        return n_steps

    def _issue_collate(self, ep, staging=None):
        eng = self.eng
        cb = self.layout.cbatch((staging if staging is not None else self.staging).data_ptr())
        _lib.check(eng.lib.cal_collate(C.byref(ep["store"].desc), ep["order"].data_ptr(), ep["n"], ep["pos"].data_ptr(),
                                       ep["B"], ep["perm"].data_ptr() if ep["with_perm"] else 0, C.byref(eng.caps),
                                       C.byref(cb), 1, eng.loss_parts_full().data_ptr(), ep["acc"].data_ptr(),
                                       eng._stream()), "cal_collate")

    def step_epoch(self):
        """One training step on the next ``graphs_per_step`` graphs of the epoch (asynchronous)."""
        self._check_alive()
        self.pipe_flush()
        ep = self._epoch
        if ep["done"] >= ep["steps"]:
            raise _lib.CalError("cal_b200: the epoch is exhausted (call begin_epoch)")
        ep["done"] += 1
        keep = self._gat_keep_for(None)
        if keep is not None:
            self._refresh_keep()
        if self.world == 1 or self.peer is not None:
            return self._step_epoch_ahead(ep, keep)
        self._prepped_ptr = None
        if not self.use_graph:
            self._issue_collate(ep)
            self._issue(self.staging.data_ptr(), keep)
            return
        key = (ep["n"], ep["with_perm"])
        if ep["graph"] is None or ep["graph"][0] != key:
            if not self._warm:
                pos = ep["pos"].clone()
                self._issue_collate(ep)
                self._warmup(self.staging, keep)
                ep["pos"].copy_(pos)
                ep["acc"].zero_()
            if self.world > 1 and self.peer is None:
                # NCCL between two captured graphs (see step()): collate + compute | all-reduce | update
                ga = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga, stream=self._capture_stream()):
                    self._issue_collate(ep)
                    self._issue(self.staging.data_ptr(), keep, part="compute")
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb, stream=self._capture_stream()):
                    self._issue(self.staging.data_ptr(), keep, part="update")
                ep["graph"] = (key, ga, gb)
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    self._issue_collate(ep)
                    self._issue(self.staging.data_ptr(), keep)
                ep["graph"] = (key, g, None)
        _, ga, gb = ep["graph"]
        ga.replay()
        if gb is not None:
            self._issue(self.staging.data_ptr(), keep, part="allreduce")
            gb.replay()

    def _step_epoch_ahead(self, ep, keep):
        """The epoch step with the look-ahead of ``step(cur, next)``: two staging buffers alternate; the NEXT batch is
        collated and prepared on the forked branch beside this step's gradient reduction and update, so the step of a
        batch that was prepared ahead starts with the forward pass.  (``cal_collate`` weighs the previous step's
        losses with the graph count it left in the cursor block, not with the dims of the buffer it is about to
        overwrite, so alternating buffers keep the epoch sums exact.)  A handful of captured graphs serve an epoch:
        first / steady (one per buffer) / last."""
        if not hasattr(self, "_staging2"):
            self._staging2 = torch.zeros_like(self.staging)
        stages = (self.staging, self._staging2)
        slot = ep.get("slot", 0)
        cur = stages[slot]
        has_next = ep["done"] < ep["steps"]                 # (ep["done"] already counts this step)
        need_collate = not ep.get("next_collated", False)
        prepped = (not need_collate) and getattr(self, "_prepped_ptr", None) == cur.data_ptr()
        nxt = stages[1 - slot] if has_next else None
        sink = self._sink()
        ready = prepped and sink is not None and self.eng.images_fresh()
        ep["next_collated"] = has_next
        ep["slot"] = 1 - slot if has_next else slot
        self._prepped_ptr = nxt.data_ptr() if nxt is not None else None

        def issue():
            if need_collate:
                self._issue_collate(ep, cur)
            self._issue(cur.data_ptr(), keep, next_ptr=nxt.data_ptr() if nxt is not None else None, prepped=prepped,
                        ready=ready, side_first=(lambda: self._issue_collate(ep, nxt)) if nxt is not None else None)

        if not self.use_graph:
            issue()
            return
        key = (ep["n"], ep["with_perm"], slot, need_collate, prepped, ready, has_next)
        graphs = ep.setdefault("graphs", {})
        g = graphs.get(key)
        if g is None:
            if not self._warm:
                pos = ep["pos"].clone()
                self._issue_collate(ep, cur)
                self._warmup(cur, keep)
                ep["pos"].copy_(pos)
                ep["acc"].zero_()
                self._prepped_ptr = nxt.data_ptr() if nxt is not None else None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream()):
                issue()
            graphs[key] = g
        g.replay()
        self.eng.images_version = self.eng.param_version() if sink is not None else None

    def end_epoch(self):
        """Epoch metrics like train_causal.py:186-196 (synchronises): dict of mean losses and accuracies."""
        ep = self._epoch
        eng = self.eng
        cb = self.layout.cbatch(self.staging.data_ptr())
        _lib.check(eng.lib.cal_collate_flush(eng.loss_parts_full().data_ptr(), cb.dims, ep["pos"].data_ptr(),
                                             ep["acc"].data_ptr(), eng._stream()), "cal_collate_flush")
        a = ep["acc"].cpu().tolist()
        self.check()
        n = max(a[7], 1.0)
        return {"graphs": int(a[7]), "loss": a[0] / n, "c_loss": a[1] / n, "o_loss": a[2] / n, "co_loss": a[3] / n,
                "acc_c": a[4] / n, "acc_o": a[5] / n, "acc_co": a[6] / n}

    def check(self):
        """Raise on a data-dependent violation of the last step (status word: node id out of range, unsorted
        ``batch``, capacity overflow) or a failed peer exchange.  Synchronises; called by ``end_epoch``."""
        self.pipe_flush()
        st = self.eng.status()
        if st != 0:
            raise _lib.CalError("cal_b200: the last step reported status bits 0x%x (1 = edge endpoint out of range, "
                                "2 = batch vector not sorted, 4 = batch exceeds the workspace capacities)" % st)
        if self.peer is not None:
            self.peer.check(self.eng)

    def metrics(self):
        """f32[7] device view: loss, c_loss, o_loss, co_loss, correct_c, correct_o, correct_co of
        the most recent step."""
        self.pipe_flush()
        return self.eng.loss_parts()

    # ---- evaluation (train_causal.py:202-223) ----
    @torch.no_grad()
    def eval_batch(self, data_dev, eval_random=False):
        self.pipe_flush()
        self._prepped_ptr = None                         # (the module path runs cal_prep on its own batch)
        was = self.model.training
        self.model.eval()
        out = self.model(data_dev, eval_random=eval_random)
        self.model.train(was)
        return out
