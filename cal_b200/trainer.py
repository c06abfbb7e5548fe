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

import numpy as np
import torch

from . import _lib

__all__ = ["PackedLayout", "Trainer", "batch_caps", "allreduce_flat_grads"]


def _up(x, m):
    return (x + m - 1) // m * m


def batch_caps(batches, slack=1.0):
    """Capacities (max nodes, max edge_index columns, max graphs) covering ``batches``."""
    n = max(int(b.batch.numel()) for b in batches)
    e = max(int(b.edge_index.size(1)) for b in batches)
    g = max(int(b.num_graphs) for b in batches)
    return _up(int(n * slack), 32), _up(max(int(e * slack), 1), 32), _up(g, 8)


def allreduce_flat_grads(flat_grad, group=None):
    """The ONE collective of the data-parallel step: sum the flat gradient
    buffer over the ranks in place and return the scale (1 / world_size) the optimizer applies.
    NCCL on the GPUs; any torch.distributed backend works (the CPU tests use gloo)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


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
                 process_group=None, use_graph=True, with_random=None, max_graphs_cached=1024):
        self.model = model
        self.eng = model.engine
        eng = self.eng
        self.device = eng.device
        eng.set_caps(*caps)
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
        self._graphs = {}
        self._max_graphs = max_graphs_cached
        self.staging = torch.zeros(self.layout.nbytes, dtype=torch.uint8, device=self.device)
        self._host_loss = torch.zeros(8, dtype=torch.float32).pin_memory()
        self.launches_per_step = None
        self.with_random = bool(model._shuffles(True)) if with_random is None else with_random
        self._is_gat = eng.is_gat
        self._warm = False
        self._update_graph = None
        self._update_launches = 0

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
    def _issue(self, base_ptr, gat_keep=None, part="all"):
        """Enqueue the step on the current stream.  part: "all", or "compute" (prep + forward + loss +
        backward) / "update" (gradient all-reduce + Adam) for the two-graph data-parallel replay."""
        eng, lib = self.eng, self.eng.lib
        if part in ("all", "compute"):
            cb = self.layout.cbatch(base_ptr)
            if gat_keep is not None:
                cb.gat_keep = gat_keep.data_ptr()
            s = eng._stream()
            d, caps = C.byref(eng.desc), C.byref(eng.caps)
            _lib.check(lib.cal_prep(d, caps, C.byref(cb), eng.ws.data_ptr(), eng.ws_bytes, s), "cal_prep")
            _lib.check(lib.cal_causal_forward(d, caps, C.byref(eng.po), C.byref(eng.bo), eng.flat.data_ptr(),
                                              eng.bn_buf.data_ptr(), eng.nbt.data_ptr(), C.byref(cb),
                                              _lib.CAL_F_TRAIN | _lib.CAL_F_LOSS, 0, eng.ws.data_ptr(), eng.ws_bytes, s),
                       "cal_causal_forward")
            _lib.check(lib.cal_causal_backward(d, caps, C.byref(eng.po), eng.flat.data_ptr(), C.byref(cb), 0,
                                               eng.flat_grad.data_ptr(), 0, eng.ws.data_ptr(), eng.ws_bytes, s),
                       "cal_causal_backward")
            eng.gen += 1
        if part in ("all", "allreduce"):
            self._scale = allreduce_flat_grads(eng.flat_grad, self.pg) if self.world > 1 else 1.0
        if part in ("all", "update"):
            eng.adam_step(0.0, self.betas, self.eps, self.weight_decay, 1.0 / self.world, lr_device=self.lr_dev)

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

    def step(self, packed_dev):
        """Enqueue one training step on a device-resident packed batch (asynchronous)."""
        if packed_dev.device != self.device:
            raise _lib.CalError("cal_b200: Trainer.step needs a device-resident packed batch (use step_host)")
        keep = self._gat_keep_for(None)
        if keep is not None:
            self._refresh_keep()
        if not self.use_graph:
            c0 = self.eng.lib.cal_launch_count()
            self._issue(packed_dev.data_ptr(), keep)
            self.launches_per_step = int(self.eng.lib.cal_launch_count() - c0) + (1 if self.world > 1 else 0)
            return
        key = packed_dev.data_ptr()
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= self._max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            if not self._warm:
                self._warmup(packed_dev, keep)
            count = self.eng.lib.cal_launch_count
            c0 = count()
            if self.world == 1:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._capture_stream()):
                    self._issue(key, keep)
                g = (g, None)
                self.launches_per_step = int(count() - c0)
            else:
                # data parallel: the NCCL all-reduce is issued eagerly between two captured graphs
                # (compute | update); the update graph is shared by every batch
                ga = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga, stream=self._capture_stream()):
                    self._issue(key, keep, part="compute")
                n_compute = int(count() - c0)
                if self._update_graph is None:
                    c1 = count()
                    self._update_graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self._update_graph, stream=self._capture_stream()):
                        self._issue(key, keep, part="update")
                    self._update_launches = int(count() - c1)
                g = (ga, self._update_graph)
                self.launches_per_step = n_compute + 1 + self._update_launches     # + the NCCL all-reduce kernel
            self._graphs[key] = (g, packed_dev)
        else:
            g = g[0]
        g[0].replay()
        if g[1] is not None:
            self._issue(key, keep, part="allreduce")
            g[1].replay()

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
        saved = self._save_state()
        keep = self._gat_keep_for(None)
        self._issue(packed_dev.data_ptr(), keep)               # every buffer holds a consistent step
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
        self.step(self.staging)
        self._host_loss.copy_(self.eng.loss_parts_full(), non_blocking=True)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
            return self._host_loss
        return None

    # ---- pipelined end-to-end path: H2D of batch i+1 overlaps the step of batch i ----
    PIPE_RING = 16

    def _pipe_init(self):
        dev = self.device
        self._pipe = {
            "i": 0,
            "stage": [torch.zeros(self.layout.nbytes, dtype=torch.uint8, device=dev) for _ in range(2)],
            "copy": torch.cuda.Stream(dev),
            "ready": [torch.cuda.Event() for _ in range(2)],
            "done": [torch.cuda.Event() for _ in range(2)],
            "ring": torch.zeros(self.PIPE_RING, 8, dtype=torch.float32).pin_memory(),
            "ring_ev": [torch.cuda.Event() for _ in range(self.PIPE_RING)],
        }

    def step_host_async(self, packed_host):
        """Enqueue one end-to-end step without synchronising the host: the packed batch is uploaded
        on a copy stream into one of two staging buffers (overlapping the previous step), the step
        runs on the current stream, and its loss parts / correct counts are copied into a pinned
        ring slot.  Returns the ring slot index; ``pipe_result(slot)`` waits for it."""
        if not hasattr(self, "_pipe"):
            self._pipe_init()
        p = self._pipe
        i = p["i"]
        b, r = i % 2, i % self.PIPE_RING
        main = torch.cuda.current_stream(self.device)
        if i >= self.PIPE_RING:
            p["ring_ev"][r].synchronize()                # bound the host run-ahead to the ring depth
        cs = p["copy"]
        if i >= 2:
            cs.wait_event(p["done"][b])                  # the step that last read this staging buffer
        with torch.cuda.stream(cs):
            p["stage"][b].copy_(packed_host, non_blocking=True)
            p["ready"][b].record(cs)
        main.wait_event(p["ready"][b])
        self.step(p["stage"][b])
        p["done"][b].record(main)
        p["ring"][r].copy_(self.eng.loss_parts_full(), non_blocking=True)
        p["ring_ev"][r].record(main)
        p["i"] = i + 1
        return r

    def pipe_result(self, slot):
        self._pipe["ring_ev"][slot].synchronize()
        return self._pipe["ring"][slot]

    def metrics(self):
        """f32[7] device view: loss, c_loss, o_loss, co_loss, correct_c, correct_o, correct_co of
        the most recent step."""
        return self.eng.loss_parts()

    # ---- evaluation (train_causal.py:202-223) ----
    @torch.no_grad()
    def eval_batch(self, data_dev, eval_random=False):
        was = self.model.training
        self.model.eval()
        out = self.model(data_dev, eval_random=eval_random)
        self.model.train(was)
        return out
