#!/usr/bin/env python
"""Benchmark of the CAL hot path: graphs/sec of one CausalGCN training step
(prep + forward + KL/NLL loss + backward + [gradient all-reduce] + Adam; train_causal.py:171-192)
on synthetic SPMotif-style batches (BASELINE.json configs[1]: bias 0.9, 3 layers, hidden 128,
batch 128 per GPU, fp32).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).  `value` = whole-job graphs/s with the packed batches resident in
HBM; `e2e` = the same through Trainer.step_host with pinned HOST batches (H2D copy of every batch
and D2H read of the loss parts inside the timed region); `roofline` = the dominant kernel's
algorithmic bytes / its live CUDA-event duration against the measured HBM peak; `cpu_baseline` =
the oracle timed on this box's host cores."""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "graphs_per_sec_train_step"
UNIT = "graphs/s"
L2_BYTES = 126 * 2 ** 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cal_b200", choices=["cal_b200", "reference"])
    ap.add_argument("--workload", default="spmotif", choices=["spmotif", "spmotif_refsize", "mutag", "large"])
    ap.add_argument("--model", default="CausalGCN", choices=["CausalGCN", "CausalGAT", "CausalGIN"])
    ap.add_argument("--batch", type=int, default=0, help="graphs per GPU per step (default: the workload's)")
    ap.add_argument("--pool", type=int, default=0, help="distinct graphs generated per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue kernels eagerly instead of CUDA-graph replay")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--stages", action="store_true", help="also print the per-stage timing table to stderr")
    ap.add_argument("--resident", type=int, default=0, help="number of distinct device-resident batches (default: > L2)")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"],
                    help="N>1 gradient exchange: NVLink peer push fused with Adam, or ncclAllReduce")
    ap.add_argument("--no-stage-timing", action="store_true", help="skip the live per-stage timing / roofline")
    ap.add_argument("--steps-per-graph", type=int, default=1,
                    help="resident loop: consecutive steps captured as one CUDA graph (Trainer.step_many); 1 = one graph per step")
    ap.add_argument("--no-prep-ahead", action="store_true",
                    help="do not overlap the next batch's structure preparation with this step's update")
    ap.add_argument("--same-batches", action="store_true",
                    help="N > 1: every rank steps the SAME batches (rank 0's): the step time then shows the pure cost of the "
                         "gradient exchange, without the max-over-ranks of data-dependent step times (an experiment, not the metric)")
    ap.add_argument("--readout-bf16", action="store_true",
                    help="readout MLP products with bf16 operands / fp32 accumulate (BASELINE.json configs[4]: "
                         "'bf16 MLP / fp32 aggregate'); everything else stays fp32")
    return ap.parse_args()


def model_args(hidden=128, layers=3, readout_bf16=False):
    return argparse.Namespace(layers=layers, hidden=hidden, with_random=True, without_node_attention=False,
                              without_edge_attention=False, fc_num="222", cat_or_add="add", c=0.5, o=1.0, co=0.5,
                              eval_random=False, readout_bf16=readout_bf16)


def build_batches(workload, batch_size, n_batches, pool, seed):
    """`n_batches` distinct batches, each `batch_size` graphs sampled (without replacement inside a
    batch) from a pool of `pool` seeded SPMotif-style graphs."""
    from cal_b200.data import CONFIGS, Batch, make_dataset
    cfg = dict(CONFIGS[workload])
    cfg.pop("batch_size")
    ds = make_dataset(pool, seed=seed, bias=0.9, **cfg)
    rng = np.random.RandomState(seed + 1)
    out = []
    for _ in range(n_batches):
        idx = rng.choice(pool, size=batch_size, replace=False)
        out.append(Batch.from_data_list([ds[i] for i in idx]))
    return out, cfg


class ClockSampler:
    """SM clock / throttle reasons sampled with NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.002):
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                names = [name for bit, name in self.REASONS.items() if r & bit and name != "gpu_idle"]
                self.sm.append((time.perf_counter(), mhz, names))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self, t0=None, t1=None):
        """Median SM clock and the throttle reasons seen inside the host-time window [t0, t1] (the timed
        region); `samples_total` counts every sample taken (warm-up included)."""
        inside = [x for x in self.sm if t0 is None or (t0 <= x[0] <= t1)]
        use = inside or self.sm
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "samples_total": 0}
        return {"sm_mhz": float(np.median([x[1] for x in use])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted({n for x in use for n in x[2]}), "samples": len(inside), "samples_total": len(self.sm),
                "window": "timed region" if inside else "warm-up + timed region (no sample fell inside the timed region)"}


def visible_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# algorithmic bytes (DESIGN.md "Kernels and rooflines")
# ------------------------------------------------------------------------------------------------

def algorithmic_bytes(stage, N, E1, B, H, F, C, L, P):
    """Bytes a stage must move at minimum for a batch with N nodes, E1 = kept edges + N appended
    self loops, B graphs.  fp32 activations, int32 CSR.  One neighbour row per edge per propagate."""
    NH, EH = 4 * N * H, 4 * E1 * H
    if stage == "prep":
        return 16 * (E1 - N) + 8 * N + 4 * 7 * E1 + 4 * 6 * N           # edge_index + batch in; 2 CSRs + norm out
    if stage == "param_prep":
        return 8 * ((L + 2) * H * H + 3 * H * H)
    if stage == "feat":
        return 4 * N * F + NH + 4 * F * H
    if stage.startswith("layer_") and not stage.endswith("_bwd"):
        return EH + NH + 12 * E1 + 4 * N + 4 * H * H
    if stage == "edge_att":
        return 16 * N + 16 * E1 + 8 * E1 + 8 * E1 + 8 * N
    if stage == "masked_convs":
        return 2 * (EH + 2 * NH + 12 * E1 + 4 * N + 4 * H * H)
    if stage == "readout":
        return 2 * NH + 8 * B * H * 4 + 3 * 4 * H * H
    if stage == "readout_bwd":
        return 12 * B * H * 4 + 3 * 8 * H * H
    if stage == "masked_gemm_bwd":
        return 2 * (3 * NH + 8 * H * H)
    if stage == "masked_gather_bwd":
        return 2 * (EH + 2 * NH + 16 * E1)
    if stage == "norm_bwd":
        return 2 * 40 * E1
    if stage == "att_bwd":
        return 4 * NH + 8 * E1
    if stage.endswith("_bwd") and stage.startswith("layer_"):
        return EH + 4 * NH + 8 * E1 + 8 * H * H
    if stage == "feat_bwd":
        return 2 * NH + 4 * N * F + 4 * F * H
    if stage == "grad_reduce":
        return 8 * P
    if stage == "adam":
        return 28 * P
    return 0


def step_algorithmic_bytes(N, E1, B, H, F, L, P):
    """A = 16(L+3)NH + 8H(L+2)E' + 8HE' + (4NF + 16E + 8N + 8B) + 12P + 32BH."""
    E = E1 - N
    return 16 * (L + 3) * N * H + 8 * H * (L + 2) * E1 + 8 * H * E1 + (4 * N * F + 16 * E + 8 * N + 8 * B) + 12 * P + 32 * B * H


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (reference op order) on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_train_steps(batches, F, C, hidden, layers, steps, warmup, seconds=None, threads=None, model="CausalGCN"):
    """Times oracle train steps (forward + loss + backward + torch Adam).  -> (graphs/s, steps, s, threads)."""
    import random
    from oracle import cal_oracle
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(666)
    random.seed(666)
    args = model_args(hidden, layers)
    net = getattr(cal_oracle, model)(F, C, args)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    i = 0
    for _ in range(warmup):
        cal_oracle.train_step(net, batches[i % len(batches)])
        opt.step()
        i += 1
    graphs, done = 0, 0
    t0 = time.perf_counter()
    while True:
        b = batches[i % len(batches)]
        cal_oracle.train_step(net, b)
        opt.step()
        graphs += int(b.num_graphs)
        done += 1
        i += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (seconds is not None and el >= seconds):
            break
    return graphs / el, done, el, torch.get_num_threads()


def run_reference(a, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path.  torch_geometric /
    torch_scatter are not installable in this image, so the timed code is the oracle
    port, which replays the reference's op sequence (index_select -> mul -> index_add_, norm
    recomputed in every GCNConv) with all host threads.  Rank 0 only."""
    if rank != 0:
        return
    from cal_b200.data import CONFIGS
    bs = a.batch or CONFIGS[a.workload]["batch_size"]
    # the first 16 batches of rank 0 of the GPU arm (same pool, same seeds)
    batches, cfg = build_batches(a.workload, bs, 16, pool_size(a, bs), 666)
    F, C = batches[0].feat.size(1), cfg["num_classes"]
    cores = os.cpu_count() or 1
    gps, done, el, thr = cpu_train_steps(batches, F, C, 128, 3, a.steps, a.warmup, threads=cores, model=a.model)
    line = {
        "impl": "reference", "metric": METRIC, "value": gps, "unit": UNIT, "n_gpus": a.gpus, "steps": done,
        "warmup": a.warmup, "ms_per_step": 1e3 * el / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, bs, batches), "sample": sample_stats(batches, "host memory (CPU run)"),
        "cpu_baseline": {"value": gps, "unit": UNIT, "cores": thr, "kind": "port",
                         "sample": "%d train steps of %d graphs, %d torch threads" % (done, bs, thr)},
        "e2e": {"value": gps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


def workload_config(a, bs, batches):
    """The workload both arms run -- byte-identical between `--impl cal_b200` and `--impl reference`
    (per-sample statistics and the residency note live in the sibling `sample` key)."""
    return {"workload": "%s SPMotif-style synthetic (bias 0.9), %s 3 layers hidden 128, batch %d per GPU"
                        % (a.workload, a.model, bs),
            "model_step": "prep + forward + KL/NLL loss + backward + grad all-reduce (N>1) + Adam",
            "graphs_per_batch": bs, "features": int(batches[0].feat.size(1)), "hidden": 128, "layers": 3,
            "generator": "cal_b200.data.make_dataset(pool=%d, seed=666 + rank, bias=0.9); batches drawn without "
                         "replacement from the pool by RandomState(667 + rank)" % pool_size(a, bs),
            "parallelism": "dp%d" % a.gpus + (" (experiment: every rank steps rank 0's batches)" if getattr(a, "same_batches", False) else "")}


def pool_size(a, bs):
    return a.pool or max(8192, 4 * bs)


def sample_stats(batches, residency):
    return {"batches": len(batches), "avg_nodes_per_batch": float(np.mean([b.batch.numel() for b in batches])),
            "avg_edge_columns_per_batch": float(np.mean([b.edge_index.size(1) for b in batches])), "cache": residency}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------

def dp_verify(tr, packed_dev, dist, dev, world):
    """Data-parallel correctness inside the driver's own multi-GPU run (SURVEY.md 8e parity rule):
    (1) every rank's flat parameter buffer has the same raw bits (xor checksum + fp64 sum, all-gathered);
    (2) one more step, recomputed independently: the ranks' local gradients are all-gathered, averaged
        in rank order and pushed through torch.optim.Adam's update formula in fp64 from the saved state;
        the parameters the fused exchange + Adam kernel produced must agree."""
    eng = tr.eng
    torch.cuda.synchronize(dev)

    def checksum():
        bits = eng.flat.view(torch.int32).clone()
        n = 1
        while n < bits.numel():
            n *= 2
        pad = torch.zeros(n, dtype=torch.int32, device=dev)
        pad[:bits.numel()] = bits
        while n > 1:
            n //= 2
            pad = pad[:n] ^ pad[n:2 * n]
        t = torch.stack([pad[0].double(), eng.flat.double().sum()])
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [[float(v) for v in o.tolist()] for o in out]

    cs0 = checksum()
    p0 = eng.flat.clone()
    m0, v0, step = [t.clone() for t in eng.opt_state]
    t_next = int(step[0]) + 1
    tr.step(packed_dev)
    torch.cuda.synchronize(dev)
    g = eng.flat_grad.clone()
    if tr.collective == "peer":                      # flat_grad holds the local gradient
        parts = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(parts, g)
        acc = torch.zeros_like(g, dtype=torch.float64)
        for q in range(world):
            acc += parts[q].double()
        gm = acc / world
    else:                                            # ncclAllReduce summed it in place
        gm = g.double() / world
    b1, b2 = tr.betas
    lr = float(tr.lr_dev.item())
    m = m0.double() + (gm - m0.double()) * (1.0 - b1)
    v = v0.double() * b2 + (1.0 - b2) * gm * gm
    bc1, bc2 = 1.0 - b1 ** t_next, 1.0 - b2 ** t_next
    want = p0.double() - (lr / bc1) * m / (v.sqrt() / (bc2 ** 0.5) + tr.eps)
    err = float((want - eng.flat.double()).abs().max())
    cs1 = checksum()
    return {"replicas_bit_identical": all(c == cs0[0] for c in cs0) and all(c == cs1[0] for c in cs1),
            "checksums_xor_sum_per_rank": cs1,
            "adam_step_abs_err_vs_fp64_recomputation": err,
            "adam_step_err_rel_to_update": err / max(float((want - p0.double()).abs().max()), 1e-30),
            "adam_step_err_rel_to_params": err / max(float(want.abs().max()), 1e-30),
            "how": "after the timed region: xor + fp64-sum checksum of the flat parameter buffer all-gathered over the "
                   "ranks; then one extra step whose update is recomputed in fp64 from the all-gathered local gradients "
                   "(rank-order mean) and the saved Adam state"}


def run_gpu(a, rank, local_rank, world):
    import cal_b200
    from cal_b200.data import CONFIGS
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the cal_b200 arm has no CPU fallback); use --impl reference")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    bs = a.batch or CONFIGS[a.workload]["batch_size"]
    # enough distinct resident batches that the inputs of the timed region exceed L2 (no batch is
    # reused within ~L2 worth of input bytes) -- the "inputs larger than L2" timing rule
    drank = 0 if a.same_batches else rank
    probe, cfg = build_batches(a.workload, bs, 2, 4 * bs, 666 + drank)
    caps_probe = cal_b200.batch_caps(probe, slack=1.15)
    lay_probe = cal_b200.PackedLayout(*caps_probe[:3], probe[0].feat.size(1))
    n_res = a.resident or int(min(max(L2_BYTES // lay_probe.nbytes + 8, 16), 1024))
    if dist is not None:                      # every rank must issue the same number of collectives
        t = torch.tensor([n_res], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_res = int(t.item())
    pool = pool_size(a, bs)
    batches, cfg = build_batches(a.workload, bs, n_res, pool, 666 + drank)
    F, C = int(batches[0].feat.size(1)), cfg["num_classes"]
    torch.manual_seed(666)
    import random
    random.seed(666 + rank)
    if a.model == "CausalGCN":
        net = cal_b200.CausalGCN(F, C, model_args(readout_bf16=a.readout_bf16)).to(dev)
    elif a.model == "CausalGIN":
        net = cal_b200.CausalGIN(F, C, model_args(readout_bf16=a.readout_bf16)).to(dev)
    else:
        net = cal_b200.CausalGAT(F, C, model_args(readout_bf16=a.readout_bf16)).to(dev)
    net.train()
    tr = cal_b200.Trainer(net, cal_b200.batch_caps(batches), lr=1e-3, process_group=True if world > 1 else None,
                          use_graph=not a.no_graph, collective=a.collective)
    hosts = [tr.pack(b) for b in batches[:min(64, n_res)]]               # pinned host copies (e2e)
    resident = [tr.pack(b).to(dev) for b in batches]
    resident_bytes = sum(int(r.numel()) for r in resident)
    graphs_per_step = bs

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # The loop knows its next batch, so each step prepares the NEXT batch's structure beside its own update
    # (Trainer.step(cur, next)): every batch is still prepared exactly once per step, inside the timed region; and it
    # issues `--steps-per-graph` consecutive steps per graph launch (Trainer.step_many: same kernels, same order).
    # The warm-up + timed sequence runs once untimed first, which captures exactly the graphs the timed run replays.
    ahead = not a.no_prep_ahead
    spg = max(1, a.steps_per_graph) if (ahead and not a.no_graph) else 1
    W = max(a.warmup, 3)

    def run(start, count):
        i = 0
        while i < count:
            k = min(spg, count - i)
            group = [resident[(start + i + j) % n_res] for j in range(k)]
            follower = resident[(start + i + k) % n_res] if ahead else None
            if k > 1:
                tr.step_many(group, follower)
            else:
                tr.step(group[0], follower)
            i += k

    run(0, W)
    run(W, a.steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # clocks are sampled (NVML, every 2 ms) from the warm-up on: a 20-step timed region is only ~3 ms long
    with ClockSampler(visible_gpu_index(local_rank)) as clk:
        run(0, W)
        barrier()
        t_begin = time.perf_counter()
        e0.record()
        run(W, a.steps)
        e1.record()
        barrier()
        t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = a.steps * graphs_per_step * world / (ms * 1e-3)
    loss_after = tr.metrics().cpu().tolist()
    lps = int(tr.launches_per_step)                  # kernels per step of the timed loop (with the look-ahead: the next batch's structure kernel included)
    launches = lps * a.steps
    dp_check = dp_verify(tr, resident[0], dist, dev, world) if dist is not None else None

    # ---- end to end: pinned host batch -> H2D -> step -> loss parts D2H, every step ----
    e2e = None
    if not a.no_e2e:
        def timed(fn):
            barrier()
            e0.record()
            last = None
            for i in range(a.steps):
                last = fn(hosts[i % len(hosts)])
            tr.pipe_flush()                              # (the pipelined path issues a batch's step one call late)
            e1.record()
            barrier()
            tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()), last
        for i in range(max(3, a.warmup // 2)):
            tr.step_host(hosts[i % len(hosts)])
        for i in range(a.steps):                         # the timed sequence once untimed: its graphs (one per staging
            tr.step_host_async(hosts[i % len(hosts)])    # buffer pair, + the first and the flushed step) get captured
        tr.pipe_reset()
        # (a) pipelined: H2D of batch i+1 on a copy stream overlaps step i; loss parts of every step
        #     land in a pinned ring; the host only waits when the ring (16 steps) is full
        ms_p, slot = timed(tr.step_host_async)
        loss_last = tr.pipe_result(slot).tolist()
        # (b) the reference loop's shape: host synchronises on every step's loss (train_causal.py:186-191)
        ms_s, _ = timed(lambda hb: tr.step_host(hb, sync=True))
        e2e = {"value": a.steps * graphs_per_step * world / (ms_p * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(tr.layout.nbytes), "d2h_bytes_per_step": 32,
               "ms_per_step": ms_p / a.steps,
               "api": "Trainer.step_host_async(pinned packed batch): cudaMemcpyAsync H2D on a copy stream into a "
                      "ring of 16 staging buffers + captured step (the follower's structure preparation rides beside the "
                      "update, the loss parts' D2H copy into the buffer's pinned result slot beside the backward pass), every step; "
                      "all K steps issued and flushed inside the timed region",
               "sync_every_step": {"value": a.steps * graphs_per_step * world / (ms_s * 1e-3),
                                   "ms_per_step": ms_s / a.steps,
                                   "api": "Trainer.step_host(..., sync=True): same copies, host waits for every step's loss"},
               "last_loss_parts": loss_last[:4]}

    # ---- live per-stage timing + roofline of the dominant kernel (rank 0) ----
    roofline, stage_tab = None, None
    if rank == 0 and not a.no_stage_timing:
        b0 = batches[0]
        st = tr.profile_stages(resident[0], reps=20)
        eng = tr.eng
        N = int(b0.batch.numel())
        ei = b0.edge_index
        E1 = int((ei[0] != ei[1]).sum()) + N
        P = int(eng.total)
        stage_tab = []
        for name, nl, t_ms in st:
            ab = algorithmic_bytes(name, N, E1, bs, eng.H, eng.F, eng.C, eng.L, P)
            stage_tab.append({"stage": name, "launches": nl, "us": 1e3 * t_ms, "alg_bytes": ab,
                              "gbs": ab / (t_ms * 1e-3) / 1e9 if t_ms > 0 else None})
        # fused small-graph path (csrc/fsg.cu, fsg_bwd.cu): ONE kernel runs the stages feat .. masked_convs
        # (masked_gemm_bwd .. feat_bwd); the stages it absorbs launch nothing -- their algorithmic bytes
        # belong to the kernel that does their work
        merged, owner = [], None
        for r in stage_tab:
            if r["launches"] == 0 and owner is not None and owner["stage"] in ("feat", "masked_gemm_bwd", "fsg_forward", "fsg_backward"):
                owner["alg_bytes"] += r["alg_bytes"]
                owner["us"] += r["us"]
                owner.setdefault("absorbs", []).append(r["stage"])
                owner["stage"] = "fsg_forward" if owner["stage"] in ("feat", "fsg_forward") else "fsg_backward"
                continue
            owner = dict(r)
            merged.append(owner)
        for r in merged:
            r["gbs"] = r["alg_bytes"] / (r["us"] * 1e-6) / 1e9 if r["us"] > 0 else None
        stage_tab = merged
        # dominant KERNEL = the kernel family with the largest share of the step (the three backbone
        # layers run the same kernel); achieved = its algorithmic bytes / its time, per launch
        fam = {}
        for r in stage_tab:
            n = r["stage"]
            key = ("k_conv_bwd" if n.startswith("layer_") and n.endswith("_bwd") else
                   "k_conv_fwd" if n.startswith("layer_") else n)
            f = fam.setdefault(key, {"stage": key, "us": 0.0, "alg_bytes": 0, "launches": 0, "n": 0})
            f["us"] += r["us"]; f["alg_bytes"] += r["alg_bytes"]; f["launches"] += r["launches"]; f["n"] += 1
        # (a stage of two launches -- a GATConv / GINConv layer -- counts as one family, timed as a whole)
        top = max(fam.values(), key=lambda f: f["us"])
        top = {"stage": top["stage"], "us": top["us"] / top["n"], "alg_bytes": top["alg_bytes"] // top["n"],
               "gbs": top["alg_bytes"] / (top["us"] * 1e-6) / 1e9, "share_of_step": top["us"] / sum(r["us"] for r in stage_tab)}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        A_step = step_algorithmic_bytes(N, E1, bs, eng.H, eng.F, eng.L, P)
        # DRAM traffic of the same kernel from the committed ncu --set full capture (profiles/ncu_traffic.json)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj.get("workload") == a.workload and a.model == "CausalGCN" and top["stage"] in tj["kernels"]:
                k = tj["kernels"][top["stage"]]
                traffic = int(k["dram_read"]) + int(k["dram_write"])
            else:                                        # the other workloads: "others": {"<workload>/<model>": {family: ...}}
                k = tj.get("others", {}).get("%s/%s" % (a.workload, a.model), {}).get(top["stage"])
                if k is not None:
                    traffic = int(k["dram_read"]) + int(k["dram_write"])
        except Exception:
            traffic = None
        roofline = {"bound": "hbm", "kernel": top["stage"], "achieved": top["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": top["gbs"] / peak, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650",
                    "kernel_us": top["us"], "kernel_alg_bytes": top["alg_bytes"], "kernel_share_of_step": top["share_of_step"],
                    "step_alg_bytes": A_step, "step_achieved_gbs": A_step * a.steps / (ms * 1e-3) / 1e9 / 1.0,
                    "step_frac": A_step * a.steps / (ms * 1e-3) / 1e9 / peak,
                    "sum_stage_us": sum(r["us"] for r in stage_tab)}
        if a.stages:
            for r in stage_tab:
                print("%-20s launches %d  %8.2f us  %10d B  %8.1f GB/s" % (r["stage"], r["launches"], r["us"],
                                                                          r["alg_bytes"], r["gbs"] or 0), file=sys.stderr)

    # ---- device-resident epochs: the dataset lives in HBM, cal_collate builds
    # every mini-batch on the GPU, one captured graph (collate + step) serves the whole epoch; per
    # epoch the host uploads the shuffled order and the random-intervention permutations, per step nothing
    if e2e is not None:
        from cal_b200.data import make_dataset
        dcfg = dict(cfg)
        ds = make_dataset(pool, seed=666 + rank, bias=0.9, **dcfg)
        store = cal_b200.GraphStore(ds, dev)
        tr2 = cal_b200.Trainer(net, store.caps(bs), lr=1e-3, process_group=True if world > 1 else None,
                               use_graph=not a.no_graph, collective=a.collective)
        rng = np.random.RandomState(777 + rank)
        per_epoch = len(ds) // bs
        def run_epochs(n_steps):
            done = 0
            while done < n_steps:
                order = rng.permutation(len(ds))[:per_epoch * bs]
                tr2.begin_epoch(store, order, bs)
                k = min(per_epoch, n_steps - done)
                for _ in range(k):
                    tr2.step_epoch()
                done += k
            return tr2.end_epoch()
        run_epochs(max(3, a.warmup, per_epoch + 2))     # a whole epoch and the start of the next: every graph of the
        #                                                  epoch path (first / steady per buffer / last step) is captured untimed
        barrier()
        e0.record()
        m_ep = run_epochs(a.steps)
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e = float(tt.item())
        # what the reference's loader does per step on the host: collate (Batch.from_data_list) + packing
        import time as _time
        from cal_b200.data import Batch as _Batch
        hc_buf = torch.zeros(tr2.layout.nbytes, dtype=torch.uint8).pin_memory()
        t0 = _time.perf_counter()
        n_hc = 0
        while n_hc < 8 or _time.perf_counter() - t0 < 0.5:
            ids = rng.permutation(len(ds))[:bs]
            tr2.pack(_Batch.from_data_list([ds[i] for i in ids]), out=hc_buf)
            n_hc += 1
        host_collate_ms = (_time.perf_counter() - t0) / n_hc * 1e3
        e2e["device_resident_epoch"] = {
            "host_collate_ms_per_batch": host_collate_ms,
            "host_collate_bound_graphs_per_s": bs / (host_collate_ms * 1e-3),
            "value": a.steps * graphs_per_step * world / (ms_e * 1e-3), "unit": UNIT, "ms_per_step": ms_e / a.steps,
            "h2d_bytes_per_step": int((4 * per_epoch * bs + (4 * per_epoch * bs if tr2.with_random else 0)) / per_epoch),
            "d2h_bytes_per_epoch": 32, "graphs_in_store": len(ds), "steps_per_epoch": per_epoch,
            "api": "Trainer.begin_epoch(store, order) + step_epoch(): cal_collate on the GPU + the step in one captured "
                   "graph (the next batch is collated and prepared on forked branches of the current step); per epoch the shuffled order and the random-intervention permutations are uploaded, epoch "
                   "metrics are accumulated on the device and read once (end_epoch)",
            "last_epoch_metrics": m_ep}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        gps, done, el, thr = cpu_train_steps(batches[:16], F, C, 128, 3, None, 3, seconds=a.cpu_seconds, threads=cores,
                                             model=a.model)
        cpu = {"value": gps, "unit": UNIT, "cores": thr, "kind": "port",
               "sample": "%d oracle train steps of %d graphs in %.1f s (same workload, %d torch threads)" % (done, bs, el, thr)}
        if cores > 1:                                    # the CPU arm at one thread as well
            try:
                g1, d1, e1_, _ = cpu_train_steps(batches[:16], F, C, 128, 3, None, 1, seconds=min(4.0, a.cpu_seconds),
                                                 threads=1, model=a.model)
                cpu["single_thread"] = {"value": g1, "cores": 1, "sample": "%d steps in %.1f s" % (d1, e1_)}
            finally:
                torch.set_num_threads(cores)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (readout MLP products: bf16 operands, fp32 accumulate)" if a.readout_bf16 else "f32", "data": "synthetic",
            "config": workload_config(a, bs, batches),
            "sample": sample_stats(batches, "inputs larger than L2: %d distinct resident batches (%.0f MB) cycled; "
                                            "workspace reused" % (n_res, resident_bytes / 2 ** 20)),
            "clocks": clk.summary(t_begin, t_end), "e2e": e2e, "gpu_launches": launches,
            "launches_per_step": lps, "cuda_graph": not a.no_graph, "steps_per_graph_launch": spg,
            "collective": {"none": "none (1 GPU)", "nccl": "ncclAllReduce of the flat gradient buffer between a compute and an update graph",
                           "peer": "NVLink peer-memory push of the gradient chunks fused with Adam (cal_dp_adam_step), one captured graph per step"}[tr.collective],
            "dp_check": dp_check, "roofline": roofline, "cpu_baseline": cpu, "stages": stage_tab,
            "loss_after": loss_after[:4],
        }
        print(json.dumps(line), file=OUT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


OUT = sys.stdout


def main():
    # stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr, so that
    # whatever a library writes to fd 1 (NCCL's "NCCL version ..." banner under torchrun) cannot precede it
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        raise SystemExit("bench.py: --gpus %d needs a torchrun launch (WORLD_SIZE=%d)" % (a.gpus, world))
    run_gpu(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
