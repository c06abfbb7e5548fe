"""Host-side logic that needs no GPU: the packed batch layout, the collate rule and rank sharding
of the DataLoader stand-in, and the data-parallel gradient exchange on 2 gloo ranks."""
import os
import socket

import numpy as np
import pytest
import torch

import cal_b200
from cal_b200.data import Batch, DataLoader, make_batches, make_dataset
from tests.util import grad_or_zero, make_args, random_case, rel_err


def test_collate_rule_matches_pyg():
    ds = make_dataset(5, seed=3)
    b = Batch.from_data_list(ds)
    off = 0
    for g, d in enumerate(ds):
        n = d.num_nodes
        assert torch.equal(b.feat[off:off + n], d.feat)
        assert torch.all(b.batch[off:off + n] == g)
        off += n
    assert b.num_graphs == 5 and b.y.numel() == 5
    assert int(b.edge_index.max()) < off and b.edge_index.size(1) == sum(d.num_edges for d in ds)
    assert torch.all(b.batch[1:] >= b.batch[:-1])                       # non-decreasing (model.py:115 relies on it)
    # every edge stays inside its graph
    assert torch.equal(b.batch[b.edge_index[0]], b.batch[b.edge_index[1]])


def test_spmotif_style_statistics():
    bs = make_batches("spmotif", num_batches=4, seed=666)
    n = np.mean([b.batch.numel() / b.num_graphs for b in bs])
    e = np.mean([b.edge_index.size(1) / b.num_graphs for b in bs])
    assert 22 <= n <= 28 and 44 <= e <= 62, (n, e)                      # BASELINE.json: avg 25 nodes, 50 edges
    b = bs[0]
    assert b.feat.shape[1] == 10 and torch.all(b.feat.sum(1) == 1)      # one-hot degree, featgen.py:19-28
    ei = b.edge_index
    fwd = set(map(tuple, ei.t().tolist()))
    assert all((c, r) in fwd for r, c in fwd)                           # stored in both directions (utils.py:55)
    assert torch.all(ei[0, 1:] >= ei[0, :-1])                           # row-sorted, as from_networkx emits


def test_packed_layout_round_trip():
    b = make_batches("spmotif", num_batches=1, seed=5, batch_size=9)[0]
    N, E, B = b.batch.numel(), b.edge_index.size(1), 9
    lay = cal_b200.PackedLayout(N + 13, E + 7, 16, 10)
    buf = torch.zeros(lay.nbytes, dtype=torch.uint8)
    perm = list(reversed(range(B)))
    lay.pack(b, buf, perm)
    a = buf.numpy()
    assert tuple(a[:16].view(np.int32)[:3]) == (N, E, B)
    assert list(a[lay.off_perm:lay.off_perm + 4 * B].view(np.int32)) == perm
    assert np.array_equal(a[lay.off_feat:lay.off_feat + 4 * N * 10].view(np.float32).reshape(N, 10), b.feat.numpy())
    ev = a[lay.off_ei:lay.off_ei + 16 * lay.Em].view(np.int64)
    assert np.array_equal(ev[:E], b.edge_index[0].numpy()) and np.array_equal(ev[lay.Em:lay.Em + E], b.edge_index[1].numpy())
    assert np.array_equal(a[lay.off_batch:lay.off_batch + 8 * N].view(np.int64), b.batch.numpy())
    assert np.array_equal(a[lay.off_y:lay.off_y + 8 * B].view(np.int64), b.y.numpy())
    for o in (lay.off_perm, lay.off_feat, lay.off_ei, lay.off_batch, lay.off_y, lay.nbytes):
        assert o % 16 == 0
    cb = lay.cbatch(4096)
    assert cb.feat == 4096 + lay.off_feat and cb.edge_stride == lay.Em
    with pytest.raises(cal_b200._lib.CalError):
        cal_b200.PackedLayout(N - 1, E, 16, 10).pack(b, buf)
    assert cal_b200.batch_caps([b])[0] >= N


def test_batch_caps_declares_what_the_batches_guarantee():
    """batch_caps: capacities + the two promises the kernels may rely on -- small graphs (<= 40 nodes, <= 320 CSR
    entries: the fused path) and edge_index columns grouped by graph (what the collate produces: per-graph structure
    preparation); a batch that breaks a promise withdraws it."""
    import copy
    small = make_batches("spmotif", num_batches=2, seed=6, batch_size=12, avg_nodes=20)
    n, e, g, is_small, grouped = cal_b200.batch_caps(small)
    assert n >= max(b.batch.numel() for b in small) and n % 32 == 0
    assert e >= max(b.edge_index.size(1) for b in small) and g >= 12 and g % 8 == 0
    assert is_small and grouped
    big = make_batches("spmotif", num_batches=1, seed=7, batch_size=6, avg_nodes=90)[0]
    assert cal_b200.batch_caps([small[0], big])[3:] == (False, True)          # a 90-node graph: not "small", still grouped
    shuffled = copy.copy(small[0])
    shuffled.edge_index = small[0].edge_index[:, torch.randperm(small[0].edge_index.size(1), generator=torch.Generator().manual_seed(1))]
    assert cal_b200.batch_caps([shuffled])[3:] == (True, False)               # same graphs, columns out of graph order
    crossing = copy.copy(small[0])
    ei = small[0].edge_index.clone()
    ei[1, 0] = small[0].batch.numel() - 1                                      # an edge from graph 0 into the last graph
    crossing.edge_index = ei
    assert cal_b200.batch_caps([crossing])[4] is False


def test_dataloader_rank_sharding_is_a_partition():
    ds = make_dataset(50, seed=1)
    for g in ds:
        g.tag = None
    seen = []
    for r in range(4):
        dl = DataLoader(ds, batch_size=8, shuffle=True, seed=7, rank=r, world_size=4)
        got = [int(b.num_graphs) for b in dl]
        assert sum(got) == len(range(r, 50, 4)) and len(got) == len(dl)
        seen.append(sum(got))
    assert sum(seen) == 50
    # same seed => same epoch permutation on every rank => disjoint shards
    ids = []
    for r in range(2):
        dl = DataLoader(list(range(10)), batch_size=100, shuffle=True, seed=3, rank=r, world_size=2)
        idx = np.arange(10)
        np.random.RandomState(3).shuffle(idx)
        ids.append(set(idx[r::2].tolist()))
    assert ids[0].isdisjoint(ids[1]) and len(ids[0] | ids[1]) == 10


def test_flat_offsets_cover_reference_state_dict_order():
    from oracle import cal_oracle
    args = make_args(hidden=128)
    net = cal_oracle.CausalGCN(10, 4, args)
    offs, total = cal_b200.flat_offsets(net)
    assert list(offs) == [n for n, _ in net.named_parameters()]
    assert all(o % 4 == 0 for o in offs.values())
    n_params = sum(p.numel() for p in net.parameters())
    assert n_params == 138660                               # 138 660 parameters at H=128, F=10, C=4
    assert total >= n_params
    ours = cal_b200.CausalGCN(10, 4, args)
    assert [(n, tuple(p.shape)) for n, p in ours.named_parameters()] == \
           [(n, tuple(p.shape)) for n, p in net.named_parameters()]
    assert list(ours.state_dict().keys()) == list(net.state_dict().keys())


def test_module_signatures_extend_the_reference_signatures():
    """SURVEY.md 8(b) "Python signature to keep": constructor and forward of the three modules take the
    reference's parameters, in the reference's order, with the reference's defaults (model.py:14-22,85,168,234,
    316-320,380; frozen by tests/golden/make_signatures.py); ours may only APPEND keyword parameters."""
    import inspect
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    frozen = json.load(open(os.path.join(here, "golden", "signatures.json")))
    if os.path.isfile("/root/reference/model.py"):                       # the fixture is what the reference says
        from tests.golden.make_signatures import reference_signatures
        assert reference_signatures("/root/reference") == frozen
    assert sorted(frozen) == sorted("%s.%s" % (k, f) for k in ("CausalGCN", "CausalGAT", "CausalGIN")
                                    for f in ("__init__", "forward"))
    for key, ref in frozen.items():
        kind, fn = key.split(".")
        sig = inspect.signature(getattr(getattr(cal_b200, kind), fn))
        ours = list(sig.parameters.values())
        n, nd = len(ref["args"]), len(ref["defaults"])
        assert [p.name for p in ours[:n]] == ref["args"], key
        assert [p.default for p in ours[n - nd:n]] == ref["defaults"], key
        assert all(p.default is inspect.Parameter.empty for p in ours[1:n - nd]), key
        assert all(p.default is not inspect.Parameter.empty for p in ours[n:]), key   # extras are optional


@pytest.mark.parametrize("kind", ["CausalGCN", "CausalGAT", "CausalGIN"])
def test_seeded_construction_reproduces_reference_init(kind):
    """gcn_conv.py:37-42 / model.py:80-83: the same seed draws the same initial parameters in the same
    order (glorot weights, zero biases, BatchNorm weight 1 / bias 1e-4) as the reference construction,
    whose restatement the golden fixtures pin (tests/test_oracle_golden.py)."""
    from oracle import cal_oracle
    args = make_args(hidden=64, layers=2)
    torch.manual_seed(666)
    ref = getattr(cal_oracle, kind)(10, 4, args)
    torch.manual_seed(666)
    ours = getattr(cal_b200, kind)(10, 4, args)
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert list(sd_r) == list(sd_o)
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_o[k]), k
    assert float(ours.bnc.bias[0]) == pytest.approx(1e-4) and float(ours.bnc.weight[0]) == 1.0
    assert float(ours.convs[0].bias.abs().max() if kind != "CausalGIN" else 0.0) == 0.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from oracle import cal_oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    ora, _, _ = random_case(seed=100, hidden=32, batch_size=4)           # same seed => same replica
    ds = make_dataset(16, seed=9)
    g = torch.Generator().manual_seed(11)
    for d in ds:
        d.feat = d.feat + 0.25 * torch.randn(d.feat.shape, generator=g)
    batch = next(iter(DataLoader(ds, batch_size=8, shuffle=True, seed=4, rank=rank, world_size=world)))
    cal_oracle.train_step(ora, batch, perm=torch.arange(batch.num_graphs))
    offs, total = cal_b200.flat_offsets(ora)
    flat = torch.zeros(total)
    for n, p in ora.named_parameters():
        flat[offs[n]:offs[n] + p.numel()] = grad_or_zero(p).reshape(-1)
    local = flat.clone()
    scale = cal_b200.allreduce_flat_grads(flat)
    torch.save(dict(local=local, reduced=flat * scale, n=batch.num_graphs), os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_gloo_world2(tmp_path):
    """2 ranks: all-reduced, 1/world-scaled flat gradients are identical on both ranks and equal
    the average of the per-rank gradients (plain-DDP semantics, BatchNorm statistics per rank)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, "r%d.pt" % r)) for r in range(2))
    assert r0["n"] == 8 and r1["n"] == 8
    assert torch.equal(r0["reduced"], r1["reduced"])
    assert not torch.equal(r0["local"], r1["local"])
    want = (r0["local"].double() + r1["local"].double()) / 2
    assert rel_err(r0["reduced"], want) < 1e-6


def test_graph_store_layout_and_caps():
    """GraphStore (the device-resident dataset behind cal_collate) built on the CPU: pointer arrays,
    graph-local edge ids, and capacities that cover every step of an order."""
    import numpy as np
    import cal_b200 as M
    ds = M.make_dataset(23, seed=4, avg_nodes=12)
    st = M.GraphStore(ds, "cpu")
    n = np.array([d.num_nodes for d in ds])
    e = np.array([d.num_edges for d in ds])
    assert st.node_ptr.tolist() == [0] + np.cumsum(n).tolist()
    assert st.edge_ptr.tolist() == [0] + np.cumsum(e).tolist()
    assert st.feat.shape == (int(n.sum()), ds[0].feat.size(1)) and st.y.tolist() == [int(d.y) for d in ds]
    for g in (0, 7, 22):                                     # edges stay graph-local
        s0, s1 = int(st.edge_ptr[g]), int(st.edge_ptr[g + 1])
        assert torch.equal(st.edge_src[s0:s1].long(), ds[g].edge_index[0])
        assert int(st.edge_dst[s0:s1].max()) < n[g]
    order = np.random.RandomState(0).permutation(23)
    B = 5
    cn, ce, cb, _small = st.caps(B, order)
    for s in range(0, 23, B):
        b = M.Batch.from_data_list([ds[i] for i in order[s:s + B]])
        assert b.batch.numel() <= cn and b.edge_index.size(1) <= ce
    wn, we, wb, _small = st.caps(B)                                  # worst case covers any order
    assert wn >= cn and we >= ce and wb == cb == 8


def _order_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = {}
    for epoch in range(2):
        o = cal_b200.epoch_order(203, epoch, seed=7, rank=rank, world_size=world, graphs_per_step=16)
        steps = torch.tensor([len(o) // 16])
        dist.all_reduce(steps, op=dist.ReduceOp.MIN)                 # what a rank would check before training
        out[epoch] = (o, int(steps))
    torch.save(out, os.path.join(out_dir, "o%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_epoch_order_shards_like_a_distributed_sampler_gloo_world2(tmp_path):
    """The per-rank epoch order of the device-resident path: same permutation on every rank, disjoint
    shards, the same number of whole steps everywhere, a new shuffle every epoch."""
    import torch.multiprocessing as mp
    mp.spawn(_order_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = [torch.load(os.path.join(tmp_path, "o%d.pt" % k), weights_only=False) for k in range(2)]
    for epoch in range(2):
        (a, sa), (b, sb) = r[0][epoch], r[1][epoch]
        assert sa == sb == len(a) // 16 == len(b) // 16 == 6       # 203 // 2 = 101 graphs per rank -> 6 steps of 16
        assert len(a) == len(b) == 96 and a.dtype == np.int32
        assert not set(a.tolist()) & set(b.tolist())                # disjoint shards
        full = np.random.RandomState((7 * 1000003 + epoch) % (2 ** 31 - 1)).permutation(203)
        assert np.array_equal(a, full[0::2][:96]) and np.array_equal(b, full[1::2][:96])
    assert not np.array_equal(r[0][0][0], r[0][1][0])               # reshuffled between epochs
    solo = cal_b200.epoch_order(50, 3, seed=1)
    assert sorted(solo.tolist()) == list(range(50))                 # one rank, no step cut: a plain permutation


def test_cosine_lr_is_the_reference_schedule():
    """train_causal.py:22,29: CosineAnnealingLR(T_max=epochs, eta_min=min_lr), stepped once per epoch."""
    for lr0, lr_min, T in ((1e-3, 1e-6, 100), (1e-2, 0.0, 7)):
        w = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([w], lr=lr0)
        sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=T, eta_min=lr_min, last_epoch=-1)
        for epoch in range(T + 1):
            assert cal_b200.cosine_lr(epoch, lr0, lr_min, T) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-9, abs=1e-15)
            opt.step()
            sch.step()


def test_bench_algorithmic_bytes_follow_the_survey_formula():
    """A = 100.75 MB per step at cfg 1 (N = 3 200, E = 6 400, F = 10, H = 128, L = 3, B = 128,
    P = 138 660) -- the figure `roofline.achieved` is computed from; the per-stage split that credits the
    kernels must not exceed the step figure by more than its documented extras (parameter re-reads, CSR words)."""
    import bench
    N, E, B, H, F, C, L, P = 3200, 6400, 128, 128, 10, 4, 3, 138660
    A = bench.step_algorithmic_bytes(N, E + N, B, H, F, L, P)
    assert A == 16 * 6 * N * H + 8 * H * 5 * (E + N) + 8 * H * (E + N) + (4 * N * F + 16 * E + 8 * N + 8 * B) + 12 * P + 32 * B * H
    assert abs(A / 1e6 - 100.75) < 0.05
    stages = (["prep", "feat"] + ["layer_%d" % l for l in range(L)] + ["edge_att", "masked_convs", "readout", "readout_bwd",
              "masked_gemm_bwd", "masked_gather_bwd", "norm_bwd", "att_bwd"] + ["layer_%d_bwd" % l for l in range(L)] +
              ["feat_bwd", "grad_reduce", "adam"])
    per = {s: bench.algorithmic_bytes(s, N, E + N, B, H, F, C, L, P) for s in stages}
    assert all(v > 0 for v in per.values())
    total = sum(per.values())
    assert 0.95 * A < total < 1.35 * A, (total, A)
