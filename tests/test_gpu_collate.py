"""Device-side mini-batch collation (cal_collate, csrc/collate.cu) against the host collate rule of
the reference's loader (torch_geometric DataLoader -> Batch.from_data_list, train_causal.py:13-15,
171-176; cal_b200.data.Batch.from_data_list restates it): bit-exact integers, verbatim features;
and a whole device-resident epoch against the same steps fed from host-collated packed batches."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.util import clone_to_cuda, random_case  # noqa: E402

DEV = "cuda:0"


def _dataset(n=70, seed=5, **kw):
    from cal_b200.data import make_dataset
    return make_dataset(n, seed=seed, avg_nodes=14, **kw)


def _collate_once(M, store, order, pos0, B, caps, perm_pool=None, advance=1):
    import cal_b200._lib as L
    lib = L.load()
    lay = M.PackedLayout(caps[0], caps[1], caps[2], store.F)
    out = torch.zeros(lay.nbytes, dtype=torch.uint8, device=DEV)
    cb = lay.cbatch(out.data_ptr())
    c = L.Caps()
    c.max_nodes, c.max_edges, c.max_graphs = caps
    od = torch.as_tensor(np.asarray(order, dtype=np.int32), device=DEV)
    pos = torch.tensor([pos0, 0, 0, 0], dtype=torch.int32, device=DEV)
    pp = None if perm_pool is None else torch.as_tensor(np.asarray(perm_pool, dtype=np.int32), device=DEV)
    rc = lib.cal_collate(C.byref(store.desc), od.data_ptr(), len(order), pos.data_ptr(), B,
                         pp.data_ptr() if pp is not None else 0, C.byref(c), C.byref(cb), advance, 0, 0,
                         torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.cal_error_string(rc)
    torch.cuda.synchronize()
    return lay, out.cpu().numpy(), pos.cpu().tolist()


def _fields(lay, a):
    N, E, B, _ = a[lay.off_dims:lay.off_dims + 16].view(np.int32)
    feat = a[lay.off_feat:lay.off_feat + 4 * N * lay.F].view(np.float32).reshape(N, lay.F)
    ei = a[lay.off_ei:lay.off_ei + 16 * lay.Em].view(np.int64)
    return dict(N=int(N), E=int(E), B=int(B), feat=feat, row=ei[:E], col=ei[lay.Em:lay.Em + E],
                batch=a[lay.off_batch:lay.off_batch + 8 * N].view(np.int64),
                y=a[lay.off_y:lay.off_y + 8 * B].view(np.int64),
                perm=a[lay.off_perm:lay.off_perm + 4 * B].view(np.int32))


@pytest.mark.parametrize("B", [1, 8, 32])
def test_collate_bit_exact_with_host_collate(B):
    import cal_b200 as M
    ds = _dataset()
    store = M.GraphStore(ds, DEV)
    rng = np.random.RandomState(3)
    order = rng.permutation(len(ds))[:67]                   # 67 is not a multiple of any B > 1: short last batch
    caps = store.caps(B)[:3]
    for start in range(0, len(order), B):
        ids = order[start:start + B]
        want = M.Batch.from_data_list([ds[i] for i in ids])
        perm = rng.permutation(len(ids))
        pool = np.zeros((-(-len(order) // B), B), dtype=np.int32)
        pool[start // B, :len(ids)] = perm
        lay, a, pos = _collate_once(M, store, order, start, B, caps, perm_pool=pool)
        f = _fields(lay, a)
        assert (f["N"], f["E"], f["B"]) == (want.batch.numel(), want.edge_index.size(1), len(ids))
        assert np.array_equal(f["feat"], want.feat.numpy())
        assert np.array_equal(f["row"], want.edge_index[0].numpy())
        assert np.array_equal(f["col"], want.edge_index[1].numpy())
        assert np.array_equal(f["batch"], want.batch.numpy())
        assert np.array_equal(f["y"], want.y.numpy())
        assert np.array_equal(f["perm"], perm)
        assert pos[0] == start + len(ids) and pos[1] == 0       # cursor advanced, arrival counter reset


def test_collate_edge_cases():
    import cal_b200 as M
    import cal_b200._lib as L
    ds = _dataset(12)
    store = M.GraphStore(ds, DEV)
    caps = store.caps(4)[:3]
    # cursor at the end of the order: an empty batch, nothing advanced
    lay, a, pos = _collate_once(M, store, list(range(12)), 12, 4, caps)
    f = _fields(lay, a)
    assert (f["N"], f["E"], f["B"]) == (0, 0, 0) and pos[0] == 12
    # no perm pool -> identity; advance = 0 keeps the cursor
    lay, a, pos = _collate_once(M, store, list(range(12)), 4, 4, caps, advance=0)
    f = _fields(lay, a)
    assert np.array_equal(f["perm"], np.arange(4)) and pos[0] == 4
    # capacities too small: dims report the true sizes (cal_prep then raises CAL_ST_CAPACITY), nothing is written
    tiny = (32, 32, 8)
    lay, a, pos = _collate_once(M, store, list(range(12)), 0, 4, tiny)
    f_n = a[lay.off_dims:lay.off_dims + 16].view(np.int32)
    want = M.Batch.from_data_list(ds[:4])
    assert int(f_n[0]) == want.batch.numel() > 32
    assert not a[lay.off_feat:].any()
    # argument errors
    lib = L.load()
    assert lib.cal_collate(None, 0, 0, 0, 4, 0, None, None, 0, 0, 0, 0) == -2          # CAL_ENULL


def test_epoch_on_device_equals_host_collated_steps():
    """Trainer.begin_epoch / step_epoch (collate + step in one captured graph, no per-step H2D) vs the
    same steps through Trainer.step_host on host-collated batches: bit-identical parameters and the
    epoch metrics of train_causal.py:186-196."""
    import cal_b200 as M
    ds = _dataset(75, seed=9)
    B = 16
    ora, _, _ = random_case(seed=41, hidden=64, batch_size=B)
    rng = np.random.RandomState(1)
    order = rng.permutation(len(ds))                          # 75 = 4 * 16 + 11: short last batch
    n_steps = -(-len(order) // B)
    perms = np.zeros((n_steps, B), dtype=np.int32)
    for s in range(n_steps):
        bn = min(B, len(order) - s * B)
        perms[s, :bn] = rng.permutation(bn)
    store = M.GraphStore(ds, DEV)
    caps = store.caps(B)
    net_d, net_h = clone_to_cuda(ora, M), clone_to_cuda(ora, M)
    tr_d = M.Trainer(net_d, caps, lr=1e-3)
    tr_h = M.Trainer(net_h, caps, lr=1e-3)
    tot = np.zeros(8)
    for epoch in range(2):
        assert tr_d.begin_epoch(store, order, B, perms=perms) == n_steps
        for s in range(n_steps):
            tr_d.step_epoch()
            ids = order[s * B:(s + 1) * B]
            hb = M.Batch.from_data_list([ds[i] for i in ids])
            res = tr_h.step_host(tr_h.pack(hb, perm=perms[s, :len(ids)].tolist())).clone().numpy()
            if epoch == 1:
                tot[:4] += res[:4] * len(ids)
                tot[4:7] += res[4:7]
                tot[7] += len(ids)
        m = tr_d.end_epoch()
    for (n, p), (_, q) in zip(net_d.named_parameters(), net_h.named_parameters()):
        assert torch.equal(p, q), n
    assert m["graphs"] == len(order) == int(tot[7])
    assert abs(m["loss"] - tot[0] / tot[7]) < 1e-5 * max(1.0, abs(tot[0] / tot[7]))
    assert abs(m["acc_o"] - tot[5] / tot[7]) < 1e-6 and abs(m["acc_co"] - tot[6] / tot[7]) < 1e-6
    assert 1 <= len(tr_d._epoch["graphs"]) <= 6               # a handful of captured graphs served both epochs (first / steady per buffer / last)


def test_trainer_freezes_capacities_eval_cannot_reallocate():
    """a larger evaluation batch must not reallocate the workspace baked into the
    Trainer's captured CUDA graphs.  Engine.ensure_caps raises a CalError instead; training goes on."""
    import cal_b200
    from cal_b200 import _lib
    ora, b, perm = random_case(seed=401, hidden=32, batch_size=8)
    big = random_case(seed=402, hidden=32, batch_size=24)[1]
    net = clone_to_cuda(ora, cal_b200)
    tr = cal_b200.Trainer(net, cal_b200.batch_caps([b]), lr=1e-3)
    host = tr.pack(b, perm=perm.tolist())
    r1 = tr.step_host(host).clone()
    ws_ptr = tr.eng.ws.data_ptr()
    with pytest.raises(_lib.CalError, match="frozen"):
        tr.eval_batch(big.to(DEV))
    assert tr.eng.ws.data_ptr() == ws_ptr                 # nothing was reallocated
    outs = tr.eval_batch(b.to(DEV))                        # an evaluation batch that fits is fine
    assert torch.isfinite(outs[0]).all()
    r2 = tr.step_host(host).clone()                        # the captured graph still replays on live memory
    assert torch.isfinite(r2).all() and float(r2[0]) != float(r1[0])
    tr.check()
    # a second Trainer on the same model takes the workspace over; the first one refuses to replay
    tr2 = cal_b200.Trainer(net, cal_b200.batch_caps([b, big]), lr=1e-3)
    with pytest.raises(_lib.CalError, match="reallocated"):
        tr.step_host(host)
    tr2.step_host(tr2.pack(big, perm=list(range(24))))
    tr2.eval_batch(big.to(DEV))
    tr2.check()
