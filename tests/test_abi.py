"""The C-ABI library loads and exports every symbol include/cal_b200.h declares; host-side
argument validation returns the documented CAL_E* codes before any launch (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cal_b200.h")


def _lib():
    from cal_b200 import _lib
    return _lib, _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cal_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L, lib = _lib()
    names = declared_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), "libcal_b200.so does not export %s" % n
    assert sorted(L.EXPORTS) == names, "cal_b200/_lib.py EXPORTS out of sync with the header"
    assert lib.cal_abi_version() == 3


def test_header_cites_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("gcn_conv.py:44-70", "model.py:85-122", "model.py:315-450", "train_causal.py:187", "train_causal.py:21,192"):
        assert cite in src, cite


def _desc(L, **kw):
    d = L.ModelDesc()
    d.model, d.num_features, d.hidden, d.num_classes, d.layers, d.heads = 0, 10, 128, 4, 3, 1
    d.bn_eps, d.bn_momentum, d.w_c, d.w_o, d.w_co = 1e-5, 0.1, 0.5, 1.0, 0.5
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def _caps(L, n=3500, e=8000, b=128):
    c = L.Caps()
    c.max_nodes, c.max_edges, c.max_graphs = n, e, b
    return c


def test_workspace_layout_host_only():
    L, lib = _lib()
    d, c = _desc(L), _caps(L)
    total = lib.cal_workspace_bytes(C.byref(d), C.byref(c))
    assert total > 0
    off, size = C.c_size_t(), C.c_size_t()
    prev_end = 0
    for r in range(len(L.WS_REGIONS)):
        assert lib.cal_workspace_region(C.byref(d), C.byref(c), r, C.byref(off), C.byref(size)) == 0
        assert off.value % 256 == 0 and off.value >= prev_end
        prev_end = off.value + size.value
    assert prev_end <= total
    assert lib.cal_workspace_region(C.byref(d), C.byref(c), len(L.WS_REGIONS), C.byref(off), C.byref(size)) == -1
    # the saved activations x_1..x_{L+1} are the largest forward region: (L+1) * Nm * H floats
    lib.cal_workspace_region(C.byref(d), C.byref(c), L.WS["X"], C.byref(off), C.byref(size))
    assert size.value == 4 * 3500 * 128 * 4
    # unsupported shapes -> 0 bytes
    for bad in (dict(hidden=48), dict(num_classes=1), dict(layers=0), dict(layers=9), dict(num_features=0), dict(model=7)):
        assert lib.cal_workspace_bytes(C.byref(_desc(L, **bad)), C.byref(c)) == 0, bad
    assert lib.cal_workspace_bytes(C.byref(d), C.byref(_caps(L, n=0))) == 0


def test_stage_names_and_error_strings():
    L, lib = _lib()
    d = _desc(L)
    nf = lib.cal_stage_count(C.byref(d), L.CAL_PASS_FORWARD)
    nb = lib.cal_stage_count(C.byref(d), L.CAL_PASS_BACKWARD)
    assert (nf, nb) == (9, 10)
    f = [lib.cal_stage_name(C.byref(d), 0, i).decode() for i in range(nf)]
    b = [lib.cal_stage_name(C.byref(d), 1, i).decode() for i in range(nb)]
    assert f == ["param_prep", "feat", "layer_0", "layer_1", "layer_2", "edge_att", "masked_convs", "readout", "copy_out"]
    assert b == ["readout_bwd", "masked_gemm_bwd", "masked_gather_bwd", "norm_bwd", "att_bwd", "layer_2_bwd",
                 "layer_1_bwd", "layer_0_bwd", "feat_bwd", "grad_reduce"]
    assert lib.cal_stage_count(C.byref(d), 5) == -1
    for code, frag in ((0, "ok"), (-1, "invalid"), (-2, "NULL"), (-3, "aligned"), (-4, "workspace"), (-5, "unsupported")):
        assert frag in lib.cal_error_string(code).decode()


def test_host_side_argument_validation_no_launch():
    """Every check below fails before the first kernel launch, so it is safe without a GPU."""
    L, lib = _lib()
    d, c = _desc(L), _caps(L)
    n0 = lib.cal_launch_count()
    # forward: NULL offsets / params
    assert lib.cal_causal_forward(C.byref(d), C.byref(c), None, None, 0, 0, 0, None, 0, 0, 0, 0, 0) == -2
    po = L.ParamOffsets()
    po.total = 0
    assert lib.cal_causal_forward(C.byref(d), C.byref(c), C.byref(po), None, 0, 0, 0, None, 0, 0, 0, 0, 0) == -1
    bad_model = _desc(L, hidden=100)
    assert lib.cal_causal_forward(C.byref(bad_model), C.byref(c), C.byref(po), None, 0, 0, 0, None, 0, 0, 0, 0, 0) == -5
    assert lib.cal_causal_backward(C.byref(bad_model), C.byref(c), C.byref(po), 0, None, 0, 0, 0, 0, 0, 0) == -5
    assert lib.cal_adam_step(0, 0, 0, 0, 10, 0, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 1.0, 0) == -2
    assert lib.cal_adam_tick(0, 0) == -2
    b = L.Batch()
    assert lib.cal_prep(C.byref(d), C.byref(c), C.byref(b), 0, 0, 0) == -2
    assert lib.cal_launch_count() == n0


def _offsets(L, layers=3):
    """Synthetic parameter offsets: every tensor 16 384 floats apart (aligned; only the host-side checks read them)."""
    po = L.ParamOffsets()
    k = 0
    for f, _t in po._fields_:
        if f == "total":
            continue
        v = getattr(po, f)
        if isinstance(v, int):
            setattr(po, f, k * 16384)
            k += 1
        else:
            for i in range(len(v)):
                v[i] = k * 16384 if i < layers or len(v) == 3 else -1
                k += 1
    po.total = k * 16384
    return po


def test_image_sink_describes_the_fused_path_images_host_only():
    """cal_image_sink_init (host only): one entry per conv matrix (forward + backward image), the input transform and the
    three fc1 matrices when the fused small-graph path is taken, none otherwise; argument errors; no launch."""
    L, lib = _lib()
    n0 = lib.cal_launch_count()
    d, po = _desc(L), _offsets(L)
    c = _caps(L)
    c.small_graphs = 1
    nbytes = lib.cal_workspace_bytes(C.byref(d), C.byref(c))
    ws = 1 << 20                                          # (never dereferenced on the host)
    sk = L.ImageSink()
    assert lib.cal_image_sink_init(C.byref(d), C.byref(c), C.byref(po), ws, nbytes, None) == -2
    assert lib.cal_image_sink_init(C.byref(d), C.byref(c), C.byref(po), 0, nbytes, C.byref(sk)) == -2
    assert lib.cal_image_sink_init(C.byref(d), C.byref(c), C.byref(po), ws, nbytes, C.byref(sk)) == 0
    assert sk.count == 3 + 2 + 1 + 3
    offs = [po.convs_w[0], po.convs_w[1], po.convs_w[2], po.context_w, po.objects_w, po.conv_feat_w,
            po.fc1_w[0], po.fc1_w[1], po.fc1_w[2]]
    img = 2 * 16384 * 4                                   # bytes of one image (hi | lo)
    seen = set()
    for q in range(sk.count):
        e = sk.entry[q]
        assert e.offset == offs[q] and e.rows == (10 if q == 5 else 128)
        for p in (e.dst_t, e.dst_n):
            if p:
                assert ws <= p and p + img <= ws + nbytes and (p - ws) % 16 == 0 and p not in seen
                seen.add(p)
        assert (e.dst_n is None) == (q == 5)               # the input transform has one image only
    assert len(seen) == 2 * 8 + 1
    c.small_graphs = 0                                    # tiled kernels: nothing to write
    nbytes0 = lib.cal_workspace_bytes(C.byref(d), C.byref(c))
    assert lib.cal_image_sink_init(C.byref(d), C.byref(c), C.byref(po), ws, nbytes0, C.byref(sk)) == 0 and sk.count == 0
    d64 = _desc(L, hidden=64)                             # the fused path is built for hidden 128
    c.small_graphs = 1
    nb64 = lib.cal_workspace_bytes(C.byref(d64), C.byref(c))
    assert lib.cal_image_sink_init(C.byref(d64), C.byref(c), C.byref(po), ws, nb64, C.byref(sk)) == 0 and sk.count == 0
    po.conv_feat_w += 2                                   # unaligned matrix: the optimizer writes four image words at a time
    c128 = _caps(L)
    c128.small_graphs = 1
    assert lib.cal_image_sink_init(C.byref(d), C.byref(c128), C.byref(po), ws, nbytes, C.byref(sk)) == 0 and sk.count == 0
    bad = L.ImageSink()
    bad.count = 17
    assert lib.cal_adam_step_images(16, 16, 16, 16, 8, 16, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 1.0, C.byref(bad), 0) == -1
    assert lib.cal_launch_count() == n0


def test_collate_and_peer_exchange_argument_validation_no_launch():
    """cal_collate / cal_dp_*: sizes and argument errors are decided on the host before any CUDA call."""
    L, lib = _lib()
    n0 = lib.cal_launch_count()
    # exchange region: header + flags + 2 parities x world slots of the padded gradient
    assert lib.cal_dp_region_bytes(0, 1000) == 0 and lib.cal_dp_region_bytes(L.CAL_MAX_WORLD + 1, 1000) == 0
    assert lib.cal_dp_region_bytes(2, 0) == 0
    n = 138660
    n_pad = -(-n // (4 * 148)) * (4 * 148)
    for w in (1, 2, 8):
        flags = -(-(2 * w * 148 * 4) // 256) * 256
        assert lib.cal_dp_region_bytes(w, n) == 256 + flags + 2 * w * n_pad * 4
    comm = L.DpComm()
    comm.world, comm.rank = 2, 0
    assert lib.cal_dp_adam_step(None, 0, 0, 0, 0, 8, 0, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 0) == -2
    assert lib.cal_dp_adam_step(C.byref(comm), 0, 0, 0, 0, 8, 0, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 0) == -2    # NULL buffers
    comm.rank = 5
    assert lib.cal_dp_adam_step(C.byref(comm), 16, 16, 16, 16, 8, 16, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 0) == -1   # rank >= world
    comm.rank = 0
    assert lib.cal_dp_adam_step(C.byref(comm), 16, 16, 16, 16, 6, 16, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 0) == -1   # n % 4 != 0
    assert lib.cal_dp_adam_step(C.byref(comm), 16, 20, 16, 16, 8, 16, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 0) == -3   # alignment
    assert lib.cal_dp_adam_step(C.byref(comm), 16, 16, 16, 16, 8, 16, 1e-3, 0, 0.9, 0.999, 1e-8, 0.0, 0) == -2   # unmapped peer region
    assert lib.cal_dp_export(None, None) == -2 and lib.cal_dp_free(None) == -2 and lib.cal_dp_unmap(None) == -2
    assert "timed out" in lib.cal_error_string(-6).decode()
    # collate
    assert lib.cal_collate(None, 0, 0, 0, 4, 0, None, None, 0, 0, 0, 0) == -2
    st, caps, out = L.GraphStoreDesc(), _caps(L), L.Batch()
    assert lib.cal_collate(C.byref(st), 16, 4, 0, 4, 0, C.byref(caps), C.byref(out), 0, 0, 0, 0) == -2      # store arrays NULL
    assert lib.cal_collate_flush(0, 0, 0, 0, 0) == -2
    assert lib.cal_launch_count() == n0


def test_product_has_no_cpu_fallback():
    """The module refuses to run off-GPU instead of silently falling back."""
    import torch
    import cal_b200
    from tests.util import make_args
    net = cal_b200.CausalGCN(10, 4, make_args())
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(cal_b200._lib.CalError):
        net.engine
    src = open(os.path.join(ROOT, "cal_b200", "model.py")).read() + open(os.path.join(ROOT, "cal_b200", "trainer.py")).read()
    assert "oracle" not in src.replace("the oracle", "")
