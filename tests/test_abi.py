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
