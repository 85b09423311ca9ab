"""The C-ABI shared library loads without a GPU and exports every symbol include/pcdet_b200.h declares."""
import ctypes
import os
import re

from pcdet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcdet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcdb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 19
    handle = ctypes.CDLL(_lib.SO_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/pcdet_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "python binding table out of sync with the header"


def test_abi_version_and_host_only_calls():
    L = _lib.lib()
    assert L.pcdb_abi_version() == 1
    # workspace sizing is pure host arithmetic
    a = L.pcdb_voxelize_workspace_bytes(20000, 1, 5, 40000)
    b = L.pcdb_voxelize_workspace_bytes(80000, 4, 5, 40000)
    assert 0 < a < b
    assert L.pcdb_rulebook_workspace_bytes(1000, 27, 8000) > 0
    assert L.pcdb_nms_workspace_bytes(4, 4096) >= 4 * 4096 * 64 * 8
    assert L.pcdb_decode_select_workspace_bytes(4, 211200, 4096) >= 4 * 211200 * 4 + 4 * 4096 * 8
    assert L.pcdb_decode_select_workspace_bytes(0, 10, 10) == 0
    assert 0 < L.pcdb_filter_points_workspace_bytes(0) < L.pcdb_filter_points_workspace_bytes(300000)


def test_argument_validation_reports_errors_instead_of_exiting():
    L = _lib.lib()
    st = L.pcdb_sparse_conv_fwd(None, 0, None, None, 0, 27, 10, None, 4, 16, 0, None, None, None, 0, None, 0, None)
    assert st == 1
    assert b"invalid argument" in L.pcdb_last_error()
    st = L.pcdb_nms(None, None, 0, 0.5, 0, None, 1, None, None, 0, None)
    assert st == 1
    st = L.pcdb_nms_counts(None, None, None, 1, 0.5, 0, None, 1, None, None, 0, None)
    assert st == 1 and b"set_counts" in L.pcdb_last_error()
    st = L.pcdb_decode_select(None, 3, None, None, None, 1, 100, 3, 2, 0.0, 0.0, 0.1, 64, 0, None, None, None, None, None, None, None, 0, None)
    assert st == 1 and b"pcdb_decode_select" in L.pcdb_last_error()
    st = L.pcdb_filter_points(None, 10, 4, None, 1, None, None, None, None, None, None, 0, None)
    assert st == 1 and b"pcdb_filter_points" in L.pcdb_last_error()
    st = L.pcdb_gather_kept(None, 1, None, 1, 1, None, None, None, None, 1, 0, 0.0, 0, None, None, None, None, None, None)
    assert st == 1


def test_host_mirrors_reject_cpu_tensors():
    """no CPU fallback anywhere: the new host-side mirrors raise on CPU tensors instead of computing something else"""
    import numpy as np
    import pytest
    import torch
    from pcdet_b200 import functional as F
    from pcdet_b200.postprocess import PostProcessor
    cls, box, anchors = torch.zeros((1, 8, 3)), torch.zeros((1, 8, 7)), torch.zeros((8, 7))
    with pytest.raises(_lib.PcdbError):
        F.decode_select(cls, box, anchors)
    with pytest.raises(_lib.PcdbError):
        PostProcessor(anchors).select(cls, box)
    with pytest.raises(_lib.PcdbError):
        F.filter_points(torch.zeros((4, 4)), torch.tensor([0, 4], dtype=torch.int32), 1)
    from pcdet_b200.ingest import calib_record, read_bin
    rec = calib_record(np.eye(3, 4), np.eye(3), np.eye(3, 4), (375, 1242))
    assert rec.shape == (26,) and rec.dtype == np.float32 and rec[24] == 375 and rec[25] == 1242


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "pcdet_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liborc" not in src, f


def test_unet_v2_state_dict_matches_reference_layout():
    """tests/golden/ref_unet_keys.npz = state-dict keys/shapes of the REFERENCE's UNetV2 (rpn_unet.py) instantiated on
    our spconv modules (tests/golden/make_golden.py unet): pcdet_b200.unet.UNetV2 has the same layout, so reference
    checkpoints load unchanged."""
    import os
    import numpy as np
    from pcdet_b200.unet import UNetV2
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_unet_keys.npz"))
    sd = UNetV2(4).state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [str(list(v.shape)) for v in sd.values()] == list(g["shapes"])


def test_round2_entry_points_validate_arguments_and_size_workspaces():
    """Training / static-shape entry points of round 2: host-side sizing and argument validation (no kernel is launched)."""
    L = _lib.lib()
    # weight-gradient workspace = one (K, c_in, c_out) fp32 partial per CTA; grows with the rows until every SM has a CTA
    small = L.pcdb_sparse_conv_wgrad_workspace_bytes(27, 128, 64, 64)
    big = L.pcdb_sparse_conv_wgrad_workspace_bytes(27, 1 << 20, 64, 64)
    assert small == 27 * 64 * 64 * 4 and big == 74 * small           # 64 -> 64 splits the offsets over two CTAs per row range
    assert L.pcdb_sparse_conv_wgrad_workspace_bytes(27, 1000, 48, 64) == 0          # unsupported channel count
    assert L.pcdb_sparse_conv_wgrad_workspace_bytes(28, 1000, 64, 64) == 0          # more than 27 offsets
    assert L.pcdb_bn_train_workspace_bytes() > 0
    st = L.pcdb_sparse_conv_wgrad(None, 0, None, None, 0, 27, 0, None, 64, 64, None, 0, None, 0, None)
    assert st == 1 and b"pcdb_sparse_conv_wgrad" in L.pcdb_last_error()
    st = L.pcdb_bn_train_fwd(None, 10, None, 64, 1, None, None, 1e-3, 0.01, None, None, 1, None, None, None, 0, None, 0, None)
    assert st == 1 and b"pcdb_bn_train_fwd" in L.pcdb_last_error()
    st = L.pcdb_bn_train_bwd(None, None, None, 10, None, 64, 1, None, None, 1, None, None, None, 0, None, 0, None)
    assert st == 1
    st = L.pcdb_bn_train_sums(None, 10, None, 64, 1, None, None, 0, None)
    assert st == 1
    st = L.pcdb_rulebook_invert(None, 0, 27, 0, None, None, 0, 0, None, None)
    assert st == 1 and b"pcdb_rulebook_invert" in L.pcdb_last_error()
    st = L.pcdb_from_dense(None, 1, None, 4, None, 16, 1, None, None, 1, None)
    assert st == 1
    st = L.pcdb_pack_conv_weights_ex(None, 0, 27, 64, 64, 0, None, None)
    assert st == 5                                                                  # PCDB_UNSUPPORTED: nothing to pack
    st = L.pcdb_roiaware_pool3d_fwd_ex(None, 1, None, 1, None, None, 0, 14, 14, 14, 128, 0, 0, None, None, None, None)
    assert st == 1
    # the residual epilogue exists on the tensor-core path only: fp32 features are refused, not silently computed without it
    import ctypes
    buf = (ctypes.c_float * 4)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    st = L.pcdb_sparse_conv_fwd_ex(p, 1, p, ctypes.cast(buf, ctypes.c_void_p), 1, 1, 1, None, 16, 16, 0, None, None, None, p, 0, p, 0, None)
    assert st == 5 and b"residual" in L.pcdb_last_error()
