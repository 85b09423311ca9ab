"""ctypes binding of libpcdet_b200.so (the C ABI declared in include/pcdet_b200.h).

There is no CPU fallback: if the shared library is missing the first call raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpcdet_b200.so")

F32, BF16 = 0, 1
EPI_RELU = 1
WEIGHT_PACKED = 2
CONV_PDL = 4
CONV_SHALLOW_RING = 8
EPI_RESIDUAL_POST = 16
PACK_TRANSPOSE = 1
PACK_FLIP = 2
RB_CLEARED = 1
RB_UNDONE = 2

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); kept in the order of include/pcdet_b200.h
SIGNATURES = {
    "pcdb_abi_version": (_i, []),
    "pcdb_last_error": (C.c_char_p, []),
    "pcdb_voxelize_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "pcdb_voxelize": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp,
                           _vp, _sz, _vp]),
    "pcdb_voxelize_sites": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp,
                           _vp, _sz, _vp]),
    "pcdb_voxelize_points": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp,
                           _vp, _sz, _vp]),
    "pcdb_vfe_mean": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp]),
    "pcdb_pillar_vfe": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "pcdb_roiaware_pool3d_fwd": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pcdb_roiaware_pool3d_fwd_ex": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pcdb_roiaware_pool3d_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pcdb_points_in_boxes": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp]),
    "pcdb_rulebook_workspace_bytes": (_sz, [_i, _i, _i]),
    "pcdb_rulebook_subm": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "pcdb_rulebook_conv_sites": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _i, _vp]),
    "pcdb_rulebook_conv_pairs": (_i, [_i, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp]),
    "pcdb_rulebook_conv_clear": (_i, [_vp, _sz, _i, _i, _i, _vp, _i, _vp]),
    "pcdb_rulebook_subm_reuse": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "pcdb_rulebook_invert": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "pcdb_rulebook_conv": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _i,
                                _vp, _sz, _vp]),
    "pcdb_rulebook_chain_workspace_bytes": (_sz, [_i, _i, _vp, _vp]),
    "pcdb_rulebook_chain_clear": (_i, [_vp, _sz, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcdb_rulebook_chain": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "pcdb_sparse_conv_fwd": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "pcdb_sparse_conv_fwd_ex": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "pcdb_conv_packed_weight_bytes": (_sz, [_i, _i, _i]),
    "pcdb_pack_conv_weights": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "pcdb_pack_conv_weights_ex": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "pcdb_sparse_conv_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "pcdb_sparse_conv_wgrad": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "pcdb_bn_train_workspace_bytes": (_sz, []),
    "pcdb_bn_train_fwd": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "pcdb_bn_train_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "pcdb_bn_train_sums": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "pcdb_bn_train_fwd_from_sums": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _f, _f, _vp, _vp, _i, _vp, _vp, _vp]),
    "pcdb_bn_train_bwd_sums": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "pcdb_bn_train_bwd_from_sums": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "pcdb_sparse_conv_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pcdb_sparse_maxpool_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "pcdb_sparse_maxpool_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "pcdb_to_dense": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "pcdb_from_dense": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "pcdb_fill_rows_i32": (_i, [_vp, _i, _i, _vp, _i, _i, _vp]),
    "pcdb_dense_clear_rows": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "pcdb_boxes_overlap_bev": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "pcdb_boxes_iou_bev": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "pcdb_boxes_iou3d": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "pcdb_nms_workspace_bytes": (_sz, [_i, _i]),
    "pcdb_nms": (_i, [_vp, _vp, _i, _f, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "pcdb_filter_points_workspace_bytes": (_sz, [_i]),
    "pcdb_filter_points": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pcdb_nms_counts": (_i, [_vp, _vp, _vp, _i, _f, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "pcdb_boxes3d_to_bev": (_i, [_vp, _i, _vp, _vp]),
    "pcdb_decode_select_workspace_bytes": (_sz, [_i, _i, _i]),
    "pcdb_decode_select": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _f, _i, _i,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pcdb_gather_kept": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_LIB = None


class PcdbError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise PcdbError(
                f"{SO_PATH} is missing: build it with `python -m pcdet_b200.build` (nvcc, sm_100a). "
                "pcdet_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().pcdb_last_error().decode("utf-8", "replace")
        raise PcdbError(f"{what or 'pcdb call'} failed with status {status}: {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def i32x3(v):
    arr = (C.c_int32 * 3)(*[int(x) for x in v])
    return arr


def f32xN(v):
    return (C.c_float * len(v))(*[float(x) for x in v])
