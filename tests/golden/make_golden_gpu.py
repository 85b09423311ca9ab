"""Runs the REFERENCE's own rotated-IoU / NMS kernels (pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu compiled into
oracle/_ref/libref_iou3d.so by oracle/build_ref.sh) on a GPU and stores their outputs as golden vectors:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/nms_ref.npz'   ->  tests/golden/nms_ref.npz

The CPU test suite then pins the C restatement (oracle/pcdet_oracle.c) against these vectors.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from pcdet_b200 import synthetic as S  # noqa: E402
from util import margin_safe_boxes  # noqa: E402


def main(dst):
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_iou3d.so"))
    out = {}
    # IoU / overlap matrices
    # one clustered set split at random so that the two halves share clusters (hundreds of overlapping pairs)
    bev = O.boxes3d_to_bev(S.nms_boxes(280, seed=101)[0])
    perm = np.random.default_rng(5).permutation(280)
    a, b = np.ascontiguousarray(bev[perm[:160]]), np.ascontiguousarray(bev[perm[160:]])
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    iou = torch.zeros((160, 120), device="cuda")
    ov = torch.zeros((160, 120), device="cuda")
    assert L.ref_boxes_iou_bev(ctypes.c_void_p(ta.data_ptr()), 160, ctypes.c_void_p(tb.data_ptr()), 120, ctypes.c_void_p(iou.data_ptr())) == 0
    assert L.ref_boxes_overlap_bev(ctypes.c_void_p(ta.data_ptr()), 160, ctypes.c_void_p(tb.data_ptr()), 120, ctypes.c_void_p(ov.data_ptr())) == 0
    out.update(iou_a=a, iou_b=b, iou=iou.cpu().numpy(), overlap=ov.cpu().numpy())
    # NMS keep lists on margin-safe, score-sorted sets
    for name, n, thresh, normal in (("nms_t001", 1200, 0.01, 0), ("nms_t07", 1200, 0.7, 0), ("nms_normal_t05", 900, 0.5, 1)):
        b3, scores = S.nms_boxes(n, seed=200 + n + int(thresh * 100))
        bev = O.boxes3d_to_bev(b3)[np.argsort(-scores, kind="stable")]
        if not normal:
            bev = margin_safe_boxes(O, bev, thresh)
        t = torch.from_numpy(bev).cuda()
        keep = np.zeros((n,), np.int64)
        cnt = L.ref_nms(ctypes.c_void_p(t.data_ptr()), n, ctypes.c_float(thresh), normal, keep.ctypes.data_as(ctypes.c_void_p), None)
        assert cnt > 0
        out[name + "_boxes"] = bev
        out[name + "_keep"] = keep[:cnt]
        out[name + "_thresh"] = np.float32(thresh)
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "nms_ref.npz"))
