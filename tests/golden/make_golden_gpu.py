"""Runs the REFERENCE's own rotated-IoU / NMS kernels (pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu compiled into
oracle/_ref/libref_iou3d.so by oracle/build_ref.sh) on a GPU and stores their outputs as golden vectors:

    gpurun -- 'python tests/golden/make_golden_gpu.py gpurun_out/nms_ref.npz'   ->  tests/golden/nms_ref.npz

    gpurun -- 'python tests/golden/make_golden_gpu.py roiaware gpurun_out/roiaware_ref.npz'  ->  tests/golden/roiaware_ref.npz
      (the reference's roiaware_pool3d_kernel.cu, oracle/_ref/libref_roiaware.so)

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


def roiaware(dst):
    """Outputs of the reference's roiaware_pool3d_kernel.cu (oracle/_ref/libref_roiaware.so) on a small scene."""
    L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_roiaware.so"))
    rng = np.random.default_rng(17)
    n, m, c, o, mp = 12, 4000, 6, 4, 8
    rois = np.zeros((n, 7), np.float32)
    rois[:, 0] = rng.uniform(5, 60, n); rois[:, 1] = rng.uniform(-30, 30, n); rois[:, 2] = rng.uniform(-2.5, -1.0, n)
    rois[:, 3] = rng.uniform(1.4, 2.2, n); rois[:, 4] = rng.uniform(3.2, 5.0, n); rois[:, 5] = rng.uniform(1.4, 2.0, n)
    rois[:, 6] = rng.uniform(-np.pi, np.pi, n)
    k = rng.integers(0, n, m)
    local = rng.uniform(-0.6, 0.6, (m, 3)) * rois[k][:, [4, 3, 5]]
    ang = rois[k, 6] + np.pi / 2
    pts = np.stack([rois[k, 0] + local[:, 0] * np.cos(ang) + local[:, 1] * np.sin(ang),
                    rois[k, 1] - local[:, 0] * np.sin(ang) + local[:, 1] * np.cos(ang),
                    rois[k, 2] + rois[k, 5] / 2 + local[:, 2]], axis=1).astype(np.float32)
    feat = rng.normal(0, 1, (m, c)).astype(np.float32)
    tr, tp, tf = (torch.from_numpy(a).cuda() for a in (rois, pts, feat))
    out = dict(rois=rois, pts=pts, feat=feat, out_size=o, max_pts=mp)
    vp = ctypes.c_void_p
    for name, code in (("max", 0), ("avg", 1)):
        pooled = torch.zeros((n, o, o, o, c), device="cuda")
        arg = torch.zeros((n, o, o, o, c), dtype=torch.int32, device="cuda")
        idx = torch.zeros((n, o, o, o, mp), dtype=torch.int32, device="cuda")
        assert L.ref_roiaware_pool3d(vp(tr.data_ptr()), n, vp(tp.data_ptr()), m, vp(tf.data_ptr()), c, o, o, o, mp, code,
                                     vp(arg.data_ptr()), vp(idx.data_ptr()), vp(pooled.data_ptr())) == 0
        out[name + "_pooled"] = pooled.cpu().numpy()
        out[name + "_idx"] = idx.cpu().numpy()
        if code == 0:
            out["max_argmax"] = arg.cpu().numpy()
    boxes = np.stack([rois, rois[::-1].copy()])
    points = np.stack([pts, pts])
    bi = torch.full((2, m), -1, dtype=torch.int32, device="cuda")
    tb, tpp = torch.from_numpy(boxes).cuda(), torch.from_numpy(points).cuda()
    assert L.ref_points_in_boxes(vp(tb.data_ptr()), 2, n, vp(tpp.data_ptr()), m, vp(bi.data_ptr())) == 0
    out["box_idx_of_points"] = bi.cpu().numpy()
    np.savez_compressed(dst, **out)
    print("wrote", dst)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "roiaware":
        roiaware(sys.argv[2])
        sys.exit(0)
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "nms_ref.npz"))
