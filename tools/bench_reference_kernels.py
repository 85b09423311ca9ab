"""The reference's OWN CUDA kernels on this B200, next to ours, on the same inputs.

    gpurun -- 'python tools/bench_reference_kernels.py --out gpurun_out/r02_reference_kernels.json'

oracle/_ref/libref_iou3d.so and libref_roiaware.so are pcdet/ops/iou3d_nms/src/iou3d_nms_kernel.cu and
pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu compiled from where they lie under /root/reference
(oracle/build_ref.sh), behind host shims that reproduce the reference's host code (iou3d_nms.cpp:79-126: cudaMalloc of
the mask, kernel, 2 MB D2H, serial host sweep, cudaFree).  They are the only pieces of the reference that run on this box;
everything else on the path is spconv, which cannot be built here.  Times: wall clock around the blocking reference call
(it synchronises itself), CUDA events around ours, median of the repetitions, L2 flushed before each.
This is a measurement tool: the oracle is executed here as the thing measured AGAINST, never by the product."""
import argparse
import ctypes
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as R

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--reps", type=int, default=15)
args = ap.parse_args()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
vp = ctypes.c_void_p


def wall(fn, reps=args.reps):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return statistics.median(ts)


def events(fn, reps=args.reps):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


res = {"device": torch.cuda.get_device_name(dev)}

# ---- rotated-BEV NMS: 4 frames x 4096 score-sorted boxes, thresh 0.01 (second.yaml:152-159) -------------------------------
B, N = 4, 4096
b3, scores = S.nms_boxes(B * N, seed=0)
bev_all = F.boxes3d_to_bev(torch.from_numpy(b3).to(dev)).cpu().numpy()
bev = np.concatenate([bev_all[b * N:(b + 1) * N][np.argsort(-scores[b * N:(b + 1) * N], kind="stable")] for b in range(B)])
bev_t = torch.from_numpy(bev).to(dev)
ours_keep = {}


def ours_nms():
    ours_keep["k"], ours_keep["n"] = F.nms_sorted_batched(bev_t, [b * N for b in range(B + 1)], 0.01, keep_stride=500)


ms_ours = events(ours_nms)
ms_ours_sync = wall(lambda: (ours_nms(), ours_keep["n"].cpu()))
entry = {"what": f"{B} x {N} score-sorted BEV boxes, thresh 0.01, keep 500 per frame", "ours_ms_device": ms_ours,
         "ours_ms_with_count_on_host": ms_ours_sync}
ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_iou3d.so")
if os.path.exists(ref_so):
    L = ctypes.CDLL(ref_so)
    L.ref_nms.restype = ctypes.c_int
    keep = np.zeros(N, np.int64)
    ref_counts = []

    def ref_nms_all():
        ref_counts.clear()
        for b in range(B):       # the reference post-processes a batch frame by frame (detector3d.py:147)
            ref_counts.append(L.ref_nms(vp(bev_t.data_ptr() + b * N * 20), N, ctypes.c_float(0.01), 0, keep.ctypes.data_as(vp), None))

    entry["reference_ms"] = wall(ref_nms_all)
    entry["reference_what"] = "iou3d_nms_kernel.cu nms_kernel + iou3d_nms.cpp host path (cudaMalloc, 2 MB D2H, serial sweep, cudaFree), frame by frame"
    entry["speedup"] = entry["reference_ms"] / ms_ours_sync
    ours_n = ours_keep["n"].cpu().numpy()
    entry["keep_counts"] = {"ours_capped_at_500": ours_n.tolist(), "reference_uncapped": ref_counts}

    # the mask kernel alone (what the reference spends on the device)
    mask = torch.empty((N, N // 64), dtype=torch.int64, device=dev)
    L.ref_boxes_iou_bev.restype = ctypes.c_int
    iou = torch.empty((N, N), dtype=torch.float32, device=dev)
    entry["reference_iou_matrix_ms"] = wall(lambda: L.ref_boxes_iou_bev(vp(bev_t.data_ptr()), N, vp(bev_t.data_ptr()), N, vp(iou.data_ptr())))
    entry["ours_iou_matrix_ms"] = events(lambda: F.boxes_iou_bev(bev_t[:N], bev_t[:N], out=iou))
else:
    entry["reference_ms"] = None
    entry["note"] = "oracle/_ref/libref_iou3d.so did not travel to this box"
res["nms"] = entry

# ---- RoI-aware pooling: the bench_rows shape (128 rois x 16384 points x 128 channels, out 14, 128 pts/voxel, max) -------------
rng = np.random.default_rng(1)
n_rois, n_pts, C = 128, 16384, 128
rois = np.zeros((n_rois, 7), np.float32)
rois[:, 0] = rng.uniform(5, 60, n_rois); rois[:, 1] = rng.uniform(-30, 30, n_rois); rois[:, 2] = rng.uniform(-2.5, -1.0, n_rois)
rois[:, 3] = rng.uniform(1.4, 2.2, n_rois); rois[:, 4] = rng.uniform(3.2, 5.0, n_rois); rois[:, 5] = rng.uniform(1.4, 2.0, n_rois)
rois[:, 6] = rng.uniform(-np.pi, np.pi, n_rois)
k = rng.integers(0, n_rois, n_pts)
local = rng.uniform(-0.6, 0.6, (n_pts, 3)) * rois[k][:, [4, 3, 5]]
ang = rois[k, 6] + np.pi / 2
pts = np.stack([rois[k, 0] + local[:, 0] * np.cos(ang) + local[:, 1] * np.sin(ang),
                rois[k, 1] - local[:, 0] * np.sin(ang) + local[:, 1] * np.cos(ang), rois[k, 2] + rois[k, 5] / 2 + local[:, 2]], axis=1).astype(np.float32)
feat = rng.normal(0, 1, (n_pts, C)).astype(np.float32)
rois_t, pts_t, feat_t = (torch.from_numpy(a).to(dev) for a in (rois, pts, feat))
pool = R.RoIAwarePool3d(14, 128)
entry = {"what": f"RoIAwarePool3d(out 14, 128 pts/voxel) max, {n_rois} rois x {n_pts} points x {C} channels (partA2_rcnn_net.py:256-295)",
         "ours_ms_device": events(lambda: pool(rois_t, pts_t, feat_t, "max"))}
ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_roiaware.so")
if os.path.exists(ref_so):
    L2 = ctypes.CDLL(ref_so)
    L2.ref_roiaware_pool3d.restype = ctypes.c_int
    pooled = torch.zeros((n_rois, 14, 14, 14, C), dtype=torch.float32, device=dev)
    argmax = torch.zeros((n_rois, 14, 14, 14, C), dtype=torch.int32, device=dev)
    idx = torch.zeros((n_rois, 14, 14, 14, 128), dtype=torch.int32, device=dev)

    def ref_pool():
        # roiaware_pool3d_utils.py:43-52: the caller zeroes the three outputs, then one launcher call
        pooled.zero_(); argmax.zero_(); idx.zero_()
        L2.ref_roiaware_pool3d(vp(rois_t.data_ptr()), n_rois, vp(pts_t.data_ptr()), n_pts, vp(feat_t.data_ptr()), C, 14, 14, 14, 128, 0,
                               vp(argmax.data_ptr()), vp(idx.data_ptr()), vp(pooled.data_ptr()))

    entry["reference_ms"] = wall(ref_pool)
    entry["reference_what"] = "roiaware_pool3d_kernel.cu launcher (mask, serial per-box collection, pooling) incl. the output zeroing of the Python wrapper"
    entry["speedup"] = entry["reference_ms"] / entry["ours_ms_device"]
    got = pool(rois_t, pts_t, feat_t, "max")
    entry["max_abs_diff"] = float((got - pooled).abs().max())
else:
    entry["reference_ms"] = None
res["roiaware_pool3d"] = entry
print(json.dumps(res, indent=1))
if args.out:
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
