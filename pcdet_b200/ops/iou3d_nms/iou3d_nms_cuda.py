"""Stand-in for the pybind module `iou3d_nms_cuda` (pcdet/ops/iou3d_nms/src/iou3d_nms.cpp:180-185):
same four entry points, same argument meaning, same return values."""
from __future__ import annotations

import torch

from ... import functional as F


def _check_cuda_contig(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDAtensor ")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous ")


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """(N,5),(M,5) -> writes ans_overlap (N,M); returns 1 (iou3d_nms.cpp:36-55)."""
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_overlap, "ans_overlap")):
        _check_cuda_contig(t, n)
    F.boxes_overlap_bev(boxes_a, boxes_b, out=ans_overlap)
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """iou3d_nms.cpp:57-76."""
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        _check_cuda_contig(t, n)
    F.boxes_iou_bev(boxes_a, boxes_b, out=ans_iou)
    return 1


def _nms(boxes, keep, thresh, normal):
    _check_cuda_contig(boxes, "boxes")
    if not keep.is_contiguous():
        raise RuntimeError("keep must be contiguous ")
    n = boxes.shape[0]
    if n == 0:
        return 0
    k, num = F.nms_sorted_batched(boxes, [0, n], thresh, normal=normal)
    cnt = int(num.item())
    # the reference contract hands the kept positions back in a CPU LongTensor (iou3d_nms.cpp:88)
    keep[:cnt] = k[0, :cnt].to(keep.device)
    return cnt


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """boxes (N,5) cuda sorted by score desc, keep (N) int64 (CPU in the reference); returns the
    number of kept boxes and fills keep[:num] with their positions (iou3d_nms.cpp:79-126)."""
    return _nms(boxes, keep, float(nms_overlap_thresh), False)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """iou3d_nms.cpp:129-177 (axis-aligned IoU)."""
    return _nms(boxes, keep, float(nms_overlap_thresh), True)
