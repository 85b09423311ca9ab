"""Same public functions as pcdet/ops/iou3d_nms/iou3d_nms_utils.py:1-100, looked up by name from
the model config (detector3d.py:264,292; proposal_layer.py:45).  The NMS variants stay on the
device end to end: no CPU `keep` tensor, no mask download."""
from __future__ import annotations

import torch

from ... import functional as F


def boxes3d_to_bevboxes_lidar_torch(boxes3d):
    """pcdet/utils/box_utils.py:237-250."""
    return F.boxes3d_to_bev(boxes3d)


def boxes_iou_bev(boxes_a, boxes_b):
    """(M,5),(N,5) -> (M,N) rotated BEV IoU (iou3d_nms_utils.py:11-24)."""
    return F.boxes_iou_bev(boxes_a.contiguous(), boxes_b.contiguous())


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7),(M,7) [x,y,z,w,l,h,ry] LiDAR -> (N,M) 3-D IoU (iou3d_nms_utils.py:27-59)."""
    boxes_a_bev = boxes3d_to_bevboxes_lidar_torch(boxes_a)
    boxes_b_bev = boxes3d_to_bevboxes_lidar_torch(boxes_b)
    boxes_a_height_max = (boxes_a[:, 2] + boxes_a[:, 5]).view(-1, 1)
    boxes_a_height_min = boxes_a[:, 2].view(-1, 1)
    boxes_b_height_max = (boxes_b[:, 2] + boxes_b[:, 5]).view(1, -1)
    boxes_b_height_min = boxes_b[:, 2].view(1, -1)
    overlaps_bev = F.boxes_overlap_bev(boxes_a_bev, boxes_b_bev)
    max_of_min = torch.max(boxes_a_height_min, boxes_b_height_min)
    min_of_max = torch.min(boxes_a_height_max, boxes_b_height_max)
    overlaps_h = torch.clamp(min_of_max - max_of_min, min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def _nms(boxes, scores, thresh, pre_maxsize, normal):
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    n = order.shape[0]
    if n == 0:
        return order
    boxes = boxes[order].contiguous()
    keep, num = F.nms_sorted_batched(boxes, [0, n], float(thresh), normal=normal)
    return order[keep[0, :int(num.item())]].contiguous()


def nms_gpu(boxes, scores, thresh, pre_maxsize=None):
    """(N,5) [x1,y1,x2,y2,ry], (N) -> kept original indices in score order (iou3d_nms_utils.py:62-78)."""
    return _nms(boxes, scores, thresh, pre_maxsize, False)


def nms_normal_gpu(boxes, scores, thresh):
    """iou3d_nms_utils.py:81-95."""
    return _nms(boxes, scores, thresh, None, True)
