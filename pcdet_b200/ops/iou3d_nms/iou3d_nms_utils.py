"""Same public functions as pcdet/ops/iou3d_nms/iou3d_nms_utils.py:1-100, looked up by name from
the model config (detector3d.py:264,292; proposal_layer.py:45).  The NMS variants stay on the
device end to end: no CPU `keep` tensor, no mask download."""
from __future__ import annotations

from ... import functional as F


def boxes3d_to_bevboxes_lidar_torch(boxes3d):
    """pcdet/utils/box_utils.py:237-250."""
    return F.boxes3d_to_bev(boxes3d)


def boxes_iou_bev(boxes_a, boxes_b):
    """(M,5),(N,5) -> (M,N) rotated BEV IoU (iou3d_nms_utils.py:11-24)."""
    return F.boxes_iou_bev(boxes_a.contiguous(), boxes_b.contiguous())


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7),(M,7) [x,y,z,w,l,h,ry] LiDAR -> (N,M) 3-D IoU (iou3d_nms_utils.py:27-59), one kernel."""
    return F.boxes_iou3d(boxes_a, boxes_b)


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
