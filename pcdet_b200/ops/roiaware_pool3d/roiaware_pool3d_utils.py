"""Drop-in for pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py on the sm_100a kernels (csrc/roiaware.cu):
RoIAwarePool3d (module + autograd Function, :7-66), points_in_boxes_gpu (:69-83), points_in_boxes_cpu (:86-98)."""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from ..._lib import check, lib, ptr
from ...functional import _require_cuda, _stream


class RoIAwarePool3d(nn.Module):
    def __init__(self, out_size, max_pts_each_voxel=128):
        super().__init__()
        self.out_size = out_size
        self.max_pts_each_voxel = max_pts_each_voxel

    def forward(self, rois, pts, pts_feature, pool_method="max"):
        assert pool_method in ["max", "avg"]
        return RoIAwarePool3dFunction.apply(rois, pts, pts_feature, self.out_size, self.max_pts_each_voxel, pool_method)


class RoIAwarePool3dFunction(Function):
    @staticmethod
    def forward(ctx, rois, pts, pts_feature, out_size, max_pts_each_voxel, pool_method):
        """rois (N,7) [x,y,z,w,l,h,ry] (z = bottom centre), pts (npoints,3), pts_feature (npoints,C)
        -> pooled_features (N, out_x, out_y, out_z, C)"""
        if isinstance(out_size, int):
            out_x = out_y = out_z = out_size
        else:
            assert len(out_size) == 3
            out_x, out_y, out_z = (int(v) for v in out_size)
        _require_cuda(rois, pts, pts_feature)
        rois, pts, feat = rois.contiguous().float(), pts.contiguous().float(), pts_feature.contiguous().float()
        n, c, m = rois.shape[0], feat.shape[-1], pts.shape[0]
        pooled = feat.new_zeros((n, out_x, out_y, out_z, c))
        argmax = torch.zeros((n, out_x, out_y, out_z, c), dtype=torch.int32, device=feat.device)
        idx = torch.zeros((n, out_x, out_y, out_z, max_pts_each_voxel), dtype=torch.int32, device=feat.device)
        method = {"max": 0, "avg": 1}[pool_method]
        check(lib().pcdb_roiaware_pool3d_fwd(ptr(rois), n, ptr(pts), m, ptr(feat), c, out_x, out_y, out_z, max_pts_each_voxel,
                                             method, ptr(argmax), ptr(idx), ptr(pooled), _stream()), "pcdb_roiaware_pool3d_fwd")
        ctx.roiaware_pool3d_for_backward = (idx, argmax, method, m, c, (out_x, out_y, out_z), max_pts_each_voxel)
        return pooled

    @staticmethod
    def backward(ctx, grad_out):
        idx, argmax, method, m, c, (ox, oy, oz), max_pts = ctx.roiaware_pool3d_for_backward
        grad_out = grad_out.contiguous().float()
        grad_in = grad_out.new_zeros((m, c))
        check(lib().pcdb_roiaware_pool3d_bwd(ptr(idx), ptr(argmax), ptr(grad_out), idx.shape[0], ox, oy, oz, c, max_pts, method,
                                             ptr(grad_in), _stream()), "pcdb_roiaware_pool3d_bwd")
        return None, None, grad_in, None, None, None


def points_in_boxes_gpu(points, boxes):
    """points (B,M,3), boxes (B,T,7) -> box_idxs_of_pts (B,M) int32, background = -1"""
    assert boxes.shape[0] == points.shape[0] and boxes.shape[2] == 7
    _require_cuda(points, boxes)
    b, m, _ = points.shape
    out = torch.full((b, m), -1, dtype=torch.int32, device=points.device)
    check(lib().pcdb_points_in_boxes(ptr(boxes.contiguous().float()), b, boxes.shape[1], ptr(points.contiguous().float()), m,
                                     ptr(out), _stream()), "pcdb_points_in_boxes")
    return out


def points_in_boxes_cpu(points, boxes):
    """points (npoints,3), boxes (N,7) -> point_indices (N,npoints) int32 (roiaware_pool3d.cpp:151-171: a HOST function
    in the reference too, used by the dataset code); the same fp32 test, vectorised over torch CPU tensors."""
    assert boxes.shape[1] == 7 and points.shape[1] == 3
    p, b = points.float(), boxes.float()
    cz = (b[:, 2].double() + b[:, 5].double() / 2.0).float()
    in_z = ((p[None, :, 2] - cz[:, None]).abs().double() <= b[:, 5].double()[:, None] / 2.0)
    rot = (b[:, 6].double() + 3.14159265358979323846 / 2).float()
    cosa, sina = torch.cos(rot)[:, None], torch.sin(rot)[:, None]
    sx, sy = p[None, :, 0] - b[:, 0:1], p[None, :, 1] - b[:, 1:2]
    lx, ly = sx * cosa + sy * (-sina), sx * sina + sy * cosa
    hl, hw = b[:, 4].double()[:, None] / 2.0, b[:, 3].double()[:, None] / 2.0
    inside = in_z & (lx.double() > -hl) & (lx.double() < hl) & (ly.double() > -hw) & (ly.double() < hw)
    return inside.to(torch.int32)
