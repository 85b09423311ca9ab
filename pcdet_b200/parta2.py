"""Part-A^2 stage 1 -> stage 2 bridge as one captured launch sequence (BASELINE config 4).

What PartA2Net.forward_rpn / forward_rcnn / PartA2RCNNNet.roiaware_pool do between the voxels and the RCNN head
(pcdet/models/detectors/PartA2_net.py:15-83, pcdet/models/rpn/rpn_unet.py:414-529,
pcdet/models/rcnn/partA2_rcnn_net.py:256-295):

    points -> voxel hash + mean VFE -> UNetV2 encoder-decoder (the UNMODIFIED module tree of pcdet_b200/unet.py, run in
    static-shape mode: SparseConvTensor.n_dev) -> [RPN head on the BEV map: dense cuDNN, out of scope -- the caller's
    cls / box / dir predictions] -> proposal layer (anchor decode + top-k + rotated NMS, PostProcessor.proposals) ->
    RoI-aware pooling of the decoder's point-wise outputs over the voxel centres: part features (sigmoid offsets masked by
    the segmentation score, + the score) averaged, segmentation features max-pooled (pcdb_roiaware_pool3d_fwd).

Every count (voxels, sites per level, candidates, kept proposals) stays on the device and every buffer is a capacity, so
the sequence is captured into ONE CUDA graph; the reference synchronises at every rulebook build and loops over the frames
in Python for the NMS and again for the pooling (boolean-mask indexing = one more sync per frame).  Here the pooling of
frame b scans the device-side row range of that frame (the voxelizer emits frames one after the other), so the in-box
arithmetic on a frame's own points is exactly the reference's; the point lists are collected once for both poolings.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import functional as F
from . import spconv
from ._lib import check, lib, ptr
from .postprocess import PostProcessConfig, PostProcessor
from .unet import UNetV2


@dataclass
class PartA2Config:
    voxel_size: Sequence[float] = (0.05, 0.05, 0.1)
    point_cloud_range: Sequence[float] = (0.0, -40.0, -3.0, 70.4, 40.0, 1.0)
    max_num_points: int = 5
    max_voxels: int = 40000
    batch_size: int = 2
    max_points_total: int = 2 * 24000
    max_voxels_total: Optional[int] = None         # row capacity of level 0 (default: min(points, B * max_voxels))
    dtype: torch.dtype = torch.bfloat16
    seg_mask_score_thresh: float = 0.3             # MODEL.RPN.BACKBONE.SEG_MASK_SCORE_THRESH (PartA2_car.yaml)
    roi_pool_size: int = 14                        # ROI_AWARE_POOL_SIZE
    max_pts_each_voxel: int = 128
    # proposal layer, TEST mode of tools/cfgs/PartA2_car.yaml:193-196
    proposals: PostProcessConfig = field(default_factory=lambda: PostProcessConfig(
        score_thresh=0.0, nms_thresh=0.7, nms_pre_maxsize=1024, nms_post_maxsize=100))


class PartA2HotPath:
    def __init__(self, cfg: PartA2Config, unet: UNetV2, anchors: torch.Tensor, device="cuda"):
        self.cfg, self.dev = cfg, torch.device(device)
        self.net = unet.to(self.dev).eval()
        if cfg.dtype == torch.bfloat16:
            self.net = self.net.to(torch.bfloat16)
        self.post = PostProcessor(anchors.to(self.dev), cfg.proposals)
        g = F.grid_size(cfg.voxel_size, cfg.point_cloud_range)
        self.sparse_shape = [int(g[2]) + 1, int(g[1]), int(g[0])]
        self.cap = int(cfg.max_voxels_total or min(cfg.max_points_total, cfg.batch_size * cfg.max_voxels))
        vs, rg = np.asarray(cfg.voxel_size, np.float32), np.asarray(cfg.point_cloud_range, np.float32)
        # voxel centre of coordinate (b, z, y, x): (x, y, z) * voxel_size + range_min + voxel_size / 2 (PartA2_net.py:97-101)
        self._vsize = torch.tensor([vs[0], vs[1], vs[2]], device=self.dev)
        self._origin = torch.tensor([rg[0] + vs[0] / 2, rg[1] + vs[1] / 2, rg[2] + vs[2] / 2], device=self.dev)
        self.graph = None
        self.out = None

    def forward(self, points: torch.Tensor, frame_offsets: torch.Tensor, rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds=None):
        """points (N, 4) f32 (N <= max_points_total, frames concatenated), frame_offsets (B+1) i32, head outputs for the
        anchors (B, A, C) / (B, A, 7) / (B, A, bins).  Returns device tensors only; rows / rois past their counts are padding."""
        c, B = self.cfg, self.cfg.batch_size
        v = F.voxelize(points, frame_offsets, B, c.voxel_size, c.point_cloud_range, c.max_num_points, c.max_voxels,
                       want_voxels=False, want_mean=True, mean_dtype=c.dtype, capacity=self.cap)
        n_dev = v["voxel_offsets"][B:B + 1]
        x = spconv.SparseConvTensor(v["mean"], v["coordinates"], self.sparse_shape, B, n_dev=n_dev)
        with torch.no_grad():
            u = self.net(x)
        overflow = x.indice_dict.get("__overflow__", [])
        # ---- point-wise stage-2 inputs (PartA2_net.py:38-48, partA2_rcnn_net.py:262-270) ----------------------------------
        coords = v["coordinates"]
        centers = coords[:, 1:4].flip(1).float() * self._vsize + self._origin
        seg_score = torch.sigmoid(u["u_seg_preds"].float().view(-1))
        part = torch.sigmoid(u["u_reg_preds"].float()) * (seg_score > c.seg_mask_score_thresh)[:, None]
        part_features = torch.cat((part, seg_score[:, None]), dim=1).contiguous()
        seg_features = u["seg_features"].float().contiguous()
        # ---- proposal layer -----------------------------------------------------------------------------------------------
        prop = self.post.proposals(rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds)
        rois = prop["rois"].contiguous()                              # (B, P, 7), zero padded (an empty box holds no point)
        P, s = rois.shape[1], c.roi_pool_size
        # ---- RoI-aware pooling (partA2_rcnn_net.py:272-290), frame by frame without leaving the device ---------------------
        pooled_part = part_features.new_zeros((B * P, s, s, s, part_features.shape[1]))
        pooled_seg = seg_features.new_zeros((B * P, s, s, s, seg_features.shape[1]))
        # the voxelizer emits the frames one after the other: frame b's centres are rows voxel_offsets[b] .. [b+1] (a device-side
        # range); the point lists of a frame are collected once and serve both poolings
        centers = centers.contiguous()
        # argmax is written for every (roi, voxel, channel) and the point lists carry their length in slot 0 (counted in shared
        # memory for pooling grids up to 48 KB of counters, i.e. 23^3): neither needs the 140 MB of zeros per frame the
        # reference's wrapper allocates (roiaware_pool3d_utils.py:35-37)
        assert s * s * s * 4 <= 48 * 1024
        argmax = torch.empty((P, s, s, s, seg_features.shape[1]), dtype=torch.int32, device=self.dev)
        for b in range(B):
            idx = torch.empty((P, s, s, s, c.max_pts_each_voxel), dtype=torch.int32, device=self.dev)
            rng = v["voxel_offsets"][b:b + 2]
            self._pool(rois[b], centers, rng, part_features, "avg", 0, argmax, idx, pooled_part[b * P:(b + 1) * P])
            self._pool(rois[b], centers, rng, seg_features, "max", 1, argmax, idx, pooled_seg[b * P:(b + 1) * P])
        return dict(rois=rois, roi_raw_scores=prop["roi_raw_scores"], roi_labels=prop["roi_labels"], num_rois=prop["num"],
                    pooled_part_features=pooled_part, pooled_rpn_features=pooled_seg, seg_features=u["seg_features"],
                    u_seg_preds=u["u_seg_preds"], u_reg_preds=u["u_reg_preds"], spatial_features=u["spatial_features"],
                    coordinates=coords, voxel_offsets=v["voxel_offsets"], voxel_centers=centers, part_features=part_features,
                    overflow=torch.cat(overflow) if overflow else torch.zeros(1, dtype=torch.int32, device=self.dev))

    def _pool(self, rois, pts, pts_range, feat, method, flags, argmax, idx, pooled):
        s, m = self.cfg.roi_pool_size, self.cfg.max_pts_each_voxel
        check(lib().pcdb_roiaware_pool3d_fwd_ex(ptr(rois), rois.shape[0], ptr(pts), pts.shape[0], ptr(pts_range), ptr(feat), feat.shape[1],
                                                s, s, s, m, {"max": 0, "avg": 1}[method], flags, ptr(argmax), ptr(idx), ptr(pooled),
                                                F._stream()), "pcdb_roiaware_pool3d_fwd_ex")

    # ------------------------------------------------------------------------------------------
    def capture(self, points, frame_offsets, rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds=None, warmup: int = 2):
        """Capture forward() on these (static) input tensors; the caller refreshes their contents and calls replay()."""
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.forward(points, frame_offsets, rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self.forward(points, frame_offsets, rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds)
        return self.out

    def replay(self):
        self.graph.replay()
        return self.out
