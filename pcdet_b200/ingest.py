"""Raw KITTI clouds -> filtered device-resident batch for the voxelizer.

Mirror of the input side of ``KittiDataset.__getitem__`` (pcdet/datasets/kitti/kitti_dataset.py:74-92, 700-717) and of
``DatasetTemplate.prepare_data``'s range mask (pcdet/datasets/dataset.py:184) for inference: the reference reads the
``.bin`` file, projects every point into the camera with numpy, masks, voxelizes on the CPU worker and ships the
VOXELS through pickling and a pageable H2D copy.  Here the raw float32 payload is the only thing that crosses PCIe
(pinned, asynchronous); FOV mask, range mask and voxelization all run on the device (``pcdb_filter_points`` ->
``pcdb_voxelize``).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import functional as F


def read_bin(path: str, num_features: int = 4) -> np.ndarray:
    """get_lidar (kitti_dataset.py:74-85): the velodyne .bin payload as (N, 4) float32."""
    return np.fromfile(path, dtype=np.float32).reshape(-1, num_features)


def calib_record(V2C, R0, P2, img_shape) -> np.ndarray:
    """26 floats of one frame for pcdb_filter_points: [V2C^T . R0^T (4,3) | P2 (3,4) | img_h, img_w] -- the matrices
    of pcdet/utils/calibration.py:27-29, combined exactly as lidar_to_rect does (calibration.py:72), in float32."""
    m = np.dot(np.asarray(V2C, np.float32).T, np.asarray(R0, np.float32).T)
    return np.concatenate([m.reshape(-1), np.asarray(P2, np.float32).reshape(-1),
                           np.asarray([img_shape[0], img_shape[1]], np.float32)]).astype(np.float32)


class KittiIngest:
    """frames (list of (N_i, C) float32 host arrays) -> (points, frame_offsets) on the device, FOV- and range-filtered."""

    def __init__(self, point_cloud_range: Optional[Sequence[float]] = None, fov_points_only: bool = True, device="cuda",
                 max_points_total: int = 1 << 20, num_features: int = 4):
        self.device = torch.device(device)
        self.fov = fov_points_only
        self.c = num_features
        r = point_cloud_range
        self.range_xy = None if r is None else torch.tensor([r[0], r[1], r[3], r[4]], dtype=torch.float32, device=self.device)
        self.host = torch.empty((max_points_total, num_features), dtype=torch.float32).pin_memory()
        self.dev = torch.empty((max_points_total, num_features), dtype=torch.float32, device=self.device)

    def __call__(self, frames, calibs=None, img_shapes=None):
        n = 0
        offs = [0]
        for f in frames:
            k = f.shape[0]
            if n + k > self.host.shape[0]:
                raise ValueError(f"{n + k} points exceed max_points_total={self.host.shape[0]}")
            self.host[n:n + k].copy_(torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32)))
            n += k
            offs.append(n)
        self.dev[:n].copy_(self.host[:n], non_blocking=True)
        offsets = torch.tensor(offs, dtype=torch.int32).pin_memory().to(self.device, non_blocking=True)
        calib = None
        if self.fov:
            if calibs is None or img_shapes is None:
                raise ValueError("fov_points_only needs one calibration and image shape per frame")
            rec = np.stack([calib_record(c["V2C"], c["R0"], c["P2"], s) for c, s in zip(calibs, img_shapes)])
            calib = torch.from_numpy(rec).pin_memory().to(self.device, non_blocking=True)
        pts, out_offs, _ = F.filter_points(self.dev[:n], offsets, len(frames), calib, self.range_xy)
        return pts, out_offs
