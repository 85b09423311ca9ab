"""spconv.utils.VoxelGenerator backed by the sm_100a voxel hash kernel."""
from __future__ import annotations

import numpy as np
import torch

from ... import functional as F


class VoxelGenerator:
    """Same constructor, attributes and `generate` contract as spconv.utils.VoxelGenerator
    (constructed at pcdet/datasets/kitti/kitti_dataset.py:674-679, called at
    pcdet/datasets/dataset.py:163): numpy in, (voxels, coordinates, num_points) numpy out.

    `generate_batch` is the B200 path: device tensors in the PCDet batch-dict layout
    (pcdet/datasets/dataset.py:266-299), no host round trip."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, overflow_break=True,
                 device="cuda"):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = int(max_num_points)
        self._max_voxels = int(max_voxels)
        self._grid_size = grid_size
        self._overflow_break = bool(overflow_break)   # spconv v1.0 `break`; False = v1.1 `continue`
        self._device = torch.device(device)

    # -- reference contract -------------------------------------------------------------------
    def generate(self, points, max_voxels=None):
        pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32)).to(self._device, non_blocking=False)
        out = self.generate_batch([pts], max_voxels=max_voxels, want_voxels=True)
        n = int(out["voxel_offsets"][1].item())
        return (out["voxels"][:n].cpu().numpy(), out["coordinates"][:n, 1:].contiguous().cpu().numpy(),
                out["num_points"][:n].cpu().numpy())

    # -- device path ----------------------------------------------------------------------------
    def generate_batch(self, points_list, max_voxels=None, want_voxels=True, want_mean=False,
                       mean_dtype=torch.float32, mean_stride=None, want_point_idx=False):
        """points_list: list of (N_b, C) float32 CUDA tensors (one per frame) or a tuple
        (points (N,C), frame_offsets (B+1) int32) already concatenated on the device."""
        if isinstance(points_list, tuple):
            pts, offs = points_list
            batch = offs.numel() - 1
        else:
            batch = len(points_list)
            sizes = [int(p.shape[0]) for p in points_list]
            pts = torch.cat(points_list, dim=0) if batch > 1 else points_list[0]
            offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=pts.device)
        pts = pts.contiguous().float()
        return F.voxelize(pts, offs, batch, self._voxel_size, self._point_cloud_range, self._max_num_points,
                          int(max_voxels or self._max_voxels), self._overflow_break, want_voxels=want_voxels,
                          want_mean=want_mean, mean_dtype=mean_dtype, mean_stride=mean_stride,
                          want_point_idx=want_point_idx)

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size


class VoxelGeneratorV2(VoxelGenerator):
    """spconv v1.1 flavour: `generate` returns a dict (accepted by pcdet/datasets/dataset.py:165-171)."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("overflow_break", False)
        super().__init__(*args, **kwargs)

    def generate(self, points, max_voxels=None):
        v, c, n = super().generate(points, max_voxels)
        return {"voxels": v, "coordinates": c, "num_points_per_voxel": n, "voxel_num": v.shape[0]}
