"""Deterministic synthetic inputs for the hot path.

There is no dataset access, so bench and tests run on LiDAR-shaped clouds produced by a small beam
model: a ground plane at z = -1.73 m, a smooth "skyline" of walls, 10 % dropout and 2 cm range noise.
The shapes follow the reference configs (tools/cfgs/second.yaml:11-51) so that voxel counts and
neighbour densities are in the range of real KITTI / nuScenes frames.
"""
from __future__ import annotations

import math

import numpy as np

# tools/cfgs/second.yaml:11,49-51 and 19,26
KITTI = dict(
    voxel_size=(0.05, 0.05, 0.1),
    point_cloud_range=(0.0, -40.0, -3.0, 70.4, 40.0, 1.0),
    max_num_points=5,
    max_voxels=40000,
)
# nuScenes-shaped 10-sweep cloud, 0.1 m voxels
NUSCENES = dict(
    voxel_size=(0.1, 0.1, 0.1),
    point_cloud_range=(-51.2, -51.2, -3.0, 51.2, 51.2, 1.0),
    max_num_points=10,
    max_voxels=160000,
)
# tools/cfgs/pointpillar.yaml (config 2)
PILLARS = dict(
    voxel_size=(0.16, 0.16, 4.0),
    point_cloud_range=(0.0, -39.68, -3.0, 69.12, 39.68, 1.0),
    max_num_points=32,
    max_voxels=12000,
)


def _sweep(rng, n_beams, az_step_deg, fov_deg):
    elev = np.deg2rad(np.linspace(-24.8, 2.0, n_beams))
    az = np.deg2rad(np.arange(fov_deg[0], fov_deg[1], az_step_deg))
    e, a = np.meshgrid(elev, az, indexing="ij")
    e, a = e.ravel(), a.ravel()
    with np.errstate(divide="ignore"):
        r_ground = np.where(e < 0, 1.73 / np.sin(-e), np.inf)
    wall = 18 + 10 * np.sin(3 * a) + 6 * np.sin(7 * a + 1) + rng.normal(0, 0.3, a.shape)
    r_wall = wall / np.cos(e)
    h_wall = r_wall * np.sin(e)
    hit_wall = (h_wall > -1.73) & (h_wall < 2.5) & (r_wall < r_ground)
    r = np.where(hit_wall, r_wall, r_ground)
    r = r + rng.normal(0, 0.02, r.shape)
    keep = np.isfinite(r) & (r < 80) & (rng.uniform(0, 1, r.shape) > 0.1)
    r, e, a = r[keep], e[keep], a[keep]
    xyz = np.stack([r * np.cos(e) * np.cos(a), r * np.cos(e) * np.sin(a), r * np.sin(e)], axis=1)
    inten = rng.uniform(0, 1, (xyz.shape[0], 1))
    return np.concatenate([xyz, inten], axis=1).astype(np.float32)


def kitti_frame(seed: int = 0) -> np.ndarray:
    """64 beams x 90 deg FOV, ~20 k points (N,4) f32 [x,y,z,intensity]."""
    return _sweep(np.random.default_rng(seed), 64, 0.26, (-45.0, 45.0))


def nuscenes_frame(seed: int = 0, sweeps: int = 10) -> np.ndarray:
    """10 jittered 32-beam 360 deg sweeps, ~314 k points."""
    rng = np.random.default_rng(seed)
    parts = []
    for _ in range(sweeps):
        p = _sweep(rng, 32, 0.33, (-180.0, 180.0))
        p[:, 0] += np.float32(rng.normal(0, 0.5))
        p[:, 1] += np.float32(rng.normal(0, 0.2))
        parts.append(p)
    return np.concatenate(parts, axis=0)


def uniform_cloud(n: int, rng_range, seed: int = 0, outside_frac: float = 0.05) -> np.ndarray:
    """Uniform random points, a fraction of them outside the range (edge-case tests)."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(rng_range[:3], np.float64)
    hi = np.asarray(rng_range[3:], np.float64)
    span = hi - lo
    pts = lo + rng.uniform(-outside_frac, 1 + outside_frac, (n, 3)) * span
    return np.concatenate([pts, rng.uniform(0, 1, (n, 1))], axis=1).astype(np.float32)


def backbone_weights(in_channels: int = 4, seed: int = 0):
    """spconv v1.0 reset_parameters: U(-1/sqrt(Cin*prod(k)), +...) per layer,
    keyed by the state-dict stem of rpn_backbone.py (e.g. 'conv2.0.0')."""
    import torch
    from .backbone import BACKBONE8X_LAYERS

    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, _kind, cin, cout, ks, _st, _pd, _key in BACKBONE8X_LAYERS:
        cin = in_channels if cin is None else cin
        kvol = ks[0] * ks[1] * ks[2]
        stdv = 1.0 / math.sqrt(cin * kvol)
        w = (torch.rand((*ks, cin, cout), generator=g, dtype=torch.float32) * 2 - 1) * stdv
        out[name] = w.numpy()
    return out


# anchor sizes (w, l, h) of second.yaml: Car, Pedestrian, Cyclist
_ANCHOR_WLH = np.array([[1.6, 3.9, 1.56], [0.6, 0.8, 1.73], [0.6, 1.76, 1.73]], dtype=np.float64)


def nms_boxes(n: int = 4096, seed: int = 0, clustered: bool = True, pc_range=KITTI["point_cloud_range"]):
    """(boxes3d (n,7) [x,y,z,w,l,h,ry] f32, scores (n,) f32 all distinct).

    clustered=True mimics detector output before NMS: many near-duplicate boxes around a few hundred
    object centres; clustered=False is the uniform stress set of SURVEY 8(d)."""
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(pc_range[:3]), np.asarray(pc_range[3:])
    cls = rng.integers(0, 3, n)
    if clustered:
        n_obj = max(n // 16, 1)
        centres = lo[:2] + rng.uniform(0, 1, (n_obj, 2)) * (hi[:2] - lo[:2])
        obj_ry = rng.uniform(-np.pi, np.pi, n_obj)
        obj_cls = rng.integers(0, 3, n_obj)
        which = rng.integers(0, n_obj, n)
        cls = obj_cls[which]
        xy = centres[which] + rng.normal(0, 0.35, (n, 2))
        ry = obj_ry[which] + rng.normal(0, 0.15, n)
    else:
        xy = lo[:2] + rng.uniform(0, 1, (n, 2)) * (hi[:2] - lo[:2])
        ry = rng.uniform(-np.pi, np.pi, n)
    wlh = _ANCHOR_WLH[cls] * rng.uniform(0.8, 1.2, (n, 3))
    z = rng.uniform(-1.8, -0.6, n)
    boxes = np.concatenate([xy, z[:, None], wlh, ry[:, None]], axis=1).astype(np.float32)
    scores = (rng.permutation(n).astype(np.float32) + 1.0) / np.float32(n + 1)
    return boxes, scores
