"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np


def sort_rows(a):
    a = np.asarray(a)
    if a.shape[0] == 0:
        return a
    return a[np.lexsort(a.T[::-1])]


def voxel_sets(voxels, coords, num_points, point_idx=None):
    """Canonical, order-insensitive form of a voxelisation: {coord tuple: (count, points bytes)}."""
    out = {}
    for v in range(coords.shape[0]):
        n = int(num_points[v])
        key = tuple(int(x) for x in coords[v])
        out[key] = (n, voxels[v, :n].tobytes(), voxels[v, n:].tobytes())
    return out


def nbr_to_pair_sets(nbr, n_out, in_ids, out_ids):
    """Neighbour map (K, ld) -> per offset sorted rows [b,z,y,x in | b,z,y,x out]."""
    res = []
    for k in range(nbr.shape[0]):
        m = nbr[k, :n_out]
        o = np.nonzero(m >= 0)[0]
        rows = np.concatenate([in_ids[m[o]], out_ids[o]], axis=1) if o.size else np.zeros((0, 8), np.int32)
        res.append(sort_rows(rows))
    return res


def margin_safe_boxes(orc, boxes_bev_sorted, thresh, margin=2e-4, max_rounds=20, rng=None):
    """Moves boxes far away until no pair has |IoU - thresh| < margin (fp64 geometry), so that any two
    correct fp32 implementations must take identical keep/suppress decisions."""
    rng = rng or np.random.default_rng(0)
    b = np.array(boxes_bev_sorted, dtype=np.float32, copy=True)
    for _ in range(max_rounds):
        iou = orc.boxes_iou_bev64(b, b)
        np.fill_diagonal(iou, -1.0)
        bad = np.argwhere(np.abs(iou - thresh) < margin)
        bad = bad[bad[:, 0] < bad[:, 1]]
        if bad.shape[0] == 0:
            return b
        for j in np.unique(bad[:, 1]):
            w, h = b[j, 2] - b[j, 0], b[j, 3] - b[j, 1]
            cx, cy = 500 + rng.uniform(0, 5000), 500 + rng.uniform(0, 5000)
            b[j, 0], b[j, 1], b[j, 2], b[j, 3] = cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2
    raise AssertionError("could not build a margin-safe box set")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
