"""GPU parity: RoI-aware point pooling (pcdet/ops/roiaware_pool3d) through the C ABI against the
oracle's C restatement and, when it travelled to the box, the reference's own kernel (oracle/_ref/libref_roiaware.so)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from pcdet_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_roiaware.so")


def scene(seed, n_rois=48, n_pts=6000, channels=16):
    rng = np.random.default_rng(seed)
    rois = np.zeros((n_rois, 7), np.float32)
    rois[:, 0] = rng.uniform(5, 60, n_rois); rois[:, 1] = rng.uniform(-30, 30, n_rois); rois[:, 2] = rng.uniform(-2.5, -1.0, n_rois)
    rois[:, 3] = rng.uniform(1.4, 2.2, n_rois); rois[:, 4] = rng.uniform(3.2, 5.0, n_rois); rois[:, 5] = rng.uniform(1.4, 2.0, n_rois)
    rois[:, 6] = rng.uniform(-np.pi, np.pi, n_rois)
    # points: most inside/around the rois (some voxels overflow max_pts_each_voxel), the rest background
    k = rng.integers(0, n_rois, n_pts)
    local = rng.uniform(-0.6, 0.6, (n_pts, 3)) * rois[k][:, [4, 3, 5]]
    ang = rois[k, 6] + np.pi / 2
    x = rois[k, 0] + local[:, 0] * np.cos(ang) + local[:, 1] * np.sin(ang)
    y = rois[k, 1] - local[:, 0] * np.sin(ang) + local[:, 1] * np.cos(ang)
    z = rois[k, 2] + rois[k, 5] / 2 + local[:, 2]
    pts = np.stack([x, y, z], axis=1).astype(np.float32)
    pts[: n_pts // 8] = rng.uniform([0, -40, -3], [70, 40, 1], (n_pts // 8, 3)).astype(np.float32)
    feat = rng.normal(0, 1, (n_pts, channels)).astype(np.float32)
    return rois, pts, feat


@pytest.mark.parametrize("method,out_size,max_pts", [("max", 14, 128), ("avg", 14, 128), ("max", (6, 5, 4), 8), ("avg", 7, 4)])
def test_roiaware_pool3d_vs_oracle(orc, method, out_size, max_pts):
    rois, pts, feat = scene(1, n_pts=6000 if max_pts >= 128 else 60000)     # the small lists must overflow
    pool = R.RoIAwarePool3d(out_size, max_pts)
    got = pool(torch.from_numpy(rois).cuda(), torch.from_numpy(pts).cuda(), torch.from_numpy(feat).cuda(), method)
    ref, ref_arg, ref_idx = orc.roiaware_pool3d(rois, pts, feat, out_size, max_pts, method)
    if method == "max":
        np.testing.assert_array_equal(got.cpu().numpy(), ref)             # a selection: bit exact
    else:
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=0, atol=1e-6)   # same summation order; FMA contraction only
    assert (ref_idx[..., 0] == max_pts - 1).any() or max_pts >= 128             # the small cases overflow their voxels


def test_roiaware_pool3d_backward_and_points_in_boxes(orc):
    rois, pts, feat = scene(2, n_rois=20, n_pts=3000, channels=8)
    tr, tp = torch.from_numpy(rois).cuda(), torch.from_numpy(pts).cuda()
    for method in ("max", "avg"):
        f = torch.from_numpy(feat).cuda().requires_grad_(True)
        out = R.RoIAwarePool3d(7, 16)(tr, tp, f, method)
        g = torch.from_numpy(np.random.default_rng(3).normal(0, 1, tuple(out.shape)).astype(np.float32)).cuda()
        out.backward(g)
        # reference semantics of the backward (roiaware_pool3d_kernel.cu:242-292) from the oracle's forward bookkeeping
        _, arg, idx = orc.roiaware_pool3d(rois, pts, feat, 7, 16, method)
        want = np.zeros_like(feat, dtype=np.float64)
        gn = g.cpu().numpy().astype(np.float64)
        if method == "max":
            sel = arg >= 0
            c_idx = np.broadcast_to(np.arange(feat.shape[1]), arg.shape)
            np.add.at(want, (arg[sel], c_idx[sel]), gn[sel])
        else:
            cnt = idx[..., 0]
            for pos in np.argwhere(cnt > 0):
                lst = idx[tuple(pos)][1:1 + cnt[tuple(pos)]]
                np.add.at(want, lst, gn[tuple(pos)] / max(cnt[tuple(pos)], 1))
        np.testing.assert_allclose(f.grad.cpu().numpy(), want, rtol=0, atol=2e-5)
    # points_in_boxes_gpu: first containing box per point, per sample
    boxes = np.stack([rois, rois[::-1].copy()])
    points = np.stack([pts, pts])
    got = R.points_in_boxes_gpu(torch.from_numpy(points).cuda(), torch.from_numpy(boxes).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, orc.points_in_boxes(points, boxes))
    assert (got >= 0).mean() > 0.3
    cpu = R.points_in_boxes_cpu(torch.from_numpy(pts), torch.from_numpy(rois)).numpy()
    first = np.where(cpu.any(axis=0), cpu.argmax(axis=0), -1)
    np.testing.assert_array_equal(first, got[0])


def test_roiaware_pool3d_vs_compiled_reference_kernel(orc):
    if not os.path.exists(REF_SO):
        pytest.skip("compiled reference kernel not present")
    L = ctypes.CDLL(REF_SO)
    rois, pts, feat = scene(5)
    tr, tp, tf = (torch.from_numpy(a).cuda() for a in (rois, pts, feat))
    for method, code in (("max", 0), ("avg", 1)):
        n, c, o, mp = rois.shape[0], feat.shape[1], 14, 128
        pooled = torch.zeros((n, o, o, o, c), device="cuda")
        arg = torch.zeros((n, o, o, o, c), dtype=torch.int32, device="cuda")
        idx = torch.zeros((n, o, o, o, mp), dtype=torch.int32, device="cuda")
        vp = ctypes.c_void_p
        assert L.ref_roiaware_pool3d(vp(tr.data_ptr()), n, vp(tp.data_ptr()), pts.shape[0], vp(tf.data_ptr()), c, o, o, o, mp, code,
                                     vp(arg.data_ptr()), vp(idx.data_ptr()), vp(pooled.data_ptr())) == 0
        got = R.RoIAwarePool3d(o, mp)(tr, tp, tf, method)
        if method == "max":
            np.testing.assert_array_equal(got.cpu().numpy(), pooled.cpu().numpy())
        else:
            np.testing.assert_allclose(got.cpu().numpy(), pooled.cpu().numpy(), rtol=0, atol=1e-6)
        ref_o, _, ref_idx = orc.roiaware_pool3d(rois, pts, feat, o, mp, method)
        np.testing.assert_array_equal(ref_idx, idx.cpu().numpy())               # the C restatement pins the point lists too
