"""GPU parity: FOV + range filter of raw clouds through the C ABI, against golden vectors made
by the reference's own KITTI dataset class / Calibration (tests/golden/ref_ingest.npz) and against the oracle.

Bar: the surviving points, their order and the frame offsets bit-exact.  A point whose projection lies within 1e-2 px
of an image border (or 1e-3 m of the camera plane) may legitimately flip between two fp32 summation orders (numpy's
BLAS vs. the kernel's FMA chain); such points are removed from the inputs first and counted."""
import os

import numpy as np
import pytest
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200.ingest import KittiIngest, calib_record

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def golden_frames():
    g = np.load(os.path.join(GOLD, "ref_ingest.npz"))
    frames, calibs, shapes, kept = [], [], [], []
    for f in range(2):
        frames.append(g[f"points_{f}"]); shapes.append(g[f"img_shape_{f}"]); kept.append(g[f"kept_{f}"])
        calibs.append(dict(V2C=g[f"V2C_{f}"], R0=g[f"R0_{f}"], P2=g[f"P2_{f}"]))
    return frames, calibs, shapes, kept, g["pc_range"]


def drop_borderline(orc, pts, calib, shape, pc_range):
    _, sh = orc.fov_flag(pts, calib["V2C"], calib["R0"], calib["P2"], shape)
    near = (np.abs(sh[:, 0]) < 1e-2) | (np.abs(sh[:, 0] - shape[1]) < 1e-2) | (np.abs(sh[:, 1]) < 1e-2) | \
           (np.abs(sh[:, 1] - shape[0]) < 1e-2) | (np.abs(sh[:, 2]) < 1e-3) | ~np.isfinite(sh).all(axis=1)
    return pts[~near], int(near.sum())


def run(frames, calibs=None, shapes=None, pc_range=None, want_index=False):
    pts = torch.from_numpy(np.concatenate(frames) if frames else np.zeros((0, 4), np.float32)).cuda()
    offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device="cuda")
    calib = None if calibs is None else torch.from_numpy(np.stack([calib_record(c["V2C"], c["R0"], c["P2"], s) for c, s in zip(calibs, shapes)])).cuda()
    rng = None if pc_range is None else torch.tensor([pc_range[0], pc_range[1], pc_range[3], pc_range[4]], dtype=torch.float32, device="cuda")
    out, o, idx = F.filter_points(pts, offs, len(frames), calib, rng, want_index=want_index)
    o = o.cpu().numpy()
    return out.cpu().numpy(), o, None if idx is None else idx.cpu().numpy()


def test_filter_matches_reference_python_golden(orc):
    frames, calibs, shapes, kept, pc_range = golden_frames()
    dropped = 0
    for f in range(2):
        frames[f], d = drop_borderline(orc, frames[f], calibs[f], shapes[f], pc_range)
        dropped += d
    assert dropped < 20
    ref = orc.filter_points(frames, calibs, shapes, pc_range)
    for f in range(2):                                                     # the oracle on the cleaned input == golden minus the dropped
        assert len(kept[f]) - 20 <= len(ref[f]) <= len(kept[f])
    out, offs, idx = run(frames, calibs, shapes, pc_range, want_index=True)
    assert offs[0] == 0
    cat = np.concatenate(frames)
    for f in range(2):
        got = out[offs[f]:offs[f + 1]]
        np.testing.assert_array_equal(got, ref[f])
        np.testing.assert_array_equal(cat[idx[offs[f]:offs[f + 1]]], got)
    # single frames one at a time: exactly the golden's kept points where no borderline point was involved
    for f in range(2):
        o1, of1, _ = run([frames[f]], [calibs[f]], [shapes[f]], pc_range)
        np.testing.assert_array_equal(o1[:of1[1]], ref[f])


@pytest.mark.parametrize("batch,fov,use_range", [(4, True, True), (4, False, True), (3, True, False), (1, False, False)])
def test_filter_then_voxelize_vs_oracle(orc, batch, fov, use_range):
    """KITTI-shaped frames widened to 360 degrees; the filtered device batch voxelizes to exactly what the oracle's
    voxel generator produces from the oracle-filtered frames."""
    frames_g, calibs_g, shapes_g, _, pc_range = golden_frames()
    rng = np.random.default_rng(3)
    frames, calibs, shapes = [], [], []
    for b in range(batch):
        f = S.kitti_frame(b)
        mirror = f.copy(); mirror[:, 0] = -mirror[:, 0]                     # points behind the car
        wide = f.copy(); wide[:, 1] *= 2.5                                   # points beyond +-40 m
        f = np.concatenate([f, mirror[::3], wide[::4]])[rng.permutation(len(f) + len(mirror[::3]) + len(wide[::4]))]
        c, s = calibs_g[b % 2], shapes_g[b % 2]
        f, _ = drop_borderline(orc, f.astype(np.float32), c, s, pc_range)
        frames.append(np.ascontiguousarray(f)); calibs.append(c); shapes.append(s)
    ref = orc.filter_points(frames, calibs if fov else None, shapes if fov else None, pc_range if use_range else None)
    out, offs, _ = run(frames, calibs if fov else None, shapes if fov else None, pc_range if use_range else None)
    for b in range(batch):
        np.testing.assert_array_equal(out[offs[b]:offs[b + 1]], ref[b])
    if fov:
        assert all(len(r) < 0.8 * len(f) for r, f in zip(ref, frames))
    # through the host-side mirror and into the voxelizer
    ing = KittiIngest(pc_range if use_range else None, fov_points_only=fov)
    pts, o = ing(frames, calibs, shapes)
    cfg = S.KITTI
    v = F.voxelize(pts, o, batch, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
    gen = orc.VoxelGenerator(cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
    vox, coords, num = orc.collate([gen.generate(r) for r in ref])
    n = int(v["voxel_offsets"][-1])
    assert n == vox.shape[0]
    np.testing.assert_array_equal(v["coordinates"][:n].cpu().numpy(), coords)
    np.testing.assert_array_equal(v["voxels"][:n].cpu().numpy(), vox)
    np.testing.assert_array_equal(v["num_points"][:n].cpu().numpy(), num)


def test_filter_edge_cases(orc):
    pc_range = S.KITTI["point_cloud_range"]
    rng = np.random.default_rng(0)
    a = rng.uniform([-10, -50, -3, 0], [80, 50, 1, 1], (1000, 4)).astype(np.float32)
    empty = np.zeros((0, 4), np.float32)
    outside = a.copy(); outside[:, 0] = -5.0
    out, offs, _ = run([empty, a, empty, outside, a[:1]], pc_range=pc_range)
    ref = orc.filter_points([empty, a, empty, outside, a[:1]], pc_range=pc_range)
    np.testing.assert_array_equal(offs, np.concatenate([[0], np.cumsum([len(r) for r in ref])]))
    np.testing.assert_array_equal(out[:offs[-1]], np.concatenate(ref))
    # inclusive ends of mask_points_by_range (common_utils.py:48-49)
    edge = np.array([[0.0, -40.0, 0, 0], [70.4, 40.0, 0, 0], [70.4001, 0, 0, 0], [0, -40.001, 0, 0]], np.float32)
    out, offs, _ = run([edge], pc_range=pc_range)
    np.testing.assert_array_equal(out[:offs[1]], edge[:2])
    # five features per point (nuScenes time channel), large frame, no filter at all = identity
    big = rng.uniform(-50, 50, (300001, 5)).astype(np.float32)
    out, offs, idx = run([big], want_index=True)
    assert offs[1] == len(big)
    np.testing.assert_array_equal(out, big)
    np.testing.assert_array_equal(idx, np.arange(len(big)))
    out, offs, _ = run([empty])
    assert offs.tolist() == [0, 0]
