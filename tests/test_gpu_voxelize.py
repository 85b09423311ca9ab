"""GPU parity: voxel hash + mean VFE (through the C ABI) against the oracle.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200.spconv.utils import VoxelGenerator
from util import voxel_sets

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_gpu(frames, cfg, overflow_break=True, max_voxels=None, want_mean=True):
    dev = torch.device("cuda")
    sizes = [f.shape[0] for f in frames]
    pts = torch.from_numpy(np.concatenate(frames, axis=0)).to(dev)
    offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=dev)
    out = F.voxelize(pts, offs, len(frames), cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"],
                     max_voxels or cfg["max_voxels"], overflow_break, want_voxels=True, want_mean=want_mean,
                     want_point_idx=True)
    torch.cuda.synchronize()
    vo = out["voxel_offsets"].cpu().numpy()
    n = int(vo[-1])
    return {k: (v[:n].cpu().numpy() if v is not None and k != "voxel_offsets" else v) for k, v in out.items()}, vo


def check_against_oracle(orc, frames, cfg, overflow_break=True, max_voxels=None):
    mv = max_voxels or cfg["max_voxels"]
    got, vo = run_gpu(frames, cfg, overflow_break, mv)
    g = orc.VoxelGenerator(cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], mv, overflow_break)
    base = 0
    for b, f in enumerate(frames):
        vox, coors, num, pidx = g.generate(f, return_point_idx=True)
        lo, hi = int(vo[b]), int(vo[b + 1])
        assert hi - lo == vox.shape[0], f"frame {b}: {hi - lo} voxels vs oracle {vox.shape[0]}"
        # ORDER is part of the contract here (first appearance), not only the set
        np.testing.assert_array_equal(got["coordinates"][lo:hi, 0], b)
        np.testing.assert_array_equal(got["coordinates"][lo:hi, 1:], coors)
        np.testing.assert_array_equal(got["num_points"][lo:hi], num)
        np.testing.assert_array_equal(got["voxels"][lo:hi].view(np.uint32), vox.view(np.uint32))
        ref_idx = np.where(pidx >= 0, pidx + base, -1)
        np.testing.assert_array_equal(got["point_idx"][lo:hi], ref_idx)
        mean = orc.vfe_mean(vox, num)
        np.testing.assert_allclose(got["mean"][lo:hi], mean, rtol=1e-6, atol=1e-6)
        # sorted-set form (north_star): identical as sets too
        assert voxel_sets(got["voxels"][lo:hi], got["coordinates"][lo:hi, 1:], got["num_points"][lo:hi]) == \
            voxel_sets(vox, coors, num)
        base += f.shape[0]
    return got, vo


def test_kitti_frame_bit_exact(orc):
    got, vo = check_against_oracle(orc, [S.kitti_frame(0)], S.KITTI)
    assert int(vo[-1]) == 16774


def test_batch_of_frames_with_empty_frame(orc):
    frames = [S.kitti_frame(1), np.zeros((0, 4), np.float32), S.kitti_frame(2)[:5000], S.kitti_frame(3)]
    check_against_oracle(orc, frames, S.KITTI)


def test_nuscenes_frame_bit_exact(orc):
    check_against_oracle(orc, [S.nuscenes_frame(0)], S.NUSCENES)


def test_pillar_config(orc):
    check_against_oracle(orc, [S.kitti_frame(4), S.kitti_frame(5)], S.PILLARS)


@pytest.mark.parametrize("overflow_break", [True, False])
def test_max_voxels_overflow_semantics(orc, overflow_break):
    cfg = dict(voxel_size=(0.5, 0.5, 0.5), point_cloud_range=(0, -8, -2, 16, 8, 2), max_num_points=3, max_voxels=300)
    frames = [S.uniform_cloud(6000, cfg["point_cloud_range"], seed=s) for s in (1, 2)]
    got, vo = check_against_oracle(orc, frames, cfg, overflow_break)
    assert np.all(np.diff(vo) == 300)


def test_dense_voxels_keep_first_p_points(orc):
    """Hundreds of points per voxel: the atomicMin cascade must keep the P smallest indices in order."""
    cfg = dict(voxel_size=(2.0, 2.0, 2.0), point_cloud_range=(0, 0, 0, 8, 8, 4), max_num_points=5, max_voxels=1000)
    check_against_oracle(orc, [S.uniform_cloud(20000, cfg["point_cloud_range"], seed=9)], cfg)
    cfg["max_num_points"] = 1
    check_against_oracle(orc, [S.uniform_cloud(5000, cfg["point_cloud_range"], seed=10)], cfg)


def test_boundary_and_nan_points(orc):
    cfg = dict(voxel_size=(0.05, 0.05, 0.1), point_cloud_range=(0, -40, -3, 70.4, 40, 1), max_num_points=5, max_voxels=100)
    pts = np.array([[0, -40, -3, 1], [70.4, 0, 0, 1], [70.39999, 39.99999, 0.99999, 1], [np.nan, 0, 0, 1],
                    [1e30, 0, 0, 1], [-1e-6, 0, 0, 1], [0.05, 0.05, 0.1, 1], [0.1, 0.1, 0.2, 1],
                    [0.15, 0.15, 0.3, 1], [35.2, 0.0, -1.0, 1]], np.float32)
    check_against_oracle(orc, [pts], cfg)


def test_more_point_features(orc):
    cfg = dict(voxel_size=(0.4, 0.4, 0.4), point_cloud_range=(0, -8, -2, 16, 8, 2), max_num_points=4, max_voxels=5000)
    rng = np.random.default_rng(3)
    pts = np.concatenate([S.uniform_cloud(4000, cfg["point_cloud_range"], seed=5), rng.normal(0, 1, (4000, 2)).astype(np.float32)], axis=1)
    check_against_oracle(orc, [np.ascontiguousarray(pts)], cfg)


def test_voxel_generator_reference_contract(orc):
    """spconv.utils.VoxelGenerator surface (kitti_dataset.py:674-688, dataset.py:163-174, second_net.py:10)."""
    vg = VoxelGenerator(voxel_size=[0.05, 0.05, 0.1], point_cloud_range=[0, -40, -3, 70.4, 40, 1], max_num_points=5,
                        max_voxels=40000)
    assert vg.grid_size.dtype == np.int64 and vg.grid_size.tolist() == [1408, 1600, 40]
    assert (vg.grid_size[::-1] + [1, 0, 0]).tolist() == [41, 1600, 1408]
    assert (vg.grid_size[:2] // 8).tolist() == [176, 200]
    assert vg.voxel_size.dtype == np.float32 and vg.point_cloud_range.dtype == np.float32
    pts = S.kitti_frame(6)
    voxels, coords, num = vg.generate(pts)
    ref = orc.VoxelGenerator([0.05, 0.05, 0.1], [0, -40, -3, 70.4, 40, 1], 5, 40000).generate(pts)
    assert voxels.dtype == np.float32 and coords.dtype == np.int32 and num.dtype == np.int32
    for a, b in zip((voxels, coords, num), ref):
        np.testing.assert_array_equal(a, b)
    centers = (coords[:, ::-1] + 0.5) * vg.voxel_size + vg.point_cloud_range[0:3]
    assert centers.shape == (coords.shape[0], 3)
    v0, c0, n0 = vg.generate(np.zeros((1, 3)))   # the constructor probe of kitti_dataset.py:681
    assert v0.shape == (1, 5, 3) and n0.tolist() == [1]


def test_vfe_mean_op(orc):
    g = np.load(os.path.join(GOLD, "ref_python.npz"))
    out = F.vfe_mean(torch.from_numpy(g["vfe_voxels"]).cuda(), torch.from_numpy(g["vfe_num"]).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), g["vfe_mean"], rtol=1e-6, atol=1e-6)   # reference Python golden
    from pcdet_b200.vfe import MeanVoxelFeatureExtractor
    m = MeanVoxelFeatureExtractor()(torch.from_numpy(g["vfe_voxels"]).cuda(), torch.from_numpy(g["vfe_num"]).cuda())
    np.testing.assert_array_equal(m.cpu().numpy(), out.cpu().numpy())


def test_errors_are_reported(orc):
    from pcdet_b200._lib import PcdbError
    pts = torch.zeros((10, 4), device="cuda")
    offs = torch.tensor([0, 10], dtype=torch.int32, device="cuda")
    with pytest.raises(PcdbError, match="32-bit hash key"):
        F.voxelize(pts, offs, 1, (0.001, 0.001, 0.001), (0, 0, 0, 100, 100, 100), 5, 100)
    with pytest.raises(PcdbError, match="CUDA tensors only"):
        F.voxelize(pts.cpu(), offs, 1, (1, 1, 1), (0, 0, 0, 4, 4, 4), 5, 100)
