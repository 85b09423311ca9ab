"""GPU parity: PointPillars pillar feature net + BEV scatter (BASELINE config 2) through the C ABI against the
oracle and against golden vectors produced by the reference's own Python."""
import os

import numpy as np
import pytest
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200.vfe import PillarFeatureNetOld2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden():
    g = np.load(os.path.join(GOLD, "ref_pillars.npz"))
    net = PillarFeatureNetOld2(4, True, (64,), False, tuple(g["voxel_size"]), tuple(g["pc_range"])).eval().cuda()
    sd = {k: torch.from_numpy(g["sd_" + k]) for k in g["state_keys"]}
    net.load_state_dict(sd)                     # the reference's state dict loads unchanged
    return g, net


def test_pillar_vfe_matches_reference_golden():
    g, net = golden()
    assert list(net.state_dict().keys()) == list(g["state_keys"])
    vox, num, co = (torch.from_numpy(g[k]).cuda() for k in ("voxels", "num_points", "coords"))
    feats = net(vox, num, co).cpu().numpy()
    assert np.abs(feats - g["features"]).max() < 2e-5 * max(1.0, np.abs(g["features"]).max())
    f2, canvas = net.forward_scatter(vox, num, co, 2, [1, 496, 432], want_features=True)
    np.testing.assert_array_equal(f2.cpu().numpy(), feats)
    canvas = canvas.cpu().numpy()
    assert canvas.shape == (2, 64, 496, 432)
    np.testing.assert_array_equal(np.argwhere(canvas != 0).astype(np.int32), g["canvas_nonzero"])
    assert np.abs(canvas[canvas != 0] - g["canvas_values"]).max() < 2e-5 * max(1.0, np.abs(g["canvas_values"]).max())


def test_pillarization_and_pfn_vs_oracle(orc):
    """BASELINE config 2 at full size: 4 synthetic frames pillarized (0.16 m, P=32, 12k pillars) by the voxel-hash
    kernel, then the fused PFN + scatter -- against the oracle's voxel generator and numpy PFN."""
    cfg = S.PILLARS
    frames = [S.kitti_frame(s) for s in range(4)]
    gen = orc.VoxelGenerator(cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
    vox, coords, num = orc.collate([gen.generate(f) for f in frames])
    pts = torch.from_numpy(np.concatenate(frames)).cuda()
    offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device="cuda")
    v = F.voxelize(pts, offs, 4, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
    n = int(v["voxel_offsets"][-1])
    assert n == vox.shape[0]
    np.testing.assert_array_equal(v["coordinates"][:n].cpu().numpy(), coords)          # bit-exact pillar order
    np.testing.assert_array_equal(v["voxels"][:n].cpu().numpy(), vox)
    rng = np.random.default_rng(0)
    w = rng.normal(0, 0.3, (64, 10)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 64).astype(np.float32)
    shift = rng.normal(0, 0.3, 64).astype(np.float32)
    ref = orc.pillar_vfe(vox, num, coords, w, scale, shift, cfg["voxel_size"], cfg["point_cloud_range"])
    vs, rg = cfg["voxel_size"], cfg["point_cloud_range"]
    off = (vs[0] / 2 + rg[0], vs[1] / 2 + rg[1], vs[2] / 2 + rg[2])
    feats, canvas = F.pillar_vfe(v["voxels"], v["num_points"], v["coordinates"], torch.from_numpy(w).cuda(),
                                 torch.from_numpy(scale).cuda(), torch.from_numpy(shift).cuda(), vs, off,
                                 canvas_shape=[1, 496, 432], batch_size=4, n_dev=v["voxel_offsets"][-1:])
    got = feats[:n].cpu().numpy()
    assert np.abs(got - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    np.testing.assert_allclose(canvas.cpu().numpy(), orc.pillar_scatter(got, coords, 4, [1, 496, 432]), atol=0)
