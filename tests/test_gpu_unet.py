"""GPU parity: Part-A^2 UNetV2 (sparse encoder + SparseInverseConv3d decoder) on the spconv
module API against the oracle's restatement of rpn_unet.py."""
import numpy as np
import pytest
import torch

from pcdet_b200 import spconv
from pcdet_b200 import synthetic as S
from pcdet_b200.unet import UNetV2
from util import rel_err

pytestmark = pytest.mark.gpu


def test_unet_v2_forward_vs_oracle(orc):
    g = orc.VoxelGenerator(S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    frames = [g.generate(S.kitti_frame(s)[::3].copy()) for s in (0, 1)]        # two thinned frames
    vox, coords, num = orc.collate(frames)
    feats = orc.vfe_mean(vox, num)
    torch.manual_seed(3)
    net = UNetV2(4).eval()
    rng = np.random.default_rng(9)
    with torch.no_grad():                          # non-trivial BatchNorm statistics
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(torch.from_numpy(rng.uniform(0.8, 1.2, m.num_features).astype(np.float32)))
                m.bias.copy_(torch.from_numpy(rng.normal(0, 0.1, m.num_features).astype(np.float32)))
                m.running_mean.copy_(torch.from_numpy(rng.normal(0, 0.1, m.num_features).astype(np.float32)))
                m.running_var.copy_(torch.from_numpy(rng.uniform(0.8, 1.2, m.num_features).astype(np.float32)))
    sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
    ref = orc.unet_v2(feats, coords, [41, 1600, 1408], 2, sd)
    net = net.cuda()
    x = spconv.SparseConvTensor(torch.from_numpy(feats).cuda(), torch.from_numpy(coords).cuda(), [41, 1600, 1408], 2)
    with torch.no_grad():
        out = net(x)
    assert out["seg_features"].shape == (coords.shape[0], 16)
    for key in ("seg_features", "u_seg_preds", "u_reg_preds", "spatial_features"):
        assert rel_err(out[key].cpu().numpy(), ref[key]) < 2e-4, key             # fp32 path, 29 layers deep
