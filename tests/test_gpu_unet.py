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


def test_unet_v2_bf16_inference_tracks_fp32(orc):
    """bf16 features and weights (the tcgen05 kernels): within 5e-2 of the fp32 module output, max-norm relative, 29
    layers deep (bf16 has 8 mantissa bits; the fp32 path above is the one pinned to the oracle)."""
    g = orc.VoxelGenerator(S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    vox, coords, num = orc.collate([g.generate(S.kitti_frame(s)[::3].copy()) for s in (0, 1)])
    feats = torch.from_numpy(orc.vfe_mean(vox, num)).cuda()
    ct = torch.from_numpy(coords).cuda()
    torch.manual_seed(3)
    net = UNetV2(4).eval().cuda()
    with torch.no_grad():
        ref = net(spconv.SparseConvTensor(feats, ct, [41, 1600, 1408], 2))
        net16 = UNetV2(4).eval().cuda()
        net16.load_state_dict(net.state_dict())
        out = net16.to(torch.bfloat16)(spconv.SparseConvTensor(feats.bfloat16(), ct, [41, 1600, 1408], 2))
    for key in ("seg_features", "u_seg_preds", "u_reg_preds", "spatial_features"):
        assert out[key].dtype == torch.bfloat16
        assert rel_err(out[key].float().cpu().numpy(), ref[key].cpu().numpy()) < 5e-2, key
