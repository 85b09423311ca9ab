"""GPU parity of the static-shape module API (SparseConvTensor.n_dev) and of the captured Part-A^2 bridge
(pcdet_b200/parta2.py; PartA2_net.py:15-83, partA2_rcnn_net.py:256-295) against the same modules run with exact shapes
(which tests/test_gpu_unet.py pins to the oracle's restatement of rpn_unet.py) and against the reference's own
roiaware_pool3d kernel compiled into oracle/_ref."""
import numpy as np
import pytest
import torch

import pcdet_b200.spconv as spconv
from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200.backbone import BackBone8x
from pcdet_b200.ops.roiaware_pool3d import roiaware_pool3d_utils as R
from pcdet_b200.parta2 import PartA2Config, PartA2HotPath
from pcdet_b200.postprocess import PostProcessor
from pcdet_b200.unet import UNetV2

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPE = [41, 1600, 1408]


def voxels(seeds, sub=2):
    frames = [S.kitti_frame(s)[::sub] for s in seeds]
    pts = torch.from_numpy(np.concatenate(frames)).to(DEV)
    offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device=DEV)
    return frames, pts, offs


def randomize_bn(net, seed=1):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.weight.data = torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.05
            m.running_mean.data = torch.randn(m.running_mean.shape, generator=g) * 0.05
            m.running_var.data = torch.rand(m.running_var.shape, generator=g) * 0.5 + 0.75
    return net


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_static_shape_mode_is_bit_identical_to_exact_shapes(dtype):
    """BackBone8x and UNetV2, unmodified module trees: capacity-sized tensors + device counts give the same bits in the
    valid rows and the same dense map as the exact-shape run (which synchronises at every strided rulebook)."""
    _frames, pts, offs = voxels((0, 1))
    v = F.voxelize(pts, offs, 2, S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000, want_mean=True, mean_dtype=dtype)
    n = int(v["voxel_offsets"][-1])
    cap = v["coordinates"].shape[0]
    assert cap > n
    torch.manual_seed(3)
    for net in (randomize_bn(BackBone8x(4)).eval().to(DEV).to(dtype), randomize_bn(UNetV2(4)).eval().to(DEV).to(dtype)):
        with torch.no_grad():
            exact = net(spconv.SparseConvTensor(v["mean"][:n].contiguous(), v["coordinates"][:n].contiguous(), SHAPE, 2))
            x = spconv.SparseConvTensor(v["mean"], v["coordinates"], SHAPE, 2, n_dev=v["voxel_offsets"][2:3])
            static = net(x)
        assert all(int(o) == 0 for o in x.indice_dict["__overflow__"])
        assert torch.equal(static["spatial_features"], exact["spatial_features"])
        if "seg_features" in exact:
            assert torch.equal(static["seg_features"][:n], exact["seg_features"])
            assert torch.equal(static["u_seg_preds"][:n], exact["u_seg_preds"])


def head_outputs(B, anchors, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = anchors.shape[0]
    cls = torch.randn((B, a, 1), device=DEV, generator=g) * 2 - 1
    box = torch.randn((B, a, 7), device=DEV, generator=g) * 0.2
    dirp = torch.randn((B, a, 2), device=DEV, generator=g)
    return cls, box, dirp


def make_anchors(n=20000, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.rand((n, 7), device=DEV, generator=g) * torch.tensor([65, 70, 0.5, 0.4, 1.0, 0.3, 3.14], device=DEV) \
        + torch.tensor([3, -35, -1.9, 1.5, 3.6, 1.4, 0], device=DEV)


def test_captured_parta2_bridge_matches_eager_modules_and_reference_pooling():
    B = 2
    frames, pts, offs = voxels((2, 3))
    cfg = PartA2Config(batch_size=B, max_points_total=pts.shape[0], dtype=torch.float32)
    torch.manual_seed(5)
    net = randomize_bn(UNetV2(4))
    anchors = make_anchors()
    hp = PartA2HotPath(cfg, net, anchors)
    cls, box, dirp = head_outputs(B, anchors)
    out = hp.capture(pts, offs, cls, box, dirp)
    hp.replay()
    torch.cuda.synchronize()
    assert int(out["overflow"].sum()) == 0
    n = int(out["voxel_offsets"][B])
    # --- the same through the exact-shape module API and the eager post-processing -----------------------------------------
    v = F.voxelize(pts, offs, B, cfg.voxel_size, cfg.point_cloud_range, 5, 40000)
    feats = F.vfe_mean(v["voxels"][:n], v["num_points"][:n])
    coords = v["coordinates"][:n].contiguous()
    with torch.no_grad():
        u = hp.net(spconv.SparseConvTensor(feats, coords, SHAPE, B))
    assert torch.equal(out["coordinates"][:n], coords)
    assert torch.equal(out["seg_features"][:n], u["seg_features"])
    assert torch.equal(out["spatial_features"], u["spatial_features"])
    prop = PostProcessor(anchors, cfg.proposals).proposals(cls, box, dirp)
    assert torch.equal(out["rois"], prop["rois"]) and torch.equal(out["num_rois"], prop["num"])
    assert int(out["num_rois"].min()) > 0
    # --- pooling: frame by frame as partA2_rcnn_net.py:272-290 does, through the module (and the reference's kernel) -------
    pool = R.RoIAwarePool3d(cfg.roi_pool_size, cfg.max_pts_each_voxel)
    centers = out["voxel_centers"][:n]
    P = out["rois"].shape[1]
    for b in range(B):
        m = coords[:, 0] == b
        want_part = pool(out["rois"][b].contiguous(), centers[m], out["part_features"][:n][m], "avg")
        want_seg = pool(out["rois"][b].contiguous(), centers[m], u["seg_features"][m].float(), "max")
        assert torch.equal(out["pooled_part_features"][b * P:(b + 1) * P], want_part)
        assert torch.equal(out["pooled_rpn_features"][b * P:(b + 1) * P], want_seg)
    assert float(out["pooled_rpn_features"].abs().sum()) > 0
    # a second, smaller batch through the same graph: nothing is re-captured, counts come from the device
    _f2, pts2, offs2 = voxels((4, 5), sub=3)
    pts.zero_()
    pts[:pts2.shape[0]].copy_(pts2)
    offs.copy_(offs2)
    hp.replay()
    torch.cuda.synchronize()
    n2 = int(out["voxel_offsets"][B])
    v2 = F.voxelize(pts2, offs2, B, cfg.voxel_size, cfg.point_cloud_range, 5, 40000)
    assert n2 == int(v2["voxel_offsets"][B]) and n2 < n
    with torch.no_grad():
        u2 = hp.net(spconv.SparseConvTensor(F.vfe_mean(v2["voxels"][:n2], v2["num_points"][:n2]), v2["coordinates"][:n2].contiguous(), SHAPE, B))
    assert torch.equal(out["seg_features"][:n2], u2["seg_features"])


@pytest.mark.parametrize("post", [False, True])
def test_residual_epilogue_of_the_tensor_core_conv(post):
    """pcdb_sparse_conv_fwd_ex: residual added before the epilogue (partial sum) or behind it (shortcut), against torch fp32."""
    torch.manual_seed(7)
    K, n_in, n_out, cin, cout = 27, 3000, 2500, 64, 64
    g = torch.Generator(device=DEV).manual_seed(1)
    nbr = torch.where(torch.rand((K, n_out), device=DEV, generator=g) < 0.3,
                      torch.randint(0, n_in, (K, n_out), device=DEV, dtype=torch.int32, generator=g),
                      torch.full((K, n_out), -1, dtype=torch.int32, device=DEV)).contiguous()
    x = torch.randn(n_in, cin, device=DEV).bfloat16()
    w = (torch.randn(K, cin, cout, device=DEV) * 0.1).bfloat16()
    r = torch.randn(n_out, cout, device=DEV).bfloat16()
    scale, shift = torch.rand(cout, device=DEV) + 0.5, torch.randn(cout, device=DEV) * 0.1
    got = F.sparse_conv_fwd(x, w, nbr, n_out, scale=scale, shift=shift, relu=True, residual=r, residual_post=post)
    y = torch.zeros(n_out, cout, device=DEV)
    for k in range(K):
        idx = nbr[k].long()
        m = idx >= 0
        y[m] += x[idx[m]].float() @ w[k].float()
    ref = torch.relu(y * scale + shift + r.float()) if post else torch.relu((y + r.float()) * scale + shift)
    assert float((got.float() - ref).abs().max() / ref.abs().max()) < 1e-2


def test_128_input_channels_run_as_two_tensor_core_launches():
    """UNetV2's merge convolutions (128 -> 64): channel halves through the residual epilogue against the FMA-pipe kernel."""
    torch.manual_seed(9)
    conv = spconv.SubMConv3d(128, 64, 3, bias=False, indice_key="k").to(DEV).to(torch.bfloat16).eval()
    _f, pts, offs = voxels((0,), sub=4)
    v = F.voxelize(pts, offs, 1, S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    n = int(v["voxel_offsets"][-1])
    coords = v["coordinates"][:n].contiguous()
    feats = torch.randn(n, 128, device=DEV).bfloat16()
    bn = torch.nn.BatchNorm1d(64).to(DEV).eval()
    bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
    with torch.no_grad():
        got = conv(spconv.SparseConvTensor(feats, coords, SHAPE, 1), fused_bn=bn, fused_relu=True)
        rb = got.indice_dict["k"]
        scale, shift = conv._folded_bn(bn)
        ref = F.sparse_conv_fwd(feats, conv._weight3d(torch.bfloat16).detach(), rb.nbr, n, scale=scale, shift=shift, relu=True, algo=1)
    assert float((got.features.float() - ref.float()).abs().max() / ref.float().abs().max()) < 1e-2


def test_graphed_sparse_module_replays_other_batches():
    """spconv.GraphedSparseModule: BackBone8x captured once through the static-shape mode, replayed on two other batches of
    different sizes -- dense maps bit-identical to the eager exact-shape runs, no overflow."""
    net = randomize_bn(BackBone8x(4)).eval().to(DEV).to(torch.bfloat16)
    runner = spconv.GraphedSparseModule(net, capacity=2 * 12000, channels=4, spatial_shape=SHAPE, batch_size=2, dtype=torch.bfloat16)
    for seeds, sub in (((0, 1), 2), ((2, 3), 3), ((4, 5), 2)):
        _f, pts, offs = voxels(seeds, sub=sub)
        v = F.voxelize(pts, offs, 2, S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000, want_mean=True, mean_dtype=torch.bfloat16)
        n = int(v["voxel_offsets"][-1])
        feats, coords = v["mean"][:n].contiguous(), v["coordinates"][:n].contiguous()
        out = runner(feats, coords)["spatial_features"]
        runner.check_overflow()
        with torch.no_grad():
            ref = net(spconv.SparseConvTensor(feats, coords, SHAPE, 2))["spatial_features"]
        assert torch.equal(out, ref)
