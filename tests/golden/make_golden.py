"""Generates tests/golden/ref_python.npz by importing the REFERENCE's own Python (from /root/reference,
which exists only in the build container) for the two torch-level functions on the hot path:
  - MeanVoxelFeatureExtractor.forward      pcdet/models/vfe/vfe_utils.py:26-34
  - boxes3d_to_bevboxes_lidar_torch        pcdet/utils/box_utils.py:237-250
and records the state-dict layout of the reference BackBone8x (pcdet/models/rpn/rpn_backbone.py)
instantiated on the pcdet_b200.spconv modules.  Run:  python tests/golden/make_golden.py
`python tests/golden/make_golden.py unet` records the state-dict layout of the reference's UNetV2 (rpn_unet.py);
`python tests/golden/make_golden.py pillars` writes ref_pillars.npz the same way for PointPillars
(PillarFeatureNetOld2, vfe_utils.py:118-215, and PointPillarsScatter, rpn/pillar_scatter.py).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("PCDET_REFERENCE", "/root/reference")


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def load_reference_module(rel_path, name, stubs):
    for k, v in stubs.items():
        sys.modules[k] = v
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def stub_packages():
    stubs = {}
    for name in ["pcdet", "pcdet.models", "pcdet.models.rpn", "pcdet.models.vfe", "pcdet.models.model_utils",
                 "pcdet.utils", "pcdet.ops", "pcdet.ops.roiaware_pool3d"]:
        m = types.ModuleType(name)
        m.__path__ = []
        stubs[name] = m
    cfgmod = types.ModuleType("pcdet.config")
    cfgmod.cfg = _Cfg(DATA_CONFIG=_Cfg(VOXEL_GENERATOR=_Cfg(VOXEL_SIZE=[0.05, 0.05, 0.1]),
                                       NUM_POINT_FEATURES={"total": 4, "use": 4}))
    stubs["pcdet.config"] = cfgmod
    pu = types.ModuleType("pcdet.models.model_utils.pytorch_utils")
    pu.Empty = torch.nn.Identity
    stubs["pcdet.models.model_utils.pytorch_utils"] = pu
    rp = types.ModuleType("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils")
    stubs["pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils"] = rp
    stubs["pcdet.ops.roiaware_pool3d"].roiaware_pool3d_utils = rp
    cu = types.ModuleType("pcdet.utils.common_utils")
    stubs["pcdet.utils.common_utils"] = cu
    stubs["pcdet.utils"].common_utils = cu
    return stubs


def main():
    import pcdet_b200.spconv as sp
    sp.install_as_spconv()
    stubs = stub_packages()
    vfe = load_reference_module("pcdet/models/vfe/vfe_utils.py", "pcdet.models.vfe.vfe_utils", stubs)
    sys.modules.setdefault("scipy", __import__("scipy"))
    box_utils = load_reference_module("pcdet/utils/box_utils.py", "pcdet.utils.box_utils", stubs)
    bb = load_reference_module("pcdet/models/rpn/rpn_backbone.py", "pcdet.models.rpn.rpn_backbone", stubs)

    rng = np.random.default_rng(1234)
    V, P, Cc = 257, 5, 4
    num = rng.integers(1, P + 1, V).astype(np.int32)
    vox = rng.normal(0, 10, (V, P, Cc)).astype(np.float32)
    for v in range(V):
        vox[v, num[v]:] = 0
    mean = vfe.MeanVoxelFeatureExtractor().forward(torch.from_numpy(vox), torch.from_numpy(num)).numpy()

    b3 = np.concatenate([rng.uniform(-40, 70, (300, 3)), rng.uniform(0.5, 5, (300, 3)),
                         rng.uniform(-np.pi, np.pi, (300, 1))], axis=1).astype(np.float32)
    bev = box_utils.boxes3d_to_bevboxes_lidar_torch(torch.from_numpy(b3)).numpy()

    net = bb.BackBone8x(4)
    keys = list(net.state_dict().keys())
    shapes = [list(v.shape) for v in net.state_dict().values()]
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_python.npz"), vfe_voxels=vox, vfe_num=num,
                        vfe_mean=mean, boxes3d=b3, boxes_bev=bev, backbone_keys=np.array(keys),
                        backbone_shapes=np.array([str(s) for s in shapes]))
    print("wrote ref_python.npz:", mean.shape, bev.shape, len(keys))


def pillars():
    """tests/golden/ref_pillars.npz: the reference's PillarFeatureNetOld2 (vfe_utils.py:118-215, eval mode, non-trivial
    BatchNorm statistics) and PointPillarsScatter (rpn/pillar_scatter.py) on a small random pillar batch."""
    stubs = stub_packages()
    vfe = load_reference_module("pcdet/models/vfe/vfe_utils.py", "pcdet.models.vfe.vfe_utils", stubs)
    ps = load_reference_module("pcdet/models/rpn/pillar_scatter.py", "pcdet.models.rpn.pillar_scatter", stubs)
    rng = np.random.default_rng(77)
    vs, rg = (0.16, 0.16, 4), (0, -39.68, -3, 69.12, 39.68, 1)
    nx, ny, B, P, V = 432, 496, 2, 32, 700
    cells = rng.choice(B * ny * nx, size=V, replace=False)
    b, rem = np.divmod(cells, ny * nx)
    y, x = np.divmod(rem, nx)
    coords = np.stack([b, np.zeros_like(b), y, x], axis=1).astype(np.int32)
    num = np.minimum(rng.geometric(0.15, V), P).astype(np.int32)
    num[:20] = P                                                       # full pillars: no padded slot in the max
    vox = np.zeros((V, P, 4), np.float32)
    for v in range(V):
        n = num[v]
        vox[v, :n, 0] = rg[0] + (x[v] + rng.uniform(0, 1, n)) * vs[0]
        vox[v, :n, 1] = rg[1] + (y[v] + rng.uniform(0, 1, n)) * vs[1]
        vox[v, :n, 2] = rng.uniform(-3, 1, n)
        vox[v, :n, 3] = rng.uniform(0, 1, n)
    torch.manual_seed(5)
    net = vfe.PillarFeatureNetOld2(num_input_features=4, use_norm=True, num_filters=(64,), with_distance=False,
                                   voxel_size=vs, pc_range=rg).eval()
    bn = net.pfn_layers[0].norm
    with torch.no_grad():
        bn.weight.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, 64).astype(np.float32)))
        bn.bias.copy_(torch.from_numpy(rng.normal(0, 0.5, 64).astype(np.float32)))
        bn.running_mean.copy_(torch.from_numpy(rng.normal(0, 0.5, 64).astype(np.float32)))
        bn.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, 64).astype(np.float32)))
        feats = net(torch.from_numpy(vox), torch.from_numpy(num), torch.from_numpy(coords))
        canvas = ps.PointPillarsScatter(64)(feats, torch.from_numpy(coords), B, output_shape=[1, ny, nx])
    sd = {k: v.numpy() for k, v in net.state_dict().items()}
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_pillars.npz"), voxels=vox, num_points=num,
                        coords=coords, features=feats.numpy(), canvas_nonzero=np.argwhere(canvas.numpy() != 0).astype(np.int32),
                        canvas_values=canvas.numpy()[canvas.numpy() != 0], voxel_size=np.array(vs), pc_range=np.array(rg),
                        state_keys=np.array(list(sd.keys())), **{"sd_" + k: v for k, v in sd.items()})
    print("wrote ref_pillars.npz:", feats.shape, canvas.shape, list(sd.keys()))


def unet():
    """tests/golden/ref_unet_keys.npz: state-dict keys and shapes of the REFERENCE's UNetV2
    (pcdet/models/rpn/rpn_unet.py:339-418, with its own SparseBasicBlock from model_utils/resnet_utils.py)
    instantiated on the pcdet_b200.spconv modules."""
    import pcdet_b200.spconv as sp
    sp.install_as_spconv()
    stubs = stub_packages()
    stubs["pcdet.config"].cfg["MODEL"] = _Cfg(RPN=_Cfg(BACKBONE=_Cfg(TARGET_CONFIG=_Cfg(GT_EXTEND_WIDTH=0.2, GENERATED_ON="dataset"))))
    lu = types.ModuleType("pcdet.utils.loss_utils")
    lu.SigmoidFocalClassificationLoss = lambda **kw: torch.nn.Identity()
    stubs["pcdet.utils.loss_utils"] = lu
    stubs["pcdet.utils"].loss_utils = lu
    ru = load_reference_module("pcdet/models/model_utils/resnet_utils.py", "pcdet.models.model_utils.resnet_utils", stubs)
    stubs["pcdet.models.model_utils"].resnet_utils = ru
    stubs["pcdet.models.model_utils.resnet_utils"] = ru
    un = load_reference_module("pcdet/models/rpn/rpn_unet.py", "pcdet.models.rpn.rpn_unet", stubs)
    net = un.UNetV2(4)
    sd = net.state_dict()
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_unet_keys.npz"), keys=np.array(list(sd.keys())),
                        shapes=np.array([str(list(v.shape)) for v in sd.values()]))
    print("wrote ref_unet_keys.npz:", len(sd))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "unet":
        unet()
    elif len(sys.argv) > 1 and sys.argv[1] == "pillars":
        pillars()
    else:
        main()
