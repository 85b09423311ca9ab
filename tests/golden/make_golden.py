"""Generates tests/golden/ref_python.npz by importing the REFERENCE's own Python (from /root/reference,
which exists only in the build container) for the two torch-level functions on the hot path:
  - MeanVoxelFeatureExtractor.forward      pcdet/models/vfe/vfe_utils.py:26-34
  - boxes3d_to_bevboxes_lidar_torch        pcdet/utils/box_utils.py:237-250
and records the state-dict layout of the reference BackBone8x (pcdet/models/rpn/rpn_backbone.py)
instantiated on the pcdet_b200.spconv modules.  Run:  python tests/golden/make_golden.py
`python tests/golden/make_golden.py unet` records the state-dict layout of the reference's UNetV2 (rpn_unet.py);
`python tests/golden/make_golden.py pillars` writes ref_pillars.npz the same way for PointPillars
(PillarFeatureNetOld2, vfe_utils.py:118-215, and PointPillarsScatter, rpn/pillar_scatter.py).
`python tests/golden/make_golden.py ingest` writes ref_ingest.npz (FOV / range filter of the KITTI dataset class);
`python tests/golden/make_golden.py postprocess` writes ref_postprocess.npz: the reference's
ResidualCoder.decode_with_head_direction_torch (box_coder_utils.py:113-144) and Detector3D.class_agnostic_nms
(detector3d.py:278-299) on seeded head outputs.
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


def postprocess():
    """tests/golden/ref_postprocess.npz.  The reference's own decode and class_agnostic_nms run on CPU tensors; the
    one thing it cannot run here is its CUDA NMS, so `iou3d_nms_utils.nms_gpu` is the oracle's NMS (itself pinned
    against the reference kernel by nms_ref.npz) and the boxes / scores the reference hands to it are recorded."""
    from oracle import oracle
    stubs = stub_packages()
    for name in ["pcdet.models.detectors", "pcdet.models.rcnn", "pcdet.models.bbox_heads", "pcdet.ops.iou3d_nms"]:
        m = types.ModuleType(name)
        m.__path__ = []
        stubs[name] = m
    for pkg, child in [("pcdet.models.vfe", "vfe_modules"), ("pcdet.models.rpn", "rpn_modules"), ("pcdet.models.rcnn", "rcnn_modules"),
                       ("pcdet.models.bbox_heads", "bbox_head_modules")]:
        m = types.ModuleType(pkg + "." + child)
        stubs[pkg + "." + child] = m
        setattr(stubs[pkg], child, m)
    del stubs["pcdet.utils.common_utils"]
    cu = load_reference_module("pcdet/utils/common_utils.py", "pcdet.utils.common_utils", stubs)
    stubs["pcdet.utils"].common_utils = cu
    sys.modules.setdefault("scipy", __import__("scipy"))
    bu = load_reference_module("pcdet/utils/box_utils.py", "pcdet.utils.box_utils", stubs)
    stubs["pcdet.utils"].box_utils = bu
    bc = load_reference_module("pcdet/utils/box_coder_utils.py", "pcdet.utils.box_coder_utils", stubs)
    handed = []

    def nms_gpu(boxes, scores, thresh):
        handed.append((boxes.numpy().copy(), scores.numpy().copy()))
        return torch.from_numpy(oracle.nms(boxes.numpy(), scores.numpy(), thresh))

    nu = types.ModuleType("pcdet.ops.iou3d_nms.iou3d_nms_utils")
    nu.nms_gpu = nms_gpu
    sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_utils"] = nu
    stubs["pcdet.ops.iou3d_nms"].iou3d_nms_utils = nu
    pre_max, post_max, score_thresh, nms_thresh = 512, 100, 0.1, 0.01
    stubs["pcdet.config"].cfg["MODEL"] = _Cfg(TEST=_Cfg(NMS_PRE_MAXSIZE_LAST=pre_max, NMS_POST_MAXSIZE_LAST=post_max))
    det = load_reference_module("pcdet/models/detectors/detector3d.py", "pcdet.models.detectors.detector3d", stubs)

    rng = np.random.default_rng(4242)
    B, H, W = 2, 25, 20
    sizes = np.array([[1.6, 3.9, 1.56], [0.6, 0.8, 1.73], [0.6, 1.76, 1.73]], np.float32)      # second.yaml anchors (w, l, h)
    zs = np.array([-1.78, -0.6, -0.6], np.float32)
    ys, xs = np.meshgrid(np.linspace(-39, 39, H), np.linspace(1, 69, W), indexing="ij")
    anchors = np.zeros((H, W, 3, 2, 7), np.float32)
    anchors[..., 0], anchors[..., 1] = xs[:, :, None, None], ys[:, :, None, None]
    anchors[..., 2] = zs[None, None, :, None]
    anchors[..., 3:6] = sizes[None, None, :, None, :]
    anchors[..., 6] = np.array([0, np.pi / 2], np.float32)[None, None, None, :]
    anchors = anchors.reshape(-1, 7)
    A = anchors.shape[0]
    cls = rng.normal(-1.5, 1.5, (B, A, 3)).astype(np.float32)
    cls[1] -= 3.5                                                        # frame 1: fewer candidates than pre_max
    box = rng.normal(0, 0.3, (B, A, 7)).astype(np.float32)
    dirp = rng.normal(0, 1, (B, A, 2)).astype(np.float32)
    kw = dict(num_dir_bins=2, dir_offset=0.78539, dir_limit_offset=0.0)
    coder = bc.ResidualCoder()
    out = {}
    for binary in (False, True):
        dec = coder.decode_with_head_direction_torch(
            box_preds=torch.from_numpy(box), anchors=torch.from_numpy(anchors).view(1, A, 7).repeat(B, 1, 1),
            dir_cls_preds=torch.from_numpy(dirp), use_binary_dir_classifier=binary, **kw)
        out["decoded_binary" if binary else "decoded"] = dec.numpy()
    dec = torch.from_numpy(out["decoded"])
    for b in range(B):
        # detector3d.py:193-197 (the caller of class_agnostic_nms; needs a dataset object, so restated in two lines)
        rank, labels = torch.max(torch.from_numpy(cls[b]), dim=-1)
        selected = det.Detector3D.class_agnostic_nms(rank_scores=rank, normalized_scores=torch.sigmoid(rank), box_preds=dec[b],
                                                     score_thresh=score_thresh, nms_thresh=nms_thresh, nms_type="nms_gpu")
        out[f"selected_{b}"] = selected.numpy()
        out[f"labels_{b}"] = (labels + 1)[selected].numpy()
        out[f"nms_in_boxes_{b}"], out[f"nms_in_scores_{b}"] = handed[-1]
    # Part-A2 bridge: the reference's proposal_layer (model_utils/proposal_layer.py) on the same decoded boxes
    nu.nms_normal_gpu = nms_gpu
    stubs["pcdet.config"].cfg["MODEL"]["TEST"].update(NMS_PRE_MAXSIZE=300, NMS_POST_MAXSIZE=64, RPN_NMS_TYPE="nms_gpu", RPN_NMS_THRESH=0.7)
    for name in ["pcdet.models.model_utils"]:
        if name not in stubs:
            m = types.ModuleType(name)
            m.__path__ = []
            stubs[name] = m
    pl = load_reference_module("pcdet/models/model_utils/proposal_layer.py", "pcdet.models.model_utils.proposal_layer", stubs)
    roi = pl.proposal_layer(B, torch.from_numpy(cls), dec, code_size=7, mode="TEST")
    out.update(prop_rois=roi["rois"].numpy(), prop_raw_scores=roi["roi_raw_scores"].numpy(), prop_labels=roi["roi_labels"].numpy(),
               prop_pre_max=300, prop_post_max=64, prop_nms_thresh=0.7)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_postprocess.npz"), cls=cls, box=box, dir=dirp, anchors=anchors,
                        pre_max=pre_max, post_max=post_max, score_thresh=score_thresh, nms_thresh=nms_thresh,
                        dir_offset=kw["dir_offset"], dir_limit_offset=kw["dir_limit_offset"], **out)
    print("wrote ref_postprocess.npz:", A, [len(out[f"selected_{b}"]) for b in range(B)], [len(out[f"nms_in_scores_{b}"]) for b in range(B)])


def ingest():
    """tests/golden/ref_ingest.npz: the reference's Calibration.lidar_to_rect / rect_to_img (utils/calibration.py:66-85),
    BaseKittiDataset.get_fov_flag (kitti_dataset.py:236-253) and mask_points_by_range (common_utils.py:47-51) on a
    synthetic KITTI-shaped cloud with a KITTI-like calibration."""
    import pcdet_b200.spconv as sp
    from pcdet_b200 import synthetic as S
    sp.install_as_spconv()
    stubs = stub_packages()
    del stubs["pcdet.utils.common_utils"]
    for name in ["skimage", "skimage.io", "cv2", "pcdet.datasets", "pcdet.datasets.data_augmentation",
                 "pcdet.datasets.data_augmentation.dbsampler", "pcdet.utils.box_utils", "pcdet.utils.object3d_utils"]:
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            m.__path__ = []
            stubs[name] = m
    stubs["skimage"].io = stubs["skimage.io"]
    stubs["pcdet.datasets.data_augmentation.dbsampler"].DataBaseSampler = object
    stubs["pcdet.datasets"].DatasetTemplate = object
    cu = load_reference_module("pcdet/utils/common_utils.py", "pcdet.utils.common_utils", stubs)
    cal = load_reference_module("pcdet/utils/calibration.py", "pcdet.utils.calibration", stubs)
    for k, m in (("common_utils", cu), ("calibration", cal), ("box_utils", stubs["pcdet.utils.box_utils"]),
                 ("object3d_utils", stubs["pcdet.utils.object3d_utils"])):
        setattr(stubs["pcdet.utils"], k, m)
    kd = load_reference_module("pcdet/datasets/kitti/kitti_dataset.py", "pcdet.datasets.kitti.kitti_dataset", stubs)
    # KITTI-like calibration (the numbers of a typical calib file, perturbed per frame)
    rng = np.random.default_rng(99)
    out = {}
    pc_range = np.array(S.KITTI["point_cloud_range"], np.float32)
    for f in range(2):
        P2 = np.array([[721.5377, 0, 609.5593, 44.85728], [0, 721.5377, 172.854, 0.2163791], [0, 0, 1, 0.002745884]], np.float32)
        R0 = np.array([[0.9999239, 0.00983776, -0.007445048], [-0.009869795, 0.9999421, -0.004278459],
                       [0.007402527, 0.004351614, 0.9999631]], np.float32)
        V2C = np.array([[0.007533745, -0.9999714, -0.000616602, -0.004069766], [0.01480249, 0.0007280733, -0.9998902, -0.07631618],
                        [0.9998621, 0.00752379, 0.01480755, -0.2717806]], np.float32)
        P2[:, 3] += rng.normal(0, 0.01, 3).astype(np.float32)
        V2C[:, 3] += rng.normal(0, 0.01, 3).astype(np.float32)
        calib = cal.Calibration({"P2": P2, "R0": R0, "Tr_velo2cam": V2C})
        img_shape = np.array([375, 1242]) if f == 0 else np.array([370, 1224])
        # full 360-degree cloud so that the FOV filter has something to remove
        n = 8000
        ang, rad = rng.uniform(-np.pi, np.pi, n), rng.uniform(2, 75, n)
        pts = np.stack([rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-2.5, 1.5, n), rng.uniform(0, 1, n)], axis=1).astype(np.float32)
        pts_rect = calib.lidar_to_rect(pts[:, 0:3])                                  # kitti_dataset.py:715
        flag = kd.BaseKittiDataset.get_fov_flag(pts_rect, img_shape, calib)          # :716
        kept = cu.mask_points_by_range(pts[flag], pc_range)                          # dataset.py:184
        pts_img, depth = calib.rect_to_img(pts_rect)
        out.update({f"P2_{f}": P2, f"R0_{f}": R0, f"V2C_{f}": V2C, f"img_shape_{f}": img_shape, f"points_{f}": pts,
                    f"fov_flag_{f}": flag, f"kept_{f}": kept, f"pts_img_{f}": pts_img.astype(np.float32), f"depth_{f}": depth.astype(np.float32)})
        print("frame", f, "points", n, "in fov", int(flag.sum()), "kept", kept.shape[0])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_ingest.npz"), pc_range=pc_range, **out)
    print("wrote ref_ingest.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "unet":
        unet()
    elif len(sys.argv) > 1 and sys.argv[1] == "ingest":
        ingest()
    elif len(sys.argv) > 1 and sys.argv[1] == "postprocess":
        postprocess()
    elif len(sys.argv) > 1 and sys.argv[1] == "pillars":
        pillars()
    else:
        main()
