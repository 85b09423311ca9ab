"""CPU tests: the oracle (oracle/pcdet_oracle.c) against independent restatements, torch's dense
conv3d, exact geometry and the golden vectors produced by the reference's own Python."""
import os

import numpy as np
import pytest
import torch

from pcdet_b200 import synthetic as S
from util import sort_rows

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---------------------------------------------------------------------------------------- voxelize
def brute_voxelize(points, vs, rng, P, max_voxels, overflow_break):
    """Dict-based first-come voxeliser written independently of the C loop."""
    vs = np.asarray(vs, np.float32)
    lo = np.asarray(rng[:3], np.float32)
    grid = np.round((np.asarray(rng[3:], np.float32) - lo) / vs).astype(np.int64)
    ids, vox, cnt = {}, [], []
    for i in range(points.shape[0]):
        c = np.floor((points[i, :3] - lo) / vs)
        if not (np.all(c >= 0) and np.all(c < grid)):
            continue
        key = (int(c[2]), int(c[1]), int(c[0]))
        if key not in ids:
            if len(ids) >= max_voxels:
                if overflow_break:
                    break
                continue
            ids[key] = len(ids)
            vox.append([])
            cnt.append(0)
        v = ids[key]
        if cnt[v] < P:
            vox[v].append(i)
            cnt[v] += 1
    return list(ids.keys()), vox


@pytest.mark.parametrize("overflow_break", [True, False])
def test_voxelize_oracle_vs_bruteforce(orc, overflow_break):
    cfg = dict(voxel_size=(0.5, 0.5, 0.5), point_cloud_range=(0, -4, -2, 8, 4, 2))
    pts = S.uniform_cloud(3000, cfg["point_cloud_range"], seed=3)
    g = orc.VoxelGenerator(cfg["voxel_size"], cfg["point_cloud_range"], 3, 200, overflow_break)
    vox, coors, num, pidx = g.generate(pts, return_point_idx=True)
    keys, lists = brute_voxelize(pts, cfg["voxel_size"], cfg["point_cloud_range"], 3, 200, overflow_break)
    assert [tuple(c) for c in coors.tolist()] == keys
    for v, l in enumerate(lists):
        assert num[v] == len(l)
        assert pidx[v, :len(l)].tolist() == l
        np.testing.assert_array_equal(vox[v, :len(l)], pts[l])
        assert np.all(vox[v, len(l):] == 0)
    # the lookup grid is restored
    assert np.all(g._lut == -1)


def test_voxelize_oracle_kitti_counts(orc):
    """the seed-0 KITTI-shaped frame has 19 953 points -> 16 774 voxels."""
    pts = S.kitti_frame(0)
    g = orc.VoxelGenerator(S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    vox, coors, num = g.generate(pts)
    assert pts.shape == (19953, 4) and vox.shape == (16774, 5, 4)
    assert g.grid_size.tolist() == [1408, 1600, 40]
    assert len({tuple(c) for c in coors.tolist()}) == coors.shape[0]


def test_voxelize_oracle_empty_and_all_outside(orc):
    g = orc.VoxelGenerator((1, 1, 1), (0, 0, 0, 4, 4, 4), 2, 10)
    v, c, n = g.generate(np.zeros((0, 4), np.float32))
    assert v.shape == (0, 2, 4) and c.shape == (0, 3)
    v, c, n = g.generate(np.full((5, 4), 9.0, np.float32))
    assert v.shape[0] == 0
    # upper boundary is exclusive, lower inclusive
    v, c, n = g.generate(np.array([[0, 0, 0, 1], [4, 0, 0, 1], [3.999, 3.999, 3.999, 1]], np.float32))
    assert c.tolist() == [[0, 0, 0], [3, 3, 3]]


def test_vfe_mean_matches_reference_python(orc):
    g = np.load(os.path.join(GOLD, "ref_python.npz"))
    out = orc.vfe_mean(g["vfe_voxels"], g["vfe_num"])
    np.testing.assert_allclose(out, g["vfe_mean"], rtol=1e-6, atol=1e-6)


def test_bev_conversion_matches_reference_python(orc):
    g = np.load(os.path.join(GOLD, "ref_python.npz"))
    np.testing.assert_array_equal(orc.boxes3d_to_bev(g["boxes3d"]), g["boxes_bev"])


# ---------------------------------------------------------------------------------------- rulebook
def random_sites(rng, n, batch, shape):
    cells = rng.choice(batch * int(np.prod(shape)), size=n, replace=False)
    b, rem = np.divmod(cells, int(np.prod(shape)))
    z, rem = np.divmod(rem, shape[1] * shape[2])
    y, x = np.divmod(rem, shape[2])
    return np.stack([b, z, y, x], axis=1).astype(np.int32)


@pytest.mark.parametrize("ks,st,pd", [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)),
                                        ((3, 1, 1), (2, 1, 1), (0, 0, 0)), ((3, 3, 3), (1, 1, 1), (1, 1, 1)),
                                        ((2, 2, 2), (2, 2, 2), (0, 0, 0))])
def test_rulebook_conv_properties(orc, ks, st, pd):
    rng = np.random.default_rng(5)
    shape = [9, 14, 12]
    idx = random_sites(rng, 400, 2, shape)
    out_ids, pairs, num, out_shape = orc.get_indice_pairs(idx, 2, shape, ks, st, pd, 1, subm=False)
    assert out_shape == orc.conv_output_size(shape, ks, st, pd, (1, 1, 1))
    assert len({tuple(r) for r in out_ids.tolist()}) == out_ids.shape[0]
    total = 0
    for k in range(pairs.shape[0]):
        kz, ky, kx = k // (ks[1] * ks[2]), (k // ks[2]) % ks[1], k % ks[2]
        i, o = pairs[k, 0, :num[k]], pairs[k, 1, :num[k]]
        assert np.all(pairs[k, :, num[k]:] == -1)
        a, b = idx[i], out_ids[o]
        assert np.all(a[:, 0] == b[:, 0])
        # out*stride - pad + k*dil == in  
        for d, kk in zip(range(3), (kz, ky, kx)):
            assert np.all(b[:, d + 1] * st[d] - pd[d] + kk == a[:, d + 1])
        assert len(set(o.tolist())) == len(o)      # an output row appears once per offset
        total += int(num[k])
    # brute-force count of (input, offset) combinations with an in-bounds output
    expect = 0
    for row in idx:
        for kz in range(ks[0]):
            for ky in range(ks[1]):
                for kx in range(ks[2]):
                    t = [row[1] + pd[0] - kz, row[2] + pd[1] - ky, row[3] + pd[2] - kx]
                    if all(v >= 0 and v % s == 0 and v // s < os_ for v, s, os_ in zip(t, st, out_shape)):
                        expect += 1
    assert total == expect
    # every output site is touched by at least one pair, first-touch order is by input row
    assert set(np.concatenate([pairs[k, 1, :num[k]] for k in range(pairs.shape[0])]).tolist()) == set(range(out_ids.shape[0]))


def test_rulebook_subm_symmetry_and_centre(orc):
    rng = np.random.default_rng(6)
    shape = [7, 11, 13]
    idx = random_sites(rng, 500, 2, shape)
    out_ids, pairs, num, _ = orc.get_indice_pairs(idx, 2, shape, 3, 1, 0, 1, subm=True)
    assert out_ids is not None and num[13] == idx.shape[0]
    np.testing.assert_array_equal(pairs[13, 0, :num[13]], pairs[13, 1, :num[13]])
    for k in range(27):
        fwd = {(int(a), int(b)) for a, b in zip(pairs[k, 0, :num[k]], pairs[k, 1, :num[k]])}
        bwd = {(int(b), int(a)) for a, b in zip(pairs[26 - k, 0, :num[26 - k]], pairs[26 - k, 1, :num[26 - k]])}
        assert fwd == bwd
    # known answer: two x-adjacent voxels
    two = np.array([[0, 1, 1, 1], [0, 1, 1, 2]], np.int32)
    _, p, n, _ = orc.get_indice_pairs(two, 1, [3, 3, 4], 3, 1, 0, 1, subm=True)
    assert n.tolist() == [0] * 12 + [1, 2, 1] + [0] * 12
    # offset 14 = (kz,ky,kx)=(1,1,2): in = out + 1 along x -> (in 1, out 0); offset 12: (in 0, out 1)
    assert (p[14, 0, 0], p[14, 1, 0]) == (1, 0) and (p[12, 0, 0], p[12, 1, 0]) == (0, 1)


# ---------------------------------------------------------------------------------------- conv vs dense conv3d
@pytest.mark.parametrize("subm", [True, False])
def test_indice_conv_matches_dense_conv3d(orc, subm):
    rng = np.random.default_rng(7)
    shape, batch, cin, cout = [6, 9, 8], 2, 5, 7
    idx = random_sites(rng, 150, batch, shape)
    feat = rng.normal(0, 1, (idx.shape[0], cin)).astype(np.float32)
    ks, st, pd = (3, 3, 3), ((1, 1, 1) if subm else (2, 2, 2)), (1, 1, 1)
    w = rng.normal(0, 0.2, (*ks, cin, cout)).astype(np.float32)
    out_ids, pairs, num, out_shape = orc.get_indice_pairs(idx, batch, shape, ks, st, pd, 1, subm=subm)
    y = orc.indice_conv(feat, w, pairs, num, out_ids.shape[0], subm=subm)
    y2 = orc.indice_conv_mm(feat, w, pairs, num, out_ids.shape[0], subm=subm)
    y64 = orc.indice_conv(feat, w, pairs, num, out_ids.shape[0], subm=subm, acc64=True)
    np.testing.assert_allclose(y, y64, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(y2, y64, rtol=1e-4, atol=1e-5)
    dense_in = torch.from_numpy(orc.to_dense(feat, idx, shape, batch))
    wt = torch.from_numpy(w).permute(4, 3, 0, 1, 2).contiguous()          # (Cout, Cin, kz, ky, kx)
    dense_out = torch.nn.functional.conv3d(dense_in.double(), wt.double(), stride=st, padding=pd).numpy()
    got = dense_out[out_ids[:, 0], :, out_ids[:, 1], out_ids[:, 2], out_ids[:, 3]]
    np.testing.assert_allclose(y64, got, rtol=1e-5, atol=1e-6)
    if not subm:
        # a regular sparse conv activates every site with a non-empty receptive field: nothing else is non-zero
        mask = np.zeros(dense_out.shape[:1] + dense_out.shape[2:], bool)
        mask[out_ids[:, 0], out_ids[:, 1], out_ids[:, 2], out_ids[:, 3]] = True
        assert np.abs(dense_out.transpose(0, 2, 3, 4, 1)[~mask]).max() == 0


@pytest.mark.parametrize("subm", [True, False])
def test_indice_conv_backward_matches_dense_conv3d_autograd(orc, subm):
    """The oracle's restatement of indiceConvBackward against torch autograd through the DENSE conv3d:
    gradient of sum(conv3d(dense(x), w)[active output sites] * g) with respect to x at the active input sites and to w."""
    rng = np.random.default_rng(11)
    shape, batch, cin, cout = [6, 9, 8], 2, 5, 7
    idx = random_sites(rng, 150, batch, shape)
    feat = rng.normal(0, 1, (idx.shape[0], cin)).astype(np.float32)
    ks, st, pd = (3, 3, 3), ((1, 1, 1) if subm else (2, 2, 2)), (1, 1, 1)
    w = rng.normal(0, 0.2, (*ks, cin, cout)).astype(np.float32)
    out_ids, pairs, num, _ = orc.get_indice_pairs(idx, batch, shape, ks, st, pd, 1, subm=subm)
    g = rng.normal(0, 1, (out_ids.shape[0], cout)).astype(np.float32)
    ib, fb = orc.indice_conv_backward(feat, w, g, pairs, num, subm=subm)
    x = torch.from_numpy(feat).double().requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    ii = torch.from_numpy(idx).long()
    dense_in = torch.zeros((batch, *shape, cin), dtype=torch.float64).index_put((ii[:, 0], ii[:, 1], ii[:, 2], ii[:, 3]), x)
    dense_in = dense_in.permute(0, 4, 1, 2, 3)
    dense_out = torch.nn.functional.conv3d(dense_in, wt.permute(4, 3, 0, 1, 2), stride=st, padding=pd)
    oo = torch.from_numpy(out_ids).long()
    (dense_out[oo[:, 0], :, oo[:, 1], oo[:, 2], oo[:, 3]] * torch.from_numpy(g).double()).sum().backward()
    np.testing.assert_allclose(ib, x.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fb, wt.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_backbone_oracle_shapes(orc):
    """Level shapes of SURVEY App. A.5 and the dense output (B, 256, 200, 176) on a tiny cloud."""
    pts = S.kitti_frame(0)[::40]
    g = orc.VoxelGenerator(S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    v, c, n = g.generate(pts)
    vox, coords, num = orc.collate([(v, c, n)])
    feat = orc.vfe_mean(vox, num)
    w = S.backbone_weights(4, 0)
    col = {}
    out = orc.backbone8x(feat, coords, [41, 1600, 1408], 1, w, conv=orc.indice_conv_mm, collect=col)
    assert out.shape == (1, 256, 200, 176)
    assert [col[k]["shape"] for k in ("conv2.0.0", "conv3.0.0", "conv4.0.0", "conv_out.0")] == \
        [[21, 800, 704], [11, 400, 352], [5, 200, 176], [2, 200, 176]]


# ---------------------------------------------------------------------------------------- IoU / NMS
def test_iou_known_answers(orc):
    a = np.array([[0, 0, 2, 2, 0.0]], np.float32)
    b = np.array([[1, 0, 3, 2, 0.0], [0, 0, 2, 2, 0.0], [5, 5, 6, 6, 0.3], [0, 0, 2, 2, np.pi / 2]], np.float32)
    iou = orc.boxes_iou_bev(a, b)[0]
    np.testing.assert_allclose(iou, [1 / 3, 1.0, 0.0, 1.0], atol=1e-5)
    ov = orc.boxes_overlap_bev(a, b)[0]
    np.testing.assert_allclose(ov, [2.0, 4.0, 0.0, 4.0], atol=1e-4)
    # 45 degree square inside a bigger square: octagon-free case, area of the rotated square
    c = np.array([[-1, -1, 1, 1, np.pi / 4]], np.float32)
    big = np.array([[-5, -5, 5, 5, 0.0]], np.float32)
    np.testing.assert_allclose(orc.boxes_overlap_bev(c, big), [[4.0]], atol=1e-4)
    # two unit-offset 45-degree squares: regular octagon of the 2x2 square, area 8*(sqrt(2)-1)
    sq = np.array([[-1, -1, 1, 1, 0.0]], np.float32)
    np.testing.assert_allclose(orc.boxes_overlap_bev(c, sq), [[8 * (np.sqrt(2) - 1)]], atol=1e-4)


def test_iou_fp32_oracle_vs_exact_geometry(orc):
    b3, _ = S.nms_boxes(400, seed=2, clustered=True)
    bev = orc.boxes3d_to_bev(b3)
    i32 = orc.boxes_iou_bev(bev, bev)
    i64 = orc.boxes_iou_bev64(bev, bev)
    assert np.abs(i32 - i64).max() < 5e-4
    assert (i64 > 0.01).sum() > 1000   # the clustered generator really produces overlapping boxes


def test_nms_oracle_greedy_semantics(orc):
    b3, scores = S.nms_boxes(600, seed=4, clustered=True)
    order = np.argsort(-scores, kind="stable")
    bev = orc.boxes3d_to_bev(b3)[order]
    for thr in (0.01, 0.7):
        keep = orc.nms_sorted(bev, thr)
        iou = orc.boxes_iou_bev(bev, bev)
        kept = np.zeros(len(bev), bool)
        kept[keep] = True
        # kept boxes do not suppress one another; every dropped box is suppressed by an earlier kept one
        sub = iou[np.ix_(keep, keep)]
        assert (np.triu(sub, 1) > thr).sum() == 0
        for j in np.nonzero(~kept)[0]:
            assert (iou[keep[keep < j], j] > thr).any()
    # wrapper returns original indices in score order
    k2 = orc.nms(orc.boxes3d_to_bev(b3), scores, 0.01)
    assert np.all(np.diff(scores[k2]) < 0)
    keep_n = orc.nms_sorted(bev, 0.3, normal=True)
    assert 0 < len(keep_n) <= len(bev)


def test_nms_oracle_edge_cases(orc):
    assert orc.nms_sorted(np.zeros((0, 5), np.float32), 0.5).shape == (0,)
    one = np.array([[0, 0, 1, 1, 0.2]], np.float32)
    assert orc.nms_sorted(one, 0.5).tolist() == [0]
    same = np.repeat(one, 70, axis=0)
    assert orc.nms_sorted(same, 0.5).tolist() == [0]


# ---------------------------------------------------------------------------------------- reference-kernel goldens
def test_iou_oracle_matches_reference_kernel_golden(orc):
    """tests/golden/nms_ref.npz holds outputs of the REFERENCE's own CUDA kernels (iou3d_nms_kernel.cu compiled
    into oracle/_ref, run on a B200 by tests/golden/make_golden_gpu.py).  The C restatement must reproduce
    them: same fp32 algorithm, libm vs CUDA sin/cos/atan2 and FMA contraction are the only differences."""
    g = np.load(os.path.join(GOLD, "nms_ref.npz"))
    iou = orc.boxes_iou_bev(g["iou_a"], g["iou_b"])
    ov = orc.boxes_overlap_bev(g["iou_a"], g["iou_b"])
    assert np.abs(iou - g["iou"]).max() < 1e-4
    assert np.abs(ov - g["overlap"]).max() < 1e-3
    assert (g["iou"] > 0).sum() > 100       # the golden really contains overlapping pairs


@pytest.mark.parametrize("name,normal", [("nms_t001", False), ("nms_t07", False), ("nms_normal_t05", True)])
def test_nms_oracle_matches_reference_kernel_golden(orc, name, normal):
    g = np.load(os.path.join(GOLD, "nms_ref.npz"))
    keep = orc.nms_sorted(g[name + "_boxes"], float(g[name + "_thresh"]), normal=normal)
    np.testing.assert_array_equal(keep, g[name + "_keep"])


def _pillar_golden():
    g = np.load(os.path.join(GOLD, "ref_pillars.npz"))
    bn_w, bn_b = g["sd_pfn_layers.0.norm.weight"], g["sd_pfn_layers.0.norm.bias"]
    mu, var = g["sd_pfn_layers.0.norm.running_mean"], g["sd_pfn_layers.0.norm.running_var"]
    scale = (bn_w / np.sqrt(var + np.float32(1e-3))).astype(np.float32)
    shift = (bn_b - mu * scale).astype(np.float32)
    return g, scale, shift


def test_pillar_vfe_oracle_matches_reference_python_golden(orc):
    """tests/golden/ref_pillars.npz was produced by the REFERENCE's PillarFeatureNetOld2 / PointPillarsScatter
    (tests/golden/make_golden.py pillars).  The numpy restatement must reproduce it to fp32 rounding."""
    g, scale, shift = _pillar_golden()
    feats = orc.pillar_vfe(g["voxels"], g["num_points"], g["coords"], g["sd_pfn_layers.0.linear.weight"], scale, shift,
                           g["voxel_size"], g["pc_range"])
    assert np.abs(feats - g["features"]).max() < 2e-5 * max(1.0, np.abs(g["features"]).max())
    canvas = orc.pillar_scatter(g["features"], g["coords"], 2, [1, 496, 432])
    nz = np.argwhere(canvas != 0).astype(np.int32)
    np.testing.assert_array_equal(nz, g["canvas_nonzero"])
    np.testing.assert_array_equal(canvas[canvas != 0], g["canvas_values"])
    # padded slots take part in the max: some pillar's channel equals relu(shift) exactly
    part = g["num_points"] < 32
    assert (np.isclose(g["features"][part], np.maximum(shift, 0)[None, :], atol=1e-6)).any()


def test_roiaware_oracle_matches_reference_kernel_golden(orc):
    """tests/golden/roiaware_ref.npz holds outputs of the REFERENCE's roiaware_pool3d_kernel.cu (compiled into
    oracle/_ref, run on a B200 by tests/golden/make_golden_gpu.py roiaware): point lists, argmax and the max pooling
    must be identical, the average within fp32 rounding (FMA contraction in the CUDA build)."""
    g = np.load(os.path.join(GOLD, "roiaware_ref.npz"))
    o, mp = int(g["out_size"]), int(g["max_pts"])
    pooled, arg, idx = orc.roiaware_pool3d(g["rois"], g["pts"], g["feat"], o, mp, "max")
    np.testing.assert_array_equal(idx, g["max_idx"])
    np.testing.assert_array_equal(arg, g["max_argmax"])
    np.testing.assert_array_equal(pooled, g["max_pooled"])
    pooled, _, idx = orc.roiaware_pool3d(g["rois"], g["pts"], g["feat"], o, mp, "avg")
    np.testing.assert_array_equal(idx, g["avg_idx"])
    np.testing.assert_allclose(pooled, g["avg_pooled"], rtol=0, atol=1e-6)
    assert (idx[..., 0] == mp - 1).any()                                         # some voxel lists are full
    boxes = np.stack([g["rois"], g["rois"][::-1].copy()])
    points = np.stack([g["pts"], g["pts"]])
    np.testing.assert_array_equal(orc.points_in_boxes(points, boxes), g["box_idx_of_points"])


def test_postprocess_oracle_matches_reference_python_golden(orc):
    """decode (box_coder_utils.py:89-144) and class_agnostic_nms (detector3d.py:278-299) of the REFERENCE, run by
    tests/golden/make_golden.py postprocess, against the oracle's restatement."""
    g = np.load(os.path.join(GOLD, "ref_postprocess.npz"))
    kw = dict(num_dir_bins=2, dir_offset=float(g["dir_offset"]), dir_limit_offset=float(g["dir_limit_offset"]))
    for binary, key in ((False, "decoded"), (True, "decoded_binary")):
        dec = orc.decode_boxes(g["box"], g["anchors"][None], g["dir"], use_binary_dir_classifier=binary, **kw)
        # torch and numpy exp differ by an ulp; everything else is the same fp32 operation sequence
        np.testing.assert_allclose(dec, g[key], rtol=2e-6, atol=2e-6)
    res = orc.post_process(g["cls"], g["box"], g["anchors"], g["dir"], score_thresh=float(g["score_thresh"]),
                           nms_thresh=float(g["nms_thresh"]), pre_max=int(g["pre_max"]), post_max=int(g["post_max"]), **kw)
    for b, r in enumerate(res):
        np.testing.assert_array_equal(r["pre_nms"]["scores"], g[f"nms_in_scores_{b}"])
        np.testing.assert_allclose(orc.boxes3d_to_bev(r["pre_nms"]["boxes"]), g[f"nms_in_boxes_{b}"], rtol=2e-6, atol=2e-6)
        np.testing.assert_array_equal(r["selected"], g[f"selected_{b}"])
        np.testing.assert_array_equal(r["labels"], g[f"labels_{b}"])
    assert len(res[1]["pre_nms"]["scores"]) < int(g["pre_max"]) == len(res[0]["pre_nms"]["scores"])
    # Part-A2 bridge: the reference's proposal_layer on its own decoded boxes
    prop = orc.proposal_layer(g["cls"], g["decoded"], int(g["prop_pre_max"]), int(g["prop_post_max"]), float(g["prop_nms_thresh"]))
    np.testing.assert_array_equal(prop["roi_raw_scores"], g["prop_raw_scores"])
    np.testing.assert_array_equal(prop["roi_labels"], g["prop_labels"])
    np.testing.assert_array_equal(prop["rois"], g["prop_rois"])


def test_ingest_oracle_matches_reference_python_golden(orc):
    """FOV flag / range mask of the reference's KITTI dataset class (tests/golden/make_golden.py ingest)."""
    g = np.load(os.path.join(GOLD, "ref_ingest.npz"))
    for f in range(2):
        pts = g[f"points_{f}"]
        flag, shadow = orc.fov_flag(pts, g[f"V2C_{f}"], g[f"R0_{f}"], g[f"P2_{f}"], g[f"img_shape_{f}"])
        np.testing.assert_array_equal(flag, g[f"fov_flag_{f}"])
        kept = orc.filter_points([pts], [dict(V2C=g[f"V2C_{f}"], R0=g[f"R0_{f}"], P2=g[f"P2_{f}"])], [g[f"img_shape_{f}"]], g["pc_range"])[0]
        np.testing.assert_array_equal(kept, g[f"kept_{f}"])
        ok = np.abs(g[f"depth_{f}"]) > 1.0                          # pixel coordinates explode next to the camera plane
        np.testing.assert_allclose(shadow[ok, :2], g[f"pts_img_{f}"][ok], rtol=1e-3, atol=2e-2)


def test_postprocess_oracle_properties(orc):
    """class_agnostic_select: threshold is inclusive on the sigmoid (detector3d.py:279), order is score-descending with
    ties to the lower anchor, labels are first-argmax + 1, at most pre_max survive; proposal_layer pads as the reference."""
    rng = np.random.default_rng(0)
    cls = np.round(rng.normal(-1, 1.5, (500, 3)) * 2).astype(np.float32) / 2          # many exact ties
    t = 0.25
    sel, sc, lab = orc.class_agnostic_select(cls, t, 64)
    rank = cls.max(axis=-1)
    assert len(sel) == min(64, int((orc.sigmoid32(rank) >= np.float32(t)).sum()))
    assert (np.diff(sc) <= 0).all()
    same = np.diff(sc) == 0
    assert (np.diff(sel)[same] > 0).all()                                            # ties: lower anchor first
    np.testing.assert_array_equal(lab, cls[sel].argmax(axis=-1) + 1)
    # everything that was left out scores no higher than the last one taken (or fails the threshold)
    left = np.setdiff1d(np.arange(500), sel)
    ok = orc.sigmoid32(rank[left]) >= np.float32(t)
    assert (rank[left][ok] <= sc[-1]).all()
    # threshold exactly on a candidate's sigmoid: kept (>=)
    x = np.float32(0.75)
    one = np.array([[x, -9, -9]], np.float32)
    assert len(orc.class_agnostic_select(one, float(orc.sigmoid32(x)), 4)[0]) == 1
    assert len(orc.class_agnostic_select(one, float(np.nextafter(orc.sigmoid32(x), np.float32(1))), 4)[0]) == 0
    # proposal_layer padding (proposal_layer.py:14-23)
    boxes = np.concatenate([rng.uniform(0, 50, (1, 10, 3)), rng.uniform(1, 2, (1, 10, 3)), rng.uniform(-3, 3, (1, 10, 1))], axis=2).astype(np.float32)
    out = orc.proposal_layer(rng.normal(0, 1, (1, 10, 3)).astype(np.float32), boxes, 8, 16, 0.7)
    n = int((out["roi_raw_scores"][0] > -100000).sum())
    assert 1 <= n <= 8 and (out["rois"][0, n:] == 0).all() and (out["roi_labels"][0, n:] == 1).all()
