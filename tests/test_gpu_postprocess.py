"""GPU parity: decode + threshold + top-k front of the post-processing through the
C ABI, against the oracle's restatement of detector3d.py:112-299 / box_coder_utils.py:89-144 and against the
reference's own Python (tests/golden/ref_postprocess.npz).

Bar: selected anchor indices, their order, scores, labels and counts bit-exact; decoded boxes within 2e-6 relative
(expf on the device vs exp on the host differ by an ulp; everything else is the same fp32 operation sequence)."""
import os

import numpy as np
import pytest
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200._lib import PcdbError
from pcdet_b200.postprocess import PostProcessConfig, PostProcessor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
BOX_TOL = dict(rtol=2e-6, atol=2e-6)


def anchors_grid(rng, n):
    a = np.zeros((n, 7), np.float32)
    a[:, 0] = rng.uniform(0, 70, n); a[:, 1] = rng.uniform(-40, 40, n); a[:, 2] = rng.uniform(-2, -0.5, n)
    a[:, 3:6] = np.array([[1.6, 3.9, 1.56], [0.6, 0.8, 1.73], [0.6, 1.76, 1.73]], np.float32)[rng.integers(0, 3, n)]
    a[:, 6] = np.array([0, np.pi / 2], np.float32)[rng.integers(0, 2, n)]
    return a


def head_outputs(seed, batch, n_anchors, n_classes=3, bins=2, mean=-1.5):
    rng = np.random.default_rng(seed)
    cls = rng.normal(mean, 1.5, (batch, n_anchors, n_classes)).astype(np.float32)
    box = rng.normal(0, 0.3, (batch, n_anchors, 7)).astype(np.float32)
    dirp = rng.normal(0, 1, (batch, n_anchors, bins)).astype(np.float32) if bins else None
    return cls, box, dirp, anchors_grid(rng, n_anchors)


def away_from_threshold(orc, cls, thresh, margin=1e-6):
    """moves the (rare) anchors whose sigmoid sits within `margin` of the threshold, where expf's last bit decides"""
    rank = cls.max(axis=-1)
    close = np.abs(orc.sigmoid32(rank).astype(np.float64) - thresh) < margin
    cls[close] += np.float32(0.01)
    return cls


def run_front(cls, box, dirp, anchors, **kw):
    out = F.decode_select(torch.from_numpy(cls).cuda(), torch.from_numpy(box).cuda(), torch.from_numpy(anchors).cuda(),
                          None if dirp is None else torch.from_numpy(dirp).cuda(), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def check_front(orc, got, cls, box, dirp, anchors, score_thresh, pre_max, **dirkw):
    for b in range(cls.shape[0]):
        sel, scores, labels = orc.class_agnostic_select(cls[b], score_thresh, pre_max)
        n = len(sel)
        assert got["count"][b] == n
        np.testing.assert_array_equal(got["anchor_index"][b, :n], sel)
        np.testing.assert_array_equal(got["scores"][b, :n], scores)
        np.testing.assert_array_equal(got["labels"][b, :n], labels)
        dec = orc.decode_boxes(box[b][sel], anchors[sel], None if dirp is None else dirp[b][sel], **dirkw)
        np.testing.assert_allclose(got["boxes3d"][b, :n], dec, **BOX_TOL)
        np.testing.assert_allclose(got["boxes_bev"][b, :n], orc.boxes3d_to_bev(dec), **BOX_TOL)
        # padding rows
        assert (got["anchor_index"][b, n:] == -1).all() and (got["labels"][b, n:] == 0).all()
        assert (got["boxes3d"][b, n:] == 0).all()
        pad = got["boxes_bev"][b, n:]
        assert (pad[:, 0] == pad[:, 2]).all() and (pad[:, 1] == pad[:, 3]).all() and (pad[:, 0] >= 1e6).all()


def test_front_matches_reference_python_golden(orc):
    g = np.load(os.path.join(GOLD, "ref_postprocess.npz"))
    kw = dict(num_dir_bins=2, dir_offset=float(g["dir_offset"]), dir_limit_offset=float(g["dir_limit_offset"]))
    got = run_front(g["cls"], g["box"], g["dir"], g["anchors"], score_thresh=float(g["score_thresh"]), pre_max=int(g["pre_max"]), **kw)
    for b in range(2):
        ref_scores, ref_boxes = g[f"nms_in_scores_{b}"], g[f"nms_in_boxes_{b}"]
        n = len(ref_scores)
        assert got["count"][b] == n
        np.testing.assert_array_equal(got["scores"][b, :n], ref_scores)          # what the reference hands to nms_gpu
        np.testing.assert_allclose(got["boxes_bev"][b, :n], ref_boxes, **BOX_TOL)
        np.testing.assert_allclose(got["boxes3d"][b, :n], g["decoded"][b][got["anchor_index"][b, :n]], **BOX_TOL)
    gotb = run_front(g["cls"], g["box"], g["dir"], g["anchors"], score_thresh=float(g["score_thresh"]), pre_max=int(g["pre_max"]),
                     use_binary_dir_classifier=True, **kw)
    for b in range(2):
        n = gotb["count"][b]
        np.testing.assert_allclose(gotb["boxes3d"][b, :n], g["decoded_binary"][b][gotb["anchor_index"][b, :n]], **BOX_TOL)


def test_post_processor_matches_reference_python_golden(orc):
    """end to end: kept anchor indices and labels of Detector3D.class_agnostic_nms"""
    g = np.load(os.path.join(GOLD, "ref_postprocess.npz"))
    cfg = PostProcessConfig(score_thresh=float(g["score_thresh"]), nms_thresh=float(g["nms_thresh"]), nms_pre_maxsize=int(g["pre_max"]),
                            nms_post_maxsize=int(g["post_max"]), dir_offset=float(g["dir_offset"]), dir_limit_offset=float(g["dir_limit_offset"]))
    pp = PostProcessor(torch.from_numpy(g["anchors"]).cuda(), cfg)
    recs = pp(torch.from_numpy(g["cls"]).cuda(), torch.from_numpy(g["box"]).cuda(), torch.from_numpy(g["dir"]).cuda())
    for b, r in enumerate(recs):
        np.testing.assert_array_equal(r["selected"].cpu().numpy(), g[f"selected_{b}"])
        np.testing.assert_array_equal(r["labels"].cpu().numpy(), g[f"labels_{b}"])
        np.testing.assert_allclose(r["boxes"].cpu().numpy(), g["decoded"][b][g[f"selected_{b}"]], **BOX_TOL)
        np.testing.assert_array_equal(r["scores"].cpu().numpy(), g["cls"][b].max(axis=-1)[g[f"selected_{b}"]])


def test_proposals_match_reference_proposal_layer_golden(orc):
    """Part-A2 stage-1 -> stage-2 bridge: decode + proposal_layer (model_utils/proposal_layer.py) of the reference"""
    g = np.load(os.path.join(GOLD, "ref_postprocess.npz"))
    cfg = PostProcessConfig(score_thresh=0.0, nms_thresh=float(g["prop_nms_thresh"]), nms_pre_maxsize=int(g["prop_pre_max"]),
                            nms_post_maxsize=int(g["prop_post_max"]), dir_offset=float(g["dir_offset"]), dir_limit_offset=float(g["dir_limit_offset"]))
    pp = PostProcessor(torch.from_numpy(g["anchors"]).cuda(), cfg)
    out = pp.proposals(torch.from_numpy(g["cls"]).cuda(), torch.from_numpy(g["box"]).cuda(), torch.from_numpy(g["dir"]).cuda())
    np.testing.assert_array_equal(out["roi_raw_scores"].cpu().numpy(), g["prop_raw_scores"])
    np.testing.assert_array_equal(out["roi_labels"].cpu().numpy(), g["prop_labels"])
    np.testing.assert_allclose(out["rois"].cpu().numpy(), g["prop_rois"], **BOX_TOL)
    assert out["roi_labels"].dtype == torch.int64 and (out["num"].cpu().numpy() <= int(g["prop_post_max"])).all()


@pytest.mark.parametrize("batch,n_anchors,n_classes,bins,pre_max,mean", [
    (2, 211200, 3, 2, 4096, -1.5),      # SECOND: 200 x 176 x 6 anchors (second.yaml), NMS_PRE_MAXSIZE_LAST 4096
    (4, 211200, 3, 2, 4096, -6.0),      # trained-network regime: a few hundred candidates per frame
    (1, 321408, 1, 2, 1000, -1.0),      # PointPillars-sized head, one class, pre_max not a power of two
    (3, 100, 3, 0, 4096, 0.0),          # fewer anchors than pre_max, no direction head
    (2, 50000, 2, 4, 16384, 1.0),       # largest supported pre_max, four direction bins
])
def test_front_vs_oracle(orc, batch, n_anchors, n_classes, bins, pre_max, mean):
    cls, box, dirp, anchors = head_outputs(7 + n_anchors, batch, n_anchors, n_classes, bins, mean)
    cls = away_from_threshold(orc, cls, 0.1)
    dirkw = dict(num_dir_bins=bins, dir_offset=0.78539, dir_limit_offset=0.0) if bins else {}
    got = run_front(cls, box, dirp, anchors, score_thresh=0.1, pre_max=pre_max, **dirkw)
    check_front(orc, got, cls, box, dirp, anchors, 0.1, pre_max, **dirkw)


def test_front_ties_go_to_the_lower_anchor(orc):
    """zero-initialised head: every anchor has the same score; and scores quantised to 1/4 (ties across the k-th)"""
    cls, box, dirp, anchors = head_outputs(3, 2, 30000)
    cls[0] = 0.0
    cls[1] = np.round(cls[1] * 4) / 4
    got = run_front(cls, box, dirp, anchors, score_thresh=0.1, pre_max=2048, dir_offset=0.78539)
    check_front(orc, got, cls, box, dirp, anchors, 0.1, 2048, num_dir_bins=2, dir_offset=0.78539, dir_limit_offset=0.0)
    np.testing.assert_array_equal(got["anchor_index"][0], np.arange(2048))


def test_front_no_candidates_and_background_column(orc):
    cls, box, dirp, anchors = head_outputs(5, 2, 5000, n_classes=4)
    cls[0] = -12.0                                                            # nothing above the threshold in frame 0
    cfg = PostProcessConfig(encode_background_as_zeros=False, nms_pre_maxsize=512, nms_post_maxsize=50, use_raw_score=False)
    pp = PostProcessor(torch.from_numpy(anchors).cuda(), cfg)
    recs = pp(torch.from_numpy(cls).cuda(), torch.from_numpy(box).cuda(), torch.from_numpy(dirp).cuda())
    assert recs[0]["boxes"].shape == (0, 7) and recs[0]["scores"].numel() == 0
    ref = orc.post_process(cls[..., 1:], box, anchors, dirp, score_thresh=0.1, nms_thresh=0.01, pre_max=512, post_max=50,
                           num_dir_bins=2, dir_offset=0.78539, dir_limit_offset=0.0)
    np.testing.assert_array_equal(recs[1]["selected"].cpu().numpy(), ref[1]["selected"])
    np.testing.assert_array_equal(recs[1]["labels"].cpu().numpy(), ref[1]["labels"])
    np.testing.assert_allclose(recs[1]["scores"].cpu().numpy(), orc.sigmoid32(ref[1]["scores"]), rtol=1e-6)   # USE_RAW_SCORE False
    np.testing.assert_allclose(recs[1]["boxes"].cpu().numpy(), ref[1]["boxes"], **BOX_TOL)


def test_post_processor_full_size_vs_oracle(orc):
    """SECOND-sized frames through decode_select + pcdb_nms.  Anchors whose decoded boxes form a pair with
    |IoU - thresh| < 1e-5 (where fp32 rounding decides) are pushed below the score threshold first."""
    cls, box, dirp, anchors = head_outputs(11, 2, 211200, mean=-4.5)
    cls = away_from_threshold(orc, cls, 0.1)
    kw = dict(score_thresh=0.1, nms_thresh=0.01, pre_max=4096, post_max=500, num_dir_bins=2, dir_offset=0.78539, dir_limit_offset=0.0)
    for _ in range(20):
        ref = orc.post_process(cls, box, anchors, dirp, **kw)
        clean = True
        for b in range(2):
            bev = orc.boxes3d_to_bev(ref[b]["pre_nms"]["boxes"])
            rows, _ = np.nonzero(np.abs(orc.boxes_iou_bev64(bev, bev) - 0.01) < 1e-5)
            if rows.size:
                cls[b, ref[b]["pre_nms"]["selected"][np.unique(rows)]] = -12.0
                clean = False
        if clean:
            break
    assert clean
    pp = PostProcessor(torch.from_numpy(anchors).cuda(), PostProcessConfig())
    recs = pp(torch.from_numpy(cls).cuda(), torch.from_numpy(box).cuda(), torch.from_numpy(dirp).cuda())
    for b in range(2):
        assert len(ref[b]["pre_nms"]["selected"]) == 4096
        np.testing.assert_array_equal(recs[b]["selected"].cpu().numpy(), ref[b]["selected"])
        np.testing.assert_array_equal(recs[b]["scores"].cpu().numpy(), ref[b]["scores"])
        np.testing.assert_array_equal(recs[b]["labels"].cpu().numpy(), ref[b]["labels"])


def test_front_argument_errors():
    cls, box, dirp, anchors = head_outputs(1, 1, 64)
    with pytest.raises(PcdbError):
        run_front(cls, box, dirp, anchors, pre_max=20000)


@pytest.mark.parametrize("nms_thresh", [0.1, [0.1, 0.3, 0.5]])
def test_multi_classes_nms_matches_the_reference_loop(orc, nms_thresh):
    """Detector3D.multi_classes_nms (detector3d.py:239-276) restated with the oracle's NMS, class by class, on boxes from
    which near-threshold IoU pairs were removed (margin_safe_boxes, for every threshold in play)."""
    from pcdet_b200.postprocess import multi_classes_nms
    from util import margin_safe_boxes
    rng = np.random.default_rng(11)
    n, C = 600, 3
    b3, _ = S.nms_boxes(n, seed=5)
    bev = orc.boxes3d_to_bev(b3)
    for t in 2 * (nms_thresh if isinstance(nms_thresh, list) else [nms_thresh]):      # twice: a move may break an earlier threshold
        bev = margin_safe_boxes(orc, bev, t, rng=rng)
    # back to LiDAR boxes with the moved footprints (x, y = centre; w, l = extents)
    b3[:, 0], b3[:, 1] = (bev[:, 0] + bev[:, 2]) / 2, (bev[:, 1] + bev[:, 3]) / 2
    rank = rng.permutation(n * C).reshape(n, C).astype(np.float32) / (n * C)          # distinct scores: no ties
    norm = 1 / (1 + np.exp(-(rank * 8 - 4)))
    score_thresh = [0.3, 0.5, 0.2]
    sel, lab = multi_classes_nms(torch.from_numpy(rank).cuda(), torch.from_numpy(norm.astype(np.float32)).cuda(), torch.from_numpy(b3).cuda(),
                                 score_thresh, nms_thresh)
    ref_sel, ref_lab = [], []
    bev_now = orc.boxes3d_to_bev(b3)
    for k in range(C):
        idx = np.nonzero(norm[:, k].astype(np.float32) >= np.float32(score_thresh[k]))[0]
        if idx.size == 0:
            continue
        t = nms_thresh[k] if isinstance(nms_thresh, list) else nms_thresh
        keep = orc.nms(bev_now[idx], rank[idx, k], t)
        ref_sel.append(idx[keep]); ref_lab.append(np.full(keep.shape[0], k + 1))
    np.testing.assert_array_equal(sel.cpu().numpy(), np.concatenate(ref_sel))
    np.testing.assert_array_equal(lab.cpu().numpy(), np.concatenate(ref_lab))
    # nothing above the thresholds: empty result
    sel0, lab0 = multi_classes_nms(torch.from_numpy(rank).cuda(), torch.zeros((n, C), device="cuda"), torch.from_numpy(b3).cuda(), 0.5, 0.1)
    assert sel0.numel() == 0 and lab0.numel() == 0
