"""GPU parity: rotated IoU matrices and NMS (through the C ABI) against the oracle and, when the
compiled reference kernel (oracle/_ref) travelled to the box, against the reference itself."""
import ctypes
import os

import numpy as np
import pytest
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200.ops.iou3d_nms import iou3d_nms_cuda, iou3d_nms_utils
from util import margin_safe_boxes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_iou3d.so")


def sorted_bev(orc, n, seed, clustered=True):
    b3, scores = S.nms_boxes(n, seed=seed, clustered=clustered)
    order = np.argsort(-scores, kind="stable")
    return orc.boxes3d_to_bev(b3)[order], b3, scores


def ref_lib():
    if not os.path.exists(REF_SO):
        pytest.skip("compiled reference kernel not present")
    L = ctypes.CDLL(REF_SO)
    return L


def ref_nms(L, boxes_t, thresh, normal=False):
    n = boxes_t.shape[0]
    keep = np.zeros((max(n, 1),), np.int64)
    cnt = L.ref_nms(ctypes.c_void_p(boxes_t.data_ptr()), n, ctypes.c_float(thresh), int(normal),
                    keep.ctypes.data_as(ctypes.c_void_p), None)
    assert cnt >= 0
    return keep[:cnt]


def test_iou_matrices_vs_oracle(orc):
    a, _, _ = sorted_bev(orc, 300, 1)
    b, _, _ = sorted_bev(orc, 200, 2)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    iou = F.boxes_iou_bev(ta, tb).cpu().numpy()
    ov = F.boxes_overlap_bev(ta, tb).cpu().numpy()
    assert np.abs(iou - orc.boxes_iou_bev(a, b)).max() < 2e-4      # two fp32 algorithms for the same polygon
    assert np.abs(iou - orc.boxes_iou_bev64(a, b)).max() < 1e-4    # exact geometry
    assert np.abs(ov - orc.boxes_overlap_bev(a, b)).max() < 2e-3
    # pybind-style entry points write into caller-allocated outputs and return 1
    ans = torch.zeros((300, 200), device="cuda")
    assert iou3d_nms_cuda.boxes_iou_bev_gpu(ta, tb, ans) == 1
    np.testing.assert_array_equal(ans.cpu().numpy(), iou)
    np.testing.assert_array_equal(iou3d_nms_utils.boxes_iou_bev(ta, tb).cpu().numpy(), iou)


def test_iou_known_answers():
    a = torch.tensor([[0, 0, 2, 2, 0.0], [-1, -1, 1, 1, np.pi / 4]], device="cuda")
    b = torch.tensor([[1, 0, 3, 2, 0.0], [0, 0, 2, 2, 0.0], [5, 5, 6, 6, 0.3], [0, 0, 2, 2, np.pi / 2],
                      [-1, -1, 1, 1, 0.0]], device="cuda")
    iou = F.boxes_iou_bev(a, b).cpu().numpy()
    np.testing.assert_allclose(iou[0, :4], [1 / 3, 1.0, 0.0, 1.0], atol=1e-5)
    ov = F.boxes_overlap_bev(a, b).cpu().numpy()
    np.testing.assert_allclose(ov[1, 4], 8 * (np.sqrt(2) - 1), atol=1e-5)


def test_iou_degenerate_geometry(orc):
    """Shared edges and corners, containment, identical boxes, edges that are parallel up to 1e-7 .. 1e-3 rad: the
    configurations where clipping one box's edges against the other is numerically delicate."""
    rng = np.random.default_rng(12)
    boxes = []
    for yaw0 in (0.0, np.pi / 2, np.pi, -np.pi / 2, np.pi / 4):
        for eps in (0.0, 1e-7, -1e-7, 1e-5, -1e-5, 1e-3):
            for _ in range(6):
                cx, cy = rng.integers(0, 4, 2).astype(np.float64)          # integer grid: edges coincide often
                w, l = rng.integers(1, 4, 2).astype(np.float64)
                boxes.append([cx - w / 2, cy - l / 2, cx + w / 2, cy + l / 2, yaw0 + eps])
    b = np.asarray(boxes, np.float32)
    t = torch.from_numpy(b).cuda()
    iou = F.boxes_iou_bev(t, t).cpu().numpy()
    ref = orc.boxes_iou_bev64(b, b)
    assert np.abs(iou - ref).max() < 2e-4, np.unravel_index(np.abs(iou - ref).argmax(), iou.shape)
    assert np.abs(np.diag(iou) - 1.0).max() < 1e-5                          # identical boxes
    ov = F.boxes_overlap_bev(t, t).cpu().numpy()
    np.testing.assert_allclose(ov, ov.T, atol=2e-4)                         # which box is clipped must not matter


def test_bev_conversion_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_python.npz"))
    out = F.boxes3d_to_bev(torch.from_numpy(g["boxes3d"]).cuda()).cpu().numpy()
    np.testing.assert_array_equal(out, g["boxes_bev"])       # produced by the reference's box_utils.py


@pytest.mark.parametrize("n,thresh,clustered", [(4096, 0.01, True), (4096, 0.7, True), (1000, 0.8, True),
                                                 (4096, 0.01, False), (777, 0.1, True), (65, 0.3, True)])
def test_nms_keep_bit_exact_vs_oracle(orc, n, thresh, clustered):
    bev, _, _ = sorted_bev(orc, n, seed=n, clustered=clustered)
    bev = margin_safe_boxes(orc, bev, thresh)
    ref = orc.nms_sorted(bev, thresh)
    t = torch.from_numpy(bev).cuda()
    keep, num = F.nms_sorted_batched(t, [0, n], thresh)
    cnt = int(num.item())
    got = keep[0, :cnt].cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    assert np.all(keep[0, cnt:].cpu().numpy() == -1)
    if os.path.exists(REF_SO):
        np.testing.assert_array_equal(ref_nms(ref_lib(), t, thresh), ref)


def test_nms_candidate_list_overflow(orc):
    """Degenerate input: 1500 boxes piled on one spot, so that nearly every pair passes the circle test and the
    global candidate list (64 pairs per box) overflows -- the strips are then redone by the resolve kernel's
    slow path and the keep list must still be exact."""
    rng = np.random.default_rng(3)
    n = 1500
    b3 = np.zeros((n, 7), np.float32)
    b3[:, 0] = 20 + rng.uniform(-1.0, 1.0, n)
    b3[:, 1] = 3 + rng.uniform(-1.0, 1.0, n)
    b3[:, 3] = rng.uniform(1.4, 1.9, n)
    b3[:, 4] = rng.uniform(3.4, 4.4, n)
    b3[:, 5] = 1.5
    b3[:, 6] = rng.uniform(-np.pi, np.pi, n)
    bev = orc.boxes3d_to_bev(b3)                     # already in "score" order
    for thresh in (0.3, 0.7):
        safe = margin_safe_boxes(orc, bev, thresh)
        keep, num = F.nms_sorted_batched(torch.from_numpy(safe).cuda(), [0, n], thresh)
        got = keep[0, :int(num.item())].cpu().numpy()
        ref = orc.nms_sorted(safe, thresh)
        np.testing.assert_array_equal(got, ref)
        assert 1 < ref.shape[0] < safe.shape[0]


def test_nms_vs_compiled_reference_kernel(orc):
    """Against the reference's own nms_kernel + host sweep (iou3d_nms.cpp:79-126) on unfiltered data:
    the mask may differ only where |IoU - thresh| is within fp32 rounding."""
    L = ref_lib()
    for seed, thresh in ((11, 0.01), (12, 0.7)):
        bev, _, _ = sorted_bev(orc, 4096, seed=seed)
        t = torch.from_numpy(bev).cuda()
        ref = ref_nms(L, t, thresh)
        keep, num = F.nms_sorted_batched(t, [0, 4096], thresh)
        got = keep[0, :int(num.item())].cpu().numpy()
        if not np.array_equal(got, ref):
            # every disagreement must trace back to a near-threshold pair
            iou = orc.boxes_iou_bev64(bev, bev)
            near = np.abs(iou - thresh) < 2e-4
            assert near.any(), "keep lists differ without any near-threshold pair"
            first = np.nonzero(got[:min(len(got), len(ref))] != ref[:min(len(got), len(ref))])[0]
            j = int(min(got[first[0]], ref[first[0]])) if first.size else int(max(got[-1], ref[-1]))
            assert near[:, j].any() or near[j, :].any()


def test_nms_normal_and_wrappers(orc):
    bev, b3, scores = sorted_bev(orc, 1500, seed=3)
    b = orc.boxes3d_to_bev(b3)
    tb, ts = torch.from_numpy(b).cuda(), torch.from_numpy(scores).cuda()
    for thresh in (0.1, 0.5):
        got = iou3d_nms_utils.nms_normal_gpu(tb, ts, thresh).cpu().numpy()
        np.testing.assert_array_equal(got, orc.nms(b, scores, thresh, normal=True))
    safe = margin_safe_boxes(orc, bev, 0.01)
    order = np.argsort(-scores, kind="stable")
    unsorted = np.empty_like(safe)
    unsorted[order] = safe
    got = iou3d_nms_utils.nms_gpu(torch.from_numpy(unsorted).cuda(), ts, 0.01)
    assert got.dtype == torch.int64 and got.is_cuda
    np.testing.assert_array_equal(got.cpu().numpy(), orc.nms(unsorted, scores, 0.01))
    got = iou3d_nms_utils.nms_gpu(torch.from_numpy(unsorted).cuda(), ts, 0.01, pre_maxsize=512)
    np.testing.assert_array_equal(got.cpu().numpy(), orc.nms(unsorted, scores, 0.01, pre_maxsize=512))
    # pybind-style: CPU keep tensor, returns the count (iou3d_nms.cpp:79-126)
    keep = torch.zeros(1500, dtype=torch.int64)
    cnt = iou3d_nms_cuda.nms_gpu(torch.from_numpy(safe).cuda(), keep, 0.01)
    np.testing.assert_array_equal(keep[:cnt].numpy(), orc.nms_sorted(safe, 0.01))
    with pytest.raises(RuntimeError, match="CUDAtensor"):
        iou3d_nms_cuda.nms_gpu(torch.from_numpy(safe), keep, 0.01)


def test_nms_batched_sets_and_edge_cases(orc):
    sets = [margin_safe_boxes(orc, sorted_bev(orc, n, seed=40 + n)[0], 0.1) if n else np.zeros((0, 5), np.float32)
            for n in (0, 1, 63, 64, 65, 500, 129)]
    offs = np.concatenate([[0], np.cumsum([s.shape[0] for s in sets])])
    t = torch.from_numpy(np.concatenate(sets)).cuda()
    keep, num = F.nms_sorted_batched(t, offs, 0.1, keep_stride=100)
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    for s, boxes in enumerate(sets):
        ref = orc.nms_sorted(boxes, 0.1)
        cnt = min(len(ref), 100)
        assert num[s] == cnt
        np.testing.assert_array_equal(keep[s, :cnt], ref[:cnt])
        assert np.all(keep[s, cnt:] == -1)
    same = np.repeat(np.array([[0, 0, 2, 1, 0.3]], np.float32), 200, axis=0)
    k, n = F.nms_sorted_batched(torch.from_numpy(same).cuda(), [0, 200], 0.5)
    assert int(n.item()) == 1 and int(k[0, 0].item()) == 0


def test_boxes_iou3d(orc):
    b3a, _ = S.nms_boxes(100, seed=1)
    b3b, _ = S.nms_boxes(80, seed=2)
    got = iou3d_nms_utils.boxes_iou3d_gpu(torch.from_numpy(b3a).cuda(), torch.from_numpy(b3b).cuda()).cpu().numpy()
    ov = orc.boxes_overlap_bev(orc.boxes3d_to_bev(b3a), orc.boxes3d_to_bev(b3b))
    hmax = np.minimum((b3a[:, 2] + b3a[:, 5])[:, None], (b3b[:, 2] + b3b[:, 5])[None])
    hmin = np.maximum(b3a[:, 2][:, None], b3b[:, 2][None])
    o3 = ov * np.clip(hmax - hmin, 0, None)
    va = (b3a[:, 3] * b3a[:, 4] * b3a[:, 5])[:, None]
    vb = (b3b[:, 3] * b3b[:, 4] * b3b[:, 5])[None]
    ref = o3 / np.clip(va + vb - o3, 1e-6, None)
    assert np.abs(got - ref).max() < 1e-3


@pytest.mark.parametrize("normal", [False, True])
def test_nms_device_side_set_counts(orc, normal):
    """pcdb_nms_counts: capacity-sized sets whose box counts live on the device (the output of pcdb_decode_select) give
    exactly the kept positions of the same sets cut to their counts; rows behind a count may hold anything."""
    cap, thresh = 2048, 0.3
    counts = [0, 1, 63, 64, 65, 1000, 2048, 2047]
    rng = np.random.default_rng(5)
    all_boxes, expect = [], []
    for i, c in enumerate(counts):
        bev, _, _ = sorted_bev(orc, cap, seed=100 + i)
        bev[:c] = margin_safe_boxes(orc, bev[:c], thresh) if c > 1 and not normal else bev[:c]
        garbage = bev[c:].copy()
        garbage[:, :4] += rng.normal(0, 0.5, garbage[:, :4].shape).astype(np.float32)     # overlapping junk behind the count
        all_boxes.append(np.concatenate([bev[:c], garbage]))
        expect.append(orc.nms_sorted(bev[:c], thresh, normal=normal) if c else np.zeros((0,), np.int64))
    boxes = torch.from_numpy(np.concatenate(all_boxes)).cuda()
    offs = [cap * i for i in range(len(counts) + 1)]
    keep, num = F.nms_sorted_batched(boxes, offs, thresh, normal=normal, keep_stride=cap,
                                     set_counts=torch.tensor(counts, dtype=torch.int32, device="cuda"))
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    for i, e in enumerate(expect):
        assert num[i] == len(e)
        np.testing.assert_array_equal(keep[i, :num[i]], e)
        assert (keep[i, num[i]:] == -1).all()
