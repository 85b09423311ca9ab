"""Detector post-processing for the class-agnostic NMS path, on the device from head outputs to kept boxes.

Mirror of ``Detector3D.predict_boxes`` / ``post_processing`` / ``class_agnostic_nms``
(pcdet/models/detectors/detector3d.py:112-128, 156-223, 278-299) for the RPN-only detectors (SECOND, PointPillars:
``rcnn_ret_dict is None``, ``MODEL.TEST.MULTI_CLASSES_NMS: False``).  The reference decodes every anchor, then loops
over the frames in Python with boolean-mask indexing (one ``nonzero`` sync per frame), ``torch.topk``, a second sort
inside ``nms_gpu`` and a CPU sweep; here ``pcdb_decode_select`` and ``pcdb_nms`` handle the whole batch in two calls
and nothing is copied to the host until the caller asks for the variable-length result.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import functional as F


@dataclass
class PostProcessConfig:
    """cfg.MODEL.TEST.* and cfg.MODEL.RPN.RPN_HEAD.ARGS.* of tools/cfgs/second.yaml:78-84,152-159."""
    score_thresh: float = 0.1
    nms_thresh: float = 0.01
    nms_pre_maxsize: int = 4096          # NMS_PRE_MAXSIZE_LAST
    nms_post_maxsize: int = 500          # NMS_POST_MAXSIZE_LAST
    use_raw_score: bool = True           # USE_RAW_SCORE
    nms_type: str = "nms_gpu"            # or "nms_normal_gpu"
    num_direction_bins: int = 2
    dir_offset: float = 0.78539
    dir_limit_offset: float = 0.0
    use_binary_dir_classifier: bool = False
    encode_background_as_zeros: bool = True


class PostProcessor:
    def __init__(self, anchors: torch.Tensor, cfg: Optional[PostProcessConfig] = None):
        """anchors: the head's anchor tensor, any shape ending in 7 (rpn_ret_dict['anchors'])."""
        self.cfg = cfg or PostProcessConfig()
        self.anchors = anchors.reshape(-1, 7).contiguous().float()

    def _front_and_nms(self, rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds):
        c = self.cfg
        a = self.anchors.shape[0]
        bsz = rpn_cls_preds.shape[0]
        cls = rpn_cls_preds.reshape(bsz, a, -1).float()                       # detector3d.py:118
        if not c.encode_background_as_zeros:
            cls = cls[..., 1:]                                                # detector3d.py:168-169
        front = F.decode_select(cls.contiguous(), rpn_box_preds.reshape(bsz, a, 7), self.anchors,
                                None if rpn_dir_cls_preds is None else rpn_dir_cls_preds.reshape(bsz, a, -1),
                                score_thresh=c.score_thresh, pre_max=c.nms_pre_maxsize, num_dir_bins=c.num_direction_bins,
                                dir_offset=c.dir_offset, dir_limit_offset=c.dir_limit_offset,
                                use_binary_dir_classifier=c.use_binary_dir_classifier)
        k = c.nms_pre_maxsize
        # the NMS reads the candidate counts on the device: padding rows are never touched
        keep, _ = F.nms_sorted_batched(front["boxes_bev"].view(-1, 5), [k * i for i in range(bsz + 1)], c.nms_thresh,
                                       normal=(c.nms_type == "nms_normal_gpu"), keep_stride=c.nms_post_maxsize,
                                       set_counts=front["count"])
        return front, keep

    def select(self, rpn_cls_preds: torch.Tensor, rpn_box_preds: torch.Tensor, rpn_dir_cls_preds: Optional[torch.Tensor] = None):
        """Device-resident result for the whole batch, no synchronisation:
        dict(boxes (B,P,7), scores (B,P), labels (B,P) i64, selected (B,P) i64 anchor indices, num (B,) i32) with
        P = nms_post_maxsize; rows >= num[b] are padding."""
        c = self.cfg
        front, keep = self._front_and_nms(rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds)
        out = F.gather_kept(keep, front, c.nms_post_maxsize, sigmoid_scores=not c.use_raw_score)
        out["front"] = front
        return out

    def proposals(self, rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds=None) -> dict:
        """``proposal_layer`` (pcdet/models/model_utils/proposal_layer.py:7-68) fused with the anchor decode in front of
        it (``PartA2Net.forward_rcnn``, detectors/PartA2_net.py:64-83): the Part-A2 stage-1 -> stage-2 bridge.  Configure
        with ``score_thresh=0`` (every anchor is a candidate), ``nms_pre_maxsize = cfg.MODEL[mode].NMS_PRE_MAXSIZE``,
        ``nms_post_maxsize = cfg.MODEL[mode].NMS_POST_MAXSIZE``, ``nms_thresh = RPN_NMS_THRESH``, ``nms_type = RPN_NMS_TYPE``.
        Returns dict(rois (B,P,7) zero padded, roi_raw_scores (B,P) padded with -100000, roi_labels (B,P) i64 padded
        with 1), on the device, no synchronisation."""
        c = self.cfg
        front, keep = self._front_and_nms(rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds)
        out = F.gather_kept(keep, front, c.nms_post_maxsize, sigmoid_scores=False, pad_score=-100000.0, pad_label=1)
        return dict(rois=out["boxes"], roi_raw_scores=out["scores"], roi_labels=out["labels"], num=out["num"])

    def __call__(self, rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds=None) -> List[dict]:
        """The reference's record dicts (detector3d.py:215-219): one dict(boxes, scores, labels) per frame, trimmed
        to the kept boxes (this is where the single device->host synchronisation happens)."""
        r = self.select(rpn_cls_preds, rpn_box_preds, rpn_dir_cls_preds)
        nums = r["num"].tolist()
        return [dict(boxes=r["boxes"][b, :n], scores=r["scores"][b, :n], labels=r["labels"][b, :n], selected=r["selected"][b, :n])
                for b, n in enumerate(nums)]


def multi_classes_nms(rank_scores: torch.Tensor, normalized_scores: torch.Tensor, box_preds: torch.Tensor, score_thresh, nms_thresh,
                      nms_type: str = "nms_gpu"):
    """``Detector3D.multi_classes_nms`` (pcdet/models/detectors/detector3d.py:239-276, ``MODEL.TEST.MULTI_CLASSES_NMS: True``)
    for one frame: per class k, the boxes with normalized_scores[:, k] >= score_thresh[k] go through the rotated NMS in
    the order of rank_scores[:, k]; the survivors of all classes are concatenated in class order.

    The reference runs one boolean-mask gather, one sort, one mask kernel, one 2 MB download and one CPU sweep per class;
    here every class is one box SET of a single ``pcdb_nms_counts`` call (set sizes stay on the device), and the only
    synchronisation is the final read of the per-class keep counts.  rank_scores / normalized_scores (N, C), box_preds
    (N, 7) LiDAR boxes; score_thresh / nms_thresh: float or per-class list.  Returns (selected (M,) int64 indices into
    the N boxes, labels (M,) int64 in 1..C); both empty when nothing survives (the reference returns [] then)."""
    n, num_classes = rank_scores.shape
    dev = box_preds.device
    st = list(score_thresh) if isinstance(score_thresh, (list, tuple)) else [score_thresh] * num_classes
    nt = list(nms_thresh) if isinstance(nms_thresh, (list, tuple)) else [nms_thresh] * num_classes
    empty = (torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int64, device=dev))
    if n == 0:
        return empty
    cand = normalized_scores >= torch.tensor(st, dtype=normalized_scores.dtype, device=dev)[None, :]          # (N, C)
    counts = cand.sum(dim=0).to(torch.int32)                                                                  # (C,) on the device
    # candidates first, by descending rank score; equal scores keep the lower index first (the reference's sort leaves
    # ties unspecified)
    keys = torch.where(cand, rank_scores.float(), torch.full_like(rank_scores, float("-inf"), dtype=torch.float32))
    order = torch.argsort(keys, dim=0, descending=True, stable=True)                                          # (N, C)
    boxes_bev = F.boxes3d_to_bev(box_preds.contiguous().float())                                              # (N, 5)
    sets = boxes_bev[order.t().reshape(-1)].contiguous()                                                      # (C * N, 5), class-major
    normal = nms_type == "nms_normal_gpu"
    if len(set(float(t) for t in nt)) == 1:
        keep, num = F.nms_sorted_batched(sets, [n * k for k in range(num_classes + 1)], float(nt[0]), normal=normal, set_counts=counts)
    else:       # per-class thresholds: one call per class (pcdb_nms takes one threshold)
        parts = [F.nms_sorted_batched(sets[n * k:n * (k + 1)], [0, n], float(nt[k]), normal=normal, set_counts=counts[k:k + 1])
                 for k in range(num_classes)]
        keep, num = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    nums = num.tolist()                                                                                       # the one synchronisation
    sel, lab = [], []
    for k, m in enumerate(nums):
        if m > 0:
            sel.append(order[keep[k, :m], k])
            lab.append(torch.full((m,), k + 1, dtype=torch.int64, device=dev))
    if not sel:
        return empty
    return torch.cat(sel), torch.cat(lab)
