"""The bench's NMS stage alone (4 box sets of 4096 score-sorted boxes), a few plain launches (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S

B, N = 4, 4096
b3, scores = S.nms_boxes(B * N, seed=0)
bev_all = F.boxes3d_to_bev(torch.from_numpy(b3).cuda()).cpu().numpy()
bev = np.concatenate([bev_all[b * N:(b + 1) * N][np.argsort(-scores[b * N:(b + 1) * N], kind="stable")] for b in range(B)])
boxes = torch.from_numpy(bev).cuda()
offs = np.arange(B + 1, dtype=np.int32) * N
for _ in range(4):
    keep, num = F.nms_sorted_batched(boxes, offs, 0.01, keep_stride=500)
torch.cuda.synchronize()
print("done", num.tolist())
