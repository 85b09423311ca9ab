"""Times the NMS stage with a library variant (PCDB_SO=path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcdet_b200 import _lib
if os.environ.get("PCDB_SO"):
    _lib.SO_PATH = os.environ["PCDB_SO"]
import numpy as np, torch
from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
B, N = 4, 4096
b3, scores = S.nms_boxes(B * N, seed=0)
bev_all = F.boxes3d_to_bev(torch.from_numpy(b3).cuda()).cpu().numpy()
bev = np.concatenate([bev_all[b * N:(b + 1) * N][np.argsort(-scores[b * N:(b + 1) * N], kind="stable")] for b in range(B)])
boxes = torch.from_numpy(bev).cuda(); offs = np.arange(B + 1, dtype=np.int32) * N
for _ in range(3): keep, num = F.nms_sorted_batched(boxes, offs, 0.01, keep_stride=500)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    F.nms_sorted_batched(boxes, offs, 0.01, keep_stride=500)
g.replay(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): g.replay()
e.record(); torch.cuda.synchronize()
print(os.environ.get("PCDB_SO", "default"), round(s.elapsed_time(e) / 20 * 1e3, 1), "us", num.tolist())
