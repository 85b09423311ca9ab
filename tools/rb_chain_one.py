"""pcdb_rulebook_chain on the KITTI batch-of-4 voxel set, a few plain calls (for ncu / timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcdet_b200 import functional as F, synthetic as S
B = 4
pts = [S.kitti_frame(i) for i in range(B)]
cat = torch.from_numpy(np.concatenate(pts)).cuda()
offs = torch.tensor(np.concatenate([[0], np.cumsum([p.shape[0] for p in pts])]), dtype=torch.int32, device="cuda")
v = F.voxelize(cat, offs, B, S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000, want_voxels=False)
n = int(v["voxel_offsets"][-1])
coords = v["coordinates"][:n].contiguous()
convs = [dict(ksize=3, stride=2, padding=1), dict(ksize=3, stride=2, padding=1), dict(ksize=3, stride=2, padding=(0, 1, 1)),
         dict(ksize=(3, 1, 1), stride=(2, 1, 1), padding=0)]
caps = [96000, 153600, 96000, 48000, 48000]
big = torch.zeros((caps[0], 4), dtype=torch.int32, device="cuda"); big[:n] = coords
nd = torch.tensor([n], dtype=torch.int32, device="cuda")
for _ in range(4):
    r = F.rulebook_chain(big, nd, B, [41, 1600, 1408], convs, [3, 3, 3, 3, None], caps=caps)
torch.cuda.synchronize()
print(n, [c.tolist() for c in r["counts"][1:]])
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    r = F.rulebook_chain(big, nd, B, [41, 1600, 1408], convs, [3, 3, 3, 3, None], caps=caps)
e.record(); torch.cuda.synchronize()
print("per call (incl. memsets, python)", s.elapsed_time(e) / 20 * 1e3, "us")
