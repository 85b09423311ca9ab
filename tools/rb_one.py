"""One strided rulebook build on KITTI-shaped level-1 coordinates (4 frames), a few plain launches (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S

cfg = S.KITTI
frames = [S.kitti_frame(s) for s in range(4)]
pts = torch.from_numpy(np.concatenate(frames)).cuda()
offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device="cuda")
v = F.voxelize(pts, offs, 4, cfg["voxel_size"], cfg["point_cloud_range"], 5, 40000, want_voxels=False)
nv = int(v["voxel_offsets"][-1])
coords = v["coordinates"][:nv].contiguous()
print(coords.shape)
for _ in range(4):
    r = F.rulebook_conv(coords, 4, [41, 1600, 1408], 3, 2, 1)
torch.cuda.synchronize()
print("n_out", r["n_out"].tolist())
