"""Kernel totals of one replay of the captured training step (torch.profiler / CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from pcdet_b200 import functional as F, synthetic as S
from pcdet_b200.backbone import BackBone8x
from pcdet_b200.train import BackboneTrainStep
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = S.NUSCENES
frames = [S.nuscenes_frame(b) for b in range(B)]
pts = torch.from_numpy(np.concatenate(frames)).to(dev)
offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device=dev)
v = F.voxelize(pts, offs, B, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
n = int(v["voxel_offsets"][-1])
feats = F.vfe_mean(v["voxels"][:n], v["num_points"][:n])
gs = F.grid_size(cfg["voxel_size"], cfg["point_cloud_range"])
net = BackBone8x(4); net.load_numpy_weights(S.backbone_weights(4, 0))
ts = BackboneTrainStep(net.to(dev).train(), B, [int(gs[2]) + 1, int(gs[1]), int(gs[0])], B * cfg["max_voxels"])
ts.set_input(feats, v["coordinates"][:n].contiguous())
ts.capture()
ts.replay(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ts.replay(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
