"""Kernel time breakdown of one config-5 training step (torch.profiler / CUPTI).  Not a bench number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
from pcdet_b200 import functional as F, spconv, synthetic as S
from pcdet_b200.backbone import BackBone8x

dev = torch.device("cuda", 0)
cfg = S.NUSCENES
frame = S.nuscenes_frame(0)
pts = torch.from_numpy(frame).to(dev)
offs = torch.tensor([0, frame.shape[0]], dtype=torch.int32, device=dev)
v = F.voxelize(pts, offs, 1, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
n = int(v["voxel_offsets"][-1])
feats = F.vfe_mean(v["voxels"][:n], v["num_points"][:n]); coords = v["coordinates"][:n].contiguous()
gs = F.grid_size(cfg["voxel_size"], cfg["point_cloud_range"])
shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]
net = BackBone8x(4); net.load_numpy_weights(S.backbone_weights(4, 0)); net = net.to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    out = net(spconv.SparseConvTensor(feats, coords, shape, 1))["spatial_features"]
    out.square().mean().backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
