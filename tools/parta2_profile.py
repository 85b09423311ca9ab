"""Kernel totals of one replay of the captured Part-A2 bridge (torch.profiler / CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from pcdet_b200 import synthetic as S
from pcdet_b200.parta2 import PartA2Config, PartA2HotPath
from pcdet_b200.unet import UNetV2
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = [S.kitti_frame(s) for s in range(B)]
pts = torch.from_numpy(np.concatenate(frames)).to(dev)
offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
A = 200 * 176 * 2
anchors = torch.rand((A, 7), device=dev, generator=g) * torch.tensor([70, 80, 0.5, 0.4, 1.0, 0.3, 1.57], device=dev) + torch.tensor([0, -40, -1.9, 1.5, 3.6, 1.4, 0], device=dev)
cls = torch.randn((B, A, 1), device=dev, generator=g) * 2 - 1
box = torch.randn((B, A, 7), device=dev, generator=g) * 0.2
dirp = torch.randn((B, A, 2), device=dev, generator=g)
torch.manual_seed(3)
hp = PartA2HotPath(PartA2Config(batch_size=B, max_points_total=int(pts.shape[0])), UNetV2(4), anchors)
hp.capture(pts, offs, cls, box, dirp)
hp.replay(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    hp.replay(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=80))
