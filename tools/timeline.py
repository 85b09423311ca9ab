"""Kernel timeline of ONE replay of the captured hot-path graph (CUPTI activity records via torch.profiler).

    gpurun -- 'python tools/timeline.py > gpurun_out/timeline.txt'

Prints every kernel of the step with its start offset, duration and stream, so that gaps and overlaps between
the three graph branches (strided rulebooks / SubM rulebooks / convolutions) are visible.  Not a bench number.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from pcdet_b200 import synthetic as S
from pcdet_b200.backbone import BackBone8x
from pcdet_b200.functional import boxes3d_to_bev
from pcdet_b200.pipeline import HotPathConfig, SecondHotPath

dev = torch.device("cuda", 0)
wl = bench.workload_cfg(sys.argv[1] if len(sys.argv) > 1 else "kitti")
B = bench.FRAMES_PER_GPU
net = BackBone8x(4).eval()
net.load_numpy_weights(S.backbone_weights(4, 0))
v = wl["vox"]
cfg = HotPathConfig(voxel_size=v["voxel_size"], point_cloud_range=v["point_cloud_range"], max_num_points=v["max_num_points"],
                    max_voxels=v["max_voxels"], batch_size=B, dtype=torch.bfloat16, max_points_total=B * wl["max_points"])
hp = SecondHotPath(cfg, net, device=dev)
frames = [wl["gen"](b) for b in range(B)]
b3, scores = S.nms_boxes(B * 4096, seed=0)
bev_all = boxes3d_to_bev(torch.from_numpy(b3).to(dev)).cpu().numpy()
bev = np.concatenate([bev_all[b * 4096:(b + 1) * 4096][np.argsort(-scores[b * 4096:(b + 1) * 4096], kind="stable")] for b in range(B)])
pts = torch.zeros((cfg.max_points_total, 4), dtype=torch.float32, device=dev)
cat = np.concatenate(frames)
pts[:cat.shape[0]] = torch.from_numpy(cat).to(dev)
offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device=dev)
g = hp.capture(pts, offs, torch.from_numpy(bev).to(dev))[0]
flush = torch.empty((bench.L2_FLUSH_BYTES,), dtype=torch.uint8, device=dev)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        flush.fill_(i)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# last replay: everything after the last FillFunctor flush kernel
last_flush = max(i for i, e in enumerate(ev) if "FillFunctor<unsigned char>" in e.name and e.time_range.elapsed_us() > 20)
step = ev[last_flush + 1:]
t0 = step[0].time_range.start
print(f"{'start_us':>9} {'dur_us':>8} {'end_us':>8}  stream  kernel")
for e in step:
    name = e.name.replace("pcdb::", "").replace("void ", "")
    name = name[:name.index("(")] if "(" in name else name
    st = e.time_range.start - t0
    print(f"{st:9.1f} {e.time_range.elapsed_us():8.1f} {st + e.time_range.elapsed_us():8.1f}  {getattr(e, 'device_index', 0)}:{getattr(e, 'stream', '?')}  {name[:70]}")
print("step span us:", step[-1].time_range.end - t0)
