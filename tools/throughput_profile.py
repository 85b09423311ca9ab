"""Kernel durations with 4 steps in flight (CUPTI via torch.profiler): which kernels stretch when the steps share the GPU."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from pcdet_b200 import synthetic as S
from pcdet_b200.backbone import BackBone8x
from pcdet_b200.functional import boxes3d_to_bev
from pcdet_b200.pipeline import HotPathConfig, SecondHotPath
dev = torch.device("cuda")
B, DEPTH = 4, int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = BackBone8x(4).eval(); net.load_numpy_weights(S.backbone_weights(4, 0))
v = S.KITTI
cfg = HotPathConfig(voxel_size=v["voxel_size"], point_cloud_range=v["point_cloud_range"], max_num_points=v["max_num_points"],
                    max_voxels=v["max_voxels"], batch_size=B, dtype=torch.bfloat16, max_points_total=B * 24000, conv_shallow_ring=DEPTH > 1)
frames = [S.kitti_frame(b) for b in range(B)]
pts = torch.zeros((cfg.max_points_total, 4), device=dev); cat = np.concatenate(frames); pts[:cat.shape[0]] = torch.from_numpy(cat).to(dev)
offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device=dev)
b3, sc = S.nms_boxes(B * 4096, seed=0)
bev = boxes3d_to_bev(torch.from_numpy(b3).to(dev))
bev = torch.cat([bev[b * 4096:(b + 1) * 4096][torch.from_numpy(np.argsort(-sc[b * 4096:(b + 1) * 4096], kind="stable")).to(dev)] for b in range(B)])
insts = []
for _ in range(DEPTH):
    hp = SecondHotPath(cfg, net, device=dev)
    p, o, x = pts.clone(), offs.clone(), bev.clone()
    g, _ = hp.capture(p, o, x)
    insts.append((g, torch.cuda.Stream(device=dev), hp))       # keep hp alive: the graph only holds addresses
def run(n):
    for i in range(n):
        g, s, _ = insts[i % DEPTH]
        with torch.cuda.stream(s): g.replay()
    torch.cuda.synchronize()
run(40)
N = 40
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(N)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[e.name[:60]]; a[0] += 1; a[1] += e.time_range.end - e.time_range.start
print(f"{DEPTH} in flight: {N} steps in {(t1 - t0):.0f} us = {(t1 - t0) / N:.1f} us/step; sum of kernel durations per step {sum(a[1] for a in agg.values()) / N:.1f} us")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
    print(f"{k:62s} {a[0] / N:5.1f}/step  avg {a[1] / a[0]:7.1f} us   per step {a[1] / N:7.1f} us")
