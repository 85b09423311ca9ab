"""BackboneTrainStep (captured) against the module API in train mode: loss, statistics, gradient direction, timing."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcdet_b200 import functional as F, spconv, synthetic as S
from pcdet_b200.backbone import BackBone8x
from pcdet_b200.train import BackboneTrainStep

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
pg = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
which = sys.argv[1] if len(sys.argv) > 1 else "kitti"
cfg = S.NUSCENES if which == "nuscenes" else S.KITTI
frame = S.nuscenes_frame(rank) if which == "nuscenes" else S.kitti_frame(rank)
pts = torch.from_numpy(frame).to(dev)
offs = torch.tensor([0, frame.shape[0]], dtype=torch.int32, device=dev)
v = F.voxelize(pts, offs, 1, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
n = int(v["voxel_offsets"][-1])
feats = F.vfe_mean(v["voxels"][:n], v["num_points"][:n])
coords = v["coordinates"][:n].contiguous()
gs = F.grid_size(cfg["voxel_size"], cfg["point_cloud_range"])
shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]


def make():
    net = BackBone8x(4)
    net.load_numpy_weights(S.backbone_weights(4, 0))
    return net.to(dev).train()


ref = make()
out = ref(spconv.SparseConvTensor(feats.bfloat16(), coords, shape, 1))["spatial_features"]
loss_ref = out.float().square().mean()
loss_ref.backward()

net = make()
ts = BackboneTrainStep(net, 1, shape, cfg["max_voxels"], grad_norm_clip=None, process_group=None, lr=0.0)
ts.set_input(feats, coords)
ts.step()
torch.cuda.synchronize()
print("loss module-API", float(loss_ref), "train step", float(ts.loss), "level counts", [int(c[0]) for c in ts.level_counts], "overflow",
      [int(c[1]) for c in ts.level_counts[1:]])
worst = 1.0
for (name, a), (_, b) in zip(net.named_parameters(), ref.named_parameters()):
    cos = float(torch.dot(a.grad.flatten(), b.grad.flatten()) / (a.grad.norm() * b.grad.norm()))
    worst = min(worst, cos)
print("worst gradient cosine against the module API:", worst)
for (name, a), (_, b) in zip(net.named_buffers(), ref.named_buffers()):
    if "running_mean" in name and float((a - b).abs().max() / b.abs().max()) > 1e-2:
        print("running stat differs", name, float((a - b).abs().max() / b.abs().max()))

# determinism + capture
net2 = make()
ts2 = BackboneTrainStep(net2, 1, shape, cfg["max_voxels"], process_group=pg, lr=1e-4)
ts2.set_input(feats, coords)
ts2.capture()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
times = []
for i in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ts2.replay(); e1.record(); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
t = torch.tensor([statistics.median(times)], device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"captured train step ({which}, world {world}): {float(t):.3f} ms, loss {float(ts2.loss):.6f}, grad norm {float(ts2.grad_norm):.4e}, voxels {n}")
if world > 1:
    dist.destroy_process_group()
