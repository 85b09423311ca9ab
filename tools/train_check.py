"""bf16 tensor-core training step of BackBone8x against the fp32 path: gradient errors and timings (GPU box)."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pcdet_b200 import functional as F, spconv, synthetic as S
from pcdet_b200.backbone import BackBone8x

dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "kitti"
cfg = S.NUSCENES if which == "nuscenes" else S.KITTI
frame = S.nuscenes_frame(0) if which == "nuscenes" else S.kitti_frame(0)
pts = torch.from_numpy(frame).to(dev)
offs = torch.tensor([0, frame.shape[0]], dtype=torch.int32, device=dev)
v = F.voxelize(pts, offs, 1, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
n = int(v["voxel_offsets"][-1])
feats = F.vfe_mean(v["voxels"][:n], v["num_points"][:n])
coords = v["coordinates"][:n].contiguous()
gs = F.grid_size(cfg["voxel_size"], cfg["point_cloud_range"])
shape = [int(gs[2]) + 1, int(gs[1]), int(gs[0])]
print("voxels", n, "shape", shape)


def make():
    net = BackBone8x(4)
    net.load_numpy_weights(S.backbone_weights(4, 0))
    return net.to(dev).train()


def step(net, x):
    net.zero_grad(set_to_none=True)
    out = net(spconv.SparseConvTensor(x, coords, shape, 1))["spatial_features"]
    if os.environ.get("LOSS") == "proj":
        g = torch.Generator(device="cuda"); g.manual_seed(5)
        loss = (out.float() - torch.randn(out.shape, device=out.device, generator=g).abs()).square().mean()
    else:
        loss = out.float().square().mean()
    loss.backward()
    return loss


class RoundBf16(torch.autograd.Function):
    """value and gradient rounded to bf16: where the mixed-precision path stores a tensor"""
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def emulate(net):
    """fp32 modules that round exactly where the tensor-core path stores bf16: conv outputs and BN+ReLU outputs"""
    for _stem, conv, _bn in net.conv_modules():
        def hook(_m, _inp, out):
            out.features = RoundBf16.apply(out.features)
            return out
        conv.register_forward_hook(hook)
    for m in net.modules():
        if isinstance(m, torch.nn.ReLU):
            m.register_forward_hook(lambda _m, _i, out: RoundBf16.apply(out))
    return net


def round_weights(net):
    with torch.no_grad():
        for _stem, conv, _bn in net.conv_modules():
            conv.weight.copy_(conv.weight.bfloat16().float())
    return net


n32, n16, nem = round_weights(make()), round_weights(make()), emulate(round_weights(make()))
l32 = step(n32, feats)
l16 = step(n16, feats.bfloat16())
lem = step(nem, feats.bfloat16().float())
torch.cuda.synchronize()
print("loss fp32", float(l32), "bf16", float(l16), "emulated", float(lem))
print(f"{'parameter':24s} tc-vs-emulation (max-norm, rms)   emulation-vs-fp32 (max-norm, rms)")
for (name, p32), (_, p16), (_, pem) in zip(n32.named_parameters(), n16.named_parameters(), nem.named_parameters()):
    a, b, c = p16.grad.float(), pem.grad.float(), p32.grad.float()
    print(f"{name:24s} {float((a - b).abs().max() / b.abs().max()):.3e} {float((a - b).norm() / b.norm()):.3e}      "
          f"{float((b - c).abs().max() / c.abs().max()):.3e} {float((b - c).norm() / c.norm()):.3e}")
for (name, b32), (_, b16) in zip(n32.named_buffers(), n16.named_buffers()):
    if "running" in name:
        a, b = b16.float(), b32.float()
        print(f"{name:24s} max-norm rel {float((a - b).abs().max() / b.abs().max()):.3e}")

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for label, net, x in (("fp32", n32, feats), ("bf16-tc", n16, feats.bfloat16())):
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
    ts, hs = [], []
    for i in range(8):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); step(net, x); opt.step(); e1.record()
        h = time.perf_counter() - t0
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1)); hs.append(h * 1e3)
    print(f"{label}: step {statistics.median(ts):.3f} ms (host issue {statistics.median(hs):.3f} ms)")

if len(sys.argv) > 2:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(n16, feats.bfloat16())
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
