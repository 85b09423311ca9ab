"""The post-processing front alone (decode + threshold + top-k, SECOND head: 4 frames x 211 200 anchors x 3 classes):
plain launches (for ncu), then a CUDA-graph replay timed with events, cold L2.  PP_MEAN = mean of the class logits
(-1.5: ~40 % of the anchors are candidates, the random-init regime; -6: a few hundred, the trained regime)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pcdet_b200 import functional as F
from pcdet_b200.postprocess import PostProcessor, PostProcessConfig

B, A, C = int(os.environ.get("PP_B", 4)), int(os.environ.get("PP_A", 211200)), 3
mean = float(os.environ.get("PP_MEAN", -1.5))
g = torch.Generator(device="cuda").manual_seed(0)
cls = torch.randn((B, A, C), device="cuda", generator=g) * 1.5 + mean
box = torch.randn((B, A, 7), device="cuda", generator=g) * 0.3
dirp = torch.randn((B, A, 2), device="cuda", generator=g)
anchors = torch.rand((A, 7), device="cuda", generator=g) * torch.tensor([70, 80, 1, 1, 3, 1, 1.5], device="cuda") \
    + torch.tensor([0, -40, -2, 0.6, 0.8, 1.5, 0], device="cuda")
pp = PostProcessor(anchors, PostProcessConfig())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=30):
    for _ in range(3):
        fn()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


out = {}
def front():
    out["f"] = F.decode_select(cls, box, anchors, dirp, score_thresh=0.1, pre_max=4096, dir_offset=0.78539)
def full():
    out["r"] = pp.select(cls, box, dirp)

m, lo = timed(front)
print(f"decode_select  B={B} A={A} mean={mean}: median {m:.1f} us  min {lo:.1f} us   counts {out['f']['count'].tolist()}")
m, lo = timed(full)
print(f"front + nms + gather: median {m:.1f} us  min {lo:.1f} us   kept {out['r']['num'].tolist()}")

if os.environ.get("PP_PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        flush.zero_(); torch.cuda.synchronize()
        full(); torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    for e in ev:
        name = e.name.replace("pcdb::", "").replace("void ", "")
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.elapsed_us():8.1f}  {name[:90]}")
