"""What a stage costs in THROUGHPUT mode (4 steps in flight): the step with one stage left out of the captured graph (its
outputs stay as the warm-up step left them), frames/s of each variant.

Read with care (round 2): leaving a stage out also removes its event edges, and several variants came out SLOWER than the full
step (voxelize, rulebooks, dense), while `nms` looked like 97 us of a 242 us step here but is 23 us when only its kernels
are skipped inside pcdb_nms in bench.py (234 -> 211 us/step; mask 15, resolve 5, sweep 3).  Kept as the starting point of
that measurement, not as a result."""
import os, sys, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcdet_b200 import synthetic as S
from pcdet_b200.backbone import BackBone8x
from pcdet_b200.functional import boxes3d_to_bev
from pcdet_b200.pipeline import HotPathConfig, SecondHotPath

dev = torch.device("cuda")
B, DEPTH, STEPS = 4, 4, 300
net = BackBone8x(4).eval(); net.load_numpy_weights(S.backbone_weights(4, 0))
v = S.KITTI
cfg = HotPathConfig(voxel_size=v["voxel_size"], point_cloud_range=v["point_cloud_range"], max_num_points=v["max_num_points"],
                    max_voxels=v["max_voxels"], batch_size=B, dtype=torch.bfloat16, max_points_total=B * 24000, conv_shallow_ring=True)
frames = [S.kitti_frame(b) for b in range(B)]
pts = torch.zeros((cfg.max_points_total, 4), device=dev); cat = np.concatenate(frames); pts[:cat.shape[0]] = torch.from_numpy(cat).to(dev)
offs = torch.tensor(np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]), dtype=torch.int32, device=dev)
b3, sc = S.nms_boxes(B * 4096, seed=0)
bev = boxes3d_to_bev(torch.from_numpy(b3).to(dev))
bev = torch.cat([bev[b * 4096:(b + 1) * 4096][torch.from_numpy(np.argsort(-sc[b * 4096:(b + 1) * 4096], kind="stable")).to(dev)] for b in range(B)])


def run(drop):
    insts = []
    for _ in range(DEPTH):
        hp = SecondHotPath(cfg, net, device=dev)
        p, o, x = pts.clone(), offs.clone(), bev.clone()
        hp.step(p, o, x); torch.cuda.synchronize()          # everything exists once
        noop = lambda *a, **k: None
        if drop == "nms": hp.nms = noop
        if drop == "voxelize": hp.voxelize = noop
        if drop == "rulebooks": hp._build_chain = noop; hp._clear_rulebook_buffers = noop if hasattr(hp, "_clear_rulebook_buffers") else None
        if drop == "convs": hp.lib = type("L", (), {"__getattr__": lambda s, n: (lambda *a: 0) if n == "pcdb_sparse_conv_fwd" else getattr(SecondHotPath.__init__.__globals__["lib"](), n)})()
        if drop == "dense":
            real = hp.lib
            hp.lib = type("L", (), {"__getattr__": lambda s, n: (lambda *a: 0) if n in ("pcdb_to_dense", "pcdb_dense_clear_rows") else getattr(real, n)})()
        g, _ = hp.capture(p, o, x)
        insts.append((hp, g, torch.cuda.Stream(device=dev)))
    torch.cuda.synchronize()
    for rep in range(2):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        t0.record(main)
        for _, _, s in insts: s.wait_event(t0)
        for i in range(STEPS):
            _, g, s = insts[i % DEPTH]
            with torch.cuda.stream(s): g.replay()
        for _, _, s in insts: main.wait_stream(s)
        t1.record(main); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / STEPS
    print(f"drop {drop:10s}: {ms * 1e3:7.1f} us/step  {B / ms * 1e3:8.0f} frames/s", flush=True)
    return ms

base = run("none")
for d in ("nms", "voxelize", "rulebooks", "convs", "dense"):
    try:
        ms = run(d)
        print(f"    -> {d} costs {1e3 * (base - ms):6.1f} us of the {base * 1e3:.1f} us step")
    except Exception as e:
        print("drop", d, "failed:", repr(e)[:200])
