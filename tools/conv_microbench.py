"""Timing model of the tcgen05 sparse-conv kernel: fixed cost vs per-offset cost vs tiles (CUDA events, warm L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcdet_b200 import functional as F

dev = "cuda"
ALGO = int(sys.argv[1]) if len(sys.argv) > 1 else 2
def bench(cin, cout, n, K, density, reps=30):
    n_in = n
    nbr = torch.where(torch.rand(K, n, device=dev) < density, torch.randint(0, n_in, (K, n), device=dev, dtype=torch.int32),
                      torch.full((K, n), -1, dtype=torch.int32, device=dev)).contiguous()
    f = torch.randn(n_in, cin, device=dev).bfloat16()
    w = torch.randn(K, cin, cout, device=dev).bfloat16()
    wt = F.pack_conv_weights(w)
    out = torch.empty(n, cout, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        F.sparse_conv_fwd(f, w, nbr, n, out=out, weight_packed=wt, algo=ALGO)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        F.sparse_conv_fwd(f, w, nbr, n, out=out, weight_packed=wt, algo=ALGO)
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3

print("cin cout tiles K density us")
SHAPES = ((64, 64), (32, 32)) if len(sys.argv) < 3 else tuple(tuple(int(v) for v in a.split("x")) for a in sys.argv[2].split(","))
print("tune", os.environ.get("PCDB_TC_TUNE"))
for cin, cout in SHAPES:
    for tiles in (296, 1184):
        for K in (27,):
            for dens in (0.5, 1.0):
                t = bench(cin, cout, tiles * 128, K, dens)
                print(f"{cin:3d} {cout:4d} {tiles:5d} {K:3d} {dens:4.1f} {t:8.2f}", flush=True)
