"""Timing model of the tcgen05 sparse-conv kernel (CUDA graph of REPS launches: no CPU launch overhead)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcdet_b200 import _lib
if os.environ.get('MB_SO'):
    _lib.SO_PATH = _lib.SO_PATH.replace('libpcdet_b200.so', os.environ['MB_SO'])
from pcdet_b200 import functional as F

dev = "cuda"
ALGO = int(sys.argv[1]) if len(sys.argv) > 1 else 3
REPS = 20


ONLY = int(os.environ.get("MB_ONLY", "-1"))      # >= 0: only that many leading offsets are populated


def bench(cin, cout, n, K, density):
    nbr = torch.where(torch.rand(K, n, device=dev) < density, torch.randint(0, n, (K, n), device=dev, dtype=torch.int32),
                      torch.full((K, n), -1, dtype=torch.int32, device=dev)).contiguous()
    if ONLY >= 0:
        nbr[ONLY:] = -1
    f = torch.randn(n, cin, device=dev).bfloat16()
    w = torch.randn(K, cin, cout, device=dev).bfloat16()
    wp = F.pack_conv_weights(w)
    out = torch.empty(n, cout, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        F.sparse_conv_fwd(f, w, nbr, n, out=out, weight_packed=wp, algo=ALGO)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            F.sparse_conv_fwd(f, w, nbr, n, out=out, weight_packed=wp, algo=ALGO)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (5 * REPS) * 1e3


SHAPES = ((64, 64), (32, 32)) if len(sys.argv) < 3 else tuple(tuple(int(v) for v in a.split("x")) for a in sys.argv[2].split(","))
TILES = (148, 296, 444, 592, 1184) if len(sys.argv) < 4 else tuple(int(v) for v in sys.argv[3].split(","))
KS = (27,) if len(sys.argv) < 5 else tuple(int(v) for v in sys.argv[4].split(","))
print("cin cout tiles density us")
for cin, cout in SHAPES:
    for tiles in TILES:
        for K in KS:
            t = bench(cin, cout, tiles * 128, K, float(os.environ.get("MB_DENSITY", "0.5")))
            print(f"{cin}x{cout} {tiles:5d} K={K:2d} {t:7.2f}", flush=True)
