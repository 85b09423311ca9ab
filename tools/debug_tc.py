"""Diagnostics for the tcgen05 sparse-conv kernel: SIMT (algo=1) vs tensor core (algo=2) on small cases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcdet_b200 import functional as F

torch.manual_seed(0)
ALGO = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = "cuda"
def run(cin, cout, n, K, density, label):
    nbr = torch.full((K, n), -1, dtype=torch.int32, device=dev)
    for k in range(K):
        m = torch.rand(n, device=dev) < density
        nbr[k] = torch.where(m, torch.randint(0, n, (n,), device=dev, dtype=torch.int32), torch.full((n,), -1, dtype=torch.int32, device=dev))
    if K > 13: nbr[13] = torch.arange(n, device=dev, dtype=torch.int32)
    f = torch.randn(n, cin, device=dev).bfloat16()
    w = (torch.randn(K, cin, cout, device=dev) / (cin * K) ** 0.5).bfloat16()
    a = F.sparse_conv_fwd(f, w, nbr, n, algo=1).float()
    b = F.sparse_conv_fwd(f, w, nbr, n, algo=ALGO).float()
    torch.cuda.synchronize()
    err = (a - b).abs().max().item() / max(a.abs().max().item(), 1e-9)
    print(f"algo {ALGO} {label} cin={cin} cout={cout} n={n} K={K}: rel err {err:.3e}", flush=True)
    if err > 1e-2:
        print("  simt[0,:8]", a[0, :8].tolist()); print("  tc  [0,:8]", b[0, :8].tolist())
        print("  simt[1,:8]", a[1, :8].tolist()); print("  tc  [1,:8]", b[1, :8].tolist())
        bad_rows = ((a - b).abs().max(dim=1).values > 1e-2 * a.abs().max()).nonzero().flatten()
        bad_cols = ((a - b).abs().max(dim=0).values > 1e-2 * a.abs().max()).nonzero().flatten()
        print("  bad rows", bad_rows[:16].tolist(), "count", bad_rows.numel(), " bad cols", bad_cols[:32].tolist(), "count", bad_cols.numel())
    return err

worst = 0
for cin in (16, 32, 64):
    for cout in (16, 32, 64, 128):
        worst = max(worst, run(cin, cout, 128, 1, 1.0, "1-offset"))
for cin, cout in ((16, 16), (32, 32), (64, 64), (64, 128), (16, 32), (32, 64)):
    worst = max(worst, run(cin, cout, 1000, 27, 0.4, "27-offset"))
    worst = max(worst, run(cin, cout, 50000, 27, 0.3, "large"))
print("WORST", worst)
