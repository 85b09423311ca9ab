"""One tcgen05 weight-gradient configuration, a few plain launches (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcdet_b200 import functional as F
cin, cout, tiles, K = (int(v) for v in sys.argv[1:5])
n = tiles * 128
dev = "cuda"
nbr = torch.where(torch.rand(K, n, device=dev) < 0.4, torch.randint(0, n, (K, n), device=dev, dtype=torch.int32),
                  torch.full((K, n), -1, dtype=torch.int32, device=dev)).contiguous()
x = torch.randn(n, cin, device=dev).bfloat16()
g = torch.randn(n, cout, device=dev).bfloat16()
for _ in range(4):
    F.sparse_conv_wgrad(x, g, nbr, n)
torch.cuda.synchronize()
print("done")
