"""One tcgen05 conv configuration, a few plain launches (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcdet_b200 import functional as F
cin, cout, tiles, K = (int(v) for v in sys.argv[1:5])
n = tiles * 128
dev = "cuda"
nbr = torch.where(torch.rand(K, n, device=dev) < 0.5, torch.randint(0, n, (K, n), device=dev, dtype=torch.int32),
                  torch.full((K, n), -1, dtype=torch.int32, device=dev)).contiguous()
f = torch.randn(n, cin, device=dev).bfloat16()
w = torch.randn(K, cin, cout, device=dev).bfloat16()
wp = F.pack_conv_weights(w)
out = torch.empty(n, cout, device=dev, dtype=torch.bfloat16)
for _ in range(6):
    F.sparse_conv_fwd(f, w, nbr, n, out=out, weight_packed=wp, algo=3)
torch.cuda.synchronize()
print("done")
