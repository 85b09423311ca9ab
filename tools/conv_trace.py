import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pass
import torch, numpy as np
from pcdet_b200 import _lib
_lib.SO_PATH = _lib.SO_PATH.replace("libpcdet_b200.so", "libpcdet_b200_trace.so")
from pcdet_b200 import functional as F
cin, cout, tiles = (int(v) for v in sys.argv[1:4])
n = tiles * 128; K = 27; dev = "cuda"
torch.manual_seed(0)
nbr = torch.where(torch.rand(K, n, device=dev) < 0.5, torch.randint(0, n, (K, n), device=dev, dtype=torch.int32), torch.full((K, n), -1, dtype=torch.int32, device=dev)).contiguous()
f = torch.randn(n, cin, device=dev).bfloat16(); w = torch.randn(K, cin, cout, device=dev).bfloat16()
wp = F.pack_conv_weights(w); out = torch.empty(n, cout, device=dev, dtype=torch.bfloat16)
for _ in range(5):
    F.sparse_conv_fwd(f, w, nbr, n, out=out, weight_packed=wp, algo=3)
torch.cuda.synchronize()
buf = np.zeros((8, 32), np.int64)
L = _lib.lib()
print(L.pcdb_debug_trace(buf.ctypes.data_as(ctypes.c_void_p)))
t0 = buf[6, 0]
print("start->acc_ready", buf[6,1]-t0, "->end", buf[6,2]-t0)
print(" k   P0.start P0.done  P3.start P3.done  MMA.full MMA.commit")
for k in range(27):
    print(f"{k:2d} " + " ".join(f"{int(buf[s,k]-t0):8d}" for s in (0,1,4,5,2,3)))
