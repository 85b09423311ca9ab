"""Numerics of the tensor-core training kernels against plain torch fp32 (run on the GPU box; prints one line per case)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcdet_b200 import functional as F

dev = "cuda"
torch.manual_seed(0)


def ref_wgrad(x, g, nbr, n_out):
    K = nbr.shape[0]
    out = torch.zeros(K, x.shape[1], g.shape[1], device=dev)
    for k in range(K):
        idx = nbr[k, :n_out].long()
        m = idx >= 0
        out[k] = x[idx[m]].float().t() @ g[:n_out][m].float()
    return out


def case(cin, cout, n_in, n_out, K, density, structured=False):
    nbr = torch.where(torch.rand(K, n_out, device=dev) < density, torch.randint(0, n_in, (K, n_out), device=dev, dtype=torch.int32),
                      torch.full((K, n_out), -1, dtype=torch.int32, device=dev)).contiguous()
    x = torch.randn(n_in, cin, device=dev).bfloat16()
    g = torch.randn(n_out, cout, device=dev).bfloat16()
    if structured:          # x[i, ci] = ci + 1 for one row, g one-hot: exposes layout permutations
        x.zero_(); g.zero_()
        x[:, :] = (torch.arange(cin, device=dev) + 1).bfloat16()
        g[:, :] = (torch.arange(cout, device=dev) + 1).bfloat16() / 16
        nbr.fill_(-1); nbr[:, 0] = 0
        nbr[1, 0] = -1; nbr[1, 17] = 1
    got = F.sparse_conv_wgrad(x, g, nbr, n_out)
    torch.cuda.synchronize()
    ref = ref_wgrad(x, g, nbr, n_out)
    err = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-9)
    print(f"wgrad {cin}->{cout} n_in={n_in} n_out={n_out} K={K} dens={density} struct={structured}: rel_err={err:.3e}", flush=True)
    if err > 1e-2 and structured:
        torch.set_printoptions(linewidth=250, precision=2, sci_mode=False)
        print("got[0][:8,:8]\n", got[0][:8, :8], "\nref[0][:8,:8]\n", ref[0][:8, :8])
        print("got[1][:4,:8]\n", got[1][:4, :8], "\nref[1]\n", ref[1][:4, :8])
    return err


worst = 0.0
for structured in (True, False):
    for cin, cout in ((64, 64), (32, 32), (16, 16), (16, 32), (32, 64), (64, 128), (64, 32), (64, 16), (16, 128)):
        worst = max(worst, case(cin, cout, 3000, 1000, 27, 0.4, structured))
worst = max(worst, case(64, 64, 50000, 45326, 27, 0.35))
worst = max(worst, case(64, 128, 20000, 9373, 3, 0.8))
worst = max(worst, case(32, 32, 100, 77, 27, 0.5))
worst = max(worst, case(64, 64, 100, 1, 1, 1.0))

# dgrad through the forward kernel with the transposed / flipped image
K, n_in, n_out, cin, cout = 27, 4000, 3000, 32, 64
nbr_t = torch.where(torch.rand(K, n_in, device=dev) < 0.3, torch.randint(0, n_out, (K, n_in), device=dev, dtype=torch.int32),
                    torch.full((K, n_in), -1, dtype=torch.int32, device=dev)).contiguous()
w = torch.randn(K, cin, cout, device=dev) * 0.1
g = torch.randn(n_out, cout, device=dev).bfloat16()
for flip in (False, True):
    wp = F.pack_conv_weights(w, transpose=True, flip=flip)
    wt = (w.flip(0) if flip else w).transpose(1, 2).contiguous().bfloat16()
    got = F.sparse_conv_fwd(g, wt, nbr_t, n_in, weight_packed=wp)
    ref = torch.zeros(n_in, cin, device=dev)
    for k in range(K):
        idx = nbr_t[k].long(); m = idx >= 0
        ref[m] += g[idx[m]].float() @ wt[k].float()
    err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"dgrad flip={flip}: rel_err={err:.3e}")
    worst = max(worst, err)

# BatchNorm train forward / backward against torch autograd
for dt in (torch.bfloat16, torch.float32):
    for c, n in ((16, 5000), (64, 45326), (128, 9373), (32, 7)):
        y = (torch.randn(n, c, device=dev) * 2 + 0.5).to(dt)
        gamma = torch.rand(c, device=dev) + 0.5
        beta = torch.randn(c, device=dev) * 0.2
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        out, stats = F.bn_train_fwd(y, gamma, beta, 1e-3, 0.01, rm, rv, relu=True)
        go = torch.randn(n, c, device=dev).to(dt)
        gy, gg, gb = F.bn_train_bwd(go, out, y, gamma, stats, relu=True)
        yr = y.float().clone().requires_grad_(True)
        gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        rm2, rv2 = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        o2 = torch.relu(torch.nn.functional.batch_norm(yr, rm2, rv2, gr, br, True, 0.01, 1e-3))
        mask = (out.float() > 0).float()        # the kernel's own mask (bf16 rounding may flip values at 0)
        (o2 * go.float() * (mask == (o2 > 0).float())).sum().backward()
        e = lambda a, b: (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-9)
        errs = dict(out=e(out, o2), gy=e(gy, yr.grad), gg=e(gg, gr.grad), gb=e(gb, br.grad), rm=e(rm, rm2), rv=e(rv, rv2))
        print(f"bn {dt} c={c} n={n}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
        worst = max(worst, *(v for v in errs.values()))
print("WORST", worst)
