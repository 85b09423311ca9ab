"""GPU parity of the mixed-precision training path: the tcgen05 weight gradient, the input gradient through the
forward kernel, train-mode BatchNorm + ReLU forward / backward, and their composition in spconv.functional.SparseConvBnReluTC,
each against plain torch fp32 (autograd) on the same inputs.

Tolerances: the kernels round activations and activation gradients to bf16 where they are stored (2^-9 per element) and
accumulate in fp32; against fp32 torch on the SAME bf16 inputs a kernel stays within 1e-2 of the largest reference value
(measured 3e-3).  Through the 12 BatchNorm layers of the whole backbone the bf16 storage noise is amplified by the
cancellations of the BatchNorm backward -- an fp32 model that only ROUNDS where the tensor-core path stores bf16 (the
`emulate` hooks below) moves just as far from pure fp32 -- so the whole-network test bounds the distance to that emulation
by the emulation's own distance to fp32 and checks the direction of every gradient."""
import numpy as np
import pytest
import torch

import pcdet_b200.spconv as spconv
from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from pcdet_b200.backbone import BackBone8x
from pcdet_b200.spconv.functional import indice_conv_bn_relu_tc

pytestmark = pytest.mark.gpu
DEV = "cuda"


class RoundBf16(torch.autograd.Function):
    """value and gradient rounded to bf16: what storing a tensor in bf16 does"""
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def rel(a, b):
    a, b = a.float(), b.float()
    return float(((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).detach())


def random_map(K, n_out, n_in, density, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    idx = torch.randint(0, n_in, (K, n_out), device=DEV, dtype=torch.int32, generator=g)
    keep = torch.rand((K, n_out), device=DEV, generator=g) < density
    return torch.where(keep, idx, torch.full_like(idx, -1)).contiguous()


def ref_conv(x, w, nbr, n_out):
    out = torch.zeros(n_out, w.shape[2], device=DEV)
    for k in range(nbr.shape[0]):
        idx = nbr[k, :n_out].long()
        m = idx >= 0
        out[m] += x[idx[m]].float() @ w[k].float()
    return out


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (64, 16), (16, 128)])
@pytest.mark.parametrize("n_out,K,density", [(1000, 27, 0.4), (129, 27, 0.05), (77, 3, 1.0)])
def test_wgrad_matches_torch(cin, cout, n_out, K, density):
    torch.manual_seed(cin * 131 + cout)
    n_in = 3000
    nbr = random_map(K, n_out, n_in, density, seed=K + n_out)
    x = torch.randn(n_in, cin, device=DEV).bfloat16()
    g = torch.randn(n_out, cout, device=DEV).bfloat16()
    got = F.sparse_conv_wgrad(x, g, nbr, n_out)
    ref = torch.zeros_like(got)
    for k in range(K):
        idx = nbr[k].long()
        m = idx >= 0
        ref[k] = x[idx[m]].float().t() @ g[m].float()
    assert rel(got, ref) < 1e-5          # exact bf16 products, fp32 accumulation: only the summation order differs
    # deterministic: per-CTA partials are added in index order
    assert torch.equal(got, F.sparse_conv_wgrad(x, g, nbr, n_out))


def test_wgrad_device_count_accumulate_and_empty():
    torch.manual_seed(3)
    n_in, cap, n, K = 2000, 4096, 3001, 27
    nbr = random_map(K, cap, n_in, 0.3, seed=9)
    x = torch.randn(n_in, 64, device=DEV).bfloat16()
    g = torch.randn(cap, 64, device=DEV).bfloat16()
    g[n:] = float("nan")                 # rows past the device-side count must never be read into the sum
    n_dev = torch.tensor([n], dtype=torch.int32, device=DEV)
    got = F.sparse_conv_wgrad(x, g, nbr, cap, n_out_dev=n_dev)
    ref = F.sparse_conv_wgrad(x, g[:n].contiguous(), nbr[:, :n].contiguous(), n)
    assert torch.equal(got, ref)
    acc = torch.ones_like(got)
    F.sparse_conv_wgrad(x, g, nbr, cap, n_out_dev=n_dev, out=acc, accumulate=True)
    assert rel(acc, ref + 1) < 1e-6
    zero = F.sparse_conv_wgrad(x, g, nbr, cap, n_out_dev=torch.zeros(1, dtype=torch.int32, device=DEV))
    assert float(zero.abs().max()) == 0.0


@pytest.mark.parametrize("flip", [False, True])
def test_input_gradient_image(flip):
    """pack_conv_weights(transpose, flip) + the forward kernel = grad_x of the layer (W[k]^T, offsets reversed for SubM)."""
    torch.manual_seed(5)
    K, n_in, n_out, cin, cout = 27, 4000, 3000, 32, 64
    nbr_t = random_map(K, n_in, n_out, 0.3, seed=11)
    w = torch.randn(K, cin, cout, device=DEV) * 0.1
    g = torch.randn(n_out, cout, device=DEV).bfloat16()
    wp = F.pack_conv_weights(w, transpose=True, flip=flip)
    got = F.sparse_conv_fwd(g, None, nbr_t, n_in, weight_packed=wp, weight_shape=(K, cout, cin))
    wt = (w.flip(0) if flip else w).transpose(1, 2).contiguous().bfloat16()
    assert rel(got, ref_conv(g, wt, nbr_t, n_in)) < 1e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("c,n", [(16, 5000), (64, 45326), (128, 9373), (32, 7), (24, 300)])
def test_bn_train_matches_torch(dtype, c, n):
    torch.manual_seed(c + n)
    y = (torch.randn(n, c, device=DEV) * 2 + 0.5).to(dtype)
    gamma, beta = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.2
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    out, stats = F.bn_train_fwd(y, gamma, beta, 1e-3, 0.01, rm, rv, relu=True)
    go = torch.randn(n, c, device=DEV).to(dtype)
    gy, gg, gb = F.bn_train_bwd(go, out, y, gamma, stats, relu=True)
    yr = y.float().clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    o2 = torch.relu(torch.nn.functional.batch_norm(yr, rm2, rv2, gr, br, True, 0.01, 1e-3))
    o2.backward(go.float())
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-5
    assert rel(out, o2) < tol and rel(gy, yr.grad) < tol
    assert rel(gg, gr.grad) < 1e-4 and rel(gb, br.grad) < 1e-4
    assert rel(rm, rm2) < 1e-5 and rel(rv, rv2) < 1e-5            # momentum update, unbiased running variance


BLOCKS = [  # (c_in of the module, c_out, kind, kernel offsets)
    (4, 16, "subm", 27), (16, 32, "strided", 27), (32, 32, "subm", 27), (64, 64, "strided", 27), (64, 64, "subm", 27),
    (64, 128, "strided", 3), (64, 128, "strided", 27)]


@pytest.mark.parametrize("cin,cout,kind,K", BLOCKS)
def test_conv_bn_relu_block_gradients(cin, cout, kind, K):
    """One post_act_block (rpn_backbone.py:79-103) in train mode: SparseConvBnReluTC against torch autograd in fp32 on the same
    bf16 inputs and the same rulebook."""
    torch.manual_seed(cin + cout)
    n_in = 2500
    n_out = n_in if kind == "subm" else 1800
    if kind == "subm":          # a centred submanifold map is its own transpose with the offsets reversed
        nbr = torch.full((K, n_in), -1, dtype=torch.int32, device=DEV)
        perm = torch.randperm(n_in, device=DEV)
        for k in range(13):
            src = perm.roll(k + 1)[: n_in // 3]
            dst = perm[: n_in // 3]
            nbr[k, dst] = src.int()
            nbr[K - 1 - k, src] = dst.int()
        nbr[13] = torch.arange(n_in, dtype=torch.int32, device=DEV)
        nbr_t, flip = None, True
    else:                       # strided: every (offset, input) reaches at most one output
        nbr = torch.full((K, n_out), -1, dtype=torch.int32, device=DEV)
        nbr_t = torch.full((K, n_in), -1, dtype=torch.int32, device=DEV)
        for k in range(K):
            ins = torch.randperm(n_in, device=DEV)[:600]
            outs = torch.randperm(n_out, device=DEV)[:600]
            nbr[k, outs] = ins.int()
            nbr_t[k, ins] = outs.int()
        flip = False
    x = torch.randn(n_in, cin, device=DEV).bfloat16().requires_grad_(True)
    w = (torch.randn(K, cin, cout, device=DEV) * 0.2).bfloat16().float().requires_grad_(True)
    gamma = (torch.rand(cout, device=DEV) + 0.5).requires_grad_(True)
    beta = (torch.randn(cout, device=DEV) * 0.2).requires_grad_(True)
    rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    out = indice_conv_bn_relu_tc(x, w, gamma, beta, nbr, n_out, nbr_t, flip, (rm, rv, 1e-3, 0.01), True)
    go = torch.randn(n_out, cout, device=DEV).bfloat16()
    out.backward(go)

    xr = x.detach().float().requires_grad_(True)
    wr, gr, br = (t.detach().clone().requires_grad_(True) for t in (w, gamma, beta))
    y = torch.zeros(n_out, cout, device=DEV)
    for k in range(K):
        idx = nbr[k].long()
        m = idx >= 0
        y = y.index_add(0, m.nonzero()[:, 0], xr[idx[m]] @ wr[k])
    y = RoundBf16.apply(y)      # the conv output is stored in bf16 (it decides the ReLU mask of values next to zero)
    o2 = torch.relu(torch.nn.functional.batch_norm(y, torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV), gr, br, True, 0.01, 1e-3))
    o2.backward(go.float())
    assert rel(out, o2) < 1e-2
    # gradients: bf16 storage of grad_y puts 2^-9 of noise on every term of the sums; norm-wise that is ~3e-3, while a
    # single element of a small-fan-in weight gradient (c_in = 4) can sit 3e-2 of the largest entry away
    nrm = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    assert nrm(w.grad, wr.grad) < 1.5e-2 and rel(w.grad, wr.grad) < 6e-2
    assert rel(gamma.grad, gr.grad) < 1e-2 and rel(beta.grad, br.grad) < 1e-2
    assert nrm(x.grad, xr.grad) < 1.5e-2 and rel(x.grad, xr.grad) < 6e-2
    assert x.grad.dtype == torch.bfloat16 and w.grad.dtype == torch.float32


def test_backbone_training_step_against_bf16_emulation():
    cfg = S.KITTI
    frame = S.kitti_frame(0)
    pts = torch.from_numpy(frame).to(DEV)
    offs = torch.tensor([0, frame.shape[0]], dtype=torch.int32, device=DEV)
    v = F.voxelize(pts, offs, 1, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_num_points"], cfg["max_voxels"])
    n = int(v["voxel_offsets"][-1])
    feats = F.vfe_mean(v["voxels"][:n], v["num_points"][:n])
    coords = v["coordinates"][:n].contiguous()
    shape = [41, 1600, 1408]
    target = torch.randn((1, 256, 200, 176), device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)).abs()

    def make(emulate=False):
        net = BackBone8x(4)
        net.load_numpy_weights(S.backbone_weights(4, 0))
        with torch.no_grad():
            for _s, conv, _bn in net.conv_modules():
                conv.weight.copy_(conv.weight.bfloat16().float())
        if emulate:     # fp32 modules that round exactly where the tensor-core path stores bf16
            def hook(_m, _i, out):
                out.features = RoundBf16.apply(out.features)
                return out
            for _s, conv, _bn in net.conv_modules():
                conv.register_forward_hook(hook)
            for m in net.modules():
                if isinstance(m, torch.nn.ReLU):
                    m.register_forward_hook(lambda _m, _i, out: RoundBf16.apply(out))
        return net.to(DEV).train()

    def step(net, x):
        out = net(spconv.SparseConvTensor(x, coords, shape, 1))["spatial_features"]
        loss = (out.float() - target).square().mean()
        loss.backward()
        return float(loss.detach())

    n32, ntc, nem = make(), make(), make(emulate=True)
    l32, ltc, lem = step(n32, feats), step(ntc, feats.bfloat16()), step(nem, feats.bfloat16().float())
    assert abs(ltc - l32) / l32 < 1e-3 and abs(ltc - lem) / lem < 1e-3
    worst_cos = 1.0
    for (name, p32), (_, ptc), (_, pem) in zip(n32.named_parameters(), ntc.named_parameters(), nem.named_parameters()):
        a, b, c = ptc.grad.float().flatten(), pem.grad.float().flatten(), p32.grad.float().flatten()
        d_tc, d_em = float((a - b).norm() / b.norm()), float((b - c).norm() / c.norm())
        cos = float(torch.dot(a, c) / (a.norm() * c.norm()))
        worst_cos = min(worst_cos, cos)
        # as close to the rounding emulation as the emulation is to fp32 (the noise floor of bf16 storage), with slack
        assert d_tc < 2.0 * d_em + 2e-2, (name, d_tc, d_em)
        assert cos > 0.9, (name, cos)
    for (name, b32), (_, btc) in zip(n32.named_buffers(), ntc.named_buffers()):
        if "running" in name:
            assert rel(btc, b32) < 1e-2, name       # batch statistics (forward) agree closely
    print("worst gradient cosine against fp32:", worst_cos)


@pytest.mark.parametrize("kind", ["subm", "strided"])
def test_backward_matches_the_oracle_restatement(orc, kind):
    """Both backward paths against the oracle's restatement of spconv's indiceConvBackward (pinned on the
    CPU to autograd through the dense conv3d) on a real rulebook: the fp32 path (pcdb_sparse_conv_bwd behind
    SparseConvFunction) within 1e-4, the tensor-core path (bf16 operands, fp32 accumulation) exact in the weight gradient
    and within bf16 output rounding in the input gradient."""
    from pcdet_b200.spconv import ops as sops
    from pcdet_b200.spconv.functional import indice_conv
    rng = np.random.default_rng(3)
    g = orc.VoxelGenerator(S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    _v, c, _n = g.generate(S.kitti_frame(1)[::3])
    coords = np.concatenate([np.zeros((c.shape[0], 1), np.int32), c], 1).astype(np.int32)
    shape, cin, cout = [41, 1600, 1408], 32, 64
    ks, st, pd = (3, 3, 3), ((1, 1, 1) if kind == "subm" else (2, 2, 2)), (1, 1, 1)
    out_ids, pairs, num, _ = orc.get_indice_pairs(coords, 1, shape, ks, st, pd, 1, subm=kind == "subm")
    n_in, n_out = coords.shape[0], out_ids.shape[0]
    x = rng.normal(0, 1, (n_in, cin)).astype(np.float32)
    w = rng.normal(0, 0.1, (*ks, cin, cout)).astype(np.float32)
    go = rng.normal(0, 1, (n_out, cout)).astype(np.float32)
    rb = sops.build_rulebook(torch.from_numpy(coords).to(DEV), 1, shape, list(ks), list(st), list(pd), [1, 1, 1], kind == "subm")
    assert rb.n_out == n_out and np.array_equal(rb.outids.cpu().numpy(), out_ids)       # same rows as the reference order
    # ---- fp32 path ----------------------------------------------------------------------------------------------------------
    ib, fb = orc.indice_conv_backward(x, w, go, pairs, num, subm=kind == "subm")
    xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
    wt = torch.from_numpy(w).to(DEV).view(27, cin, cout).requires_grad_(True)
    y = indice_conv(xt, wt, rb.nbr, n_out, None if kind == "subm" else rb.nbr_inv, kind == "subm")
    y.backward(torch.from_numpy(go).to(DEV))
    assert rel(xt.grad, torch.from_numpy(ib).to(DEV)) < 1e-4
    assert rel(wt.grad.view(-1), torch.from_numpy(fb).to(DEV).view(-1)) < 1e-4
    # ---- tensor-core path on bf16-rounded operands ------------------------------------------------------------------------------
    xb, wb, gb = (torch.from_numpy(a).to(DEV).bfloat16() for a in (x, w, go))
    ib16, fb16 = orc.indice_conv_backward(xb.float().cpu().numpy(), wb.float().cpu().numpy(), gb.float().cpu().numpy(), pairs, num,
                                          subm=kind == "subm")
    gw = F.sparse_conv_wgrad(xb, gb, rb.nbr, n_out)
    assert rel(gw.view(-1), torch.from_numpy(fb16).to(DEV).view(-1)) < 1e-5
    wimg = F.pack_conv_weights(wb.view(27, cin, cout).float().contiguous(), transpose=True, flip=kind == "subm")
    gx = F.sparse_conv_fwd(gb, None, rb.nbr if kind == "subm" else rb.nbr_inv, n_in, weight_packed=wimg, weight_shape=(27, cout, cin))
    assert rel(gx, torch.from_numpy(ib16).to(DEV)) < 1e-2


def test_sync_batchnorm_path_with_one_rank():
    """The SyncBatchNorm halves (sums -> all-reduce -> finalize / apply, forward and backward) through a one-rank NCCL group:
    the same numbers as the single-call path.  Across ranks (ragged row counts, against ONE BatchNorm over all rows) it is
    checked by tools/syncbn_check.py under torchrun (2 GPUs: fp32 1e-7, bf16 one rounding flip)."""
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        import socket
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        torch.manual_seed(2)
        n, c = 7001, 64
        y = (torch.randn(n, c, device=DEV) * 2 + 0.5).bfloat16()
        go = torch.randn(n, c, device=DEV).bfloat16()
        gamma, beta = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.2
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        rm2, rv2 = rm.clone(), rv.clone()
        out, stats, sums = F.bn_train_fwd(y, gamma, beta, 1e-3, 0.01, rm, rv, relu=True, process_group=dist.group.WORLD)
        gy, gg, gb = F.bn_train_bwd(go, out, y, gamma, stats, relu=True, process_group=dist.group.WORLD, fwd_sums=sums)
        o2, st2 = F.bn_train_fwd(y, gamma, beta, 1e-3, 0.01, rm2, rv2, relu=True)
        gy2, gg2, gb2 = F.bn_train_bwd(go, o2, y, gamma, st2, relu=True)
        assert int(sums[2 * c]) == n
        assert rel(out, o2) < 1e-2 and rel(gy, gy2) < 1e-2
        assert rel(stats, st2) < 1e-6 and rel(gg, gg2) < 1e-5 and rel(gb, gb2) < 1e-5 and rel(rm, rm2) < 1e-6 and rel(rv, rv2) < 1e-6
    finally:
        if created:
            dist.destroy_process_group()
