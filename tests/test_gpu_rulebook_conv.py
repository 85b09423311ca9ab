"""GPU parity: rulebook construction and sparse convolution (through the C ABI) against the oracle."""
import numpy as np
import pytest
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from util import nbr_to_pair_sets, rel_err, sort_rows

pytestmark = pytest.mark.gpu


def kitti_coords(orc, seeds=(0,)):
    g = orc.VoxelGenerator(S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    frames = [g.generate(S.kitti_frame(s)) for s in seeds]
    vox, coords, num = orc.collate(frames)
    return vox, coords, num


def random_sites(rng, n, batch, shape):
    cells = rng.choice(batch * int(np.prod(shape)), size=n, replace=False)
    b, rem = np.divmod(cells, int(np.prod(shape)))
    z, rem = np.divmod(rem, shape[1] * shape[2])
    y, x = np.divmod(rem, shape[2])
    return np.stack([b, z, y, x], axis=1).astype(np.int32)


def check_subm(orc, coords, batch, shape, ks=3):
    nbr = F.rulebook_subm(torch.from_numpy(coords).cuda(), batch, shape, ks).cpu().numpy()
    out_ids, pairs, num, _ = orc.get_indice_pairs(coords, batch, shape, ks, 1, 0, 1, subm=True)
    got = nbr_to_pair_sets(nbr, coords.shape[0], coords, coords)
    ref = orc.pairs_to_sets(coords, coords, pairs, num)
    for k, (a, b) in enumerate(zip(got, ref)):
        np.testing.assert_array_equal(a, b, err_msg=f"offset {k}")
    assert (nbr[:, :coords.shape[0]] >= 0).sum() == num.sum()
    return nbr, pairs, num


def check_conv(orc, coords, batch, shape, ks, st, pd):
    r = F.rulebook_conv(torch.from_numpy(coords).cuda(), batch, shape, ks, st, pd)
    n_out, overflow = r["n_out"].tolist()
    assert overflow == 0
    out_ids, pairs, num, out_shape = orc.get_indice_pairs(coords, batch, shape, ks, st, pd, 1, subm=False)
    assert r["out_shape"] == out_shape and n_out == out_ids.shape[0]
    got_ids = r["out_indices"][:n_out].cpu().numpy()
    # same ORDER as the serial reference loop, hence also the same set
    np.testing.assert_array_equal(got_ids, out_ids)
    nbr = r["nbr"].cpu().numpy()
    got = nbr_to_pair_sets(nbr, n_out, coords, got_ids)
    ref = orc.pairs_to_sets(out_ids, coords, pairs, num)
    for k, (a, b) in enumerate(zip(got, ref)):
        np.testing.assert_array_equal(a, b, err_msg=f"offset {k}")
    # inverse map lists the same pairs from the input side
    inv = r["nbr_inv"].cpu().numpy()[:, :coords.shape[0]]
    for k in range(nbr.shape[0]):
        i = np.nonzero(inv[k] >= 0)[0]
        np.testing.assert_array_equal(nbr[k, inv[k, i]], i)
        assert i.size == (nbr[k, :n_out] >= 0).sum()
    return r, out_ids, pairs, num


def test_rulebook_backbone_levels_kitti(orc):
    """All 8 rulebooks of one BackBone8x forward on a 2-frame KITTI-shaped batch: pairs bit-exact as sets,
    output ids in reference order."""
    _, coords, _ = kitti_coords(orc, (0, 1))
    shape, batch = [41, 1600, 1408], 2
    check_subm(orc, coords, batch, shape)
    for ks, st, pd in [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (1, 1, 1)),
                       ((3, 3, 3), (2, 2, 2), (0, 1, 1)), ((3, 1, 1), (2, 1, 1), (0, 0, 0))]:
        r, out_ids, _, _ = check_conv(orc, coords, batch, shape, ks, st, pd)
        coords, shape = out_ids, r["out_shape"]
        if ks == (3, 3, 3):
            check_subm(orc, coords, batch, shape)


def test_rulebook_subm_reuses_strided_site_table(orc):
    """SubM rulebook of a level built from the hash table the strided build of that level left behind
    (pcdb_rulebook_subm_reuse): identical to the stand-alone build, including with a clamped capacity."""
    _, coords, _ = kitti_coords(orc, (2,))
    shape, batch = [41, 1600, 1408], 1
    for ks, st, pd in [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1))]:
        r = F.rulebook_conv(torch.from_numpy(coords).cuda(), batch, shape, ks, st, pd, keep_table=True)
        n_out = int(r["n_out"][0])
        out_ids = r["out_indices"][:n_out].contiguous()
        alone = F.rulebook_subm(out_ids, batch, r["out_shape"], 3)
        # the pipeline's form: capacity-sized coordinate buffer + device-side count
        reused = F.rulebook_subm(r["out_indices"], batch, r["out_shape"], 3, n_dev=r["n_out"], site_table=r["site_table"])
        np.testing.assert_array_equal(reused[:, :n_out].cpu().numpy(), alone.cpu().numpy())
        check_subm(orc, out_ids.cpu().numpy(), batch, r["out_shape"])
        coords, shape = out_ids.cpu().numpy(), r["out_shape"]


def test_rulebook_subm_even_and_dilated_kernels(orc):
    """Even kernel sizes have no k <-> K-1-k symmetry (all offsets probed); dilation keeps it."""
    rng = np.random.default_rng(5)
    shape = [8, 11, 13]
    coords = random_sites(rng, 500, 2, shape)
    check_subm(orc, coords, 2, shape, ks=(2, 2, 2))
    check_subm(orc, coords, 2, shape, ks=(1, 3, 3))
    nbr = F.rulebook_subm(torch.from_numpy(coords).cuda(), 2, shape, 3, dilation=2).cpu().numpy()
    _, pairs, num, _ = orc.get_indice_pairs(coords, 2, shape, 3, 1, 0, 2, subm=True)
    got = nbr_to_pair_sets(nbr, coords.shape[0], coords, coords)
    for k, (a, b) in enumerate(zip(got, orc.pairs_to_sets(coords, coords, pairs, num))):
        np.testing.assert_array_equal(a, b, err_msg=f"offset {k}")


@pytest.mark.parametrize("ks,st,pd", [((3, 3, 3), (1, 1, 1), (1, 1, 1)), ((2, 2, 2), (2, 2, 2), (0, 0, 0)),
                                        ((3, 3, 3), (3, 2, 1), (1, 0, 2)), ((1, 3, 3), (1, 2, 2), (0, 1, 1))])
def test_rulebook_conv_generic_geometry(orc, ks, st, pd):
    rng = np.random.default_rng(11)
    shape = [9, 17, 12]
    coords = random_sites(rng, 900, 3, shape)
    check_conv(orc, coords, 3, shape, ks, st, pd)


def test_rulebook_edge_cases(orc):
    shape = [5, 6, 7]
    one = np.array([[0, 0, 0, 0]], np.int32)
    check_subm(orc, one, 1, shape)
    check_conv(orc, one, 1, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    full = random_sites(np.random.default_rng(1), 5 * 6 * 7, 1, shape)     # every site active
    nbr, _, num = check_subm(orc, full, 1, shape)
    assert num[13] == 210
    check_conv(orc, full, 1, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    # device-side count smaller than the host bound: rows beyond it are ignored
    coords = random_sites(np.random.default_rng(2), 300, 2, shape)
    n_dev = torch.tensor([200], dtype=torch.int32, device="cuda")
    nbr = F.rulebook_subm(torch.from_numpy(coords).cuda(), 2, shape, 3, n_dev=n_dev).cpu().numpy()
    ref = F.rulebook_subm(torch.from_numpy(coords[:200].copy()).cuda(), 2, shape, 3).cpu().numpy()
    np.testing.assert_array_equal(nbr[:, :200], ref[:, :200])


def test_get_indice_pairs_compat(orc):
    """spconv.ops.get_indice_pairs form: (outids, indice_pairs (K,2,N) -1 padded, indice_pair_num (K))."""
    from pcdet_b200.spconv import ops
    rng = np.random.default_rng(3)
    shape = [8, 10, 12]
    coords = random_sites(rng, 500, 2, shape)
    t = torch.from_numpy(coords).cuda()
    for subm, st, pd in [(True, 1, 0), (False, 2, 1)]:
        outids, pairs, num = ops.get_indice_pairs(t, 2, shape, 3, st, pd, 1, subm=subm)
        o_ids, o_pairs, o_num, _ = orc.get_indice_pairs(coords, 2, shape, 3, st, pd, 1, subm=subm)
        np.testing.assert_array_equal(num.cpu().numpy(), o_num)
        got = orc.pairs_to_sets(outids.cpu().numpy(), coords, pairs.cpu().numpy(), num.cpu().numpy())
        ref = orc.pairs_to_sets(o_ids, coords, o_pairs, o_num)
        for a, b in zip(got, ref):
            np.testing.assert_array_equal(a, b)
        assert pairs.shape[0] == 27 and pairs.shape[1] == 2 and pairs.dtype == torch.int32


# ---------------------------------------------------------------------------------------------- conv
CHANNELS = [(4, 16), (16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (5, 7), (20, 130)]


@pytest.mark.parametrize("cin,cout", CHANNELS)
@pytest.mark.parametrize("subm", [True, False])
def test_conv_fwd_fp32(orc, cin, cout, subm):
    """fp32 path: <= 1e-4 relative to the fp64-accumulated oracle (north_star tolerance)."""
    rng = np.random.default_rng(cin * 131 + cout)
    shape, batch = [9, 20, 24], 2
    coords = random_sites(rng, 1500, batch, shape)
    feat = rng.normal(0, 1, (coords.shape[0], cin)).astype(np.float32)
    ks, st, pd = (3, 3, 3), ((1, 1, 1) if subm else (2, 2, 2)), (1, 1, 1)
    w = (rng.uniform(-1, 1, (*ks, cin, cout)) / np.sqrt(cin * 27)).astype(np.float32)
    out_ids, pairs, num, _ = orc.get_indice_pairs(coords, batch, shape, ks, st, pd, 1, subm=subm)
    ref = orc.indice_conv(feat, w, pairs, num, out_ids.shape[0], subm=subm, acc64=True)
    t = torch.from_numpy(coords).cuda()
    if subm:
        nbr, n_out = F.rulebook_subm(t, batch, shape, ks), coords.shape[0]
        perm = np.arange(n_out)
    else:
        r = F.rulebook_conv(t, batch, shape, ks, st, pd)
        nbr, n_out = r["nbr"], int(r["n_out"][0].item())
    wt = torch.from_numpy(w.reshape(27, cin, cout)).cuda()
    got = F.sparse_conv_fwd(torch.from_numpy(feat).cuda(), wt, nbr, n_out, algo=1).cpu().numpy()
    assert rel_err(got, ref) < 1e-4
    # fused epilogue: scale/shift/bias/relu
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)).cuda()
    shift = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.normal(0, 0.1, cout).astype(np.float32)).cuda()
    got2 = F.sparse_conv_fwd(torch.from_numpy(feat).cuda(), wt, nbr, n_out, scale=scale, shift=shift, bias=bias,
                             relu=True, algo=1).cpu().numpy()
    ref2 = np.maximum((ref + bias.cpu().numpy()) * scale.cpu().numpy() + shift.cpu().numpy(), 0)
    assert rel_err(got2, ref2) < 1e-4


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128)])
def test_conv_fwd_bf16(orc, cin, cout):
    """bf16 storage, fp32 accumulate: <= 1e-2 relative to the fp32 oracle; SIMT and tensor-core
    kernels agree with each other far more tightly (same operands, different summation order)."""
    rng = np.random.default_rng(cin * 17 + cout)
    shape, batch = [9, 30, 40], 2
    coords = random_sites(rng, 5000, batch, shape)
    feat = rng.normal(0, 1, (coords.shape[0], cin)).astype(np.float32)
    w = (rng.uniform(-1, 1, (3, 3, 3, cin, cout)) / np.sqrt(cin * 27)).astype(np.float32)
    out_ids, pairs, num, _ = orc.get_indice_pairs(coords, batch, shape, 3, 1, 0, 1, subm=True)
    ref = orc.indice_conv_mm(feat, w, pairs, num, coords.shape[0], subm=True)
    nbr = F.rulebook_subm(torch.from_numpy(coords).cuda(), batch, shape, 3)
    fb = torch.from_numpy(feat).cuda().bfloat16()
    wb = torch.from_numpy(w.reshape(27, cin, cout)).cuda().bfloat16()
    simt = F.sparse_conv_fwd(fb, wb, nbr, coords.shape[0], algo=1).float().cpu().numpy()
    auto = F.sparse_conv_fwd(fb, wb, nbr, coords.shape[0], algo=0).float().cpu().numpy()
    assert rel_err(simt, ref) < 1e-2
    assert rel_err(auto, ref) < 1e-2
    assert rel_err(auto, simt) < 8e-3    # two bf16 ulps of the largest value (different summation order)


@pytest.mark.parametrize("cin,cout", [(64, 64), (64, 128), (32, 32)])
def test_conv_fwd_tma_gather4_variant(cin, cout):
    """algo=2: the TMA tile::gather4 variant of the tcgen05 kernel (c_in = 64; other widths run the cp.async kernel) must
    give the same result as the default algorithm: same operands, same accumulation order per tile."""
    rng = np.random.default_rng(cin + cout)
    n, K = 3000, 27
    nbr = torch.where(torch.rand(K, n, device="cuda") < 0.4, torch.randint(0, n, (K, n), device="cuda", dtype=torch.int32),
                      torch.full((K, n), -1, dtype=torch.int32, device="cuda")).contiguous()
    f = torch.from_numpy(rng.normal(0, 1, (n, cin)).astype(np.float32)).cuda().bfloat16()
    w = torch.from_numpy((rng.uniform(-1, 1, (K, cin, cout)) / np.sqrt(cin * K)).astype(np.float32)).cuda().bfloat16()
    a = F.sparse_conv_fwd(f, w, nbr, n, algo=3).float()
    b = F.sparse_conv_fwd(f, w, nbr, n, algo=2).float()
    c = F.sparse_conv_fwd(f, w, nbr, n, algo=1).float()
    assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < 1e-6
    assert rel_err(a.cpu().numpy(), c.cpu().numpy()) < 8e-3


def _autograd_reference(feat, w, nbr, n_out, go):
    """torch formulation of the same sum (gather with a zero row for -1) and its autograd gradients"""
    n_in, cin = feat.shape
    f = feat.detach().clone().requires_grad_(True)
    ww = w.detach().clone().requires_grad_(True)
    fpad = torch.cat([f, f.new_zeros((1, cin))], dim=0)
    idx = nbr[:, :n_out].long()
    idx = torch.where(idx < 0, torch.full_like(idx, n_in), idx)
    y = torch.einsum("koc,kcd->od", fpad[idx], ww)
    y.backward(go)
    return y.detach(), f.grad, ww.grad


@pytest.mark.parametrize("cin,cout", [(16, 32), (4, 16), (64, 64), (64, 128), (6, 10)])
@pytest.mark.parametrize("mode", ["scatter", "transposed_map", "subm_flip"])
def test_conv_bwd_matches_autograd(orc, cin, cout, mode):
    """pcdb_sparse_conv_bwd (tiled and generic weight gradient; scattered input gradient) and the input gradient
    computed as a forward convolution over the rulebook read the other way round -- all against torch autograd."""
    from pcdet_b200.spconv.functional import indice_conv
    rng = np.random.default_rng(21 + cin)
    shape, batch = [7, 12, 14], 2
    coords = random_sites(rng, 800, batch, shape)
    ct = torch.from_numpy(coords).cuda()
    if mode == "subm_flip":
        nbr, n_out, nbr_t, flip = F.rulebook_subm(ct, batch, shape, 3, 1), coords.shape[0], None, True
    else:
        r = F.rulebook_conv(ct, batch, shape, 3, 2, 1)
        nbr, n_out = r["nbr"], int(r["n_out"][0].item())
        nbr_t, flip = (r["nbr_inv"] if mode == "transposed_map" else None), False
    feat = torch.from_numpy(rng.normal(0, 1, (coords.shape[0], cin)).astype(np.float32)).cuda()
    w = torch.from_numpy(rng.normal(0, 0.1, (27, cin, cout)).astype(np.float32)).cuda()
    go = torch.from_numpy(rng.normal(0, 1, (n_out, cout)).astype(np.float32)).cuda()
    y_ref, gf_ref, gw_ref = _autograd_reference(feat, w, nbr, n_out, go)
    f2, w2 = feat.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y = indice_conv(f2, w2, nbr, n_out, nbr_t, flip)
    assert rel_err(y.detach().cpu().numpy(), y_ref.cpu().numpy()) < 1e-4
    y.backward(go)
    assert rel_err(f2.grad.cpu().numpy(), gf_ref.cpu().numpy()) < 1e-4
    assert rel_err(w2.grad.cpu().numpy(), gw_ref.cpu().numpy()) < 1e-4


def test_module_backward_uses_the_transposed_rulebook(orc):
    """SubMConv3d, SparseConv3d and SparseInverseConv3d in train mode: gradients through the module API equal the
    autograd reference of every layer's own rulebook (the inverse conv's transposed map is the strided conv's nbr)."""
    import pcdet_b200.spconv as spconv
    rng = np.random.default_rng(4)
    shape, batch = [9, 16, 16], 2
    coords = random_sites(rng, 900, batch, shape)
    net = spconv.SparseSequential(spconv.SubMConv3d(8, 16, 3, bias=False, indice_key="s1"),
                                  spconv.SparseConv3d(16, 32, 3, 2, 1, bias=False, indice_key="d1"),
                                  spconv.SparseInverseConv3d(32, 8, 3, indice_key="d1", bias=False)).cuda().train()
    feat = torch.from_numpy(rng.normal(0, 1, (coords.shape[0], 8)).astype(np.float32)).cuda().requires_grad_(True)
    x = spconv.SparseConvTensor(feat, torch.from_numpy(coords).cuda(), shape, batch)
    out = net(x)
    go = torch.from_numpy(rng.normal(0, 1, tuple(out.features.shape)).astype(np.float32)).cuda()
    out.features.backward(go)
    # the same three layers written with torch ops on the rulebooks the modules stored
    rb_s, rb_d = x.indice_dict["s1"], x.indice_dict["d1"]
    f = feat.detach().clone().requires_grad_(True)
    ws = [m.weight.detach().clone().view(27, m.in_channels, m.out_channels).requires_grad_(True) for m in net]

    def conv(inp, w, nbr, n_out):
        pad = torch.cat([inp, inp.new_zeros((1, inp.shape[1]))], dim=0)
        idx = nbr[:, :n_out].long()
        idx = torch.where(idx < 0, torch.full_like(idx, inp.shape[0]), idx)
        return torch.einsum("koc,kcd->od", pad[idx], w)

    y = conv(conv(conv(f, ws[0], rb_s.nbr, rb_s.n_out), ws[1], rb_d.nbr, rb_d.n_out), ws[2], rb_d.nbr_inv, rb_d.n_in)
    assert rel_err(out.features.detach().cpu().numpy(), y.detach().cpu().numpy()) < 1e-4
    y.backward(go)
    assert rel_err(feat.grad.cpu().numpy(), f.grad.cpu().numpy()) < 1e-4
    for m, w in zip(net, ws):
        assert rel_err(m.weight.grad.view(27, m.in_channels, m.out_channels).cpu().numpy(), w.grad.cpu().numpy()) < 1e-4


def test_to_dense(orc):
    rng = np.random.default_rng(5)
    shape, batch, c = [2, 20, 17], 3, 128
    coords = random_sites(rng, 400, batch, shape)
    feat = rng.normal(0, 1, (400, c)).astype(np.float32)
    ref = orc.to_dense(feat, coords, shape, batch)
    got = F.to_dense(torch.from_numpy(feat).cuda(), torch.from_numpy(coords).cuda(), shape, batch)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    gb = F.to_dense(torch.from_numpy(feat).cuda().bfloat16(), torch.from_numpy(coords).cuda(), shape, batch)
    assert gb.dtype == torch.bfloat16 and gb.shape == (batch, c, *shape)


def test_sparse_maxpool(orc):
    """SparseMaxPool3d(2, 2) as used by the Part-A2 RCNN head (partA2_rcnn_net.py:165)."""
    import pcdet_b200.spconv as spconv
    rng = np.random.default_rng(8)
    shape, batch, c = [14, 14, 14], 3, 32
    coords = random_sites(rng, 900, batch, shape)
    feat = rng.normal(0, 1, (coords.shape[0], c)).astype(np.float32)
    out_ids, pairs, num, out_shape = orc.get_indice_pairs(coords, batch, shape, 2, 2, 0, 1, subm=False)
    ref = orc.indice_maxpool(feat, pairs, num, out_ids.shape[0])
    x = spconv.SparseConvTensor(torch.from_numpy(feat).cuda(), torch.from_numpy(coords).cuda(), shape, batch)
    y = spconv.SparseMaxPool3d(2, 2)(x)
    assert list(y.spatial_shape) == out_shape == [7, 7, 7]
    np.testing.assert_array_equal(y.indices.cpu().numpy(), out_ids)
    np.testing.assert_array_equal(y.features.cpu().numpy(), ref)
    yb = spconv.SparseMaxPool3d(3, 2, 1)(spconv.SparseConvTensor(torch.from_numpy(feat).cuda().bfloat16(),
                                                               torch.from_numpy(coords).cuda(), shape, batch))
    assert yb.features.dtype == torch.bfloat16


def test_sparse_maxpool_backward(orc):
    """Gradient of SparseMaxPool3d against torch autograd on the dense equivalent: the pooled maximum's gradient goes to the
    input that attained it (distinct random values: no ties; all-negative windows pool to 0 and pass nothing back)."""
    import pcdet_b200.spconv as spconv
    rng = np.random.default_rng(9)
    shape, batch, c = [8, 8, 8], 2, 8
    coords = random_sites(rng, 300, batch, shape)
    feat = torch.from_numpy(rng.normal(0, 1, (coords.shape[0], c)).astype(np.float32)).cuda().requires_grad_(True)
    idx = torch.from_numpy(coords).cuda()
    y = spconv.SparseMaxPool3d(2, 2)(spconv.SparseConvTensor(feat, idx, shape, batch))
    g = torch.from_numpy(rng.normal(0, 1, tuple(y.features.shape)).astype(np.float32)).cuda()
    y.features.backward(g)
    # dense reference: inactive cells hold -inf, the pooled value is max(0, max over the window)
    f2 = feat.detach().clone().requires_grad_(True)
    dense = torch.full((batch, *shape, c), -float("inf"), device="cuda")
    li = idx.long()
    dense = dense.index_put((li[:, 0], li[:, 1], li[:, 2], li[:, 3]), f2).permute(0, 4, 1, 2, 3)
    pooled = torch.clamp_min(torch.nn.functional.max_pool3d(dense, 2, 2), 0.0)
    oi = y.indices.long()
    ref = pooled[oi[:, 0], :, oi[:, 1], oi[:, 2], oi[:, 3]]
    torch.testing.assert_close(y.features.detach(), ref.detach())
    ref.backward(g)
    torch.testing.assert_close(feat.grad, f2.grad)


def test_extent_clears(orc):
    """pcdb_fill_rows_i32 touches exactly the first min(*rows, cap) entries of every map; pcdb_dense_clear_rows undoes a
    pcdb_to_dense exactly."""
    import ctypes as C
    from pcdet_b200._lib import F32, check, i32x3, lib, ptr
    L = lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    K, ld = 5, 1000
    for rows, cap in ((0, 1000), (1, 1000), (333, 1000), (1000, 1000), (5000, 777)):
        buf = torch.arange(K * ld, dtype=torch.int32, device="cuda").view(K, ld).contiguous()
        n_dev = torch.tensor([rows, 0], dtype=torch.int32, device="cuda")
        check(L.pcdb_fill_rows_i32(ptr(buf), ld, K, ptr(n_dev), cap, -1, st), "fill_rows")
        want = torch.arange(K * ld, dtype=torch.int32).view(K, ld).clone()
        want[:, :min(rows, cap)] = -1
        assert torch.equal(buf.cpu(), want), (rows, cap)
    rng = np.random.default_rng(5)
    shape, batch, c = [2, 20, 17], 3, 128
    coords = torch.from_numpy(random_sites(rng, 400, batch, shape)).cuda()
    feat = torch.from_numpy(rng.normal(0, 1, (400, c)).astype(np.float32)).cuda()
    dense = F.to_dense(feat, coords, shape, batch)
    assert dense.abs().sum() > 0
    n_dev = torch.tensor([400, 0], dtype=torch.int32, device="cuda")
    check(L.pcdb_dense_clear_rows(ptr(coords), 400, ptr(n_dev), c, batch, i32x3(shape), ptr(dense), F32, st), "dense_clear_rows")
    assert dense.abs().sum().item() == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_1x1x1_runs_on_the_library_kernels(orc, dtype):
    """kernel_size 1 (spconv: torch.mm on the features): identity rulebook through pcdb_sparse_conv_fwd, with bias, and -- in
    fp32 -- with gradients through the module."""
    import pcdet_b200.spconv as spconv
    rng = np.random.default_rng(21)
    shape, batch = [10, 12, 14], 2
    coords = random_sites(rng, 777, batch, shape)
    conv = spconv.SubMConv3d(16, 32, 1, bias=True, indice_key="p").cuda()
    feat = torch.from_numpy(rng.normal(0, 1, (coords.shape[0], 16)).astype(np.float32)).cuda()
    x = spconv.SparseConvTensor(feat.to(dtype), torch.from_numpy(coords).cuda(), shape, batch)
    with torch.no_grad():
        y = (conv.to(dtype) if dtype == torch.bfloat16 else conv)(x)
    w = conv.weight.detach().float().view(16, 32)
    ref = feat.to(dtype).float() @ w + conv.bias.detach().float()
    assert list(y.spatial_shape) == shape and torch.equal(y.indices, x.indices)
    assert rel_err(y.features.float().cpu().numpy(), ref.cpu().numpy()) < (1e-5 if dtype == torch.float32 else 1e-2)
    if dtype == torch.float32:
        f2 = feat.clone().requires_grad_(True)
        out = conv(spconv.SparseConvTensor(f2, torch.from_numpy(coords).cuda(), shape, batch)).features
        g = torch.randn_like(out)
        out.backward(g)
        f3 = feat.clone().requires_grad_(True)
        (f3 @ conv.weight.view(16, 32) + conv.bias).backward(g)
        assert rel_err(f2.grad.cpu().numpy(), f3.grad.cpu().numpy()) < 1e-5
