"""GPU parity: pcdb_rulebook_chain (every rulebook of a strided backbone in four launches) against the oracle's
restatement of spconv getIndicePair, level by level.  The chain numbers the rows of levels >= 1 in ascending (b,z,y,x)
order -- the order of the reference's CUDA rulebook -- while the oracle restates the CPU loop (first touch), so site sets
and per-offset pair sets are compared through the coordinates (sorted), as north_star prescribes where the reference's
own order is implementation-defined."""
import numpy as np
import pytest
import torch

from pcdet_b200 import functional as F
from pcdet_b200 import synthetic as S
from util import nbr_to_pair_sets, sort_rows

pytestmark = pytest.mark.gpu

BACKBONE_CONVS = [dict(ksize=3, stride=2, padding=1), dict(ksize=3, stride=2, padding=1),
                  dict(ksize=3, stride=2, padding=(0, 1, 1)), dict(ksize=(3, 1, 1), stride=(2, 1, 1), padding=0)]
BACKBONE_SUBM = [3, 3, 3, 3, None]


def kitti_coords(orc, seeds):
    g = orc.VoxelGenerator(S.KITTI["voxel_size"], S.KITTI["point_cloud_range"], 5, 40000)
    _vox, coords, _num = orc.collate([g.generate(S.kitti_frame(s)) for s in seeds])
    return coords


def random_sites(rng, n, batch, shape):
    cells = rng.choice(batch * int(np.prod(shape)), size=n, replace=False)
    b, rem = np.divmod(cells, int(np.prod(shape)))
    z, rem = np.divmod(rem, shape[1] * shape[2])
    y, x = np.divmod(rem, shape[2])
    return np.stack([b, z, y, x], axis=1).astype(np.int32)


def check_chain(orc, coords0, batch, shape, convs, subm, caps=None):
    r = F.rulebook_chain(torch.from_numpy(coords0).cuda(), None, batch, shape, convs, subm, caps=caps)
    torch.cuda.synchronize()
    ref_ids = coords0
    got_prev = coords0
    for l in range(len(convs) + 1):
        if l > 0:
            c = convs[l - 1]
            out_ids, pairs, num, out_shape = orc.get_indice_pairs(ref_ids, batch, r["shapes"][l - 1], c["ksize"], c["stride"],
                                                                  c["padding"], 1, subm=False)
            assert r["shapes"][l] == out_shape
            n_out, overflow = r["counts"][l].tolist()
            assert overflow == 0 and n_out == out_ids.shape[0], (l, n_out, out_ids.shape)
            got_ids = r["coords"][l][:n_out].cpu().numpy()
            np.testing.assert_array_equal(sort_rows(got_ids), sort_rows(out_ids), err_msg=f"site set of level {l}")
            got = nbr_to_pair_sets(r["nbr_conv"][l].cpu().numpy(), n_out, got_prev, got_ids)
            ref = orc.pairs_to_sets(out_ids, ref_ids, pairs, num)
            for k, (a, b) in enumerate(zip(got, ref)):
                np.testing.assert_array_equal(a, b, err_msg=f"conv {l} offset {k}")
            ref_ids, got_prev = out_ids, got_ids
        if subm[l] is not None:
            n = got_prev.shape[0]
            _o, pairs, num, _s = orc.get_indice_pairs(ref_ids, batch, r["shapes"][l], subm[l], 1, 0, 1, subm=True)
            got = nbr_to_pair_sets(r["nbr_subm"][l].cpu().numpy(), n, got_prev, got_prev)
            ref = orc.pairs_to_sets(ref_ids, ref_ids, pairs, num)
            for k, (a, b) in enumerate(zip(got, ref)):
                np.testing.assert_array_equal(a, b, err_msg=f"SubM level {l} offset {k}")
    return r


def test_chain_backbone8x_kitti(orc):
    coords = kitti_coords(orc, (0, 1))
    r = check_chain(orc, coords, 2, [41, 1600, 1408], BACKBONE_CONVS, BACKBONE_SUBM)
    assert r["shapes"] == [[41, 1600, 1408], [21, 800, 704], [11, 400, 352], [5, 200, 176], [2, 200, 176]]


@pytest.mark.parametrize("convs,subm", [
    ([dict(ksize=2, stride=2, padding=0), dict(ksize=3, stride=1, padding=1)], [3, None, 3]),
    ([dict(ksize=3, stride=1, padding=0), dict(ksize=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1)), dict(ksize=3, stride=3, padding=1)],
     [None, (1, 3, 3), 3, None]),
    ([dict(ksize=(3, 2, 3), stride=(2, 2, 1), padding=(1, 0, 2))], [(3, 1, 3), (1, 1, 3)]),
    ([], [3]),
])
def test_chain_generic_geometry(orc, convs, subm):
    rng = np.random.default_rng(3)
    shape = [9, 14, 17]
    coords = random_sites(rng, 700, 2, shape)
    check_chain(orc, coords, 2, shape, convs, subm)


def test_chain_rows_are_sorted_and_independent_of_the_input_order(orc):
    """The rows of every level >= 1 come in ascending (b, z, y, x) order (the reference's CUDA rulebook order), so they and the
    maps between them cannot depend on the order in which the level-0 rows arrive."""
    coords = kitti_coords(orc, (2,))
    rng = np.random.default_rng(0)
    perm = rng.permutation(coords.shape[0])
    a = F.rulebook_chain(torch.from_numpy(coords).cuda(), None, 1, [41, 1600, 1408], BACKBONE_CONVS, BACKBONE_SUBM)
    b = F.rulebook_chain(torch.from_numpy(coords[perm].copy()).cuda(), None, 1, [41, 1600, 1408], BACKBONE_CONVS, BACKBONE_SUBM)
    for l in range(1, 5):
        n = int(a["counts"][l][0])
        assert n == int(b["counts"][l][0])
        assert torch.equal(a["coords"][l][:n], b["coords"][l][:n])
        got = a["coords"][l][:n].cpu().numpy()
        np.testing.assert_array_equal(got, sort_rows(got))
        if l >= 2:          # both inputs and outputs of these maps are level >= 1 rows
            assert torch.equal(a["nbr_conv"][l][:, :n], b["nbr_conv"][l][:, :n])
        if BACKBONE_SUBM[l] is not None:
            assert torch.equal(a["nbr_subm"][l][:, :n], b["nbr_subm"][l][:, :n])


def test_chain_edge_cases(orc):
    shape = [9, 14, 17]
    # empty level 0 (device-side count 0 with a non-empty buffer)
    coords = random_sites(np.random.default_rng(1), 50, 1, shape)
    r = F.rulebook_chain(torch.from_numpy(coords).cuda(), torch.zeros(1, dtype=torch.int32, device="cuda"), 1, shape,
                         [dict(ksize=3, stride=2, padding=1)], [3, 3])
    assert r["counts"][1].tolist() == [0, 0]
    assert (r["nbr_subm"][0] == -1).all() and (r["nbr_conv"][1] == -1).all()
    # device-side count smaller than the buffer: only the first rows exist
    r = F.rulebook_chain(torch.from_numpy(coords).cuda(), torch.tensor([20], dtype=torch.int32, device="cuda"), 1, shape,
                         [dict(ksize=3, stride=2, padding=1)], [None, None])
    out_ids, _p, _n, _s = orc.get_indice_pairs(coords[:20], 1, shape, 3, 2, 1, 1, subm=False)
    n = r["counts"][1].tolist()
    assert n == [out_ids.shape[0], 0]
    np.testing.assert_array_equal(sort_rows(r["coords"][1][:n[0]].cpu().numpy()), sort_rows(out_ids))
    # a capacity that is too small raises the overflow flag and truncates, never writes out of bounds
    r = F.rulebook_chain(torch.from_numpy(coords).cuda(), None, 1, shape, [dict(ksize=3, stride=2, padding=1)], [None, 3],
                         caps=[50, 10])
    cnt, overflow = r["counts"][1].tolist()
    assert overflow == 1 and cnt == 10
    assert int(r["nbr_conv"][1].max()) < 50 and int(r["nbr_subm"][1][:, :cnt].max()) < 10
    # kernel < stride leaves holes between the windows: not a box, rejected
    with pytest.raises(Exception):
        F.rulebook_chain(torch.from_numpy(coords).cuda(), None, 1, shape, [dict(ksize=1, stride=2, padding=0)], [None, None])
