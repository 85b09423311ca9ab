"""spconv.ops compatible functions on top of the pcdet_b200 rulebook kernels."""
from __future__ import annotations

import torch

from .. import functional as F


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    return F.conv_output_size(input_size, kernel_size, stride, padding, dilation)


def get_deconv_output_size(input_size, kernel_size, stride, padding, dilation, output_padding):
    return [(i - 1) * s - 2 * p + k + op
            for i, k, s, p, op in zip(input_size, kernel_size, stride, padding, output_padding)]


def nbr_to_pairs(nbr: torch.Tensor, n_out: int, n_in: int):
    """Reference form of a rulebook from the output-stationary neighbour map:
    indice_pairs (K, 2, n_in) int32 padded with -1 and indice_pair_num (K) int32, listing for every
    kernel offset the (input row, output row) pairs -- the contract of spconv.ops.get_indice_pairs.
    Pair order inside an offset is by output row (the reference GPU path leaves it to atomics)."""
    K = nbr.shape[0]
    m = nbr[:, :n_out]
    valid = m >= 0
    num = valid.sum(dim=1).to(torch.int32)
    width = max(int(n_in), int(num.max().item()) if K > 0 and n_out > 0 else 0, 1)
    pairs = torch.full((K, 2, width), -1, dtype=torch.int32, device=nbr.device)
    if n_out > 0:
        k_idx, o_idx = valid.nonzero(as_tuple=True)
        pos = (torch.cumsum(valid, dim=1) - 1)[k_idx, o_idx]
        pairs[k_idx, 0, pos] = m[k_idx, o_idx]
        pairs[k_idx, 1, pos] = o_idx.to(torch.int32)
    return pairs, num


class Rulebook:
    """What SparseConvTensor.indice_dict holds per indice_key.  Iterates / indexes like the
    reference's 5-tuple (outids, indices, indice_pairs, indice_pair_num, spatial_shape) -- the pair
    tensors are materialised lazily -- and carries the neighbour maps the kernels consume."""

    def __init__(self, outids, indices, nbr, nbr_inv, n_out, n_in, spatial_shape, out_spatial_shape, subm,
                 n_out_dev=None, n_in_dev=None, overflow=None):
        self.outids = outids
        self.indices = indices
        self.nbr = nbr                # (K, ld) int32: input row per (offset, output row)
        self.nbr_inv = nbr_inv        # (K, ld_in) int32: output row per (offset, input row); None for subm
        self.n_out = int(n_out)
        self.n_in = int(n_in)
        self.spatial_shape = spatial_shape
        self.out_spatial_shape = out_spatial_shape
        self.subm = subm
        # static-shape mode: n_out / n_in are capacities, the counts live on the device; overflow (1,) i32 is raised
        # when a strided build found more sites than its capacity
        self.n_out_dev, self.n_in_dev, self.overflow = n_out_dev, n_in_dev, overflow
        self._pairs = None

    def pairs(self):
        if self._pairs is None:
            self._pairs = nbr_to_pairs(self.nbr, self.n_out, self.n_in)
        return self._pairs

    def _as_tuple(self):
        p, n = self.pairs()
        return (self.outids, self.indices, p, n, self.spatial_shape)

    def __iter__(self):
        return iter(self._as_tuple())

    def __getitem__(self, i):
        return self._as_tuple()[i]

    def __len__(self):
        return 5


# Static-shape mode (SparseConvTensor.n_dev): row capacity of the output of a strided convolution as a multiple of its
# input's capacity.  LiDAR surfaces dilate by ~1.4x under the first stride-2 3x3x3 convolution and shrink afterwards
# (SURVEY App. B), the worst case is 8x; an overflow raises the rulebook's flag and drops the sites beyond the capacity.
CAPACITY_GROWTH = {0: 1.6}
CAPACITY_GROWTH_DEFAULT = 1.0


def build_rulebook(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, subm, n_dev=None, depth=0) -> Rulebook:
    spatial_shape = [int(s) for s in spatial_shape]
    indices = indices.contiguous()
    n = indices.shape[0]
    if subm:
        nbr = F.rulebook_subm(indices, batch_size, spatial_shape, ksize, dilation, n_dev=n_dev)
        return Rulebook(indices, indices, nbr, None, n, n, spatial_shape, spatial_shape, True, n_out_dev=n_dev, n_in_dev=n_dev)
    if n_dev is not None:
        cap = max(int(n * CAPACITY_GROWTH.get(depth, CAPACITY_GROWTH_DEFAULT)), 64)
        r = F.rulebook_conv(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, n_dev=n_dev, out_capacity=cap)
        return Rulebook(r["out_indices"], indices, r["nbr"], r["nbr_inv"], cap, n, spatial_shape, r["out_shape"], False,
                        n_out_dev=r["n_out"][:1], n_in_dev=n_dev, overflow=r["n_out"][1:])
    r = F.rulebook_conv(indices, batch_size, spatial_shape, ksize, stride, padding, dilation)
    count, overflow = r["n_out"].tolist()     # the module API needs exact shapes: one host sync per build
    assert overflow == 0, "rulebook_conv output capacity exceeded"
    outids = r["out_indices"][:count]
    return Rulebook(outids, indices, r["nbr"], r["nbr_inv"], count, n, spatial_shape, r["out_shape"], False)


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1, out_padding=0,
                     subm=False, transpose=False, grid=None):
    """spconv.ops.get_indice_pairs: returns (outids, indice_pairs, indice_pair_num)."""
    assert not transpose, "transposed sparse convolution is not used by PCDet and not implemented"
    ndim = indices.shape[1] - 1
    assert ndim == 3

    def tri(v):
        return [v] * 3 if isinstance(v, int) else list(v)

    rb = build_rulebook(indices.int(), batch_size, spatial_shape, tri(ksize), tri(stride), tri(padding),
                        tri(dilation), subm)
    p, n = rb.pairs()
    return rb.outids, p, n
