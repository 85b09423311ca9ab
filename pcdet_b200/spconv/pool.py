"""spconv.SparseMaxPool3d: rulebook of a strided convolution, running maximum instead of a GEMM.
Forward and backward (the reference uses it in the Part-A2 RCNN head, pcdet/models/rcnn/partA2_rcnn_net.py:165)."""
from __future__ import annotations

from .. import functional as F
from . import ops
from .conv import _ntuple
from .modules import SparseModule
from .tensor import SparseConvTensor


class SparseMaxPool(SparseModule):
    def __init__(self, ndim, kernel_size, stride=1, padding=0, dilation=1, subm=False):
        super().__init__()
        assert ndim == 3
        self.ndim = ndim
        self.kernel_size = _ntuple(kernel_size, ndim)
        self.stride = _ntuple(stride, ndim)
        self.padding = _ntuple(padding, ndim)
        self.dilation = _ntuple(dilation, ndim)
        self.subm = subm

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        spatial_shape = [int(s) for s in input.spatial_shape]
        rb = ops.build_rulebook(input.indices, input.batch_size, spatial_shape, self.kernel_size, self.stride,
                                self.padding, self.dilation, self.subm)
        feats = F.sparse_maxpool(input.features.contiguous(), rb.nbr, rb.n_out)
        out = SparseConvTensor(feats, rb.outids, rb.out_spatial_shape, input.batch_size)
        out.indice_dict = input.indice_dict
        out.grid = input.grid
        return out


class SparseMaxPool3d(SparseMaxPool):
    def __init__(self, kernel_size, stride=1, padding=0, dilation=1):
        super().__init__(3, kernel_size, stride, padding, dilation)
