"""spconv.SparseConvTensor."""
from __future__ import annotations

import numpy as np
import torch

from .. import functional as F


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, n_dev=None):
        """features (N, C) float cuda; indices (N, 4) int32 [b, z, y, x]; spatial_shape zyx.

        n_dev (optional, (1,) int32 on the device): STATIC-SHAPE mode.  N is then a capacity, only the first n_dev[0] rows
        are sites, and every module downstream keeps its row counts on the device too (rulebooks at capacity, see
        spconv.ops.build_rulebook) -- no device->host copy anywhere, so a forward pass of an unmodified module tree can be
        captured into a CUDA graph (pcdet_b200/parta2.py does that with UNetV2).  Rows past the count hold garbage."""
        self.features = features
        self.n_dev = n_dev
        self.depth = 0              # strided convolutions passed so far (static-shape mode: picks the capacity growth)
        self.indices = indices
        if self.indices.dtype != torch.int32:
            self.indices = self.indices.int()
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def dense(self, channels_first=True):
        shape = [int(s) for s in self.spatial_shape]
        if torch.is_grad_enabled() and self.features.requires_grad:
            out = F.to_dense_autograd(self.features, self.indices.contiguous(), shape, int(self.batch_size))
        else:
            out = F.to_dense(self.features.contiguous(), self.indices.contiguous(), shape, int(self.batch_size), n_dev=self.n_dev)
        if not channels_first:
            ndim = len(self.spatial_shape)
            return out.permute(0, *range(2, ndim + 2), 1).contiguous()
        return out

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size
