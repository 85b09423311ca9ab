"""Autograd wrappers (spconv/functional.py in the reference dependency)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import functional as F


class SparseConvFunction(Function):
    """y = sum_k x[nbr[k]] @ W[k]; backward through pcdb_sparse_conv_bwd (fp32)."""

    @staticmethod
    def forward(ctx, features, weight3d, nbr, n_out):
        ctx.save_for_backward(features, weight3d, nbr)
        ctx.n_out = n_out
        return F.sparse_conv_fwd(features.contiguous(), weight3d.contiguous(), nbr, n_out)

    @staticmethod
    def backward(ctx, grad_out):
        features, weight3d, nbr = ctx.saved_tensors
        gf, gw = F.sparse_conv_bwd(features, weight3d, grad_out, nbr, ctx.n_out,
                                   need_input_grad=ctx.needs_input_grad[0],
                                   need_weight_grad=ctx.needs_input_grad[1])
        if gf is not None:
            gf = gf.to(features.dtype)
        if gw is not None:
            gw = gw.to(weight3d.dtype)
        return gf, gw, None, None


indice_conv = SparseConvFunction.apply
