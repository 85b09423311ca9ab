"""Autograd wrappers (spconv/functional.py in the reference dependency)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import functional as F


class SparseConvFunction(Function):
    """y = sum_k x[nbr[k]] @ W[k]; backward in fp32.

    grad_weight comes from pcdb_sparse_conv_bwd.  grad_features is itself a sparse convolution -- of grad_out, with the
    transposed weights, over the rulebook read the other way round -- so when the caller can name that map it runs
    through the forward kernel (register-tiled, no atomics, fixed summation order):
      * nbr_t given (strided conv: the rulebook's input-stationary map; inverse conv: its output-stationary one):
        grad_x[i] = sum_k grad_y[nbr_t[k][i]] @ W[k]^T;
      * flip (submanifold conv with centred offsets): nbr[k][o] = i  <=>  nbr[K-1-k][i] = o, so the same map serves
        with the offsets reversed: grad_x[i] = sum_k grad_y[nbr[k][i]] @ W[K-1-k]^T.
    Otherwise pcdb_sparse_conv_bwd scatters with atomics (spconv's indiceConvBackward order)."""

    @staticmethod
    def forward(ctx, features, weight3d, nbr, n_out, nbr_t=None, flip=False):
        ctx.save_for_backward(features, weight3d, nbr, nbr_t)
        ctx.n_out = n_out
        ctx.flip = flip
        return F.sparse_conv_fwd(features.contiguous(), weight3d.contiguous(), nbr, n_out)

    @staticmethod
    def backward(ctx, grad_out):
        features, weight3d, nbr, nbr_t = ctx.saved_tensors
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_out = grad_out.contiguous().float()
        gf = None
        if need_x and (nbr_t is not None or ctx.flip):
            wt = weight3d.detach().float()
            if ctx.flip:
                wt = wt.flip(0)
            wt = wt.transpose(1, 2).contiguous()
            gf = F.sparse_conv_fwd(grad_out, wt, nbr if ctx.flip else nbr_t, features.shape[0])
            need_x = False
        gf2, gw = F.sparse_conv_bwd(features, weight3d, grad_out, nbr, ctx.n_out, need_input_grad=need_x,
                                    need_weight_grad=need_w)
        gf = gf if gf is not None else gf2
        if gf is not None:
            gf = gf.to(features.dtype)
        if gw is not None:
            gw = gw.to(weight3d.dtype)
        return gf, gw, None, None, None, None


indice_conv = SparseConvFunction.apply
