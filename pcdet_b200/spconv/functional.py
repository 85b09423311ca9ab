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


class SparseConvBnReluTC(Function):
    """conv [-> BatchNorm1d (train mode) [-> ReLU]] in mixed precision on the tensor cores: bf16 activations
    and activation gradients, fp32 accumulation, fp32 parameters and parameter gradients.

      forward   y = conv(x) (tcgen05, weight image packed from the fp32 parameter) -> batch statistics, running
                statistics, x' = relu(bn(y))                                           (pcdb_bn_train_fwd)
      backward  grad_y from grad_x' (ReLU mask, BatchNorm backward: pcdb_bn_train_bwd)
                grad_W = x[nbr]^T @ grad_y on tcgen05, fp32                            (pcdb_sparse_conv_wgrad)
                grad_x = the forward kernel over the rulebook read the other way round with W^T (see SparseConvFunction)

    bn_state: None (convolution only) or (running_mean, running_var, eps, momentum[, process_group]); a process group makes it a
    SyncBatchNorm (statistics of every rank's rows, torch.nn.SyncBatchNorm's forward / backward protocol)."""

    @staticmethod
    def forward(ctx, features, weight3d, gamma, beta, nbr, n_out, nbr_t, flip, bn_state, relu):
        K, c_w, c_out = weight3d.shape
        c_in = next(c for c in (16, 32, 64) if c >= c_w)       # the tensor-core kernels take 16 / 32 / 64 input channels
        w32 = weight3d.detach().float().contiguous()
        x = features.contiguous()
        if c_in != c_w:                                         # e.g. the 4 point features of conv_input: zero channels
            w32 = torch.nn.functional.pad(w32, (0, 0, 0, c_in - c_w)).contiguous()
            x = torch.nn.functional.pad(x, (0, c_in - c_w))
        wp = F.pack_conv_weights(w32)
        y = F.sparse_conv_fwd(x, None, nbr, n_out, weight_packed=wp, weight_shape=(K, c_in, c_out),
                              relu=relu and bn_state is None)
        pg = sums = None
        if bn_state is not None:
            rm, rv, eps, momentum = bn_state[:4]
            pg = bn_state[4] if len(bn_state) > 4 else None
            res = F.bn_train_fwd(y, gamma, beta, eps, momentum, rm, rv, relu=relu, process_group=pg)
            out, stats = res[0], res[1]
            sums = res[2] if pg is not None else None
        else:
            out, stats = y, None
        ctx.pg, ctx.sums = pg, sums
        ctx.save_for_backward(x, w32, y, out, gamma, stats, nbr, nbr_t)
        ctx.n_out, ctx.flip, ctx.relu, ctx.has_bn, ctx.c_feat = n_out, flip, relu, bn_state is not None, features.shape[1]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, w32, y, out, gamma, stats, nbr, nbr_t = ctx.saved_tensors
        K, c_in, c_out = w32.shape
        g = grad_out.contiguous().to(torch.bfloat16)
        gg = gb = None
        if ctx.has_bn:
            gy, gg, gb = F.bn_train_bwd(g, out, y, gamma, stats, relu=ctx.relu, process_group=ctx.pg, fwd_sums=ctx.sums)
            if gamma is None:
                gg = gb = None
        elif ctx.relu:
            gy = g * (out > 0)
        else:
            gy = g
        gw = F.sparse_conv_wgrad(x, gy, nbr, ctx.n_out) if ctx.needs_input_grad[1] else None
        if gw is not None and ctx.c_feat != c_in:
            gw = gw[:, :ctx.c_feat].contiguous()
        gx = None
        if ctx.needs_input_grad[0]:
            nb = nbr if ctx.flip else nbr_t
            if c_out <= 64:
                wpt = F.pack_conv_weights(w32, transpose=True, flip=ctx.flip)
                gx = F.sparse_conv_fwd(gy, None, nb, x.shape[0], weight_packed=wpt, weight_shape=(K, c_out, c_in))
            elif 2 * K <= 27:
                # 128 gradient channels (conv_out): the forward kernel gathers 64-channel rows, so every offset becomes two --
                # grad_y viewed as (2n, 64), offset 2k+h reading row 2*nbr+h against the h-th half of W[k]^T
                wk = w32.flip(0) if ctx.flip else w32
                w2 = wk.view(K, c_in, 2, 64).permute(0, 2, 1, 3).reshape(2 * K, c_in, 64).contiguous()
                nb2 = torch.stack([torch.where(nb >= 0, nb * 2, nb), torch.where(nb >= 0, nb * 2 + 1, nb)], 1)
                wpt = F.pack_conv_weights(w2, transpose=True)
                gx = F.sparse_conv_fwd(gy.view(-1, 64), None, nb2.view(2 * K, -1).contiguous(), x.shape[0], weight_packed=wpt,
                                       weight_shape=(2 * K, 64, c_in))
            else:
                # ... and with more than 13 offsets the two channel halves are two convolutions whose results are added
                for h in range(c_out // 64):
                    wpt = F.pack_conv_weights(w32[:, :, 64 * h:64 * h + 64].contiguous(), transpose=True, flip=ctx.flip)
                    part = F.sparse_conv_fwd(gy[:, 64 * h:64 * h + 64].contiguous(), None, nb, x.shape[0], weight_packed=wpt,
                                             weight_shape=(K, 64, c_in))
                    gx = part if h == 0 else gx + part
            if ctx.c_feat != c_in:
                gx = gx[:, :ctx.c_feat].contiguous()
        return gx, gw, gg, gb, None, None, None, None, None, None


indice_conv = SparseConvFunction.apply
indice_conv_bn_relu_tc = SparseConvBnReluTC.apply
