"""spconv.conv: SparseConvolution and the 3-D variants PCDet instantiates.

Parameter layout, names and initialisation follow spconv v1.0 so that PCDet checkpoints
(`rpn_net.conv_input.0.weight`, shape (kz,ky,kx,Cin,Cout)) load unchanged."""
from __future__ import annotations

import math

import numpy as np
import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from .. import functional as F
from . import ops
from .functional import indice_conv, indice_conv_bn_relu_tc
from .modules import SparseModule
from .tensor import SparseConvTensor


def _ntuple(v, n):
    if isinstance(v, (list, tuple)):
        assert len(v) == n
        return [int(x) for x in v]
    return [int(v)] * n


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None):
        super().__init__()
        assert groups == 1
        assert ndim == 3, "only the 3-D variants used by PCDet are implemented"
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _ntuple(kernel_size, ndim)
        self.conv1x1 = int(np.prod(self.kernel_size)) == 1
        self.stride = _ntuple(stride, ndim)
        self.padding = _ntuple(padding, ndim)
        self.dilation = _ntuple(dilation, ndim)
        self.transposed = transposed
        self.inverse = inverse
        self.output_padding = _ntuple(output_padding, ndim)
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        assert not transposed, "transposed sparse convolution is not used by PCDet"
        self.weight = Parameter(torch.Tensor(*self.kernel_size, in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self._cache = {}
        self.reset_parameters()

    def reset_parameters(self):
        n = self.in_channels
        for k in self.kernel_size:
            n *= k
        stdv = 1.0 / math.sqrt(n)
        init.uniform_(self.weight, -stdv, stdv)
        if self.bias is not None:
            init.uniform_(self.bias, -stdv, stdv)

    # -- helpers ----------------------------------------------------------------------------------
    def _weight3d(self, dtype):
        """(K, Cin, Cout) view of the parameter; a cached cast when the features are bf16."""
        w = self.weight
        if dtype == w.dtype:
            return w.view(-1, self.in_channels, self.out_channels)
        key = ("w", dtype, w._version, w.data_ptr())
        hit = self._cache.get("w")
        if hit is None or hit[0] != key:
            hit = (key, w.detach().to(dtype).view(-1, self.in_channels, self.out_channels).contiguous())
            self._cache["w"] = hit
        return hit[1]

    def _weight_packed(self, dtype, c_in=None):
        """Cached tensor-core operand image of the weight (F.pack_conv_weights); c_in > in_channels pads with zero channels."""
        w = self.weight
        key = ("wp", dtype, w._version, w.data_ptr(), c_in)
        hit = self._cache.get("wp")
        if hit is None or hit[0] != key:
            w3 = self._weight3d(dtype).detach()
            if c_in is not None and c_in != self.in_channels:
                w3 = torch.nn.functional.pad(w3, (0, 0, 0, c_in - self.in_channels))
            hit = (key, F.pack_conv_weights(w3.contiguous()))
            self._cache["wp"] = hit
        return hit[1]

    def _weight_packed_halves(self, dtype):
        """Operand images of the two 64-channel halves of a 128-input-channel weight (see forward)."""
        w = self.weight
        key = ("wph", dtype, w._version, w.data_ptr())
        hit = self._cache.get("wph")
        if hit is None or hit[0] != key:
            w3 = self._weight3d(dtype).detach()
            hit = (key, F.pack_conv_weights(w3[:, :64].contiguous()), F.pack_conv_weights(w3[:, 64:].contiguous()))
            self._cache["wph"] = hit
        return hit[1], hit[2]

    def tc_inference(self, dtype, K=27):
        """True when forward() runs this layer on the tensor cores for features of `dtype` without gradients."""
        c_in = 16 if self.in_channels < 16 else (64 if self.in_channels == 128 else self.in_channels)
        return F.tc_eligible(dtype, c_in, self.out_channels, K)

    def _folded_bn(self, bn):
        key = ("bn", bn.weight._version if bn.weight is not None else -1,
               bn.bias._version if bn.bias is not None else -1, bn.running_mean._version,
               bn.running_var._version, bn.running_mean.data_ptr())
        hit = self._cache.get("bn")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                scale = torch.rsqrt(bn.running_var.float() + bn.eps)
                if bn.weight is not None:
                    scale = scale * bn.weight.float()
                shift = -bn.running_mean.float() * scale
                if bn.bias is not None:
                    shift = shift + bn.bias.float()
            hit = (key, scale.contiguous(), shift.contiguous())
            self._cache["bn"] = hit
        return hit[1], hit[2]

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, input, fused_bn=None, fused_relu=False, train_bn=None, fused_residual=None):
        assert isinstance(input, SparseConvTensor)
        features = input.features
        indices = input.indices
        spatial_shape = [int(s) for s in input.spatial_shape]
        batch_size = input.batch_size
        if not self.subm:
            out_spatial_shape = ops.get_conv_output_size(spatial_shape, self.kernel_size, self.stride, self.padding,
                                                         self.dilation)
        else:
            out_spatial_shape = spatial_shape

        scale = shift = None
        if fused_bn is not None:
            scale, shift = self._folded_bn(fused_bn)
        bias = self.bias.float() if self.bias is not None else None

        datas = input.find_indice_pair(self.indice_key)
        if self.conv1x1:
            # 1x1x1 kernel (spconv: a plain torch.mm on the features): the same kernels over the identity rulebook, so that
            # the fused bias / BatchNorm / ReLU epilogue and the tensor-core path apply here too
            n_rows = features.shape[0]
            nbr = torch.arange(n_rows, dtype=torch.int32, device=features.device).view(1, n_rows)
            outids, n_out, nbr_t = indices, n_rows, nbr
            out_spatial_shape = spatial_shape
            n_out_dev, depth = input.n_dev, input.depth
        elif self.inverse:
            assert datas is not None and self.indice_key is not None
            rb = datas
            assert not rb.subm and rb.nbr_inv is not None, "inverse convolution needs a strided rulebook"
            assert rb.n_out == indices.shape[0], "inverse conv input must be the output of the paired conv"
            outids, nbr, n_out = rb.indices, rb.nbr_inv, rb.n_in
            nbr_t = rb.nbr
            out_spatial_shape = rb.spatial_shape
            n_out_dev, depth = rb.n_in_dev, input.depth - 1
        elif self.indice_key is not None and datas is not None:
            rb = datas
            outids, nbr, n_out = rb.outids, rb.nbr, rb.n_out
            nbr_t = rb.nbr_inv
            n_out_dev, depth = rb.n_out_dev, input.depth + (0 if rb.subm else 1)
        else:
            rb = ops.build_rulebook(indices, batch_size, spatial_shape, self.kernel_size, self.stride, self.padding,
                                    self.dilation, self.subm, n_dev=input.n_dev, depth=input.depth)
            if self.indice_key is not None:
                input.indice_dict[self.indice_key] = rb
            if rb.overflow is not None:
                input.indice_dict.setdefault("__overflow__", []).append(rb.overflow)
            outids, nbr, n_out = rb.outids, rb.nbr, rb.n_out
            nbr_t = rb.nbr_inv
            n_out_dev, depth = rb.n_out_dev, input.depth + (0 if rb.subm else 1)

        features = features.contiguous()
        needs_grad = torch.is_grad_enabled() and (features.requires_grad or self.weight.requires_grad)
        centred = self.subm and all(k % 2 == 1 for k in self.kernel_size) and all(d == 1 for d in self.dilation)
        if nbr_t is not None and nbr_t.shape[1] < features.shape[0]:
            nbr_t = None
        if needs_grad and features.dtype == torch.bfloat16:
            # mixed-precision training on the tensor cores: bf16 activations, fp32 parameters (SparseConvBnReluTC)
            assert self.in_channels <= 64 and self.out_channels in (16, 32, 64, 128) and nbr.shape[0] <= 27, \
                "bf16 training takes c_in <= 64, c_out in {16,32,64,128}, <= 27 kernel offsets; train this layer in fp32"
            assert centred or nbr_t is not None or not features.requires_grad, \
                "bf16 training needs the rulebook's transposed map for the input gradient (even / dilated SubM: use fp32)"
            bn_state = gamma = beta = None
            if train_bn is not None:
                assert self.bias is None, "a conv bias in front of a train-mode BatchNorm is not fused; pass bias=False"
                bn = train_bn
                momentum = bn.momentum
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked.add_(1)
                    if momentum is None:
                        momentum = 1.0 / float(bn.num_batches_tracked)
                bn_state = (bn.running_mean, bn.running_var, bn.eps, 0.0 if momentum is None else momentum)
                if isinstance(bn, torch.nn.SyncBatchNorm) and torch.distributed.is_available() and torch.distributed.is_initialized():
                    pg = bn.process_group or torch.distributed.group.WORLD
                    if torch.distributed.get_world_size(pg) > 1:
                        bn_state = bn_state + (pg,)
                gamma, beta = bn.weight, bn.bias
            out_features = indice_conv_bn_relu_tc(features, self.weight.view(-1, self.in_channels, self.out_channels), gamma, beta,
                                                  nbr, n_out, None if self.subm else nbr_t, centred, bn_state,
                                                  fused_relu and (train_bn is not None or bias is None))
            if bias is not None:
                out_features = out_features + bias.to(out_features.dtype)
                if fused_relu:
                    out_features = torch.relu(out_features)
        elif needs_grad:
            assert train_bn is None
            assert features.dtype == torch.float32, "training runs in fp32, or in bf16 on the tensor cores"
            # the input gradient is a convolution over the rulebook read the other way round (see SparseConvFunction)
            out_features = indice_conv(features, self.weight.view(-1, self.in_channels, self.out_channels), nbr, n_out,
                                       None if self.subm else nbr_t, centred)
            if bias is not None:
                out_features = out_features + bias.to(out_features.dtype)
            if scale is not None:
                out_features = (out_features.float() * scale + shift).to(features.dtype)
            if fused_relu:
                out_features = torch.relu(out_features)
        else:
            K = nbr.shape[0]
            res = dict(residual=fused_residual, residual_post=True) if fused_residual is not None else {}
            assert fused_residual is None or self.tc_inference(features.dtype, K), "the shortcut epilogue needs a tensor-core layer"
            if self.in_channels == 128 and F.tc_eligible(features.dtype, 64, self.out_channels, K):
                # 128 input channels (UNetV2's merge convolutions): two launches over the channel halves, the first without
                # epilogue, its bf16 output entering the accumulator of the second (pcdb_sparse_conv_fwd_ex) -- instead of the
                # FMA-pipe kernel (measured: 320 us -> 2 x 16 us for 45k rows)
                assert fused_residual is None
                wa, wb = self._weight_packed_halves(features.dtype)
                part = F.sparse_conv_fwd(features[:, :64].contiguous(), None, nbr, n_out, n_out_dev=n_out_dev, weight_packed=wa,
                                         weight_shape=(K, 64, self.out_channels))
                out_features = F.sparse_conv_fwd(features[:, 64:].contiguous(), None, nbr, n_out, n_out_dev=n_out_dev, scale=scale,
                                                 shift=shift, bias=bias, relu=fused_relu, weight_packed=wb,
                                                 weight_shape=(K, 64, self.out_channels), residual=part)
            elif self.in_channels < 16 and F.tc_eligible(features.dtype, 16, self.out_channels, K):
                # e.g. the 4 point features of conv_input in bf16: zero channels up to the MMA's K step instead of the FMA pipe
                out_features = F.sparse_conv_fwd(torch.nn.functional.pad(features, (0, 16 - self.in_channels)), None, nbr, n_out,
                                                 n_out_dev=n_out_dev, scale=scale, shift=shift, bias=bias, relu=fused_relu,
                                                 weight_packed=self._weight_packed(features.dtype, 16),
                                                 weight_shape=(K, 16, self.out_channels), **res)
            else:
                wp = self._weight_packed(features.dtype) if F.tc_eligible(features.dtype, self.in_channels, self.out_channels, K) else None
                out_features = F.sparse_conv_fwd(features, self._weight3d(features.dtype).detach(), nbr, n_out, n_out_dev=n_out_dev,
                                                 scale=scale, shift=shift, bias=bias, relu=fused_relu, weight_packed=wp, **res)
        assert n_out_dev is None or not needs_grad, "static-shape mode (SparseConvTensor.n_dev) is an inference mode"
        out = SparseConvTensor(out_features, outids, out_spatial_shape, batch_size, n_dev=n_out_dev)
        out.depth = depth
        out.indice_dict = input.indice_dict
        out.grid = input.grid
        return out


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         indice_key=indice_key)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key=indice_key)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True, indice_key=indice_key)
