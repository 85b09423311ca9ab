"""spconv.SparseModule / SparseSequential.

SparseSequential keeps the reference behaviour (spconv modules see the SparseConvTensor, plain
nn.Modules see `.features`) and adds one thing: in eval mode a `sparse conv -> BatchNorm1d -> ReLU`
run, which is how every block of pcdet/models/rpn/rpn_backbone.py:79-103 is written, is executed as
ONE kernel (BN folded into a per-channel scale/shift applied in the convolution epilogue)."""
from __future__ import annotations

from collections import OrderedDict

import torch
from torch import nn

from .tensor import SparseConvTensor


def is_spconv_module(module):
    return isinstance(module, SparseModule)


def is_sparse_conv(module):
    from .conv import SparseConvolution
    return isinstance(module, SparseConvolution)


class SparseModule(nn.Module):
    """Marker base class: modules deriving from it receive the SparseConvTensor itself."""
    pass


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)
        self.fuse_bn_relu = True

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for _ in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            module = mods[i]
            if is_spconv_module(module):
                # conv -> BatchNorm1d(eval) [-> ReLU] in one kernel
                if (self.fuse_bn_relu and is_sparse_conv(module) and i + 1 < len(mods)
                        and isinstance(mods[i + 1], nn.BatchNorm1d) and not mods[i + 1].training
                        and mods[i + 1].track_running_stats and not torch.is_grad_enabled()
                        and isinstance(input, SparseConvTensor)):
                    relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                    input = module(input, fused_bn=mods[i + 1], fused_relu=relu)
                    i += 3 if relu else 2
                    continue
                # conv -> BatchNorm1d(train) [-> ReLU] with bf16 features: the mixed-precision tensor-core training path
                # (batch statistics, ReLU and their backward in pcdet_b200's own kernels; spconv/functional.py)
                if (self.fuse_bn_relu and is_sparse_conv(module) and i + 1 < len(mods)
                        and type(mods[i + 1]) in (nn.BatchNorm1d, nn.SyncBatchNorm) and mods[i + 1].training and torch.is_grad_enabled()
                        and isinstance(input, SparseConvTensor) and input.features.dtype == torch.bfloat16
                        and module.bias is None and input.indices.shape[0] != 0):
                    relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                    input = module(input, train_bn=mods[i + 1], fused_relu=relu)
                    i += 3 if relu else 2
                    continue
                input = module(input)
            else:
                if isinstance(input, SparseConvTensor):
                    if input.indices.shape[0] != 0:
                        input.features = module(input.features)
                else:
                    input = module(input)
            i += 1
        return input
