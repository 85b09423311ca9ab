"""BackBone8x ("VoxelBackBone8x") assembled from the pcdet_b200.spconv modules.

Same module tree -- hence the same state-dict keys (`conv_input.0.weight`, `conv2.1.1.running_mean`,
...) and the same arithmetic -- as pcdet/models/rpn/rpn_backbone.py:7-103, but driven by a layer
table instead of PCDet's global `cfg`.  The unmodified reference file also runs on these modules
(tests/test_reference_modules.py does exactly that where /root/reference is present, and compares state-dict
layout and forward output with this class)."""
from __future__ import annotations

from functools import partial

import torch
from torch import nn

from . import spconv

# (state-dict stem of the conv, kind, c_in (None = input_channels), c_out, ksize, stride, padding, indice_key)
# rpn_backbone.py:12-51; SubMConv3d built by post_act_block keeps the default padding=0 (spconv forces k//2).
BACKBONE8X_LAYERS = [
    ("conv_input.0", "subm", None, 16, (3, 3, 3), (1, 1, 1), (1, 1, 1), "subm1"),
    ("conv1.0.0", "subm", 16, 16, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm1"),
    ("conv2.0.0", "spconv", 16, 32, (3, 3, 3), (2, 2, 2), (1, 1, 1), "spconv2"),
    ("conv2.1.0", "subm", 32, 32, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm2"),
    ("conv2.2.0", "subm", 32, 32, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm2"),
    ("conv3.0.0", "spconv", 32, 64, (3, 3, 3), (2, 2, 2), (1, 1, 1), "spconv3"),
    ("conv3.1.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm3"),
    ("conv3.2.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm3"),
    ("conv4.0.0", "spconv", 64, 64, (3, 3, 3), (2, 2, 2), (0, 1, 1), "spconv4"),
    ("conv4.1.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm4"),
    ("conv4.2.0", "subm", 64, 64, (3, 3, 3), (1, 1, 1), (0, 0, 0), "subm4"),
    ("conv_out.0", "spconv", 64, 128, (3, 1, 1), (2, 1, 1), (0, 0, 0), "spconv_down2"),
]


def _block(kind, c_in, c_out, ksize, stride, padding, key, norm_fn):
    if kind == "subm":
        conv = spconv.SubMConv3d(c_in, c_out, ksize, padding=padding, bias=False, indice_key=key)
    else:
        conv = spconv.SparseConv3d(c_in, c_out, ksize, stride=stride, padding=padding, bias=False, indice_key=key)
    return spconv.SparseSequential(conv, norm_fn(c_out), nn.ReLU())


class BackBone8x(nn.Module):
    def __init__(self, input_channels: int = 4, last_pad=0):
        """last_pad: 0 for 0.1/0.2 m voxel height, (1,0,0) otherwise (rpn_backbone.py:44)."""
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        groups = {}
        for stem, kind, c_in, c_out, ks, st, pd, key in BACKBONE8X_LAYERS:
            c_in = input_channels if c_in is None else c_in
            if stem == "conv_out.0":
                pd = last_pad if isinstance(last_pad, (tuple, list)) else (last_pad,) * 3
            groups.setdefault(stem.split(".")[0], []).append(_block(kind, c_in, c_out, ks, st, pd, key, norm_fn))
        # conv_input / conv_out are a single (conv, bn, relu) sequence; conv1..4 nest their blocks
        self.conv_input = groups["conv_input"][0]
        self.conv1 = spconv.SparseSequential(*groups["conv1"])
        self.conv2 = spconv.SparseSequential(*groups["conv2"])
        self.conv3 = spconv.SparseSequential(*groups["conv3"])
        self.conv4 = spconv.SparseSequential(*groups["conv4"])
        self.conv_out = groups["conv_out"][0]

    def forward(self, input_sp_tensor, **kwargs):
        x = self.conv_input(input_sp_tensor)
        x = self.conv1(x)
        x = self.conv2(x)
        x = self.conv3(x)
        x = self.conv4(x)
        out = self.conv_out(x)
        dense = out.dense()
        n, c, d, h, w = dense.shape
        return {"spatial_features": dense.view(n, c * d, h, w)}

    def conv_modules(self):
        """[(stem, SparseConvolution, BatchNorm1d)] in execution order."""
        res = []
        for stem, *_ in BACKBONE8X_LAYERS:
            seq = self
            parts = stem.split(".")
            for p in parts[:-1]:
                seq = getattr(seq, p) if not p.isdigit() else seq[int(p)]
            res.append((stem, seq[0], seq[1]))
        return res

    def load_numpy_weights(self, weights):
        with torch.no_grad():
            for stem, conv, _bn in self.conv_modules():
                conv.weight.copy_(torch.as_tensor(weights[stem]))
