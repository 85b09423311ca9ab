"""UNetV2 (Part-A^2 sparse encoder-decoder, pcdet/models/rpn/rpn_unet.py:339-529) and SparseBasicBlock
(pcdet/models/model_utils/resnet_utils.py:17-48) assembled from the pcdet_b200.spconv modules: same attribute
names, hence the same state-dict keys, and the same inference arithmetic; driven by constructor arguments
instead of PCDet's global `cfg`.  Target assignment and losses (rpn_unet.py:13-336, training only) are not
part of the hot path and are not mirrored.  SURVEY §8(f) rank 1."""
from __future__ import annotations

from functools import partial

import torch
from torch import nn

from . import spconv


class SparseBasicBlock(spconv.SparseModule):
    """resnet_utils.py:17-48: conv3x3 - bn - relu - conv3x3 - bn - (+identity) - relu, all SubM on one indice_key."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, indice_key=None, norm_fn=None):
        super().__init__()
        self.conv1 = spconv.SubMConv3d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False,
                                       indice_key=indice_key)
        self.bn1 = norm_fn(planes)
        self.relu = nn.ReLU()
        self.conv2 = spconv.SubMConv3d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False, indice_key=indice_key)
        self.bn2 = norm_fn(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x.features
        assert x.features.dim() == 2
        if not torch.is_grad_enabled() and not self.bn1.training and not self.bn2.training:
            # inference: both BatchNorms (and the first ReLU) ride in the convolution epilogues, as SparseSequential does;
            # on the tensor-core path the shortcut and the last ReLU do too (one kernel per convolution, nothing in between)
            out = self.conv1(x, fused_bn=self.bn1, fused_relu=True)
            if self.downsample is None and self.conv2.tc_inference(identity.dtype) and identity.is_contiguous():
                return self.conv2(out, fused_bn=self.bn2, fused_relu=True, fused_residual=identity)
            out = self.conv2(out, fused_bn=self.bn2)
        else:
            out = self.conv1(x)
            out.features = self.relu(self.bn1(out.features))
            out = self.conv2(out)
            out.features = self.bn2(out.features)
        if self.downsample is not None:
            identity = self.downsample(x)
        out.features = self.relu(out.features + identity)
        return out


class UNetV2(nn.Module):
    def __init__(self, input_channels: int = 4, last_pad=0):
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        block = partial(self.post_act_block, norm_fn=norm_fn)
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key="subm1"), norm_fn(16), nn.ReLU())
        self.conv1 = spconv.SparseSequential(block(16, 16, 3, padding=1, indice_key="subm1"))
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, stride=2, padding=1, indice_key="spconv2", conv_type="spconv"),
            block(32, 32, 3, padding=1, indice_key="subm2"), block(32, 32, 3, padding=1, indice_key="subm2"))
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, stride=2, padding=1, indice_key="spconv3", conv_type="spconv"),
            block(64, 64, 3, padding=1, indice_key="subm3"), block(64, 64, 3, padding=1, indice_key="subm3"))
        self.conv4 = spconv.SparseSequential(
            block(64, 64, 3, stride=2, padding=(0, 1, 1), indice_key="spconv4", conv_type="spconv"),
            block(64, 64, 3, padding=1, indice_key="subm4"), block(64, 64, 3, padding=1, indice_key="subm4"))
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(64, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key="spconv_down2"), norm_fn(128), nn.ReLU())
        # decoder (rpn_unet.py:395-418)
        self.conv_up_t4 = SparseBasicBlock(64, 64, indice_key="subm4", norm_fn=norm_fn)
        self.conv_up_m4 = block(128, 64, 3, padding=1, indice_key="subm4")
        self.inv_conv4 = block(64, 64, 3, indice_key="spconv4", conv_type="inverseconv")
        self.conv_up_t3 = SparseBasicBlock(64, 64, indice_key="subm3", norm_fn=norm_fn)
        self.conv_up_m3 = block(128, 64, 3, padding=1, indice_key="subm3")
        self.inv_conv3 = block(64, 32, 3, indice_key="spconv3", conv_type="inverseconv")
        self.conv_up_t2 = SparseBasicBlock(32, 32, indice_key="subm2", norm_fn=norm_fn)
        self.conv_up_m2 = block(64, 32, 3, indice_key="subm2")
        self.inv_conv2 = block(32, 16, 3, indice_key="spconv2", conv_type="inverseconv")
        self.conv_up_t1 = SparseBasicBlock(16, 16, indice_key="subm1", norm_fn=norm_fn)
        self.conv_up_m1 = block(32, 16, 3, indice_key="subm1")
        self.conv5 = spconv.SparseSequential(block(16, 16, 3, padding=1, indice_key="subm1"))
        self.seg_cls_layer = nn.Linear(16, 1, bias=True)
        self.seg_reg_layer = nn.Linear(16, 3, bias=True)

    @staticmethod
    def post_act_block(in_channels, out_channels, kernel_size, indice_key, stride=1, padding=0, conv_type="subm", norm_fn=None):
        if conv_type == "subm":
            conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
        elif conv_type == "spconv":
            conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False,
                                       indice_key=indice_key)
        elif conv_type == "inverseconv":
            conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, indice_key=indice_key, bias=False)
        else:
            raise NotImplementedError
        return spconv.SparseSequential(conv, norm_fn(out_channels), nn.ReLU())

    @staticmethod
    def channel_reduction(x, out_channels):
        """Folds the channels of x down to out_channels by summing groups of in_channels / out_channels neighbours
        (rpn_unet.py:424-437); in place, like the reference."""
        n, c = x.features.shape
        assert c % out_channels == 0 and c >= out_channels
        x.features = x.features.reshape(n, out_channels, c // out_channels).sum(dim=2)
        return x

    def _decoder_stage(self, lateral, below, transform, merge, up):
        """One decoder stage (the reference's "UR block", rpn_unet.py:414-422): the lateral encoder tensor goes through a
        residual SubM block, is appended behind the tensor coming from below, a SubM conv merges the pair, the
        channel-folded pair is added as a shortcut, and `up` (inverse conv; a SubM conv at the last stage) leaves the level."""
        t = transform(lateral)
        pair = torch.cat((below.features, t.features), dim=1)
        t.features = pair
        merged = merge(t).features
        t.features = merged + pair.reshape(pair.shape[0], merged.shape[1], -1).sum(dim=2)
        return up(t)

    # the reference's name for _decoder_stage (rpn_unet.py:414), kept for code written against it
    UR_block_forward = _decoder_stage

    def forward(self, input_sp_tensor, **kwargs):
        # encoder: keep every level for the decoder's lateral connections
        levels = []
        x = self.conv_input(input_sp_tensor)
        for stage in (self.conv1, self.conv2, self.conv3, self.conv4):
            x = stage(x)
            levels.append(x)
        dense = self.conv_out(x).dense()
        n, c, d, h, w = dense.shape
        ret = {"spatial_features": dense.view(n, c * d, h, w)}
        # decoder, coarse to fine (rpn_unet.py:470-496); the coarsest stage takes the encoder output from both sides
        stages = ((self.conv_up_t4, self.conv_up_m4, self.inv_conv4), (self.conv_up_t3, self.conv_up_m3, self.inv_conv3),
                  (self.conv_up_t2, self.conv_up_m2, self.inv_conv2), (self.conv_up_t1, self.conv_up_m1, self.conv5))
        below = levels[-1]
        for lateral, (transform, merge, up) in zip(reversed(levels), stages):
            below = self._decoder_stage(lateral, below, transform, merge, up)
        seg = below.features
        ret.update({"u_seg_preds": self.seg_cls_layer(seg), "u_reg_preds": self.seg_reg_layer(seg), "seg_features": seg})
        return ret
