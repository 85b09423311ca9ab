"""Voxel feature extractors of pcdet/models/vfe/vfe_utils.py on the sm_100a kernels: MeanVoxelFeatureExtractor (:19-34)
and PillarFeatureNetOld2 (:118-215, PointPillars)."""
from __future__ import annotations

from torch import nn

from . import functional as F


class MeanVoxelFeatureExtractor(nn.Module):
    def __init__(self, num_point_features: int = 4, **kwargs):
        super().__init__()
        self.num_point_features = num_point_features

    def get_output_feature_dim(self):
        return self.num_point_features

    def forward(self, features, num_voxels, **kwargs):
        """features (N, P, C) zero padded; num_voxels (N) -> (N, C) mean over the real points."""
        return F.vfe_mean(features, num_voxels)


class PFNLayer(nn.Module):
    """Parameter container with the reference's names (vfe_utils.py:61-100): linear.weight (F, Cin), norm.*"""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=True):
        super().__init__()
        assert last_layer, "the fused kernel implements the single, last PFN layer PointPillars uses"
        self.units = out_channels
        self.linear = nn.Linear(in_channels, out_channels, bias=not use_norm)
        self.norm = nn.BatchNorm1d(out_channels, eps=1e-3, momentum=0.01) if use_norm else None


class PillarFeatureNetOld2(nn.Module):
    """PillarFeatureNetOld2 (pcdet/models/vfe/vfe_utils.py:118-215) in eval mode on the fused sm_100a kernel: same
    constructor, same state-dict keys (pfn_layers.0.linear.weight, pfn_layers.0.norm.*), same forward signature.
    `forward_scatter` additionally writes the BEV canvas of PointPillarsScatter in the same launch."""

    def __init__(self, num_input_features=4, use_norm=True, num_filters=(64,), with_distance=False,
                 voxel_size=(0.2, 0.2, 4), pc_range=(0, -40, -3, 70.4, 40, 1)):
        super().__init__()
        assert len(num_filters) == 1, "one PFN layer (the PointPillars configuration, tools/cfgs/pointpillar.yaml:56-62)"
        self.with_distance = with_distance
        self.num_filters = list(num_filters)
        n_in = num_input_features + 6 + (1 if with_distance else 0)
        self.pfn_layers = nn.ModuleList([PFNLayer(n_in, num_filters[0], use_norm, last_layer=True)])
        self.vx, self.vy, self.vz = (float(v) for v in voxel_size)
        self.x_offset = self.vx / 2 + pc_range[0]
        self.y_offset = self.vy / 2 + pc_range[1]
        self.z_offset = self.vz / 2 + pc_range[2]

    def get_output_feature_dim(self):
        return self.num_filters[-1]

    def _affine(self):
        pfn = self.pfn_layers[0]
        if pfn.norm is None:
            return None, pfn.linear.bias
        assert not self.training, "the fused pillar VFE folds BatchNorm: call .eval() (training uses the reference module)"
        bn = pfn.norm
        scale = bn.weight / (bn.running_var + bn.eps).sqrt()
        return scale, bn.bias - bn.running_mean * scale

    def forward(self, features, num_voxels, coords, **kwargs):
        scale, shift = self._affine()
        out, _ = F.pillar_vfe(features, num_voxels, coords, self.pfn_layers[0].linear.weight, scale, shift,
                              (self.vx, self.vy, self.vz), (self.x_offset, self.y_offset, self.z_offset), self.with_distance)
        return out

    def forward_scatter(self, features, num_voxels, coords, batch_size, output_shape, want_features=False):
        scale, shift = self._affine()
        return F.pillar_vfe(features, num_voxels, coords, self.pfn_layers[0].linear.weight, scale, shift,
                            (self.vx, self.vy, self.vz), (self.x_offset, self.y_offset, self.z_offset), self.with_distance,
                            want_features=want_features, canvas_shape=output_shape, batch_size=batch_size)
