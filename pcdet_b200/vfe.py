"""MeanVoxelFeatureExtractor (pcdet/models/vfe/vfe_utils.py:19-34) on the sm_100a kernel."""
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
