"""Op callers re-bound to the B200 ops: R-FCN head and correlation tracker.

Mirrors detect_to_track/models/rfcn.py (:10-84) and
detect_to_track/models/correlation_tracker.py (:13-87): same constructor
arguments, attribute names (`sm_conv`, `roi_pool`, `channel_reduce`, `cls_head`,
`reg_head`, `point_corr`, `pool`, `reg_fc`, `fc_channels`) and tensor contracts, so a
reference state_dict loads unchanged.  Only the three imports differ.  Backbone
and RPN are stock torch modules and out of this package's scope.
"""
from typing import Dict, Tuple

import torch
from torch import Tensor, nn

from .pointwise_correlation import PointwiseCorrelation, TrackFeaturesFunction
from .ps_roipool import PSROIPool, PSROIPoolVoteFunction
from .roipool import ROIPool
from .track_head import TrackHeadFunction


class _RFCNHead(nn.Module):
    """R-FCN head (rfcn.py:10-43): 1x1 score-map conv -> PSROIPool -> vote (mean over the k x k grid)."""

    def __init__(self, in_channels: int, n_targets: int, k: int, fused: bool = False) -> None:
        super().__init__()
        self.sm_conv = nn.Conv2d(in_channels, n_targets * k ** 2, kernel_size=1)
        self.roi_pool = PSROIPool(n_targets, k)
        self.n_targets = n_targets
        self.fused = fused

    def pool_and_vote(self, score_map: Tensor, regions: Tensor) -> Tensor:
        """(n_targets*k^2, H, W), (|R|, 4) -> (|R|, n_targets)   (rfcn.py:40-41)."""
        if self.fused and score_map.dtype == torch.float32:   # PSROIPool + vote as one operator
            return PSROIPoolVoteFunction.apply(score_map, regions, self.n_targets, self.roi_pool.r_hw)
        pooled = self.roi_pool(score_map, regions)
        return pooled.mean(-1).mean(-1)

    def forward(self, x: Tensor, regions: Tensor) -> Tensor:
        x = x[None, :, :, :]
        # (.contiguous() is a no-op for the reference's NCHW stack; it converts once when the convolutions above run
        # channels_last -- the op itself, like the reference's, takes contiguous maps only)
        score_map = self.sm_conv(x).squeeze(0).contiguous()
        return self.pool_and_vote(score_map, regions)


class RFCN(nn.Module):
    """R-FCN (rfcn.py:46-84).  `fused=True` (extension, default off): the heads pool and vote in one operator."""

    def __init__(self, in_channels: int, n_classes: int, k: int, fused: bool = False) -> None:
        super().__init__()
        self.channel_reduce = nn.Conv2d(in_channels, 512, kernel_size=3, dilation=6, padding=6)
        self.cls_head = _RFCNHead(512, n_classes + 1, k, fused)
        self.reg_head = _RFCNHead(512, 4, k, fused)
        self.relu = nn.ReLU(inplace=True)
        self.softmax = nn.Softmax(dim=1)

    def forward(self, x: Tensor, regions: Tensor) -> Tuple[Tensor, Tensor]:
        x = x[None, :, :, :]
        x = self.relu(self.channel_reduce(x)).squeeze(0)
        c_hat = self.softmax(self.cls_head(x, regions))
        b_hat = self.reg_head(x, regions)
        return c_hat, b_hat


class CorrelationTracker(nn.Module):
    """correlation tracker (correlation_tracker.py:13-87).  `fused=True` (extension, default off) replaces
    pool -> view -> reg_fc by the fused track head; parameters and state_dict are unchanged."""

    def __init__(self, d_max: int, r_hw: int, reg_channels: int, stride: int = 1, fused: bool = False) -> None:
        super().__init__()
        self.fused = fused
        self.point_corr = PointwiseCorrelation(d_max, stride)
        self.pool = ROIPool(r_hw)
        self.fc_channels = (3 * pow(2 * d_max + 1, 2) + 2 * reg_channels) * pow(r_hw, 2)
        self.reg_fc = nn.Linear(self.fc_channels, 4)

    def correlation_features(self, fm_pyr_0: Dict[str, Tensor], fm_pyr_1: Dict[str, Tensor]):
        """three correlation maps, each ((2d+1)^2, H, W)   (correlation_tracker.py:56-72)."""
        fm_keys = ["c3", "c4", "c5"]
        c3_0, c4_0, c5_0 = [fm_pyr_0[key][None, :, :, :] for key in fm_keys]
        c3_1, c4_1, c5_1 = [fm_pyr_1[key][None, :, :, :] for key in fm_keys]
        c3_0 = nn.functional.interpolate(c3_0, scale_factor=1 / 2).contiguous()
        c3_1 = nn.functional.interpolate(c3_1, scale_factor=1 / 2).contiguous()
        return [
            cf.squeeze(0).view(cf.size(1), cf.size(2), -1).permute(2, 0, 1)
            for cf in [self.point_corr(c3_0, c3_1), self.point_corr(c4_0, c4_1), self.point_corr(c5_0, c5_1)]
        ]

    def forward(self, fm_pyr_0, fm_pyr_1, reg_fm_0: Tensor, reg_fm_1: Tensor, rois: Tensor) -> Tensor:
        if self.fused:
            # same function, two fusions: the correlations write channel-major maps
            # straight into the concatenated buffer, and ROIPool -> view -> Linear is one operator (track_head.py)
            down = lambda t: nn.functional.interpolate(t[None, :, :, :], scale_factor=1 / 2).contiguous()
            track_feats = TrackFeaturesFunction.apply(
                reg_fm_0, reg_fm_1, down(fm_pyr_0["c3"]), down(fm_pyr_1["c3"]), fm_pyr_0["c4"][None], fm_pyr_1["c4"][None],
                fm_pyr_0["c5"][None], fm_pyr_1["c5"][None], self.point_corr.d_max, self.point_corr.stride)
            return TrackHeadFunction.apply(track_feats, rois, self.reg_fc.weight, self.reg_fc.bias, self.pool.r_hw)
        corr_feats = self.correlation_features(fm_pyr_0, fm_pyr_1)
        track_feats = torch.cat([reg_fm_0, reg_fm_1, *corr_feats])
        pooled_feats = self.pool(track_feats, rois)
        pooled_feats = pooled_feats.view(pooled_feats.size(0), self.fc_channels)
        return self.reg_fc(pooled_feats)
