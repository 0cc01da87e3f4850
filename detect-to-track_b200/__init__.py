"""detect-to-track_b200: B200-native (sm_100a) PointwiseCorrelation / ROIPool / PSROIPool.

Drop-in for the three custom ops of jfc4050/detect-to-track
(`from detect_to_track.models import PointwiseCorrelation, ROIPool, PSROIPool`):
identical Module / Function names, constructor and forward signatures, output
layouts and numerical quirks.  Kernels: csrc/*.cu behind the C ABI of
include/d2t_b200.h.  No Triton, no CPU fallback.
"""
from .pointwise_correlation import PointwiseCorrelation, PointwiseCorrelationFunction, TrackFeaturesFunction
from .roipool import ROIPool, ROIPoolFunction
from .ps_roipool import PSROIPool, PSROIPoolFunction, PSROIPoolBatched, PSROIPoolBatchedFunction, PSROIPoolVoteFunction
from .track_head import TrackHeadFunction
from .models import RFCN, CorrelationTracker

__all__ = [
    "PointwiseCorrelation", "PointwiseCorrelationFunction",
    "ROIPool", "ROIPoolFunction",
    "PSROIPool", "PSROIPoolFunction", "PSROIPoolBatched", "PSROIPoolBatchedFunction", "PSROIPoolVoteFunction",
    "TrackHeadFunction", "TrackFeaturesFunction", "RFCN", "CorrelationTracker",
]
