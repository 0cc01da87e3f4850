"""TEST INFRASTRUCTURE ONLY: the three ops as CPU autograd Functions backed by the C oracle (oracle/d2t_oracle.c).

Used to run the reference's UNMODIFIED callers (rfcn.py, correlation_tracker.py) in a container without a GPU
(tools/make_golden_models.py, tests/test_models.py).  Same constructor / forward signatures as the reference modules
(models/{pointwise_correlation,roipool,ps_roipool}/*.py).  The product never imports this.
"""
import numpy as np
import torch
from torch import nn
from torch.autograd import Function

import oracle


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy())


class _Corr(Function):
    @staticmethod
    def forward(ctx, fm0, fm1, d_max, stride):
        ctx.save_for_backward(fm0, fm1)
        ctx.cfg = (d_max, stride)
        return torch.from_numpy(oracle.corr_fwd(_np(fm0), _np(fm1), d_max, stride))

    @staticmethod
    def backward(ctx, go):
        fm0, fm1 = ctx.saved_tensors
        g0, g1 = oracle.corr_bwd(_np(go), _np(fm0), _np(fm1), *ctx.cfg)
        return torch.from_numpy(g0), torch.from_numpy(g1), None, None


class _ROIPool(Function):
    @staticmethod
    def forward(ctx, fm, rois, r_hw):
        ctx.save_for_backward(rois)
        ctx.hw = tuple(fm.shape[-2:])
        return torch.from_numpy(oracle.roipool_fwd(_np(fm), _np(rois), r_hw))

    @staticmethod
    def backward(ctx, go):
        rois, = ctx.saved_tensors
        return torch.from_numpy(oracle.roipool_bwd(_np(go), _np(rois), *ctx.hw)), None, None


class _PSROIPool(Function):
    @staticmethod
    def forward(ctx, fm, rois, n_targets, r_hw):
        ctx.save_for_backward(rois)
        ctx.hw = tuple(fm.shape[-2:])
        if fm.size(0) != n_targets * r_hw ** 2:
            raise ValueError("channel mismatch")
        return torch.from_numpy(oracle.psroipool_fwd(_np(fm), _np(rois), n_targets, r_hw))

    @staticmethod
    def backward(ctx, go):
        rois, = ctx.saved_tensors
        return torch.from_numpy(oracle.psroipool_bwd(_np(go), _np(rois), *ctx.hw)), None, None, None


class PointwiseCorrelation(nn.Module):
    def __init__(self, d_max, stride):
        super().__init__()
        self.d_max, self.stride = d_max, stride

    def forward(self, fm0, fm1):
        return _Corr.apply(fm0, fm1, self.d_max, self.stride)


class ROIPool(nn.Module):
    def __init__(self, r_hw):
        super().__init__()
        self.r_hw = r_hw

    def forward(self, fm, rois):
        return _ROIPool.apply(fm, rois, self.r_hw)


class PSROIPool(nn.Module):
    def __init__(self, n_targets, r_hw):
        super().__init__()
        self.n_targets, self.r_hw = n_targets, r_hw

    def forward(self, fm, rois):
        return _PSROIPool.apply(fm, rois, self.n_targets, self.r_hw)
