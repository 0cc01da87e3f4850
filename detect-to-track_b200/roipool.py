"""RoI Pooling function and module.

Host-side mirror of detect_to_track/models/roipool/roipool.py (Function :22-57,
Module :60-81).  Average pooling over fractional-ijhw RoIs, one image per call
(SURVEY.md F2/F3).  No gradient with respect to `rois`, like the reference.
"""
from typing import Tuple

import torch
from torch import Tensor
from torch.autograd import Function
from torch.nn import Module

from . import _lib


def _check_rois(FM_dtype, FM_device, rois: Tensor) -> None:
    _lib.check_input(rois, "rois")
    if rois.dim() != 2 or rois.size(1) != 4:
        raise RuntimeError(f"rois must be (|R|, 4); got {tuple(rois.shape)}")
    if rois.dtype != FM_dtype:
        # the reference raises RuntimeError here too (rois.data<scalar_t>(), roipool_cuda.cu:150)
        raise RuntimeError(f"expected rois of dtype {FM_dtype}, got {rois.dtype}")
    if rois.device != FM_device:
        raise RuntimeError("FM and rois must be on the same device")


def roipool_forward(FM: Tensor, rois: Tensor, r_hw: int, exact: bool = False) -> Tensor:
    """replaces `_ext.roipool_forward` (roipool.cpp:22-31).

    exact=True (float32 only) keeps the reference's left-to-right summation order and is bit-identical to
    its kernel; the default sums bins through per-row prefix sums (rounding-level differences, much faster)."""
    _lib.check_input(FM, "FM")
    if FM.dim() != 3:
        raise RuntimeError(f"FM must be (C, H, W); got {tuple(FM.shape)}")
    _check_rois(FM.dtype, FM.device, rois)
    sfx = _lib.suffix(FM.dtype)
    C, H, W = FM.shape
    R = rois.size(0)
    lib = _lib.lib()
    with torch.cuda.device(FM.device):
        out = torch.empty((R, C, r_hw, r_hw), dtype=FM.dtype, device=FM.device)
        name = "d2t_roipool_fwd_f32_exact" if (exact and sfx == "f32") else f"d2t_roipool_fwd_{sfx}"
        rc = getattr(lib, name)(
            FM.data_ptr(), rois.data_ptr(), out.data_ptr(), R, C, H, W, r_hw, None, 0, _lib.stream_ptr(FM.device))
        _lib.check(rc, "roipool_forward")
    return out


def roipool_backward(grad_out: Tensor, rois: Tensor, i_h: int, i_w: int) -> Tensor:
    """replaces `_ext.roipool_backward` (roipool.cpp:34-45): R, C, r_hw come from grad_out."""
    _lib.check_input(grad_out, "gradOut")
    if grad_out.dim() != 4 or grad_out.size(2) != grad_out.size(3):
        raise RuntimeError(f"grad_out must be (|R|, C, r_hw, r_hw); got {tuple(grad_out.shape)}")
    _check_rois(grad_out.dtype, grad_out.device, rois)
    sfx = _lib.suffix(grad_out.dtype)
    R, C, r_hw, _ = grad_out.shape
    if rois.size(0) != R:
        raise RuntimeError(f"grad_out has {R} RoIs but rois has {rois.size(0)}")
    lib = _lib.lib()
    with torch.cuda.device(grad_out.device):
        grad_fm = torch.empty((C, i_h, i_w), dtype=grad_out.dtype, device=grad_out.device)
        rc = getattr(lib, f"d2t_roipool_bwd_{sfx}")(
            grad_out.data_ptr(), rois.data_ptr(), grad_fm.data_ptr(), R, C, i_h, i_w, r_hw, None, 0,
            _lib.stream_ptr(grad_out.device))
        _lib.check(rc, "roipool_backward")
    return grad_fm


def pool_bins(rois: Tensor, H: int, W: int, r_hw: int, clamp_start: bool) -> Tensor:
    """Integer bin edges (R, r_hw, 4) = (I0, I1, J0, J1), computed by the device code the
    pooling kernels use.  clamp_start=True: ROIPool rule; False: PSROIPool rule."""
    _lib.check_input(rois, "rois")
    sfx = _lib.suffix(rois.dtype)
    R = rois.size(0)
    with torch.cuda.device(rois.device):
        edges = torch.empty((R, r_hw, 4), dtype=torch.int32, device=rois.device)
        rc = getattr(_lib.lib(), f"d2t_pool_bins_{sfx}")(
            rois.data_ptr(), edges.data_ptr(), R, H, W, r_hw, int(bool(clamp_start)), _lib.stream_ptr(rois.device))
        _lib.check(rc, "pool_bins")
    return edges


class ROIPoolFunction(Function):
    """autograd node of the reference's AVERAGE RoI pooling (roipool.py:22-57): same `apply` signature."""

    @staticmethod
    def forward(ctx: object, FM: Tensor, rois: Tensor, r_hw: int) -> Tensor:
        """FM (C, H, W), rois (R, 4) = fractional (centre_i, centre_j, height, width) of the same dtype ->
        (R, C, r_hw, r_hw) bin means.  The RoI start is clamped to the map (a box crossing the top/left border is
        shifted, F7); an empty bin gives 0/0 = NaN like the reference."""
        ctx.i_h, ctx.i_w = FM.shape[-2:]
        ctx.save_for_backward(rois)
        return roipool_forward(FM, rois, r_hw)

    @staticmethod
    def backward(ctx: object, grad_out: Tensor) -> Tuple[Tensor, None, None]:
        """grad_out (R, C, r_hw, r_hw) -> (grad_FM, None, None); there is no gradient for `rois`."""
        grad_out = grad_out.contiguous()
        rois, = ctx.saved_tensors
        grad_fm = roipool_backward(grad_out, rois, ctx.i_h, ctx.i_w)
        return grad_fm, None, None


class ROIPool(Module):
    """nn.Module face of `ROIPoolFunction`; constructor `(r_hw)` and attribute `.r_hw` as in the reference
    (roipool.py:60-81)."""

    def __init__(self, r_hw: int) -> None:
        super().__init__()
        self.r_hw = r_hw

    def forward(self, FM: Tensor, rois: Tensor) -> Tensor:
        """(C, H, W), (R, 4) -> (R, C, r_hw, r_hw); see `ROIPoolFunction.forward`."""
        return ROIPoolFunction.apply(FM, rois, self.r_hw)
