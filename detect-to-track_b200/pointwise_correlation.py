"""pointwise correlation function and module.

Host-side mirror of the reference's
detect_to_track/models/pointwise_correlation/pointwise_correlation.py
(Function :25-67, Module :70-95): same class names, argument order, saved
tensors, output layout and `None` grads for the integer arguments.  The
kernels are in csrc/ (sm_100a), reached through the C ABI in include/d2t_b200.h.
"""
from typing import Tuple

import torch
from torch import Tensor
from torch.autograd import Function
from torch.nn import Module

from . import _lib


def pointwise_correlation_forward(FM0: Tensor, FM1: Tensor, d_max: int, stride: int) -> Tensor:
    """replaces `_ext.pointwise_correlation_forward` (pointwise_correlation.cpp:23-33)."""
    _lib.check_input(FM0, "FM0")
    _lib.check_input(FM1, "FM1")
    if FM0.shape != FM1.shape or FM0.dim() != 4:
        raise RuntimeError(f"FM0 and FM1 must both be (|B|, C, H, W); got {tuple(FM0.shape)} and {tuple(FM1.shape)}")
    if FM0.dtype != FM1.dtype or FM0.device != FM1.device:
        raise RuntimeError("FM0 and FM1 must share dtype and device")
    sfx = _lib.suffix(FM0.dtype)
    B, C, H, W = FM0.shape
    k = 2 * d_max + 1
    lib = _lib.lib()
    with torch.cuda.device(FM0.device):
        out = torch.empty((B, H, W, k, k), dtype=FM0.dtype, device=FM0.device)
        nbytes = lib.d2t_corr_fwd_workspace_bytes(B, C, H, W, d_max, stride, FM0.element_size())
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, FM0.device)
        rc = getattr(lib, f"d2t_corr_fwd_{sfx}")(
            FM0.data_ptr(), FM1.data_ptr(), out.data_ptr(), B, C, H, W, d_max, stride,
            ws_ptr, ws_n, _lib.stream_ptr(FM0.device))
        _lib.check(rc, "pointwise_correlation_forward")
    return out


def pointwise_correlation_backward(
    grad_out: Tensor, FM0: Tensor, FM1: Tensor, d_max: int, stride: int
) -> Tuple[Tensor, Tensor]:
    """replaces `_ext.pointwise_correlation_backward` (pointwise_correlation.cpp:36-48)."""
    _lib.check_input(grad_out, "gradOut")
    _lib.check_input(FM0, "FM0")
    _lib.check_input(FM1, "FM1")
    sfx = _lib.suffix(FM0.dtype)
    B, C, H, W = FM0.shape
    k = 2 * d_max + 1
    if tuple(grad_out.shape) != (B, H, W, k, k) or grad_out.dtype != FM0.dtype:
        raise RuntimeError(f"grad_out must be {(B, H, W, k, k)} {FM0.dtype}; got {tuple(grad_out.shape)} {grad_out.dtype}")
    lib = _lib.lib()
    with torch.cuda.device(FM0.device):
        grad_FM0 = torch.empty_like(FM0)
        grad_FM1 = torch.empty_like(FM1)
        nbytes = lib.d2t_corr_bwd_workspace_bytes(B, C, H, W, d_max, stride, FM0.element_size())
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, FM0.device)
        rc = getattr(lib, f"d2t_corr_bwd_{sfx}")(
            grad_out.data_ptr(), FM0.data_ptr(), FM1.data_ptr(), grad_FM0.data_ptr(), grad_FM1.data_ptr(),
            B, C, H, W, d_max, stride, ws_ptr, ws_n, _lib.stream_ptr(FM0.device))
        _lib.check(rc, "pointwise_correlation_backward")
    return grad_FM0, grad_FM1


class PointwiseCorrelationFunction(Function):
    """autograd node of the D&T cross-frame correlation layer (arXiv 1710.03958, eq. 4); same `apply` signature as
    the reference Function (pointwise_correlation.py:25-67)."""

    @staticmethod
    def forward(ctx, FM0: Tensor, FM1: Tensor, d_max: int, stride: int) -> Tensor:
        """FM0, FM1: (B, C, H, W) CUDA, contiguous, same dtype (frames t and t+tau).
        Returns (B, H, W, 2*d_max+1, 2*d_max+1): entry (ci, cj) of position (i, j) is the channel dot product of
        FM0[:, :, i, j] with FM1[:, :, i-d_max+ci, j-d_max+cj] for the displacements the reference samples
        (the last row/column is never sampled; stride phase follows the clamped start), 0 elsewhere."""
        ctx.save_for_backward(FM0, FM1)
        ctx.d_max = d_max
        ctx.stride = stride
        return pointwise_correlation_forward(FM0, FM1, d_max, stride)

    @staticmethod
    def backward(ctx, grad_out: Tensor) -> Tuple[Tensor, Tensor, None, None]:
        """grad_out (B, H, W, k, k) -> (grad_FM0, grad_FM1, None, None); deterministic (no atomics)."""
        grad_out = grad_out.contiguous()
        FM0, FM1 = ctx.saved_tensors
        grad_FM0, grad_FM1 = pointwise_correlation_backward(grad_out, FM0, FM1, ctx.d_max, ctx.stride)
        return grad_FM0, grad_FM1, None, None


class PointwiseCorrelation(Module):
    """nn.Module face of `PointwiseCorrelationFunction`; constructor `(d_max, stride)` and attributes `.d_max`,
    `.stride` as in the reference (pointwise_correlation.py:70-95)."""

    def __init__(self, d_max: int, stride: int) -> None:
        super().__init__()
        self.d_max = d_max
        self.stride = stride

    def forward(self, FM0: Tensor, FM1: Tensor) -> Tensor:
        """(B, C, H, W) x 2 -> (B, H, W, 2d+1, 2d+1); see `PointwiseCorrelationFunction.forward`."""
        return PointwiseCorrelationFunction.apply(FM0, FM1, self.d_max, self.stride)
