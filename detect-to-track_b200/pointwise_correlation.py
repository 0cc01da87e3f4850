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


def pointwise_correlation_forward_channel_major(FM0: Tensor, FM1: Tensor, d_max: int, stride: int, out: Tensor) -> Tensor:
    """Tracker glue fusion (correlation_tracker.py:64-80): the correlation of ONE frame pair (B = 1) written directly as
    the channel-major ((2d+1)^2, H, W) map into `out`, a contiguous slice of the tracker's concatenated feature buffer.
    Bit-identical to `pointwise_correlation_forward(...).squeeze(0).view(H, W, -1).permute(2, 0, 1)`; no permute copy,
    no torch.cat."""
    _lib.check_input(FM0, "FM0")
    _lib.check_input(FM1, "FM1")
    _lib.check_input(out, "out")
    if FM0.shape != FM1.shape or FM0.dim() != 4 or FM0.size(0) != 1:
        raise RuntimeError(f"FM0 and FM1 must both be (1, C, H, W); got {tuple(FM0.shape)} and {tuple(FM1.shape)}")
    sfx = _lib.suffix(FM0.dtype)
    _, C, H, W = FM0.shape
    k = 2 * d_max + 1
    if tuple(out.shape) != (k * k, H, W) or out.dtype != FM0.dtype or out.device != FM0.device:
        raise RuntimeError(f"out must be a contiguous {(k * k, H, W)} {FM0.dtype} tensor on {FM0.device}")
    lib = _lib.lib()
    with torch.cuda.device(FM0.device):
        nbytes = lib.d2t_corr_fwd_workspace_bytes(1, C, H, W, d_max, stride, FM0.element_size())
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, FM0.device)
        rc = getattr(lib, f"d2t_corr_fwd_strided_{sfx}")(
            FM0.data_ptr(), FM1.data_ptr(), out.data_ptr(), 1, C, H, W, d_max, stride, 0, 1, H * W,
            ws_ptr, ws_n, _lib.stream_ptr(FM0.device))
        _lib.check(rc, "pointwise_correlation_forward_channel_major")
    return out


class TrackFeaturesFunction(Function):
    """track_feats = cat([reg_fm_0, reg_fm_1, corr(c3), corr(c4), corr(c5)]) as the reference builds it
    (correlation_tracker.py:64-80), with the three correlations written channel-major straight into their slices of
    the output: no (1, H, W, k, k) -> (k^2, H, W) permute copies and no torch.cat.  Extension used by
    `CorrelationTracker(fused=True)`; gradients flow to all eight inputs."""

    @staticmethod
    def forward(ctx, reg_fm_0: Tensor, reg_fm_1: Tensor, c3_0: Tensor, c3_1: Tensor, c4_0: Tensor, c4_1: Tensor,
                c5_0: Tensor, c5_1: Tensor, d_max: int, stride: int) -> Tensor:
        pairs = [(c3_0, c3_1), (c4_0, c4_1), (c5_0, c5_1)]
        ctx.save_for_backward(c3_0, c3_1, c4_0, c4_1, c5_0, c5_1)
        ctx.cfg = (d_max, stride, reg_fm_0.size(0), reg_fm_1.size(0))
        kk = (2 * d_max + 1) ** 2
        Cr0, Cr1 = reg_fm_0.size(0), reg_fm_1.size(0)
        H, W = reg_fm_0.shape[-2:]
        out = torch.empty((Cr0 + Cr1 + 3 * kk, H, W), dtype=reg_fm_0.dtype, device=reg_fm_0.device)
        out[:Cr0].copy_(reg_fm_0)
        out[Cr0:Cr0 + Cr1].copy_(reg_fm_1)
        base = Cr0 + Cr1
        for n, (a, b) in enumerate(pairs):
            pointwise_correlation_forward_channel_major(a.contiguous(), b.contiguous(), d_max, stride,
                                                        out[base + n * kk: base + (n + 1) * kk])
        return out

    @staticmethod
    def backward(ctx, grad: Tensor):
        d_max, stride, Cr0, Cr1 = ctx.cfg
        saved = ctx.saved_tensors
        k = 2 * d_max + 1
        kk = k * k
        H, W = grad.shape[-2:]
        grads = [grad[:Cr0], grad[Cr0:Cr0 + Cr1]]
        base = Cr0 + Cr1
        for n in range(3):
            a, b = saved[2 * n], saved[2 * n + 1]
            go = grad[base + n * kk: base + (n + 1) * kk].permute(1, 2, 0).contiguous().view(1, H, W, k, k)
            g0, g1 = pointwise_correlation_backward(go, a.contiguous(), b.contiguous(), d_max, stride)
            grads += [g0, g1]
        return (*grads, None, None)


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
