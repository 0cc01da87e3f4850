"""Fused track-regression head: ROIPool -> view -> Linear as one operator (csrc/track_head.cu).

Extension beside the API-parity ops.  The reference composes (correlation_tracker.py:82-85)

    pooled = self.pool(track_feats, rois)                       # (|R|, C, r_hw, r_hw): 111 MB at the D&T size
    t_hat  = self.reg_fc(pooled.view(pooled.size(0), -1))       # (|R|, 4)

`TrackHeadFunction.apply(track_feats, rois, weight, bias, r_hw)` returns the same `t_hat` and the same gradients for
`track_feats`, `weight` and `bias` (FP32 rounding apart) without ever forming `pooled`.  `CorrelationTracker(fused=True)`
uses it; the default stays the reference composition.

Batched form (extension): `FM` (N, C, H, W) and `rois` (N, |R|, 4) -- the track features of N frame pairs that share
`weight` / `bias` -- give `t_hat` (N, |R|, n_out) with ONE set of launches (the GEMMs run over all N*H*W positions;
`grad_weight` / `grad_bias` are the sums over the pairs, as autograd would accumulate them).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib
from .roipool import _check_rois


def _dims(FM: Tensor, rois: Tensor, weight: Tensor, r_hw: int):
    """-> (N, R, C, H, W, n_out, batched); N = 1 and batched = False for the reference's (C, H, W) / (|R|, 4) form"""
    _lib.check_input(FM, "FM")
    _lib.check_input(weight, "weight")
    if FM.dim() not in (3, 4):
        raise RuntimeError(f"FM must be (C, H, W) or (N, C, H, W); got {tuple(FM.shape)}")
    if FM.dtype != torch.float32:
        raise RuntimeError("the fused track head is float32 only")
    batched = FM.dim() == 4
    if batched:
        _lib.check_input(rois, "rois")
        if rois.dim() != 3 or rois.size(2) != 4 or rois.size(0) != FM.size(0) or rois.dtype != FM.dtype or rois.device != FM.device:
            raise RuntimeError(f"rois must be (N, |R|, 4) {FM.dtype} with N = {FM.size(0)} on {FM.device}; got {tuple(rois.shape)}")
    else:
        _check_rois(FM.dtype, FM.device, rois)
    C, H, W = FM.shape[-3:]
    if weight.dim() != 2 or weight.size(1) != C * r_hw * r_hw or weight.dtype != FM.dtype or weight.device != FM.device:
        raise RuntimeError(f"weight must be (n_out, {C * r_hw * r_hw}) {FM.dtype} on {FM.device}; got {tuple(weight.shape)}")
    return (FM.size(0) if batched else 1), rois.size(-2), C, H, W, weight.size(0), batched


def track_head_forward(FM: Tensor, rois: Tensor, weight: Tensor, bias: Optional[Tensor], r_hw: int) -> Tensor:
    N, R, C, H, W, n_out, batched = _dims(FM, rois, weight, r_hw)
    if bias is not None:
        _lib.check_input(bias, "bias")
        if tuple(bias.shape) != (n_out,) or bias.dtype != FM.dtype:
            raise RuntimeError(f"bias must be ({n_out},) {FM.dtype}")
    lib = _lib.lib()
    with torch.cuda.device(FM.device):
        out = torch.empty((N, R, n_out) if batched else (R, n_out), dtype=FM.dtype, device=FM.device)
        if N == 0:
            return out
        nbytes = lib.d2t_trackhead_fwd_batched_workspace_bytes(N, R, C, H, W, r_hw, n_out)
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, FM.device)
        rc = lib.d2t_trackhead_fwd_batched_f32(FM.data_ptr(), rois.data_ptr(), weight.data_ptr(),
                                               bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                               N, R, C, H, W, r_hw, n_out, ws_ptr, ws_n, _lib.stream_ptr(FM.device))
        _lib.check(rc, "track_head_forward")
    return out


def track_head_backward(grad_out: Tensor, FM: Tensor, rois: Tensor, weight: Tensor, r_hw: int,
                        need_fm: bool = True, need_weight: bool = True, need_bias: bool = True
                        ) -> Tuple[Optional[Tensor], Optional[Tensor], Optional[Tensor]]:
    N, R, C, H, W, n_out, batched = _dims(FM, rois, weight, r_hw)
    _lib.check_input(grad_out, "gradOut")
    want = (N, R, n_out) if batched else (R, n_out)
    if tuple(grad_out.shape) != want or grad_out.dtype != FM.dtype:
        raise RuntimeError(f"grad_out must be {want} {FM.dtype}; got {tuple(grad_out.shape)} {grad_out.dtype}")
    lib = _lib.lib()
    with torch.cuda.device(FM.device):
        g_fm = torch.empty_like(FM) if need_fm else None
        g_w = torch.empty_like(weight) if need_weight else None
        g_b = torch.empty((n_out,), dtype=FM.dtype, device=FM.device) if need_bias else None
        if N == 0:
            for t in (g_w, g_b):
                if t is not None:
                    t.zero_()
            return g_fm, g_w, g_b
        nbytes = lib.d2t_trackhead_bwd_batched_workspace_bytes(N, R, C, H, W, r_hw, n_out)
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, FM.device)
        ptr = lambda t: t.data_ptr() if t is not None else None
        rc = lib.d2t_trackhead_bwd_batched_f32(grad_out.data_ptr(), FM.data_ptr(), rois.data_ptr(), weight.data_ptr(),
                                               ptr(g_fm), ptr(g_w), ptr(g_b), N, R, C, H, W, r_hw, n_out, ws_ptr, ws_n,
                                               _lib.stream_ptr(FM.device))
        _lib.check(rc, "track_head_backward")
    return g_fm, g_w, g_b


class TrackHeadFunction(Function):
    """t_hat = Linear(weight, bias)(ROIPool(r_hw)(FM, rois).view(|R|, -1)), fused; no gradient for `rois`.
    FM (C, H, W) / rois (|R|, 4) as the reference's tracker calls it, or the batched (N, C, H, W) / (N, |R|, 4) form."""

    @staticmethod
    def forward(ctx, FM: Tensor, rois: Tensor, weight: Tensor, bias: Optional[Tensor], r_hw: int) -> Tensor:
        ctx.save_for_backward(FM, rois, weight)
        ctx.r_hw = r_hw
        ctx.has_bias = bias is not None
        return track_head_forward(FM, rois, weight, bias, r_hw)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        FM, rois, weight = ctx.saved_tensors
        need = ctx.needs_input_grad
        g_fm, g_w, g_b = track_head_backward(grad_out.contiguous(), FM, rois, weight, ctx.r_hw,
                                             need_fm=need[0], need_weight=need[2], need_bias=ctx.has_bias and need[3])
        return g_fm, None, g_w, g_b, None
