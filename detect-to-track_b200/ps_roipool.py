"""position-sensitive ROI pooling function and module.

Host-side mirror of detect_to_track/models/ps_roipool/ps_roipool.py (Function
:24-72, Module :75-99).  The channel map is the reference's
(t+1)*(i*r_hw+j) (ps_roipool_cuda.cu:58); `canonical_map=True`
is an opt-in extension selecting the textbook R-FCN map t*r_hw^2 + i*r_hw + j.
"""
from typing import Tuple

import torch
from torch import Tensor
from torch.autograd import Function
from torch.nn import Module

from . import _lib
from .roipool import _check_rois

_CANONICAL = 1  # D2T_PS_CANONICAL_MAP


def ps_roipool_forward(FM: Tensor, rois: Tensor, n_targets: int, r_hw: int, canonical_map: bool = False) -> Tensor:
    """replaces `_ext.ps_roipool_forward` (ps_roipool.cpp:23-33)."""
    _lib.check_input(FM, "FM")
    if FM.dim() != 3:
        raise RuntimeError(f"FM must be (n_targets*r_hw^2, H, W); got {tuple(FM.shape)}")
    _check_rois(FM.dtype, FM.device, rois)
    sfx = _lib.suffix(FM.dtype)
    _, H, W = FM.shape
    R = rois.size(0)
    lib = _lib.lib()
    with torch.cuda.device(FM.device):
        out = torch.empty((R, n_targets, r_hw, r_hw), dtype=FM.dtype, device=FM.device)
        nbytes = lib.d2t_psroipool_fwd_workspace_bytes(R, n_targets, H, W, r_hw, FM.element_size())
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, FM.device)
        rc = getattr(lib, f"d2t_psroipool_fwd_{sfx}")(
            FM.data_ptr(), rois.data_ptr(), out.data_ptr(), R, n_targets, H, W, r_hw,
            _CANONICAL if canonical_map else 0, ws_ptr, ws_n, _lib.stream_ptr(FM.device))
        _lib.check(rc, "ps_roipool_forward")
    return out


def ps_roipool_backward(grad_out: Tensor, rois: Tensor, fm_h: int, fm_w: int, canonical_map: bool = False) -> Tensor:
    """replaces `_ext.ps_roipool_backward` (ps_roipool.cpp:36-47): R, n_targets, r_hw come from grad_out."""
    _lib.check_input(grad_out, "gradOut")
    if grad_out.dim() != 4 or grad_out.size(2) != grad_out.size(3):
        raise RuntimeError(f"grad_out must be (|R|, n_targets, r_hw, r_hw); got {tuple(grad_out.shape)}")
    _check_rois(grad_out.dtype, grad_out.device, rois)
    sfx = _lib.suffix(grad_out.dtype)
    R, n_targets, r_hw, _ = grad_out.shape
    if rois.size(0) != R:
        raise RuntimeError(f"grad_out has {R} RoIs but rois has {rois.size(0)}")
    lib = _lib.lib()
    with torch.cuda.device(grad_out.device):
        grad_FM = torch.empty((n_targets * r_hw * r_hw, fm_h, fm_w), dtype=grad_out.dtype, device=grad_out.device)
        nbytes = lib.d2t_psroipool_bwd_workspace_bytes(R, n_targets, fm_h, fm_w, r_hw, grad_out.element_size())
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, grad_out.device)
        rc = getattr(lib, f"d2t_psroipool_bwd_{sfx}")(
            grad_out.data_ptr(), rois.data_ptr(), grad_FM.data_ptr(), R, n_targets, fm_h, fm_w, r_hw,
            _CANONICAL if canonical_map else 0, ws_ptr, ws_n, _lib.stream_ptr(grad_out.device))
        _lib.check(rc, "ps_roipool_backward")
    return grad_FM


def _check_batched(x: Tensor, rois: Tensor, what: str) -> None:
    _lib.check_input(x, what)
    _lib.check_input(rois, "rois")
    if x.dtype != torch.float32:
        raise RuntimeError("the batched PSROIPool is float32 only")
    if rois.dtype != x.dtype or rois.device != x.device:
        raise RuntimeError("rois must have the dtype and device of the feature map")
    if rois.dim() != 3 or rois.size(2) != 4 or rois.size(0) != x.size(0):
        raise RuntimeError(f"rois must be (N, |R|, 4) with N = {x.size(0)}; got {tuple(rois.shape)}")


def ps_roipool_forward_batched(FM: Tensor, rois: Tensor, n_targets: int, r_hw: int, canonical_map: bool = False) -> Tensor:
    """N frames in one set of launches: FM (N, n_targets*r_hw^2, H, W), rois (N, |R|, 4) ->
    (N, |R|, n_targets, r_hw, r_hw).  Bit-identical to N calls of `ps_roipool_forward`.  Extension: the reference
    pools one frame per call (rfcn.py:36-41)."""
    if FM.dim() != 4:
        raise RuntimeError(f"FM must be (N, n_targets*r_hw^2, H, W); got {tuple(FM.shape)}")
    _check_batched(FM, rois, "FM")
    N, C, H, W = FM.shape
    if C != n_targets * r_hw ** 2:
        raise ValueError(f"expected {n_targets * r_hw ** 2} feature map channels, recieved feature map of shape {tuple(FM.shape)}")
    R = rois.size(1)
    lib = _lib.lib()
    with torch.cuda.device(FM.device):
        out = torch.empty((N, R, n_targets, r_hw, r_hw), dtype=FM.dtype, device=FM.device)
        nbytes = lib.d2t_psroipool_fwd_batched_workspace_bytes(N, R, n_targets, H, W, r_hw, 4)
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, FM.device)
        rc = lib.d2t_psroipool_fwd_batched_f32(
            FM.data_ptr(), rois.data_ptr(), out.data_ptr(), N, R, n_targets, H, W, r_hw,
            _CANONICAL if canonical_map else 0, ws_ptr, ws_n, _lib.stream_ptr(FM.device))
        _lib.check(rc, "ps_roipool_forward_batched")
    return out


def ps_roipool_backward_batched(grad_out: Tensor, rois: Tensor, fm_h: int, fm_w: int, canonical_map: bool = False) -> Tensor:
    """grad_out (N, |R|, n_targets, r_hw, r_hw), rois (N, |R|, 4) -> grad_FM (N, n_targets*r_hw^2, H, W)."""
    if grad_out.dim() != 5 or grad_out.size(3) != grad_out.size(4):
        raise RuntimeError(f"grad_out must be (N, |R|, n_targets, r_hw, r_hw); got {tuple(grad_out.shape)}")
    _check_batched(grad_out, rois, "gradOut")
    N, R, n_targets, r_hw, _ = grad_out.shape
    if rois.size(1) != R:
        raise RuntimeError(f"grad_out has {R} RoIs per frame but rois has {rois.size(1)}")
    lib = _lib.lib()
    with torch.cuda.device(grad_out.device):
        grad_FM = torch.empty((N, n_targets * r_hw * r_hw, fm_h, fm_w), dtype=grad_out.dtype, device=grad_out.device)
        nbytes = lib.d2t_psroipool_bwd_batched_workspace_bytes(N, R, n_targets, fm_h, fm_w, r_hw, 4)
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, grad_out.device)
        rc = lib.d2t_psroipool_bwd_batched_f32(
            grad_out.data_ptr(), rois.data_ptr(), grad_FM.data_ptr(), N, R, n_targets, fm_h, fm_w, r_hw,
            _CANONICAL if canonical_map else 0, ws_ptr, ws_n, _lib.stream_ptr(grad_out.device))
        _lib.check(rc, "ps_roipool_backward_batched")
    return grad_FM


def ps_roipool_vote_supported(N: int, R: int, n_targets: int, H: int, W: int, r_hw: int) -> bool:
    return bool(_lib.lib().d2t_psroipool_vote_supported(N, R, n_targets, H, W, r_hw))


class PSROIPoolVoteFunction(Function):
    """PSROIPool followed by the R-FCN vote `pooled.mean(-1).mean(-1)` (rfcn.py:40-41) as ONE operator:
    FM (n_targets*r_hw^2, H, W) [or (N, ...)], rois (|R|, 4) [or (N, |R|, 4)] -> (|R|, n_targets) [or (N, |R|, n_targets)].
    The pooled (|R|, n_targets, r_hw, r_hw) tensor and the reductions never exist.  float32; shapes the fused kernels do
    not cover fall back to the composition of the API-parity ops (same result)."""

    @staticmethod
    def forward(ctx, FM: Tensor, rois: Tensor, n_targets: int, r_hw: int, canonical_map: bool = False) -> Tensor:
        single = FM.dim() == 3
        FMb, roisb = (FM[None], rois[None]) if single else (FM, rois)
        if FMb.dim() != 4:
            raise RuntimeError(f"FM must be (n_targets*r_hw^2, H, W) or (N, n_targets*r_hw^2, H, W); got {tuple(FM.shape)}")
        _check_batched(FMb, roisb, "FM")
        N, C, H, W = FMb.shape
        if C != n_targets * r_hw ** 2:
            raise ValueError(f"expected {n_targets * r_hw ** 2} feature map channels, recieved feature map of shape {tuple(FM.shape)}")
        R = roisb.size(1)
        ctx.save_for_backward(roisb)
        ctx.cfg = (n_targets, r_hw, H, W, canonical_map, single)
        ctx.fused = R > 0 and ps_roipool_vote_supported(N, R, n_targets, H, W, r_hw)
        if not ctx.fused:
            out = ps_roipool_forward_batched(FMb, roisb, n_targets, r_hw, canonical_map).mean(-1).mean(-1)
            return out[0] if single else out
        with torch.cuda.device(FM.device):
            out = torch.empty((N, R, n_targets), dtype=FM.dtype, device=FM.device)
            rc = _lib.lib().d2t_psroipool_vote_fwd_f32(FMb.data_ptr(), roisb.data_ptr(), out.data_ptr(), N, R, n_targets, H, W,
                                                       r_hw, _CANONICAL if canonical_map else 0, _lib.stream_ptr(FM.device))
            _lib.check(rc, "ps_roipool_vote_forward")
        return out[0] if single else out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        roisb, = ctx.saved_tensors
        n_targets, r_hw, H, W, canonical_map, single = ctx.cfg
        g = (grad_out[None] if single else grad_out).contiguous()
        N, R = roisb.shape[:2]
        if not ctx.fused:
            gp = (g / (r_hw * r_hw))[..., None, None].expand(N, R, n_targets, r_hw, r_hw).contiguous()
            grad_FM = ps_roipool_backward_batched(gp, roisb, H, W, canonical_map)
        else:
            _lib.check_input(g, "gradOut")
            with torch.cuda.device(g.device):
                grad_FM = torch.empty((N, n_targets * r_hw * r_hw, H, W), dtype=g.dtype, device=g.device)
                nbytes = _lib.lib().d2t_psroipool_vote_bwd_workspace_bytes(N, R, n_targets, H, W, r_hw)
                ws, ws_ptr, ws_n = _lib.workspace(nbytes, g.device)
                rc = _lib.lib().d2t_psroipool_vote_bwd_f32(g.data_ptr(), roisb.data_ptr(), grad_FM.data_ptr(), N, R, n_targets, H, W,
                                                           r_hw, _CANONICAL if canonical_map else 0, ws_ptr, ws_n,
                                                           _lib.stream_ptr(g.device))
                _lib.check(rc, "ps_roipool_vote_backward")
        return (grad_FM[0] if single else grad_FM), None, None, None, None


class PSROIPoolBatchedFunction(Function):
    """PSROIPoolFunction over a leading frame dimension (one set of kernel launches for all frames)."""

    @staticmethod
    def forward(ctx, FM: Tensor, rois: Tensor, n_targets: int, r_hw: int, canonical_map: bool = False) -> Tensor:
        ctx.save_for_backward(rois)
        ctx.fm_h, ctx.fm_w = FM.shape[-2:]
        ctx.canonical_map = canonical_map
        return ps_roipool_forward_batched(FM, rois, n_targets, r_hw, canonical_map)

    @staticmethod
    def backward(ctx: object, grad_out: Tensor) -> Tuple[Tensor, None, None, None, None]:
        rois, = ctx.saved_tensors
        grad_FM = ps_roipool_backward_batched(grad_out.contiguous(), rois, ctx.fm_h, ctx.fm_w, ctx.canonical_map)
        return grad_FM, None, None, None, None


class PSROIPoolFunction(Function):
    """autograd node of R-FCN's position-sensitive RoI pooling (arXiv 1605.06409) as the reference implements it
    (ps_roipool.py:24-72): same `apply` signature plus the optional trailing `canonical_map`."""

    @staticmethod
    def forward(ctx, FM: Tensor, rois: Tensor, n_targets: int, r_hw: int, canonical_map: bool = False) -> Tensor:
        """FM (n_targets*r_hw^2, H, W), rois (R, 4) fractional ijhw -> (R, n_targets, r_hw, r_hw): cell (i, j) of
        target t is the mean of channel (t+1)*(i*r_hw+j) (the reference's map, F6) over the cell's pixels; the RoI
        start is NOT clamped and an empty cell gives 0 (F7).  Raises ValueError on a channel-count mismatch."""
        ctx.save_for_backward(rois)
        _, ctx.fm_h, ctx.fm_w = FM.shape
        ctx.canonical_map = canonical_map

        expected_channels = n_targets * r_hw ** 2
        if FM.size(0) != expected_channels:
            raise ValueError(
                f"expected {expected_channels} feature map channels, "
                f"recieved feature map of shape {tuple(FM.shape)}"
            )

        return ps_roipool_forward(FM, rois, n_targets, r_hw, canonical_map)

    @staticmethod
    def backward(ctx: object, grad_out: Tensor) -> Tuple[Tensor, None, None, None, None]:
        """grad_out (R, n_targets, r_hw, r_hw) -> (grad_FM (n_targets*r_hw^2, H, W), None, ...)."""
        grad_out = grad_out.contiguous()
        rois, = ctx.saved_tensors
        grad_FM = ps_roipool_backward(grad_out, rois, ctx.fm_h, ctx.fm_w, ctx.canonical_map)
        return grad_FM, None, None, None, None


class PSROIPool(Module):
    """nn.Module face of `PSROIPoolFunction`; constructor `(n_targets, r_hw)` and attributes `.n_targets`, `.r_hw` as
    in the reference (ps_roipool.py:75-99).  `canonical_map=True` is an opt-in extension (module docstring)."""

    def __init__(self, n_targets: int, r_hw: int, canonical_map: bool = False) -> None:
        super().__init__()
        self.n_targets = n_targets
        self.r_hw = r_hw
        self.canonical_map = canonical_map

    def forward(self, FM: Tensor, rois: Tensor) -> Tensor:
        """(n_targets*r_hw^2, H, W), (R, 4) -> (R, n_targets, r_hw, r_hw); see `PSROIPoolFunction.forward`."""
        if self.canonical_map:
            return PSROIPoolFunction.apply(FM, rois, self.n_targets, self.r_hw, True)
        return PSROIPoolFunction.apply(FM, rois, self.n_targets, self.r_hw)


class PSROIPoolBatched(Module):
    """PSROIPool over a batch of frames: FM (N, n_targets*r_hw^2, H, W), rois (N, |R|, 4) ->
    (N, |R|, n_targets, r_hw, r_hw).  Same attributes as `PSROIPool`; float32 only."""

    def __init__(self, n_targets: int, r_hw: int, canonical_map: bool = False) -> None:
        super().__init__()
        self.n_targets = n_targets
        self.r_hw = r_hw
        self.canonical_map = canonical_map

    def forward(self, FM: Tensor, rois: Tensor) -> Tensor:
        if self.canonical_map:
            return PSROIPoolBatchedFunction.apply(FM, rois, self.n_targets, self.r_hw, True)
        return PSROIPoolBatchedFunction.apply(FM, rois, self.n_targets, self.r_hw)
