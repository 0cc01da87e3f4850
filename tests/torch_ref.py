"""Independent torch re-expressions of the three ops (second witness for the C oracle).

Correlation: zero-pad + unfold + einsum, times the closed-form live mask of SURVEY.md F4/F5.
Pools: dense membership matrices built from integer bin edges, then einsum -- so autograd
provides an independent backward.
"""
import torch


def corr_live_mask(n: int, d: int, stride: int) -> torch.Tensor:
    """mask[i, ci]: entry ci of position i is sampled by the reference's loop
    `for (di = max(0,i-d); di < min(i+d,n); di += stride)` (pointwise_correlation_cuda.cu:92)."""
    k = 2 * d + 1
    m = torch.zeros(n, k, dtype=torch.bool)
    for i in range(n):
        for di in range(max(0, i - d), min(i + d, n), stride):
            m[i, di - i + d] = True
    return m


def corr_fwd(fm0: torch.Tensor, fm1: torch.Tensor, d: int, stride: int) -> torch.Tensor:
    B, C, H, W = fm0.shape
    k = 2 * d + 1
    pad = torch.nn.functional.pad(fm1, (d, d, d, d))
    win = pad.unfold(2, k, 1).unfold(3, k, 1)            # (B, C, H, W, k, k): win[..,i,j,ci,cj] = fm1[.., i-d+ci, j-d+cj]
    out = torch.einsum("bchw,bchwkl->bhwkl", fm0, win)
    mask = corr_live_mask(H, d, stride)[:, None, :, None] & corr_live_mask(W, d, stride)[None, :, None, :]
    return out * mask.to(out.dtype).to(out.device)


def membership(edges0: torch.Tensor, edges1: torch.Tensor, n: int) -> torch.Tensor:
    """(R, k, n) 0/1 matrix: pixel y in [e0, e1)."""
    y = torch.arange(n)[None, None, :]
    return ((y >= edges0[:, :, None]) & (y < edges1[:, :, None]))


def roipool_fwd_from_edges(fm: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """edges: (R, k, 4) int (I0, I1, J0, J1).  out (R, C, k, k) = mean over bin; empty bin -> NaN."""
    C, H, W = fm.shape
    A = membership(edges[:, :, 0], edges[:, :, 1], H).to(fm.dtype)    # (R, k, H)
    Bm = membership(edges[:, :, 2], edges[:, :, 3], W).to(fm.dtype)   # (R, k, W)
    s = torch.einsum("rih,chw,rjw->rcij", A, fm, Bm)
    numel = A.sum(-1)[:, :, None] * Bm.sum(-1)[:, None, :]            # (R, k, k)
    return s / numel[:, None, :, :]


def psroipool_fwd_from_edges(fm: torch.Tensor, edges: torch.Tensor, nT: int, canonical: bool = False) -> torch.Tensor:
    ch, H, W = fm.shape
    R, k, _ = edges.shape
    A = membership(edges[:, :, 0], edges[:, :, 1], H).to(fm.dtype)
    Bm = membership(edges[:, :, 2], edges[:, :, 3], W).to(fm.dtype)
    t = torch.arange(nT)[:, None, None]
    i = torch.arange(k)[None, :, None]
    j = torch.arange(k)[None, None, :]
    chan = (t * k * k + i * k + j) if canonical else (t + 1) * (i * k + j)   # (nT, k, k)
    sel = fm[chan]                                                      # (nT, k, k, H, W)
    s = torch.einsum("rih,tijhw,rjw->rtij", A, sel, Bm)
    numel = A.sum(-1)[:, :, None] * Bm.sum(-1)[:, None, :]
    numel = torch.where(numel > 0, numel, torch.ones_like(numel))
    return s / numel[:, None, :, :]
