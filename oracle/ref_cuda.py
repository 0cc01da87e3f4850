"""TEST INFRASTRUCTURE ONLY: ctypes door onto oracle/_ref/libd2t_ref_cuda.so.

That library is the reference's OWN three *_cuda.cu files, compiled unmodified
from /root/reference by oracle/Makefile against oracle/ref_shim/ (it needs a
GPU to run, so it is used on the GPU box by `-m gpu` tests, tools/make_golden.py
and bench.py's "reference kernel on B200" row).  Operates on torch CUDA tensors.
The reference launches on the legacy default stream; callers
must be on torch's default stream too.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

import torch

SO_PATH = Path(__file__).resolve().parent / "_ref" / "libd2t_ref_cuda.so"
_lib = None


def available() -> bool:
    return SO_PATH.exists()


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(SO_PATH))
    return _lib


def _sfx(t: torch.Tensor) -> str:
    return {torch.float32: "f32", torch.float64: "f64"}[t.dtype]


def _p(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def _ck(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"reference {what} failed with CUDA error {rc}")


def corr_fwd(fm0, fm1, d, stride):
    B, C, H, W = fm0.shape
    k = 2 * d + 1
    out = torch.empty((B, H, W, k, k), dtype=fm0.dtype, device=fm0.device)
    _ck(getattr(lib(), f"ref_corr_fwd_{_sfx(fm0)}")(_p(fm0), _p(fm1), _p(out), B, C, H, W, d, stride), "corr_fwd")
    return out


def corr_bwd(go, fm0, fm1, d, stride):
    B, C, H, W = fm0.shape
    g0 = torch.empty_like(fm0)
    g1 = torch.empty_like(fm1)
    _ck(getattr(lib(), f"ref_corr_bwd_{_sfx(fm0)}")(_p(go), _p(fm0), _p(fm1), _p(g0), _p(g1), B, C, H, W, d, stride),
        "corr_bwd")
    return g0, g1


def roipool_fwd(fm, rois, k):
    C, H, W = fm.shape
    R = rois.shape[0]
    out = torch.empty((R, C, k, k), dtype=fm.dtype, device=fm.device)
    _ck(getattr(lib(), f"ref_roipool_fwd_{_sfx(fm)}")(_p(fm), _p(rois), _p(out), R, C, H, W, k), "roipool_fwd")
    return out


def roipool_bwd(go, rois, H, W):
    R, C, k, _ = go.shape
    gin = torch.empty((C, H, W), dtype=go.dtype, device=go.device)
    _ck(getattr(lib(), f"ref_roipool_bwd_{_sfx(go)}")(_p(go), _p(rois), _p(gin), R, C, H, W, k), "roipool_bwd")
    return gin


def psroipool_fwd(fm, rois, nT, k):
    _, H, W = fm.shape
    R = rois.shape[0]
    out = torch.empty((R, nT, k, k), dtype=fm.dtype, device=fm.device)
    _ck(getattr(lib(), f"ref_psroipool_fwd_{_sfx(fm)}")(_p(fm), _p(rois), _p(out), R, nT, H, W, k), "psroipool_fwd")
    return out


def psroipool_bwd(go, rois, H, W):
    R, nT, k, _ = go.shape
    gin = torch.empty((nT * k * k, H, W), dtype=go.dtype, device=go.device)
    _ck(getattr(lib(), f"ref_psroipool_bwd_{_sfx(go)}")(_p(go), _p(rois), _p(gin), R, nT, H, W, k), "psroipool_bwd")
    return gin
