"""numpy/ctypes face of oracle/d2t_oracle.c (TEST INFRASTRUCTURE ONLY).

The C file is the restatement of the reference's algorithm (it cites the
reference file:line it follows); this module only builds it with gcc on first
use, loads it, and marshals numpy arrays.  Nothing under detect-to-track_b200/
imports it.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "d2t_oracle.c"
_BUILD = _HERE / "_build"
_LIB = None

__all__ = [
    "build", "lib", "max_threads", "set_threads",
    "corr_fwd", "corr_bwd", "roipool_fwd", "roipool_bwd", "psroipool_fwd", "psroipool_bwd",
    "bins", "corr_live_pairs",
]


def _so_path() -> Path:
    h = hashlib.sha1(_SRC.read_bytes()).hexdigest()[:12]
    return _BUILD / f"libd2t_oracle_{h}.so"


def build(force: bool = False) -> Path:
    """Compile d2t_oracle.c (gcc -O3 -march=native -ffp-contract=off -fopenmp)."""
    so = _so_path()
    if so.exists() and not force:
        return so
    _BUILD.mkdir(exist_ok=True)
    tmp = so.with_suffix(f".tmp{os.getpid()}.so")
    cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
           "-o", str(tmp), str(_SRC), "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, so)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = _so_path()
        if not so.exists():
            build()
        try:
            _LIB = ctypes.CDLL(str(so))
        except OSError:
            build(force=True)
            _LIB = ctypes.CDLL(str(so))
        _LIB.d2t_oracle_max_threads.restype = ctypes.c_int
    return _LIB


def max_threads() -> int:
    return int(lib().d2t_oracle_max_threads())


def set_threads(n: int) -> None:
    lib().d2t_oracle_set_threads(ctypes.c_int(int(n)))


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _c(a: np.ndarray, dtype=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def corr_fwd(fm0: np.ndarray, fm1: np.ndarray, d_max: int, stride: int) -> np.ndarray:
    fm0 = _c(fm0); fm1 = _c(fm1, fm0.dtype)
    B, C, H, W = fm0.shape
    k = 2 * d_max + 1
    out = np.empty((B, H, W, k, k), dtype=fm0.dtype)
    getattr(lib(), f"d2t_oracle_corr_fwd_{_sfx(fm0.dtype)}")(
        _p(fm0), _p(fm1), _p(out), B, C, H, W, int(d_max), int(stride))
    return out


def corr_bwd(grad_out: np.ndarray, fm0: np.ndarray, fm1: np.ndarray, d_max: int, stride: int):
    fm0 = _c(fm0); fm1 = _c(fm1, fm0.dtype); grad_out = _c(grad_out, fm0.dtype)
    B, C, H, W = fm0.shape
    g0 = np.empty_like(fm0); g1 = np.empty_like(fm1)
    getattr(lib(), f"d2t_oracle_corr_bwd_{_sfx(fm0.dtype)}")(
        _p(grad_out), _p(fm0), _p(fm1), _p(g0), _p(g1), B, C, H, W, int(d_max), int(stride))
    return g0, g1


def roipool_fwd(fm: np.ndarray, rois: np.ndarray, r_hw: int) -> np.ndarray:
    fm = _c(fm); rois = _c(rois, fm.dtype)
    C, H, W = fm.shape
    R = rois.shape[0]
    out = np.empty((R, C, r_hw, r_hw), dtype=fm.dtype)
    getattr(lib(), f"d2t_oracle_roipool_fwd_{_sfx(fm.dtype)}")(
        _p(fm), _p(rois), _p(out), R, C, H, W, int(r_hw))
    return out


def roipool_bwd(grad_out: np.ndarray, rois: np.ndarray, H: int, W: int) -> np.ndarray:
    grad_out = _c(grad_out); rois = _c(rois, grad_out.dtype)
    R, C, k, _ = grad_out.shape
    gin = np.empty((C, H, W), dtype=grad_out.dtype)
    getattr(lib(), f"d2t_oracle_roipool_bwd_{_sfx(grad_out.dtype)}")(
        _p(grad_out), _p(rois), _p(gin), R, C, int(H), int(W), k)
    return gin


def psroipool_fwd(fm: np.ndarray, rois: np.ndarray, n_targets: int, r_hw: int,
                  canonical_map: bool = False) -> np.ndarray:
    fm = _c(fm); rois = _c(rois, fm.dtype)
    ch, H, W = fm.shape
    assert ch == n_targets * r_hw * r_hw
    R = rois.shape[0]
    out = np.empty((R, n_targets, r_hw, r_hw), dtype=fm.dtype)
    getattr(lib(), f"d2t_oracle_psroipool_fwd_{_sfx(fm.dtype)}")(
        _p(fm), _p(rois), _p(out), R, int(n_targets), H, W, int(r_hw), int(bool(canonical_map)))
    return out


def psroipool_bwd(grad_out: np.ndarray, rois: np.ndarray, H: int, W: int,
                  canonical_map: bool = False) -> np.ndarray:
    grad_out = _c(grad_out); rois = _c(rois, grad_out.dtype)
    R, nT, k, _ = grad_out.shape
    gin = np.empty((nT * k * k, H, W), dtype=grad_out.dtype)
    getattr(lib(), f"d2t_oracle_psroipool_bwd_{_sfx(grad_out.dtype)}")(
        _p(grad_out), _p(rois), _p(gin), R, nT, int(H), int(W), k, int(bool(canonical_map)))
    return gin


def bins(rois: np.ndarray, H: int, W: int, r_hw: int, clamp_start: bool) -> np.ndarray:
    """Integer bin edges, shape (R, r_hw, 4) = (I0, I1, J0, J1) of row-bin / column-bin b.

    clamp_start=True -> ROIPool rule (roipool_cuda.cu:38-50); False -> PSROIPool rule
    (ps_roipool_cuda.cu:42-54).
    """
    rois = _c(rois)
    R = rois.shape[0]
    e = np.empty((R, r_hw, 4), dtype=np.int32)
    getattr(lib(), f"d2t_oracle_bins_{_sfx(rois.dtype)}")(
        _p(rois), _p(e), R, int(H), int(W), int(r_hw), int(bool(clamp_start)))
    return e


def corr_live_pairs(B: int, H: int, W: int, d_max: int, stride: int) -> int:
    """P of SURVEY.md section 8: number of live (position, displacement) pairs."""
    def v(n):
        return sum(len(range(max(0, i - d_max), min(i + d_max, n), stride)) for i in range(n))
    return B * v(H) * v(W)
