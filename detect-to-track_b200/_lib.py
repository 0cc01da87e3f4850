"""ctypes binding of libd2t_b200.so (the C ABI declared in include/d2t_b200.h).

There is NO fallback of any kind: if the shared library is missing or a symbol is
absent, importing / calling raises.  PyTorch is used only for device memory,
streams and autograd plumbing; every kernel lives in the .so.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent
SO_PATH = _PKG / "libd2t_b200.so"
HEADER = _ROOT / "include" / "d2t_b200.h"
ABI_VERSION = 2

_c_int, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/d2t_b200.h one to one
_P = _c_void_p
_CORR_FWD = [_P, _P, _P] + [_c_int] * 6 + [_P, _c_size_t, _P]
_CORR_BWD = [_P] * 5 + [_c_int] * 6 + [_P, _c_size_t, _P]
_POOL = [_P, _P, _P] + [_c_int] * 5 + [_P, _c_size_t, _P]
_PSPOOL = [_P, _P, _P] + [_c_int] * 6 + [_P, _c_size_t, _P]
_PSPOOL_B = [_P, _P, _P] + [_c_int] * 7 + [_P, _c_size_t, _P]
_WS6 = [_c_int] * 6
_WS7 = [_c_int] * 7
SIGNATURES = {
    "d2t_abi_version": (_c_int, []),
    "d2t_last_error": (ctypes.c_char_p, []),
    "d2t_launch_count": (ctypes.c_ulonglong, []),
    "d2t_corr_fwd_workspace_bytes": (_c_size_t, _WS7),
    "d2t_corr_bwd_workspace_bytes": (_c_size_t, _WS7),
    "d2t_corr_fwd_f32": (_c_int, _CORR_FWD),
    "d2t_corr_fwd_f64": (_c_int, _CORR_FWD),
    "d2t_corr_fwd_simt_workspace_bytes": (_c_size_t, _WS6),
    "d2t_corr_fwd_f32_simt": (_c_int, _CORR_FWD),
    "d2t_corr_fwd_f32_tc": (_c_int, _CORR_FWD),
    "d2t_corr_fwd_strided_f32": (_c_int, [_P, _P, _P] + [_c_int] * 6 + [ctypes.c_longlong] * 3 + [_P, _c_size_t, _P]),
    "d2t_corr_fwd_strided_f64": (_c_int, [_P, _P, _P] + [_c_int] * 6 + [ctypes.c_longlong] * 3 + [_P, _c_size_t, _P]),
    "d2t_corr_bwd_f32": (_c_int, _CORR_BWD),
    "d2t_corr_bwd_f64": (_c_int, _CORR_BWD),
    "d2t_corr_bwd_f32_simt": (_c_int, _CORR_BWD),
    "d2t_corr_bwd_tc_workspace_bytes": (_c_size_t, _WS6),
    "d2t_corr_bwd_f32_tc": (_c_int, _CORR_BWD),
    "d2t_roipool_fwd_workspace_bytes": (_c_size_t, _WS6),
    "d2t_roipool_bwd_workspace_bytes": (_c_size_t, _WS6),
    "d2t_roipool_fwd_f32": (_c_int, _POOL),
    "d2t_roipool_fwd_f64": (_c_int, _POOL),
    "d2t_roipool_fwd_f32_exact": (_c_int, _POOL),
    "d2t_roipool_bwd_f32": (_c_int, _POOL),
    "d2t_roipool_bwd_f64": (_c_int, _POOL),
    "d2t_psroipool_fwd_workspace_bytes": (_c_size_t, _WS6),
    "d2t_psroipool_bwd_workspace_bytes": (_c_size_t, _WS6),
    "d2t_psroipool_fwd_f32": (_c_int, _PSPOOL),
    "d2t_psroipool_fwd_f64": (_c_int, _PSPOOL),
    "d2t_psroipool_bwd_f32": (_c_int, _PSPOOL),
    "d2t_psroipool_bwd_f64": (_c_int, _PSPOOL),
    "d2t_psroipool_fwd_batched_workspace_bytes": (_c_size_t, _WS7),
    "d2t_psroipool_bwd_batched_workspace_bytes": (_c_size_t, _WS7),
    "d2t_psroipool_fwd_batched_f32": (_c_int, _PSPOOL_B),
    "d2t_psroipool_bwd_batched_f32": (_c_int, _PSPOOL_B),
    "d2t_psroipool_vote_supported": (_c_int, _WS6),
    "d2t_psroipool_vote_fwd_f32": (_c_int, [_P, _P, _P] + [_c_int] * 7 + [_P]),
    "d2t_psroipool_vote_bwd_workspace_bytes": (_c_size_t, _WS6),
    "d2t_psroipool_vote_bwd_f32": (_c_int, [_P, _P, _P] + [_c_int] * 7 + [_P, _c_size_t, _P]),
    "d2t_trackhead_fwd_workspace_bytes": (_c_size_t, _WS6),
    "d2t_trackhead_bwd_workspace_bytes": (_c_size_t, _WS6),
    "d2t_trackhead_fwd_f32": (_c_int, [_P] * 5 + [_c_int] * 6 + [_P, _c_size_t, _P]),
    "d2t_trackhead_bwd_f32": (_c_int, [_P] * 7 + [_c_int] * 6 + [_P, _c_size_t, _P]),
    "d2t_trackhead_fwd_batched_workspace_bytes": (_c_size_t, _WS7),
    "d2t_trackhead_bwd_batched_workspace_bytes": (_c_size_t, _WS7),
    "d2t_trackhead_fwd_batched_f32": (_c_int, [_P] * 5 + [_c_int] * 7 + [_P, _c_size_t, _P]),
    "d2t_trackhead_bwd_batched_f32": (_c_int, [_P] * 7 + [_c_int] * 7 + [_P, _c_size_t, _P]),
    "d2t_gemm_tf32x3_f32": (_c_int, [_P, _P, _P] + [_c_int] * 9 + [_P]),
    "d2t_roi_nms_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "d2t_roi_decode_filter_f32": (_c_int, [_P] * 5 + [_c_int, ctypes.c_float, _P]),
    "d2t_roi_nms_f32": (_c_int, [_P] * 5 + [_c_int] * 3 + [ctypes.c_float, _P, _c_size_t, _P]),
    "d2t_pool_bins_f32": (_c_int, [_P, _P] + [_c_int] * 5 + [_P]),
    "d2t_pool_bins_f64": (_c_int, [_P, _P] + [_c_int] * 5 + [_P]),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function name include/d2t_b200.h declares (used by the export test)."""
    text = HEADER.read_text()
    return sorted(set(re.findall(r"D2T_API\s+[\w\s\*]+?\b(d2t_\w+)\s*\(", text)))


def build(verbose: bool = False) -> Path:
    """Compile csrc/*.cu into libd2t_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(_PKG / "csrc")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building libd2t_b200.so failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout)
    return SO_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not SO_PATH.exists():
            raise RuntimeError(
                f"{SO_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        handle = ctypes.CDLL(str(SO_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        got = handle.d2t_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libd2t_b200.so ABI version {got} != expected {ABI_VERSION}")
        _lib = handle
    return _lib


def launch_count() -> int:
    """kernels launched by libd2t_b200.so so far in this process."""
    return int(lib().d2t_launch_count())


def last_error() -> str:
    msg = lib().d2t_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def suffix(dtype: torch.dtype) -> str:
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise RuntimeError(f"only float32 and float64 are implemented (got {dtype})")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def workspace(nbytes: int, device: torch.device):
    """Scratch from torch's caching allocator; returns (tensor_or_None, ptr, nbytes)."""
    if nbytes <= 0:
        return None, None, 0
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return buf, buf.data_ptr(), nbytes


def check_input(x: torch.Tensor, name: str) -> None:
    """Same conditions and exception type as the reference's CHECK_INPUT (common/cpp_common.hpp:1-3)."""
    if not x.is_cuda:
        raise RuntimeError("CPU op not implemented")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
