"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/d2t_b200.h declares, validates arguments, and the Python mirror raises the
reference's exception types.  No kernel is launched here."""
import ctypes
import subprocess

import pytest
import torch

from detect_to_track_b200 import _lib
import detect_to_track_b200 as d2t


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _lib.header_symbols()
    assert len(declared) == 52
    assert sorted(_lib.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.d2t_abi_version() == _lib.ABI_VERSION


def test_no_torch_or_python_dependency_in_the_so():
    out = subprocess.run(["ldd", str(_lib.SO_PATH)], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out and "c10" not in out


def test_bad_arguments_are_rejected_before_any_launch():
    lib = _lib.lib()
    rc = lib.d2t_corr_fwd_f32(None, None, None, 1, 1, 4, 4, 2, 0, None, 0, None)      # stride 0
    assert rc == 1 and b"stride" in lib.d2t_last_error()
    rc = lib.d2t_corr_fwd_f32(None, None, None, 1, 1, 4, 4, 2, 1, None, 0, None)      # null pointers
    assert rc == 1 and b"null" in lib.d2t_last_error()
    rc = lib.d2t_roipool_fwd_f32(None, None, None, 3, 2, 4, 4, 0, None, 0, None)      # r_hw 0
    assert rc == 1 and b"roipool_fwd" in lib.d2t_last_error()
    rc = lib.d2t_psroipool_fwd_f64(None, None, None, 3, 0, 4, 4, 7, 0, None, 0, None)  # n_targets 0
    assert rc == 1


def test_empty_problem_is_a_noop():
    lib = _lib.lib()
    assert lib.d2t_corr_fwd_f32(None, None, None, 0, 4, 8, 8, 2, 1, None, 0, None) == 0
    assert lib.d2t_roipool_fwd_f32(None, None, None, 0, 4, 8, 8, 7, None, 0, None) == 0


def test_workspace_queries_need_no_gpu():
    lib = _lib.lib()
    assert lib.d2t_psroipool_bwd_workspace_bytes(300, 31, 38, 63, 7, 4) > 300 * 49 * 32 * 4   # transposed gradients + edges
    assert lib.d2t_psroipool_bwd_workspace_bytes(300, 31, 38, 63, 7, 8) > 0       # float64: per-pixel gather kernels
    assert lib.d2t_roipool_fwd_workspace_bytes(300, 1891, 38, 63, 7, 4) == 0
    assert lib.d2t_corr_fwd_workspace_bytes(1, 8, 10, 10, 3, 2, 8) == 0


def test_cpu_tensors_raise_like_the_reference():
    """common/cpp_common.hpp:1 -> AT_ASSERTM(..., "CPU op not implemented") -> RuntimeError."""
    fm = torch.rand(1, 2, 6, 6)
    with pytest.raises(RuntimeError, match="CPU op not implemented"):
        d2t.PointwiseCorrelation(2, 1)(fm, fm)
    with pytest.raises(RuntimeError, match="CPU op not implemented"):
        d2t.ROIPool(3)(torch.rand(2, 6, 6), torch.rand(2, 4))
    with pytest.raises(RuntimeError, match="CPU op not implemented"):
        d2t.PSROIPool(2, 3)(torch.rand(18, 6, 6), torch.rand(2, 4))


def test_psroipool_channel_mismatch_is_a_value_error():
    """ps_roipool.py:44-49 raises ValueError before touching the extension."""
    with pytest.raises(ValueError, match="expected 18 feature map channels"):
        d2t.PSROIPool(2, 3)(torch.rand(17, 6, 6), torch.rand(2, 4))


def test_module_attributes_match_the_reference():
    pc = d2t.PointwiseCorrelation(8, 1)
    assert (pc.d_max, pc.stride) == (8, 1)
    assert d2t.ROIPool(7).r_hw == 7
    ps = d2t.PSROIPool(31, 7)
    assert (ps.n_targets, ps.r_hw) == (31, 7)
    assert list(pc.parameters()) == []


def test_product_does_not_import_the_oracle():
    import pathlib
    pkg = pathlib.Path(d2t.__file__).resolve().parent
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = p.read_text()
        assert "oracle" not in text.lower(), f"{p} mentions the oracle"
