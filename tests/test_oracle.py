"""CPU tests (no GPU): pin the C oracle.

 - against an independent torch re-expression (forward values and, through autograd, backward);
 - against the reference's own test content: the one known-answer test
   (tests/test_ps_roipool.py:33-44) and the float64 gradient checks
   (tests/test_pointwise_correlation.py:8-22, test_roipool.py:10-27, test_ps_roipool.py:8-30),
   restated as adjoint identities / numeric Jacobians since the ops are linear in the feature maps;
 - against tests/golden/*.npz, outputs of the reference's own kernels run on a B200.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

import cases
import torch_ref
import oracle

GOLDEN = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------------------------ correlation
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("B,C,H,W,d,stride", [
    (1, 2, 10, 10, 3, 1), (2, 2, 11, 10, 3, 2), (2, 2, 10, 11, 3, 1), (1, 2, 11, 11, 3, 2),   # reference grid
    (1, 5, 7, 9, 4, 1), (1, 3, 20, 21, 8, 1), (2, 4, 9, 14, 2, 3), (1, 3, 5, 4, 8, 1), (1, 1, 1, 1, 1, 1),
])
def test_corr_fwd_matches_closed_form(B, C, H, W, d, stride, dtype):
    fm0, fm1, _ = cases.corr_inputs(B, C, H, W, d, seed=1, dtype=dtype)
    got = oracle.corr_fwd(fm0, fm1, d, stride)
    want = torch_ref.corr_fwd(torch.from_numpy(fm0).double(), torch.from_numpy(fm1).double(), d, stride).numpy()
    tol = 1e-5 if dtype == np.float32 else 1e-12
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol)
    # last row / column of every map is dead
    assert np.all(got[..., 2 * d, :] == 0) and np.all(got[..., :, 2 * d] == 0)


@pytest.mark.parametrize("B,C,H,W,d,stride", [(1, 2, 10, 10, 3, 1), (2, 2, 11, 10, 3, 2), (1, 3, 12, 9, 4, 1), (1, 2, 9, 14, 2, 3)])
def test_corr_bwd_matches_autograd(B, C, H, W, d, stride):
    fm0, fm1, go = cases.corr_inputs(B, C, H, W, d, seed=2, dtype=np.float64)
    t0 = torch.from_numpy(fm0).requires_grad_(True)
    t1 = torch.from_numpy(fm1).requires_grad_(True)
    torch_ref.corr_fwd(t0, t1, d, stride).backward(torch.from_numpy(go))
    g0, g1 = oracle.corr_bwd(go, fm0, fm1, d, stride)
    np.testing.assert_allclose(g0, t0.grad.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g1, t1.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_corr_live_pairs_matches_survey():
    # P = 115,200 for config 1 and 513,536 per pair at 38x63, d=8
    assert oracle.corr_live_pairs(2, 32, 32, 4, 1) == 115200
    assert oracle.corr_live_pairs(1, 38, 63, 8, 1) == 513536


# ------------------------------------------------------------------ bin edges
def _edges_python(rois, H, W, k, clamp_start, dtype):
    """Independent restatement of roipool_cuda.cu:38-50 / ps_roipool_cuda.cu:42-54 in numpy scalars."""
    T = dtype
    cl = lambda x: max(T(0), min(T(1), x))
    out = np.zeros((len(rois), k, 4), np.int32)
    for r, (rI, rJ, rH, rW) in enumerate(rois.astype(dtype)):
        for b in range(k):
            for col, (c, ln, n) in enumerate(((rI, rH, H), (rJ, rW, W))):
                bl = T(ln / T(k))
                start = cl(T(c - T(ln / T(2)))) if clamp_start else T(c - T(ln / T(2)))
                if dtype == np.float32:
                    centre = T(np.float64(start) + (np.float64(T(b)) + 0.5) * np.float64(bl))
                else:
                    import math
                    centre = T(math.fma(float(b) + 0.5, float(bl), float(start))) if hasattr(math, "fma") else None
                if centre is None:
                    pytest.skip("math.fma needs Python >= 3.13")
                e0 = int(np.floor(T(cl(T(centre - T(bl / T(2)))) * T(n))))
                e1 = int(np.ceil(T(cl(T(centre + T(bl / T(2)))) * T(n))))
                out[r, b, 2 * col] = e0
                out[r, b, 2 * col + 1] = e1
    return out


@pytest.mark.parametrize("clamp_start", [True, False])
def test_bins_match_python_restatement_f32(clamp_start):
    H, W, k = 38, 63, 7
    rois = np.concatenate([cases.rois_edge_cases(H, W), cases.rois_random(200, 7), cases.ROIS_OOB.astype(np.float32)])
    with np.errstate(all="ignore"):
        want = _edges_python(rois, H, W, k, clamp_start, np.float32)
    got = oracle.bins(rois, H, W, k, clamp_start)
    np.testing.assert_array_equal(got, want)


def test_bins_pixel_aligned():
    # 7x7-pixel box starting at pixel (4,6): bin b covers exactly pixel row 4+b (ROIPool and PSROIPool alike)
    H, W, k = 38, 63, 7
    roi = np.asarray([[7.5 / H, 9.5 / W, 7.0 / H, 7.0 / W]], np.float64)
    for clamp in (True, False):
        e = oracle.bins(roi, H, W, k, clamp)[0]
        assert np.all(e[:, 1] - e[:, 0] >= 1) and np.all(e[:, 3] - e[:, 2] >= 1)
        assert e[0, 0] == 4 and e[-1, 1] == 11 and e[0, 2] == 6 and e[-1, 3] == 13


# ------------------------------------------------------------------ ROIPool
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("C,H,W,k", [(2, 10, 10, 5), (2, 11, 10, 6), (3, 38, 63, 7)])
def test_roipool_fwd_matches_membership_einsum(C, H, W, k, dtype):
    rois = np.concatenate([cases.rois_edge_cases(H, W, dtype), cases.rois_random(12, 3, dtype)])
    fm, _ = cases.pool_inputs(C, H, W, (1,), 4, dtype)
    got = oracle.roipool_fwd(fm, rois, k)
    edges = torch.from_numpy(oracle.bins(rois, H, W, k, True)).long()
    want = torch_ref.roipool_fwd_from_edges(torch.from_numpy(fm).double(), edges).numpy()
    tol = 2e-5 if dtype == np.float32 else 1e-12
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol, equal_nan=True)
    assert np.isnan(got).sum() == np.isnan(want).sum()


@pytest.mark.parametrize("r_hw", [5, 6])
@pytest.mark.parametrize("fm_h", [10, 11])
@pytest.mark.parametrize("fm_w", [10, 11])
def test_roipool_gradients(r_hw, fm_h, fm_w):
    """tests/test_roipool.py:10-27 restated: analytic backward == Jacobian of the op's own forward.
    The op is linear in FM, so the Jacobian test is the adjoint identity <fwd(x), g> == <x, bwd(g)>
    plus bwd == autograd of the independent re-expression."""
    rois = np.asarray([[0.5, 0.5, 0.5, 0.5], [0.1, 0.1, 0.2, 0.3]], np.float64)      # tests/test_roipool.py:19
    rng = np.random.default_rng(5)
    fm = rng.random((2, fm_h, fm_w))
    go = rng.standard_normal((2, 2, r_hw, r_hw))
    out = oracle.roipool_fwd(fm, rois, r_hw)
    gin = oracle.roipool_bwd(go, rois, fm_h, fm_w)
    assert abs((out * go).sum() - (fm * gin).sum()) < 1e-10
    t = torch.from_numpy(fm).requires_grad_(True)
    edges = torch.from_numpy(oracle.bins(rois, fm_h, fm_w, r_hw, True)).long()
    torch_ref.roipool_fwd_from_edges(t, edges).backward(torch.from_numpy(go))
    np.testing.assert_allclose(gin, t.grad.numpy(), rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------ PSROIPool
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("canonical", [False, True])
@pytest.mark.parametrize("nT,H,W,k", [(1, 10, 10, 6), (2, 11, 10, 7), (4, 38, 63, 7)])
def test_psroipool_fwd_matches_membership_einsum(nT, H, W, k, canonical, dtype):
    rois = np.concatenate([cases.rois_edge_cases(H, W, dtype), cases.rois_random(12, 3, dtype), cases.ROIS_OOB.astype(dtype)])
    fm, _ = cases.pool_inputs(nT * k * k, H, W, (1,), 4, dtype)
    got = oracle.psroipool_fwd(fm, rois, nT, k, canonical)
    edges = torch.from_numpy(oracle.bins(rois, H, W, k, False)).long()
    want = torch_ref.psroipool_fwd_from_edges(torch.from_numpy(fm).double(), edges, nT, canonical).numpy()
    tol = 2e-5 if dtype == np.float32 else 1e-12
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol)


@pytest.mark.parametrize("n_targets", [1, 2])
@pytest.mark.parametrize("r_hw", [6, 7])
@pytest.mark.parametrize("fm_h", [10, 11])
@pytest.mark.parametrize("fm_w", [10, 11])
def test_ps_roipool_gradients(n_targets, r_hw, fm_h, fm_w):
    """tests/test_ps_roipool.py:8-30 restated (same RoIs, the last fully out of bounds)."""
    rois = np.asarray([[0.5, 0.5, 0.1, 0.1], [0.1, 0.1, 0.2, 0.3], [1.5, 1.5, 0.2, 0.2]], np.float64)
    rng = np.random.default_rng(6)
    fm = rng.random((n_targets * r_hw ** 2, fm_h, fm_w))
    go = rng.standard_normal((3, n_targets, r_hw, r_hw))
    out = oracle.psroipool_fwd(fm, rois, n_targets, r_hw)
    gin = oracle.psroipool_bwd(go, rois, fm_h, fm_w)
    assert abs((out * go).sum() - (fm * gin).sum()) < 1e-10
    t = torch.from_numpy(fm).requires_grad_(True)
    edges = torch.from_numpy(oracle.bins(rois, fm_h, fm_w, r_hw, False)).long()
    torch_ref.psroipool_fwd_from_edges(t, edges, n_targets).backward(torch.from_numpy(go))
    np.testing.assert_allclose(gin, t.grad.numpy(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("n_targets", [1, 2])
@pytest.mark.parametrize("r_hw", [6, 7])
@pytest.mark.parametrize("fm_h", [10, 11])
@pytest.mark.parametrize("fm_w", [10, 11])
def test_ps_roipool_can_handle_oob(n_targets, r_hw, fm_h, fm_w):
    """tests/test_ps_roipool.py:33-44, the reference's only known-answer test."""
    fm = np.full((n_targets * r_hw ** 2, fm_h, fm_w), 10, np.float32)
    rois = np.asarray([[3.0, 3.0, 0.5, 0.5]], np.float32)
    ans = oracle.psroipool_fwd(fm, rois, n_targets, r_hw)
    assert np.allclose(ans, np.zeros((1, n_targets, r_hw, r_hw)))


def test_psroipool_channel_map_f6():
    """with the reference map only 608 of 1519 (nT=31) / 117 of 196 (nT=4) channels are read."""
    for nT, n_live, max_ch in ((31, 608, 1488), (4, 117, 192)):
        k, H, W = 7, 6, 6
        rois = np.asarray([[0.5, 0.5, 1.0, 1.0]], np.float64)
        go = np.ones((1, nT, k, k))
        gin = oracle.psroipool_bwd(go, rois, H, W)
        live = np.nonzero(np.abs(gin).sum(axis=(1, 2)))[0]
        assert len(live) == n_live and live.max() == max_ch


# ------------------------------------------------------------------ golden vectors from the reference kernels
def _golden(name):
    p = GOLDEN / f"{name}.npz"
    if not p.exists():
        pytest.skip(f"{p.name} not generated yet (tools/make_golden.py on a GPU box)")
    return np.load(p)


@pytest.mark.parametrize("case", cases.GOLDEN_CORR, ids=lambda c: c[0])
def test_golden_corr(case):
    name, B, C, H, W, d, s, dt = case
    g = _golden(name)
    fm0, fm1, go = cases.corr_inputs(B, C, H, W, d, seed=sum(map(ord, name)), dtype=np.dtype(dt))
    np.testing.assert_array_equal(fm0, g["fm0"])          # the seeded inputs are reproducible
    tol = 1e-5 if dt == "float32" else 1e-12
    np.testing.assert_allclose(oracle.corr_fwd(fm0, fm1, d, s), g["out"], rtol=tol, atol=tol)
    g0, g1 = oracle.corr_bwd(go, fm0, fm1, d, s)
    np.testing.assert_allclose(g0, g["g0"], rtol=tol, atol=tol)
    np.testing.assert_allclose(g1, g["g1"], rtol=tol * 10, atol=tol * 10)     # reference: atomicAdd order


@pytest.mark.parametrize("case", cases.GOLDEN_ROIPOOL, ids=lambda c: c[0])
def test_golden_roipool(case):
    name, C, H, W, k, R, dt = case
    g = _golden(name)
    out = oracle.roipool_fwd(g["fm"], g["rois"], k)
    if dt == "float32":
        np.testing.assert_array_equal(out, g["out"])      # same summation order => bit-identical (NaNs included)
    else:
        np.testing.assert_allclose(out, g["out"], rtol=1e-13, atol=1e-13, equal_nan=True)
    tol = 1e-5 if dt == "float32" else 1e-12
    np.testing.assert_allclose(oracle.roipool_bwd(g["go"], g["rois"], H, W), g["gin"], rtol=tol, atol=tol, equal_nan=True)


@pytest.mark.parametrize("case", cases.GOLDEN_PSROIPOOL, ids=lambda c: c[0])
def test_golden_psroipool(case):
    name, nT, H, W, k, R, dt = case
    g = _golden(name)
    out = oracle.psroipool_fwd(g["fm"], g["rois"], nT, k)
    if dt == "float32":
        np.testing.assert_array_equal(out, g["out"])
    else:
        np.testing.assert_allclose(out, g["out"], rtol=1e-13, atol=1e-13)
    tol = 1e-5 if dt == "float32" else 1e-12
    np.testing.assert_allclose(oracle.psroipool_bwd(g["go"], g["rois"], H, W), g["gin"], rtol=tol, atol=tol)
