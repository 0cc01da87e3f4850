"""Seeded inputs shared by the golden-vector generator (tools/make_golden.py) and the tests.

Everything is derived from numpy's PCG64 with fixed seeds, so the GPU box and this
container build bit-identical inputs.
"""
from __future__ import annotations

import numpy as np


def rois_random(R: int, seed: int, dtype=np.float32) -> np.ndarray:
    """centres U(0.1,0.9)^2, h,w ~ U(0.05,0.6)."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.1, 0.9, size=(R, 2))
    s = rng.uniform(0.05, 0.6, size=(R, 2))
    return np.concatenate([c, s], axis=1).astype(dtype)


def rois_edge_cases(H: int, W: int, dtype=np.float32) -> np.ndarray:
    """Out-of-bounds, border-crossing, pixel-aligned, degenerate and the reference tests' own RoIs."""
    r = [
        [0.5, 0.5, 0.5, 0.5],                 # tests/test_roipool.py:19
        [0.1, 0.1, 0.2, 0.3],                 # tests/test_roipool.py:19, tests/test_ps_roipool.py:22
        [0.5, 0.5, 0.1, 0.1],                 # tests/test_ps_roipool.py:22
        [0.05, 0.05, 0.3, 0.3],               # crosses top/left (ROIPool shifts, PSROIPool crops: F7)
        [0.95, 0.9, 0.3, 0.4],                # crosses bottom/right
        [0.5, 0.5, 1.0, 1.0],                 # whole map
        [0.5, 0.5, 2.0, 2.0],                 # larger than the map
        [4.0 / H + 3.5 / H, 6.0 / W + 3.5 / W, 7.0 / H, 7.0 / W],   # pixel-aligned 7x7 box at (4,6)
        [8.0 / H, 8.0 / W, 14.0 / H, 14.0 / W],                     # pixel-aligned 14x14 box at (1,1)
        [0.5, 0.5, 1.0 / H, 1.0 / W],         # one pixel: bins thinner than a pixel
        [0.3, 0.7, 0.0, 0.0],                 # zero size
        [0.25, 0.75, 0.5, 0.02],              # thin
    ]
    return np.asarray(r, dtype=dtype)


ROIS_OOB = np.asarray([[1.5, 1.5, 0.2, 0.2], [3.0, 3.0, 0.5, 0.5], [-1.0, 0.5, 0.3, 0.3]])  # tests/test_ps_roipool.py:22,37


def corr_inputs(B, C, H, W, d, seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    k = 2 * d + 1
    fm0 = rng.standard_normal((B, C, H, W)).astype(dtype)
    fm1 = rng.standard_normal((B, C, H, W)).astype(dtype)
    go = rng.standard_normal((B, H, W, k, k)).astype(dtype)
    return fm0, fm1, go


def pool_inputs(C, H, W, R_out_shape, seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    fm = rng.standard_normal((C, H, W)).astype(dtype)
    go = rng.standard_normal(R_out_shape).astype(dtype)
    return fm, go


# (name, B, C, H, W, d, stride, dtype)
GOLDEN_CORR = [
    ("corr_b2c2h10w11d3s1_f64", 2, 2, 10, 11, 3, 1, "float64"),    # the reference test grid
    ("corr_b2c2h11w10d3s2_f64", 2, 2, 11, 10, 3, 2, "float64"),
    ("corr_b1c24h12w13d4s1_f32", 1, 24, 12, 13, 4, 1, "float32"),
    ("corr_b2c40h20w21d8s1_f32", 2, 40, 20, 21, 8, 1, "float32"),
    ("corr_b1c8h9w14d2s3_f32", 1, 8, 9, 14, 2, 3, "float32"),
]
# (name, C, H, W, k, R_random, dtype)
GOLDEN_ROIPOOL = [
    ("roipool_c3h10w11k5_f64", 3, 10, 11, 5, 6, "float64"),
    ("roipool_c5h38w63k7_f32", 5, 38, 63, 7, 20, "float32"),
]
# (name, nT, H, W, k, R_random, dtype)
GOLDEN_PSROIPOOL = [
    ("psroipool_t2h10w11k6_f64", 2, 10, 11, 6, 6, "float64"),
    ("psroipool_t4h38w63k7_f32", 4, 38, 63, 7, 20, "float32"),
]


def golden_rois(H, W, R_random, seed, dtype, include_oob):
    parts = [rois_edge_cases(H, W, dtype), rois_random(R_random, seed, dtype)]
    if include_oob:
        parts.append(ROIS_OOB.astype(dtype))
    return np.concatenate(parts, axis=0)
