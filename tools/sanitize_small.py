"""Small end-to-end run of every kernel for compute-sanitizer (memcheck): odd sizes, ragged channel counts,
multi-tile widths, both tuned displacements, float32 + float64."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
from detect_to_track_b200 import pointwise_correlation as pc, roipool as rp, ps_roipool as ps  # noqa: E402

dev = torch.device("cuda:0")
for (B, C, H, W, d, s, dt) in [(2, 13, 19, 35, 8, 1, np.float32), (1, 21, 17, 18, 4, 1, np.float32), (1, 3, 9, 10, 3, 2, np.float64),
                               (1, 37, 38, 63, 8, 1, np.float32)]:
    fm0, fm1, go = (torch.from_numpy(a).to(dev) for a in cases.corr_inputs(B, C, H, W, d, 5, dt))
    o = pc.pointwise_correlation_forward(fm0, fm1, d, s)
    g0, g1 = pc.pointwise_correlation_backward(go, fm0, fm1, d, s)
# tensor-core backward (tcgen05 / TMEM, default from 128 channels): ragged channel counts, maps smaller than one tile
for (B, C, H, W) in [(1, 137, 38, 63), (2, 300, 9, 17), (1, 520, 20, 21)]:
    fm0, fm1, go = (torch.from_numpy(a).to(dev) for a in cases.corr_inputs(B, C, H, W, 8, 6, np.float32))
    o = pc.pointwise_correlation_forward(fm0, fm1, 8, 1)
    g0, g1 = pc.pointwise_correlation_backward(go, fm0, fm1, 8, 1)
for (C, H, W, k, dt) in [(37, 38, 63, 7, np.float32), (5, 11, 10, 6, np.float64), (18, 20, 30, 3, np.float32)]:
    rois = np.concatenate([cases.rois_edge_cases(H, W, dt), cases.rois_random(70, 3, dt), cases.ROIS_OOB.astype(dt)])
    fm, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), 9, dt)
    r = torch.from_numpy(rois).to(dev)
    rp.roipool_forward(torch.from_numpy(fm).to(dev), r, k)
    rp.roipool_backward(torch.from_numpy(go).to(dev), r, H, W)
for (nT, H, W, k, dt) in [(31, 38, 63, 7, np.float32), (2, 11, 10, 6, np.float64)]:
    rois = np.concatenate([cases.rois_edge_cases(H, W, dt), cases.rois_random(70, 3, dt), cases.ROIS_OOB.astype(dt)])
    fm, go = cases.pool_inputs(nT * k * k, H, W, (rois.shape[0], nT, k, k), 9, dt)
    r = torch.from_numpy(rois).to(dev)
    ps.ps_roipool_forward(torch.from_numpy(fm).to(dev), r, nT, k)
    ps.ps_roipool_backward(torch.from_numpy(go).to(dev), r, H, W)
torch.cuda.synchronize()
print("sanitize run ok")
