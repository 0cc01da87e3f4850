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
# round 2: batched PSROIPool (ordered forward, targets-on-lanes backward in both column-block widths, box-head lane split,
# row-list fallback, r_hw = 9 / 33 targets on the older kernels), pool + vote, fused track head (incl. the spill path)
from detect_to_track_b200 import track_head as th  # noqa: E402
import detect_to_track_b200 as d2t  # noqa: E402
for (N, nT, H, W, k, R) in [(16, 31, 38, 63, 7, 300), (2, 31, 38, 63, 7, 40), (3, 4, 38, 63, 7, 50), (2, 33, 12, 13, 3, 20),
                            (2, 3, 20, 21, 9, 30), (1, 5, 20, 21, 3, 700), (2, 2, 12, 300, 7, 25)]:
    rois = torch.stack([torch.from_numpy(np.concatenate([cases.rois_edge_cases(H, W), cases.rois_random(R, 11 + n),
                                                         cases.ROIS_OOB.astype(np.float32)])) for n in range(N)]).to(dev)
    Rt = rois.shape[1]
    g = torch.Generator(device="cpu").manual_seed(3)
    fm = torch.randn(N, nT * k * k, H, W, generator=g).to(dev)
    go = torch.randn(N, Rt, nT, k, k, generator=g).to(dev)
    ps.ps_roipool_forward_batched(fm, rois, nT, k)
    ps.ps_roipool_backward_batched(go, rois, H, W)
    ps.ps_roipool_backward(go[0], rois[0], H, W, True)
    fm.requires_grad_(True)
    d2t.PSROIPoolVoteFunction.apply(fm, rois, nT, k, False).sum().backward()
for (N, C, H, W, R, k, nO) in [(2, 37, 20, 21, 40, 7, 4), (8, 40, 38, 63, 30, 7, 4), (1, 130, 38, 63, 77, 7, 4)]:
    g = torch.Generator(device="cpu").manual_seed(4)
    fm = torch.randn(N, C, H, W, generator=g).to(dev)
    rois = torch.stack([torch.from_numpy(cases.rois_random(R, 21 + n)) for n in range(N)]).to(dev)
    w = (torch.randn(nO, C * k * k, generator=g) / 50).to(dev)
    b = torch.zeros(nO, device=dev)
    tgo = torch.randn(N, R, nO, generator=g).to(dev)
    th.track_head_forward(fm, rois, w, b, k)
    th.track_head_backward(tgo, fm, rois, w, k)
torch.cuda.synchronize()
print("sanitize run (round 2) ok")
