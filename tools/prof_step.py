"""Tiny driver for ncu: one eager pass over the bench step's ops (optionally only some), for launch lists.

    python tools/prof_step.py [passes] [corr,ps,track]
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
from detect_to_track_b200 import pointwise_correlation as pc, roipool as rp, ps_roipool as ps  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["corr", "ps", "track"]
dev = torch.device("cuda:0")
inp = bench.build_device_inputs(torch, dev, 1234)
H, W, K, D = bench.H, bench.W, bench.K, bench.D
for _ in range(n):
    keep = []
    if "corr" in which:
        for fm0, fm1, go in inp["corr"]:
            keep.append(pc.pointwise_correlation_forward(fm0, fm1, D, 1))
            keep.append(pc.pointwise_correlation_backward(go, fm0, fm1, D, 1))
    if "ps" in which:
        for nT, fm, rois, go in inp["ps"]:
            keep.append(ps.ps_roipool_forward_batched(fm, rois, nT, K))
            keep.append(ps.ps_roipool_backward_batched(go, rois, H, W))
    if "ps1" in which:   # single-frame calls
        for nT, fm, rois, go in inp["ps"]:
            keep.append(ps.ps_roipool_forward(fm[0], rois[0], nT, K))
            keep.append(ps.ps_roipool_backward(go[0], rois[0], H, W))
    if "track" in which:
        for fm, rois, go in inp["track"][:2]:
            keep.append(rp.roipool_forward(fm, rois, K))
            keep.append(rp.roipool_backward(go, rois, H, W))
    torch.cuda.synchronize()
print("ok")
