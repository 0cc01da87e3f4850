"""Tiny driver for ncu: the pooling ops at the BASELINE config-2 / config-4 shapes."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
from detect_to_track_b200 import roipool as rp, ps_roipool as ps  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1238)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
C, H, W, k, R = 1891, 38, 63, 7, 300
rois = torch.from_numpy(cases.rois_random(R, 1238)).to(dev)
fm = torch.randn(C, H, W, generator=g).to(dev)
go = torch.randn(R, C, k, k, generator=g).to(dev)
nT = 31
sfm = torch.randn(nT * k * k, H, W, generator=g).to(dev)
sgo = torch.randn(R, nT, k, k, generator=g).to(dev)
for _ in range(n):
    o = rp.roipool_forward(fm, rois, k)
    gi = rp.roipool_backward(go, rois, H, W)
    po = ps.ps_roipool_forward(sfm, rois, nT, k)
    pg = ps.ps_roipool_backward(sgo, rois, H, W)
torch.cuda.synchronize()
print("ok", float(torch.nan_to_num(o).sum()), float(gi.sum()), float(po.sum()), float(pg.sum()))
