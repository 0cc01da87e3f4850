"""Driver for ncu: the PSROIPool backward at the R-FCN head sizes -- class head (31 targets) and box head (4 targets), 16 frames
batched and one frame.  Two passes; profile the second.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ps_launches.csv python tools/prof_ps.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
from detect_to_track_b200 import ps_roipool as ps  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
H, W, K, R, NF = 38, 63, 7, 300, 16
brois = torch.stack([torch.from_numpy(cases.rois_random(R, 1237 + f)) for f in range(NF)]).to(dev)
gos = {nT: torch.randn(NF, R, nT, K, K, generator=g).to(dev) for nT in (31, 4)}
for _ in range(2):
    for nT in (31, 4):
        ps.ps_roipool_backward_batched(gos[nT], brois, H, W)
        ps.ps_roipool_backward(gos[nT][0], brois[0], H, W)
torch.cuda.synchronize()
print("ok")
