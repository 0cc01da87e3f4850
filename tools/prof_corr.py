"""Tiny driver for ncu: a few forward+backward correlation calls at the config-3 c5 shape (C=2048, 38x63, d=8)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from detect_to_track_b200 import pointwise_correlation as pc  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1234)
fm0 = (torch.randn(B, 2048, 38, 63, generator=g).relu_() / 16).to(dev)
fm1 = (torch.randn(B, 2048, 38, 63, generator=g).relu_() / 16).to(dev)
go = torch.randn(B, 38, 63, 17, 17, generator=g).to(dev)
for _ in range(n):
    out = pc.pointwise_correlation_forward(fm0, fm1, 8, 1)
    g0, g1 = pc.pointwise_correlation_backward(go, fm0, fm1, 8, 1)
torch.cuda.synchronize()
print("ok", float(out.sum()), float(g0.sum()), float(g1.sum()))
