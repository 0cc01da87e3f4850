"""ncu driver: correlation forward c5 (B = 8 and B = 1), default dispatch (tensor-core kernel)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from detect_to_track_b200 import pointwise_correlation as pc
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1234)
for B in (8, 1):
    fm0 = (torch.randn(B, 2048, 38, 63, generator=g).relu_() / 16).to(dev)
    fm1 = (torch.randn(B, 2048, 38, 63, generator=g).relu_() / 16).to(dev)
    for _ in range(2):
        o = pc.pointwise_correlation_forward(fm0, fm1, 8, 1)
torch.cuda.synchronize()
print("ok")
