"""Driver for ncu: every DEFAULT kernel of the library once (after one warm-up pass) at the bench shapes --
correlation c5 B=8 (SIMT forward + finalize, flip + tcgen05 backward), track-head ROIPool fwd/bwd, PSROIPool class head
batched over 16 frames fwd/bwd and single-frame, fused track head fwd/bwd.

    python tools/prof_all.py [passes]
    ncu --set full --clock-control none --import-source on -k regex:'corr_|roipool_|psb_|psroipool_|gemm_|th_' \
        -s <launches of the warm-up pass> -c <launches of one pass> -o gpurun_out/prof_all python tools/prof_all.py 2
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
from detect_to_track_b200 import _lib, pointwise_correlation as pc, roipool as rp, ps_roipool as ps, track_head as th  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1234)
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, H, W, D, K, R = 8, 38, 63, 8, 7, 300
fm0 = (torch.randn(B, 2048, H, W, generator=g).relu_() / 16).to(dev)
fm1 = (torch.randn(B, 2048, H, W, generator=g).relu_() / 16).to(dev)
cgo = torch.randn(B, H, W, 17, 17, generator=g).to(dev)
C = 1891
rois = torch.from_numpy(cases.rois_random(R, 1238)).to(dev)
fm = torch.randn(C, H, W, generator=g).to(dev)
go = torch.randn(R, C, K, K, generator=g).to(dev)
nT, NF = 31, 16
brois = torch.stack([torch.from_numpy(cases.rois_random(R, 1237 + f)) for f in range(NF)]).to(dev)
sfm = torch.randn(NF, nT * K * K, H, W, generator=g).to(dev)
sgo = torch.randn(NF, R, nT, K, K, generator=g).to(dev)
w = (torch.randn(4, C * K * K, generator=g) / 300).to(dev)
b = torch.zeros(4, device=dev)
tgo = torch.randn(R, 4, generator=g).to(dev)
for n in range(passes):
    l0 = _lib.launch_count()
    pc.pointwise_correlation_forward(fm0, fm1, D, 1)
    pc.pointwise_correlation_backward(cgo, fm0, fm1, D, 1)
    rp.roipool_forward(fm, rois, K)
    rp.roipool_backward(go, rois, H, W)
    ps.ps_roipool_forward_batched(sfm, brois, nT, K)
    ps.ps_roipool_backward_batched(sgo, brois, H, W)
    ps.ps_roipool_forward(sfm[0], brois[0], nT, K)
    ps.ps_roipool_backward(sgo[0], brois[0], H, W)
    th.track_head_forward(fm, rois, w, b, K)
    th.track_head_backward(tgo, fm, rois, w, K)
    torch.cuda.synchronize()
    print("pass", n, "library launches:", _lib.launch_count() - l0)
