"""Where does the end-to-end step go?  (a) H2D of one step's pinned inputs alone, (b) the eager module step with the
inputs already on the device, (c) the bench's e2e step (uploads overlapped with compute)."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import bench  # noqa: E402
import cases  # noqa: E402
import detect_to_track_b200 as d2t  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
H, W, K, R, D = bench.H, bench.W, bench.K, bench.R, bench.D
g = torch.Generator(device="cpu").manual_seed(1)
pin = lambda *s: torch.randn(*s, generator=g).pin_memory()
pairs = []
for pr in range(bench.PAIRS_PER_GPU):
    pairs.append({
        "c3": [pin(512, 2 * H, 2 * W).relu_() for _ in range(2)], "c4": [pin(1024, H, W).relu_() for _ in range(2)],
        "c5": [pin(2048, H, W).relu_() for _ in range(2)], "reg": [pin(bench.REG_CH, H, W) for _ in range(2)],
        "cls_map": [pin(bench.N_CLS * K * K, H, W) for _ in range(2)], "reg_map": [pin(bench.N_REG * K * K, H, W) for _ in range(2)],
        "rois": [torch.from_numpy(cases.rois_random(R, 2000 + 2 * pr + f)).pin_memory() for f in range(2)]})
nbytes = sum(t.numel() * t.element_size() for it in pairs for v in it.values() for t in v)
tracker = d2t.CorrelationTracker(D, K, bench.REG_CH).to(dev)
cls_pool, reg_pool = d2t.PSROIPool(bench.N_CLS, K), d2t.PSROIPool(bench.N_REG, K)


def compute(d):
    for k in ("c3", "c4", "c5", "reg", "cls_map", "reg_map"):
        for t in d[k]:
            t.requires_grad_(True)
            t.grad = None
    p0 = {"c3": d["c3"][0], "c4": d["c4"][0], "c5": d["c5"][0]}
    p1 = {"c3": d["c3"][1], "c4": d["c4"][1], "c5": d["c5"][1]}
    loss = tracker(p0, p1, d["reg"][0], d["reg"][1], d["rois"][0]).square().mean()
    for f in range(2):
        loss = loss + cls_pool(d["cls_map"][f], d["rois"][f]).mean(-1).mean(-1).square().mean()
        loss = loss + reg_pool(d["reg_map"][f], d["rois"][f]).mean(-1).mean(-1).square().mean()
    loss.backward()
    return loss.detach()


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


dst = [{k: [torch.empty_like(t, device=dev) for t in v] for k, v in it.items()} for it in pairs[:2]]


def h2d_only():
    for n, it in enumerate(pairs):
        for k, v in it.items():
            for i, t in enumerate(v):
                dst[n & 1][k][i].copy_(t, non_blocking=True)


res = {k: [t.to(dev) for t in v] for k, v in pairs[0].items()}


def compute_only():
    for _ in pairs:
        compute(res)


def cpu_only_issue():
    t0 = time.perf_counter()
    for _ in pairs:
        compute(res)
    return (time.perf_counter() - t0) * 1e3


ms_h2d = timed(h2d_only)
ms_cmp = timed(compute_only)
torch.cuda.synchronize()
ms_issue = cpu_only_issue()
torch.cuda.synchronize()
print(f"bytes/step {nbytes/1e6:.1f} MB; H2D alone {ms_h2d:.2f} ms ({nbytes/ms_h2d/1e6:.1f} GB/s); "
      f"eager compute, inputs resident {ms_cmp:.2f} ms (host issue time {ms_issue:.2f} ms) for {len(pairs)} pairs")
