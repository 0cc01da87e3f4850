"""Where does the config-5 train step spend its time?  torch.profiler kernel table of one step (2 pairs), and wall time of
variants: cudnn.benchmark, channels_last backbone.   python tools/prof_train_step.py [pairs]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from detect_to_track_b200 import train_step as ts  # noqa: E402

dev = torch.device("cuda:0")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
R = 300


def build(fast=False, fused=False, batched=False):
    torch.manual_seed(1239)
    model = ts.DetectTrackModule("resnet101", 3, fused_tracker=fused, fast_backbone=fast).to(dev)
    stepm = ts.DetectTrackTrainStep(model, batch_backbone=batched)
    opt = ts.make_optimizer(stepm)
    batch = ts.synthetic_batch(pairs, 608, 1008, R, 30, seed=1239, device=dev)
    return stepm, opt, batch


def step(stepm, opt, batch):
    opt.zero_grad(set_to_none=True)
    loss, _ = stepm(batch)
    loss.backward()
    opt.step()
    return loss


def timeit(stepm, opt, batch, n=3):
    for _ in range(2):
        step(stepm, opt, batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        loss = step(stepm, opt, batch)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, float(loss.detach())


print("allow_tf32 conv:", torch.backends.cudnn.allow_tf32, " matmul:", torch.backends.cuda.matmul.allow_tf32, " benchmark:", torch.backends.cudnn.benchmark)
for fast, fused, batched in ((False, False, False), (True, False, False), (True, False, True), (True, True, True)):
    stepm, opt, batch = build(fast, fused, batched)
    ms, loss = timeit(stepm, opt, batch)
    print(f"fast_backbone={fast!s:5s} fused_tracker={fused!s:5s} batch_backbone={batched!s:5s} {ms:8.1f} ms / step of {pairs} pairs   loss {loss:.4f}")
    if fast and not fused and batched:
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
            step(stepm, opt, batch)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=70))
        torch.backends.cudnn.benchmark = True
        ms, loss = timeit(stepm, opt, batch)
        print(f"  ... with cudnn.benchmark      {ms:8.1f} ms   loss {loss:.4f}")
        torch.backends.cudnn.benchmark = False
    del stepm, opt, batch
    torch.cuda.empty_cache()
