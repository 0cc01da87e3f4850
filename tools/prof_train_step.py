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


def build(channels_last=False):
    torch.manual_seed(1239)
    model = ts.DetectTrackModule("resnet101", 3, fused_tracker=False).to(dev)
    if channels_last:
        model.backbone = model.backbone.to(memory_format=torch.channels_last)
    stepm = ts.DetectTrackTrainStep(model)
    opt = ts.make_optimizer(stepm)
    batch = ts.synthetic_batch(pairs, 608, 1008, R, 30, seed=1239, device=dev)
    if channels_last:
        for it in batch:
            it["x"] = it["x"].contiguous(memory_format=torch.channels_last)
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
    return (time.perf_counter() - t0) / n * 1e3, float(loss)


print("allow_tf32 conv:", torch.backends.cudnn.allow_tf32, " matmul:", torch.backends.cuda.matmul.allow_tf32, " benchmark:", torch.backends.cudnn.benchmark)
stepm, opt, batch = build()
ms, loss = timeit(stepm, opt, batch)
print(f"baseline                      {ms:8.1f} ms / step of {pairs} pairs   loss {loss:.4f}")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    step(stepm, opt, batch)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
torch.backends.cudnn.benchmark = True
ms, loss = timeit(stepm, opt, batch)
print(f"cudnn.benchmark               {ms:8.1f} ms   loss {loss:.4f}")
del stepm, opt, batch
torch.cuda.empty_cache()
stepm, opt, batch = build(channels_last=True)
ms, loss = timeit(stepm, opt, batch)
print(f"benchmark + channels_last     {ms:8.1f} ms   loss {loss:.4f}")
torch.backends.cudnn.allow_tf32 = False
ms, loss = timeit(stepm, opt, batch)
print(f"  ... with conv TF32 disabled {ms:8.1f} ms   loss {loss:.4f}")
