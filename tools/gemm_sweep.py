"""Timing sweep of the 3xTF32 GEMM building block (d2t_gemm_tf32x3_f32): fixed vs per-k-block cost."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from detect_to_track_b200 import _lib

lib = _lib.lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(M, N, K, splits, bn, col=0, iters=15):
    A = torch.randn(M, K + (-K) % 4, device=dev)
    B = torch.randn(N, K + (-K) % 4, device=dev)
    ldo = M if col else N + (-N) % 4
    out = torch.empty(splits * (N if col else M) * ldo, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: lib.d2t_gemm_tf32x3_f32(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, A.size(1), B.size(1), ldo, col, splits, bn, st)
    for _ in range(3):
        assert call() == 0, _lib.last_error()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    ref = (A[:, :K].double() @ B[:, :K].double().t())
    got = out.view(splits, -1).sum(0)
    got = got.view(N, M).t() if col else got.view(M, ldo)[:, :N]
    err = float((got.double() - ref).abs().max() / ref.abs().max())
    print(f"M={M:5d} N={N:5d} K={K:5d} splits={splits:2d} bn={bn} col={col}: {ts[len(ts)//2]:7.1f} us  ctas={-(-M//128)*-(-N//bn)*splits:4d}  kb/cta={-(-K//32)/splits:5.1f}  relerr={err:.2e}")


for splits in (1, 2, 4, 7):
    run(2394, 196, 1891, splits, 208)
for K in (32, 64, 128, 256, 512):
    run(128 * 148, 196, K, 1, 208)
run(2394, 1891, 196, 1, 208, col=1)
run(2394, 1891, 196, 1, 256, col=1)
run(1891, 196, 2394, 9, 208)
