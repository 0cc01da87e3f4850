"""Per-op device timing (CUDA events, L2 flushed between iterations) for the BASELINE shapes.

    python tools/time_ops.py [--ref] [--iters N]

Prints one line per op: median ms, algorithmic GFLOP/s or GB/s (SURVEY.md section 8d figures).
--ref also times the reference's own kernels (oracle/_ref) on the same inputs.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cases  # noqa: E402
from detect_to_track_b200 import pointwise_correlation as pc, roipool as rp, ps_roipool as ps  # noqa: E402


def live_pairs(B, H, W, d):
    v = lambda n: sum(len(range(max(0, i - d), min(i + d, n))) for i in range(n))
    return B * v(H) * v(W)


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[len(ts) // 10], ts[(len(ts) * 9) // 10]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--skip-bwd", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device="cpu").manual_seed(1234)
    rows = []

    def report(name, ms, flops=None, nbytes=None):
        med, p10, p90 = ms
        s = f"{name:44s} {med*1e3:9.1f} us  (p10 {p10*1e3:8.1f}, p90 {p90*1e3:8.1f})"
        if flops:
            s += f"  {flops/med*1e-9:8.2f} TFLOP/s"
        if nbytes:
            s += f"  {nbytes/med*1e-6:8.1f} GB/s"
        print(s, flush=True)
        rows.append({"name": name, "us": med * 1e3, "tflops": flops / med * 1e-9 if flops else None,
                     "gbs": nbytes / med * 1e-6 if nbytes else None})

    ref = None
    if args.ref:
        from oracle import ref_cuda
        ref = ref_cuda if ref_cuda.available() else None

    corr_cfgs = [("cfg1 C=256 32x32 d=4 B=2", 2, 256, 32, 32, 4)]
    for B in (1, 8):
        for nm, C in (("c3", 512), ("c4", 1024), ("c5", 2048)):
            corr_cfgs.append((f"cfg3 {nm} C={C} 38x63 d=8 B={B}", B, C, 38, 63, 8))
    for name, B, C, H, W, d in corr_cfgs:
        if args.only and args.only not in "corr" + name:
            continue
        fm0 = (torch.randn(B, C, H, W, generator=g).relu_() / 16).to(dev)
        fm1 = (torch.randn(B, C, H, W, generator=g).relu_() / 16).to(dev)
        go = torch.randn(B, H, W, 2 * d + 1, 2 * d + 1, generator=g).to(dev)
        P = live_pairs(B, H, W, d)
        k2 = (2 * d + 1) ** 2
        fb = 2 * B * C * H * W * 4 + B * H * W * k2 * 4
        bb = B * H * W * k2 * 4 + 4 * B * C * H * W * 4
        report("corr fwd " + name, timeit(lambda: pc.pointwise_correlation_forward(fm0, fm1, d, 1), args.iters, flush), 2.0 * C * P, fb)
        if d == 8 and C >= 128:   # the FP32-pipe kernel the default replaced
            from detect_to_track_b200 import _lib
            lib = _lib.lib()
            n = lib.d2t_corr_fwd_simt_workspace_bytes(B, C, H, W, d, 1)
            wsb = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
            o2 = torch.empty((B, H, W, 2 * d + 1, 2 * d + 1), device=dev)
            st = torch.cuda.current_stream().cuda_stream
            report("  corr fwd (FP32-pipe kernel) " + name, timeit(lambda: lib.d2t_corr_fwd_f32_simt(
                fm0.data_ptr(), fm1.data_ptr(), o2.data_ptr(), B, C, H, W, d, 1, wsb.data_ptr(), n, st), args.iters, flush), 2.0 * C * P, fb)
        if not args.skip_bwd:
          report("corr bwd " + name, timeit(lambda: pc.pointwise_correlation_backward(go, fm0, fm1, d, 1), args.iters, flush), 4.0 * C * P, bb)
        if ref is not None and B * C <= 2048:
            n = max(2, args.iters // 5)
            report("  REF corr fwd " + name, timeit(lambda: ref.corr_fwd(fm0, fm1, d, 1), n, flush), 2.0 * C * P, fb)
            report("  REF corr bwd " + name, timeit(lambda: ref.corr_bwd(go, fm0, fm1, d, 1), n, flush), 4.0 * C * P, bb)

    if not args.only or "roipool" in args.only:
        C, H, W, k, R = 1891, 38, 63, 7, 300
        rois = torch.from_numpy(cases.rois_random(R, 1238)).to(dev)
        fm = torch.randn(C, H, W, generator=g).to(dev)
        go = torch.randn(R, C, k, k, generator=g).to(dev)
        nb = C * H * W * 4 + R * C * k * k * 4
        report("roipool fwd cfg4 C=1891 R=300", timeit(lambda: rp.roipool_forward(fm, rois, k), args.iters, flush), None, nb)
        report("roipool bwd cfg4 C=1891 R=300", timeit(lambda: rp.roipool_backward(go, rois, H, W), args.iters, flush), None, nb)
        if ref is not None:
            report("  REF roipool fwd", timeit(lambda: ref.roipool_fwd(fm, rois, k), 4, flush), None, nb)
            report("  REF roipool bwd", timeit(lambda: ref.roipool_bwd(go, rois, H, W), 4, flush), None, nb)

    if not args.only or "trackhead" in args.only:
        from detect_to_track_b200 import track_head as th
        C, H, W, k, R, n_out = 1891, 38, 63, 7, 300, 4
        rois = torch.from_numpy(cases.rois_random(R, 1238)).to(dev)
        fm = torch.randn(C, H, W, generator=g).to(dev)
        weight = (torch.randn(n_out, C * k * k, generator=g) / 300).to(dev)
        bias = torch.zeros(n_out, device=dev)
        go = torch.randn(R, n_out, generator=g).to(dev)
        report("trackhead fused fwd C=1891 R=300", timeit(lambda: th.track_head_forward(fm, rois, weight, bias, k), args.iters, flush))
        report("trackhead fused bwd (fm+W+b grads)", timeit(lambda: th.track_head_backward(go, fm, rois, weight, k), args.iters, flush))
        report("trackhead fused bwd (fm grad only)", timeit(lambda: th.track_head_backward(go, fm, rois, weight, k, True, False, False), args.iters, flush))
        lin = torch.nn.Linear(C * k * k, n_out).to(dev)
        torch.backends.cuda.matmul.allow_tf32 = False

        def unfused():
            x = fm.detach().requires_grad_(True)
            o = lin(rp.ROIPoolFunction.apply(x, rois, k).view(R, -1))
            o.backward(go)
        report("trackhead UNFUSED fwd+bwd (ROIPool+Linear, autograd)", timeit(unfused, max(3, args.iters // 2), flush))

    if not args.only or "psroi" in args.only:
        H, W, k, R = 38, 63, 7, 300
        rois = torch.from_numpy(cases.rois_random(R, 1237)).to(dev)
        for nm, nT, live in (("cls nT=31", 31, 608), ("reg nT=4", 4, 117)):
            fm = torch.randn(nT * k * k, H, W, generator=g).to(dev)
            go = torch.randn(R, nT, k, k, generator=g).to(dev)
            fb = live * H * W * 4 + R * nT * k * k * 4 + R * 16
            bb = nT * k * k * H * W * 4 + R * nT * k * k * 4 + R * 16
            report(f"psroipool fwd cfg2 {nm} R=300", timeit(lambda: ps.ps_roipool_forward(fm, rois, nT, k), args.iters, flush), None, fb)
            report(f"psroipool bwd cfg2 {nm} R=300", timeit(lambda: ps.ps_roipool_backward(go, rois, H, W), args.iters, flush), None, bb)
            if ref is not None:
                report("  REF psroipool fwd", timeit(lambda: ref.psroipool_fwd(fm, rois, nT, k), 4, flush), None, fb)
                report("  REF psroipool bwd", timeit(lambda: ref.psroipool_bwd(go, rois, H, W), 4, flush), None, bb)
    if not args.only or "psbatch" in args.only:
        H, W, k, R, NF = 38, 63, 7, 300, 16
        rois = torch.stack([torch.from_numpy(cases.rois_random(R, 1237 + f)) for f in range(NF)]).to(dev)
        for nm, nT, live in (("cls nT=31", 31, 608), ("reg nT=4", 4, 117)):
            fm = torch.randn(NF, nT * k * k, H, W, generator=g).to(dev)
            go = torch.randn(NF, R, nT, k, k, generator=g).to(dev)
            fb = NF * (live * H * W * 4 + R * nT * k * k * 4 + R * 16)
            bb = NF * (nT * k * k * H * W * 4 + R * nT * k * k * 4 + R * 16)
            report(f"psroipool batched fwd {nm} {NF} frames", timeit(lambda: ps.ps_roipool_forward_batched(fm, rois, nT, k), args.iters, flush), None, fb)
            report(f"psroipool batched bwd {nm} {NF} frames", timeit(lambda: ps.ps_roipool_backward_batched(go, rois, H, W), args.iters, flush), None, bb)
    if args.json:
        Path(args.json).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
