"""Times the tensor-core correlation kernels of the product library beside the staging-ablated builds
(tools/experiments/staging_ablation.sh) at the bench shapes.  The ablated builds compute garbage; their kernel times bound
what a TMA-fed variant could reach (see the headers of csrc/corr_umma_fwd.cu / corr_umma_bwd.cu).

    tools/experiments/staging_ablation.sh && python tools/time_staging_ablation.py
"""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
P, I, Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
FWD = [P, P, P] + [I] * 6 + [P, Z, P]
BWD = [P] * 5 + [I] * 6 + [P, Z, P]


def load(path):
    lib = ctypes.CDLL(str(path))
    lib.d2t_corr_fwd_f32_tc.argtypes, lib.d2t_corr_fwd_f32_tc.restype = FWD, I
    lib.d2t_corr_bwd_f32_tc.argtypes, lib.d2t_corr_bwd_f32_tc.restype = BWD, I
    lib.d2t_corr_bwd_tc_workspace_bytes.argtypes, lib.d2t_corr_bwd_tc_workspace_bytes.restype = [I] * 6, Z
    lib.d2t_last_error.restype = ctypes.c_char_p
    return lib


def time_us(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    libs = [("product", ROOT / "detect-to-track_b200" / "libd2t_b200.so")]
    for m in (1, 2, 3):
        p = ROOT / "tools" / "_build" / f"libd2t_ablate{m}.so"
        if p.exists():
            libs.append((f"ablate{m}", p))
    what = {
        "product": "as shipped",
        "ablate1": "no global loads of the staged operands (stores unchanged)",
        "ablate2": "no global loads; LDS.128 raw -> STS.128 lo only (TMA + raw-as-hi design)",
        "ablate3": "no global loads; LDS.128 raw -> STS.128 hi in place + STS.128 lo (TMA + rounded split)",
    }
    B, H, W, d = 8, 38, 63, 8
    stream = P(torch.cuda.current_stream().cuda_stream)
    print(f"# tensor-core correlation kernels, B = {B}, {H}x{W}, d = {d}; median of 20, CUDA events; us per call")
    print(f"# {'build':9s} {'C':>5s} {'fwd':>8s} {'bwd (flip + 2 kernels)':>24s}   what")
    for C in (512, 1024, 2048):
        g = torch.Generator(device=dev).manual_seed(1)
        fm0 = torch.randn(B, C, H, W, device=dev, generator=g)
        fm1 = torch.randn(B, C, H, W, device=dev, generator=g)
        go = torch.randn(B, H, W, 17, 17, device=dev, generator=g)
        out = torch.empty(B, H, W, 17, 17, device=dev)
        g0, g1 = torch.empty_like(fm0), torch.empty_like(fm1)
        for name, path in libs:
            lib = load(path)
            wsb = lib.d2t_corr_bwd_tc_workspace_bytes(B, C, H, W, d, 1)
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)

            def fwd():
                rc = lib.d2t_corr_fwd_f32_tc(fm0.data_ptr(), fm1.data_ptr(), out.data_ptr(), B, C, H, W, d, 1, None, 0, stream)
                assert rc == 0, lib.d2t_last_error()

            def bwd():
                rc = lib.d2t_corr_bwd_f32_tc(go.data_ptr(), fm0.data_ptr(), fm1.data_ptr(), g0.data_ptr(), g1.data_ptr(), B, C, H, W,
                                             d, 1, ws.data_ptr(), wsb, stream)
                assert rc == 0, lib.d2t_last_error()

            print(f"  {name:9s} {C:5d} {time_us(fwd):8.1f} {time_us(bwd):24.1f}   {what[name]}")
    # what padding both maps to a 16-byte-aligned pitch would cost (the pre-pass a TMA-fed kernel needs): one strided copy each
    for C in (512, 1024, 2048):
        fm = torch.randn(B, C, H, W, device=dev)
        pad = torch.empty(B, C, H, 64, device=dev)
        t = time_us(lambda: pad[..., :W].copy_(fm))
        print(f"  pad copy  {C:5d} {t:8.1f} us per map (torch strided copy {H}x{W} -> pitch 64; two maps per call)")


if __name__ == "__main__":
    sys.exit(main())
