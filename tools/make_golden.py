"""Generate tests/golden/*.npz from the reference's OWN kernels, on a GPU box.

    gpurun -- python tools/make_golden.py gpurun_out/golden
    cp gpurun_out/golden/*.npz tests/golden/

Runs the unmodified reference kernels (oracle/_ref/libd2t_ref_cuda.so, built by
oracle/Makefile from /root/reference) on the seeded inputs of tests/cases.py and
stores inputs + outputs.  ROIPool golden RoIs exclude fully out-of-bounds boxes
only where the reference itself yields NaN (kept, compared with equal_nan).
grad_FM1 / pooled grads come from atomicAdd, so their last bits depend on the
run; tests compare them with a tolerance.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import cases  # noqa: E402
from oracle import ref_cuda  # noqa: E402


def main(outdir):
    out = Path(outdir)
    out.mkdir(parents=True, exist_ok=True)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)

    for name, B, C, H, W, d, s, dt in cases.GOLDEN_CORR:
        fm0, fm1, go = cases.corr_inputs(B, C, H, W, d, seed=sum(map(ord, name)), dtype=np.dtype(dt))
        o = ref_cuda.corr_fwd(t(fm0), t(fm1), d, s)
        g0, g1 = ref_cuda.corr_bwd(t(go), t(fm0), t(fm1), d, s)
        torch.cuda.synchronize()
        np.savez_compressed(out / f"{name}.npz", fm0=fm0, fm1=fm1, go=go, d=d, stride=s,
                            out=o.cpu().numpy(), g0=g0.cpu().numpy(), g1=g1.cpu().numpy())
        print("wrote", name)

    for name, C, H, W, k, R, dt in cases.GOLDEN_ROIPOOL:
        seed = sum(map(ord, name))
        rois = cases.golden_rois(H, W, R, seed, np.dtype(dt), include_oob=True)
        fm, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), seed, np.dtype(dt))
        o = ref_cuda.roipool_fwd(t(fm), t(rois), k)
        g = ref_cuda.roipool_bwd(t(go), t(rois), H, W)
        torch.cuda.synchronize()
        np.savez_compressed(out / f"{name}.npz", fm=fm, rois=rois, go=go, k=k, out=o.cpu().numpy(), gin=g.cpu().numpy())
        print("wrote", name)

    for name, nT, H, W, k, R, dt in cases.GOLDEN_PSROIPOOL:
        seed = sum(map(ord, name))
        rois = cases.golden_rois(H, W, R, seed, np.dtype(dt), include_oob=True)
        fm, go = cases.pool_inputs(nT * k * k, H, W, (rois.shape[0], nT, k, k), seed, np.dtype(dt))
        o = ref_cuda.psroipool_fwd(t(fm), t(rois), nT, k)
        g = ref_cuda.psroipool_bwd(t(go), t(rois), H, W)
        torch.cuda.synchronize()
        np.savez_compressed(out / f"{name}.npz", fm=fm, rois=rois, go=go, nT=nT, k=k, out=o.cpu().numpy(),
                            gin=g.cpu().numpy())
        print("wrote", name)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
