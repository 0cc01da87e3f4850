"""Golden vectors for the op CALLERS: the reference's UNMODIFIED models/rfcn.py and models/correlation_tracker.py,
imported from /root/reference into a stub `detect_to_track.models` package whose three ops are CPU stand-ins backed
by the oracle (tests/oracle_ops.py), run forward + backward on seeded inputs.

    python tools/make_golden_models.py            # writes tests/golden/models_{rfcn,tracker}.npz

Runs in THIS container (needs /root/reference, no GPU).  tests/test_models.py re-runs it when the reference is mounted
(fixture provenance) and compares detect_to_track_b200.models on the GPU against the fixtures.
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
REF_MODELS = Path("/root/reference/detect_to_track/models")

RFCN_CFG = dict(in_channels=8, n_classes=3, k=7)
TRACKER_CFG = dict(d_max=3, r_hw=7, reg_channels=6, stride=1)
H, W = 20, 21


def load_reference_callers():
    """stub package detect_to_track.models = {oracle-backed ops}; then exec the reference's two caller files inside it"""
    import oracle_ops
    for name in [m for m in sys.modules if m == "detect_to_track" or m.startswith("detect_to_track.")]:
        del sys.modules[name]
    top = types.ModuleType("detect_to_track")
    top.__path__ = []
    pkg = types.ModuleType("detect_to_track.models")
    pkg.__path__ = []
    pkg.PSROIPool, pkg.PointwiseCorrelation, pkg.ROIPool = oracle_ops.PSROIPool, oracle_ops.PointwiseCorrelation, oracle_ops.ROIPool
    sys.modules["detect_to_track"], sys.modules["detect_to_track.models"] = top, pkg
    mods = {}
    for stem in ("rfcn", "correlation_tracker"):
        spec = importlib.util.spec_from_file_location(f"detect_to_track.models.{stem}", REF_MODELS / f"{stem}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)   # the file is executed as it lies under /root/reference
        mods[stem] = mod
    return mods["rfcn"].RFCN, mods["correlation_tracker"].CorrelationTracker


def rois(seed, n):
    import cases
    r = np.concatenate([cases.rois_random(n, seed), cases.rois_edge_cases(H, W)[:6]], 0).astype(np.float32)
    half = r[:, 2:] / 2          # keep the boxes inside the frame: ROIPool's empty bins are NaN (F7) and would poison the loss
    r[:, :2] = np.minimum(np.maximum(r[:, :2], half), 1.0 - half)
    return torch.from_numpy(r)


def run_rfcn(RFCN):
    torch.manual_seed(101)
    net = RFCN(**RFCN_CFG)
    x = torch.randn(RFCN_CFG["in_channels"], H, W, requires_grad=True)
    regions = rois(7, 10)
    c_hat, b_hat = net(x, regions)
    wc, wb = torch.randn(c_hat.shape), torch.randn(b_hat.shape)
    ((c_hat * wc).sum() + (b_hat * wb).sum()).backward()
    out = {"x": x, "regions": regions, "wc": wc, "wb": wb, "c_hat": c_hat, "b_hat": b_hat, "grad_x": x.grad}
    for k, v in net.state_dict().items():
        out["sd." + k] = v
    for k, v in net.named_parameters():
        out["grad." + k] = v.grad
    return {k: v.detach().numpy().copy() for k, v in out.items()}


def run_tracker(Tracker):
    torch.manual_seed(202)
    net = Tracker(**TRACKER_CFG)
    Cr = TRACKER_CFG["reg_channels"]
    inp = {
        "c3_0": torch.randn(8, 2 * H, 2 * W), "c3_1": torch.randn(8, 2 * H, 2 * W),
        "c4_0": torch.randn(12, H, W), "c4_1": torch.randn(12, H, W),
        "c5_0": torch.randn(16, H, W), "c5_1": torch.randn(16, H, W),
        "reg_0": torch.randn(Cr, H, W), "reg_1": torch.randn(Cr, H, W),
    }
    for v in inp.values():
        v.requires_grad_(True)
    r = rois(8, 9)
    t_hat = net({"c3": inp["c3_0"], "c4": inp["c4_0"], "c5": inp["c5_0"]},
                {"c3": inp["c3_1"], "c4": inp["c4_1"], "c5": inp["c5_1"]}, inp["reg_0"], inp["reg_1"], r)
    wt = torch.randn(t_hat.shape)
    (t_hat * wt).sum().backward()
    out = {"rois": r, "wt": wt, "t_hat": t_hat}
    for k, v in inp.items():
        out["in." + k] = v
        out["gin." + k] = v.grad
    for k, v in net.state_dict().items():
        out["sd." + k] = v
    for k, v in net.named_parameters():
        out["grad." + k] = v.grad
    return {k: v.detach().numpy().copy() for k, v in out.items()}


def generate():
    RFCN, Tracker = load_reference_callers()
    return run_rfcn(RFCN), run_tracker(Tracker)


if __name__ == "__main__":
    rf, tr = generate()
    gold = ROOT / "tests" / "golden"
    np.savez_compressed(gold / "models_rfcn.npz", **rf)
    np.savez_compressed(gold / "models_tracker.npz", **tr)
    print("wrote", gold / "models_rfcn.npz", gold / "models_tracker.npz")
