"""The op CALLERS: detect_to_track_b200.models against the reference's UNMODIFIED
models/rfcn.py and models/correlation_tracker.py.

The reference files cannot travel to the GPU box and need the CUDA ops to run, so the comparison is made through golden
vectors: tools/make_golden_models.py imports the two files from /root/reference into a stub `detect_to_track.models`
package whose ops are the CPU oracle (tests/oracle_ops.py), runs forward + backward and stores the state_dict, inputs,
outputs and every gradient.  Here:
  * (CPU, when /root/reference is mounted) the generator is re-run and must reproduce the committed fixtures;
  * (GPU) RFCN / CorrelationTracker of this package load the SAME state_dict (strict) and must give the same outputs,
    input gradients and parameter gradients on the CUDA ops, rtol 1e-4 -- eager, and with the fused track head.
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(ROOT / "tools"))


def _close(got, want, what):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got
    scale = float(np.abs(want).max()) or 1.0
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 * scale, err_msg=what)


@pytest.mark.skipif(not Path("/root/reference/detect_to_track/models/rfcn.py").exists(), reason="reference not mounted")
def test_fixtures_come_from_the_unmodified_reference_callers():
    import make_golden_models as mg
    rf, tr = mg.generate()
    for got, name in ((rf, "models_rfcn"), (tr, "models_tracker")):
        want = np.load(GOLDEN / f"{name}.npz")
        assert sorted(got) == sorted(want.files)
        for k in want.files:
            np.testing.assert_allclose(got[k], want[k], rtol=1e-6, atol=1e-7, err_msg=f"{name}:{k}")


def _state_dict(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_rfcn_matches_the_reference_wiring(cuda, fused):
    import make_golden_models as mg
    import detect_to_track_b200 as d2t
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(GOLDEN / "models_rfcn.npz")
    net = d2t.RFCN(**mg.RFCN_CFG, fused=fused).to(cuda)
    net.load_state_dict(_state_dict(g), strict=True)      # same parameter names and shapes as rfcn.py
    x = torch.from_numpy(g["x"]).to(cuda).requires_grad_(True)
    regions = torch.from_numpy(g["regions"]).to(cuda)
    c_hat, b_hat = net(x, regions)
    _close(c_hat, g["c_hat"], "c_hat")
    _close(b_hat, g["b_hat"], "b_hat")
    ((c_hat * torch.from_numpy(g["wc"]).to(cuda)).sum() + (b_hat * torch.from_numpy(g["wb"]).to(cuda)).sum()).backward()
    _close(x.grad, g["grad_x"], "grad_x")
    for name, prm in net.named_parameters():
        _close(prm.grad, g["grad." + name], name)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_correlation_tracker_matches_the_reference_wiring(cuda, fused):
    import make_golden_models as mg
    import detect_to_track_b200 as d2t
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(GOLDEN / "models_tracker.npz")
    net = d2t.CorrelationTracker(**mg.TRACKER_CFG, fused=fused).to(cuda)
    net.load_state_dict(_state_dict(g), strict=True)
    inp = {k[3:]: torch.from_numpy(g[k]).to(cuda).requires_grad_(True) for k in g.files if k.startswith("in.")}
    rois = torch.from_numpy(g["rois"]).to(cuda)
    t_hat = net({"c3": inp["c3_0"], "c4": inp["c4_0"], "c5": inp["c5_0"]},
                {"c3": inp["c3_1"], "c4": inp["c4_1"], "c5": inp["c5_1"]}, inp["reg_0"], inp["reg_1"], rois)
    _close(t_hat, g["t_hat"], "t_hat")
    (t_hat * torch.from_numpy(g["wt"]).to(cuda)).sum().backward()
    for k, v in inp.items():
        _close(v.grad, g["gin." + k], "grad " + k)
    for name, prm in net.named_parameters():
        _close(prm.grad, g["grad." + name], name)
