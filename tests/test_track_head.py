"""GPU parity tests of the fused track head (ROIPool -> view -> Linear, csrc/track_head.cu + gemm_tf32x3.cu) against the
reference COMPOSITION: the API-parity ROIPool (bit-identical `exact` kernel; at full size also the reference's own CUDA
kernel, oracle/_ref) followed by the Linear layer evaluated in float64.  Tolerance as for every float32 op: rtol 1e-4,
atol 1e-5 * max|ref| (the GEMMs are 3xTF32 on tcgen05, stated |err| <= 2e-6 * sum|a||b|)."""
import numpy as np
import pytest
import torch

import cases
from oracle import ref_cuda
from detect_to_track_b200 import roipool as rp_mod, track_head as th

pytestmark = pytest.mark.gpu


def close(got, want, what):
    got, want = got.detach().double().cpu().numpy(), want.detach().double().cpu().numpy()
    scale = float(np.nanmax(np.abs(want))) or 1.0
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 * scale, equal_nan=True, err_msg=what)


def composition(fm, rois, weight, bias, k, go, pooled=None):
    """reference composition with the Linear layer in float64; returns t_hat, grad_fm, grad_weight, grad_bias"""
    R, C = rois.size(0), fm.size(0)
    if pooled is None:
        pooled = rp_mod.roipool_forward(fm, rois, k, exact=True)
    flat = pooled.view(R, -1).double()
    t_hat = flat @ weight.double().t() + bias.double()
    g_pooled = (go.double() @ weight.double()).float().view(R, C, k, k).contiguous()
    g_fm = rp_mod.roipool_backward(g_pooled, rois, fm.size(1), fm.size(2))
    g_w = go.double().t() @ torch.nan_to_num(flat, nan=0.0)
    return t_hat, g_fm, g_w, go.double().sum(0)


def inside(rois):
    half = rois[:, 2:] / 2
    rois[:, :2] = np.minimum(np.maximum(rois[:, :2], half), 1.0 - half)
    return rois


@pytest.mark.parametrize("C,H,W,R,k,n_out", [(37, 20, 21, 40, 7, 4), (130, 38, 63, 77, 7, 4), (8, 9, 10, 5, 3, 2),
                                             (300, 12, 70, 33, 5, 8), (1, 1, 1, 2, 7, 4)])
def test_track_head_matches_the_composition(cuda, C, H, W, R, k, n_out):
    g = torch.Generator(device="cpu").manual_seed(1000 + C)
    fm = torch.randn(C, H, W, generator=g).to(cuda)
    rois = torch.from_numpy(inside(cases.rois_random(R, 77))).to(cuda)
    weight = (torch.randn(n_out, C * k * k, generator=g) / (C * k * k) ** 0.5).to(cuda)
    bias = torch.randn(n_out, generator=g).to(cuda)
    go = torch.randn(R, n_out, generator=g).to(cuda)
    want = composition(fm, rois, weight, bias, k, go)
    out = th.track_head_forward(fm, rois, weight, bias, k)
    close(out, want[0], "t_hat")
    got = th.track_head_backward(go, fm, rois, weight, k)
    for a, b, nm in zip(got, want[1:], ("grad_fm", "grad_weight", "grad_bias")):
        close(a, b, nm)
    again = th.track_head_backward(go, fm, rois, weight, k)
    assert torch.equal(out, th.track_head_forward(fm, rois, weight, bias, k))
    assert all(torch.equal(a, b) for a, b in zip(got, again))          # bitwise reproducible


@pytest.mark.parametrize("N,C,H,W,R,k,n_out", [(3, 37, 20, 21, 40, 7, 4), (2, 130, 38, 63, 77, 7, 4), (4, 8, 9, 10, 5, 3, 2),
                                               (2, 300, 12, 70, 33, 5, 8), (8, 257, 38, 63, 50, 7, 4)])
def test_track_head_batched_matches_per_image_composition(cuda, N, C, H, W, R, k, n_out):
    """the batched form (N images, shared weight / bias, one set of launches; H*W not a multiple of 4 in most cases, so the
    padded K blocks of the weight-gradient GEMM are exercised): t_hat and grad_fm per image, grad_weight / grad_bias summed
    over the images -- against the reference composition per image, and bitwise reproducible; N = 1 batched == unbatched."""
    g = torch.Generator(device="cpu").manual_seed(2000 + C)
    fm = torch.randn(N, C, H, W, generator=g).to(cuda)
    rois = torch.stack([torch.from_numpy(inside(cases.rois_random(R, 90 + n))) for n in range(N)]).to(cuda)
    weight = (torch.randn(n_out, C * k * k, generator=g) / (C * k * k) ** 0.5).to(cuda)
    bias = torch.randn(n_out, generator=g).to(cuda)
    go = torch.randn(N, R, n_out, generator=g).to(cuda)
    out = th.track_head_forward(fm, rois, weight, bias, k)
    got = th.track_head_backward(go, fm, rois, weight, k)
    assert tuple(out.shape) == (N, R, n_out) and tuple(got[0].shape) == tuple(fm.shape)
    gw, gb = 0, 0
    for n in range(N):
        want = composition(fm[n], rois[n], weight, bias, k, go[n])
        close(out[n], want[0], f"t_hat[{n}]")
        close(got[0][n], want[1], f"grad_fm[{n}]")
        gw, gb = gw + want[2], gb + want[3]
    close(got[1], gw, "grad_weight")
    close(got[2], gb, "grad_bias")
    assert torch.equal(out, th.track_head_forward(fm, rois, weight, bias, k))
    assert all(torch.equal(a, b) for a, b in zip(got, th.track_head_backward(go, fm, rois, weight, k)))
    one = th.track_head_forward(fm[:1], rois[:1], weight, bias, k)
    assert torch.equal(one[0], th.track_head_forward(fm[0], rois[0], weight, bias, k))
    # autograd through the batched Function
    fm_r, w_r, b_r = fm.clone().requires_grad_(True), weight.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    (th.TrackHeadFunction.apply(fm_r, rois, w_r, b_r, k) * go).sum().backward()
    assert torch.equal(fm_r.grad, got[0]) and torch.equal(w_r.grad, got[1]) and torch.equal(b_r.grad, got[2])


def test_track_head_full_size_vs_reference_kernels(cuda):
    """BASELINE config 4 (C = 1891, 38x63, 300 RoIs, k = 7, Linear(92659, 4)): the fused head against the reference's own
    ROIPool kernels (oracle/_ref) + float64 Linear, on the WHOLE tensors, including RoIs with empty bins (NaN rows)."""
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    C, H, W, R, k, n_out = 1891, 38, 63, 300, 7, 4
    g = torch.Generator(device="cpu").manual_seed(1238)
    fm = torch.randn(C, H, W, generator=g).to(cuda)
    rois = torch.from_numpy(cases.rois_random(R, 1238)).to(cuda)
    weight = (torch.randn(n_out, C * k * k, generator=g) / (C * k * k) ** 0.5).to(cuda)
    bias = torch.randn(n_out, generator=g).to(cuda)
    go = torch.randn(R, n_out, generator=g).to(cuda)
    pooled = ref_cuda.roipool_fwd(fm, rois, k)
    flat = pooled.view(R, -1).double()
    want = flat @ weight.double().t() + bias.double()
    out = th.track_head_forward(fm, rois, weight, bias, k)
    assert 0 < int(torch.isnan(want).any(1).sum()) < R                 # some RoIs cross the border: empty bins -> NaN (F7)
    assert torch.equal(torch.isnan(out), torch.isnan(want))
    close(out, want, "t_hat")
    g_pooled = (go.double() @ weight.double()).float().view(R, C, k, k).contiguous()
    want_fm = ref_cuda.roipool_bwd(g_pooled, rois, H, W)
    g_fm, g_w, g_b = th.track_head_backward(go, fm, rois, weight, k)
    close(g_fm, want_fm, "grad_fm")
    close(g_w, go.double().t() @ torch.nan_to_num(flat, nan=0.0), "grad_weight")   # empty bins carry no gradient
    close(g_b, go.double().sum(0), "grad_bias")


def test_track_head_autograd_and_optional_grads(cuda):
    C, H, W, R, k, n_out = 21, 14, 15, 12, 7, 4
    g = torch.Generator(device="cpu").manual_seed(5)
    fm = torch.randn(C, H, W, generator=g).to(cuda).requires_grad_(True)
    rois = torch.from_numpy(inside(cases.rois_random(R, 78))).to(cuda)
    lin = torch.nn.Linear(C * k * k, n_out).to(cuda)
    out = th.TrackHeadFunction.apply(fm, rois, lin.weight, lin.bias, k)
    wt = torch.randn(out.shape, generator=g).to(cuda)
    (out * wt).sum().backward()
    fm2 = fm.detach().clone().requires_grad_(True)
    lin2 = torch.nn.Linear(C * k * k, n_out).to(cuda)
    lin2.load_state_dict(lin.state_dict())
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = lin2(rp_mod.ROIPoolFunction.apply(fm2, rois, k).view(R, -1))
    (ref * wt).sum().backward()
    close(out, ref, "t_hat")
    close(fm.grad, fm2.grad, "grad_fm")
    close(lin.weight.grad, lin2.weight.grad, "grad_weight")
    close(lin.bias.grad, lin2.bias.grad, "grad_bias")
    # no bias, frozen weights: only grad_fm is computed
    fm3 = fm.detach().clone().requires_grad_(True)
    o3 = th.TrackHeadFunction.apply(fm3, rois, lin.weight.detach(), None, k)
    o3.sum().backward()
    assert fm3.grad is not None
    # no RoIs
    empty = torch.zeros(0, 4, device=cuda)
    assert tuple(th.track_head_forward(fm.detach(), empty, lin.weight.detach(), None, k).shape) == (0, n_out)
    gz = th.track_head_backward(torch.zeros(0, n_out, device=cuda), fm.detach(), empty, lin.weight.detach(), k)
    assert all(not bool(t.any()) for t in gz)


@pytest.mark.parametrize("M,N,K,splits,bn,col", [(2394, 196, 1891, 7, 208, 0), (300, 50, 70, 1, 64, 0), (1000, 1891, 196, 1, 208, 1),
                                                 (129, 256, 33, 2, 256, 0), (1891, 196, 2394, 9, 208, 0)])
def test_gemm_building_block_matches_float64(cuda, M, N, K, splits, bn, col):
    """d2t_gemm_tf32x3_f32 (TMA + tcgen05, 3xTF32) against a float64 matmul: ragged M / N / K (TMA zero-fills), split-K slabs,
    row- and column-major epilogues; stated bound |err| <= 4e-6 * sum_k |a||b| (measured 2.5e-6 at K = 1891)."""
    from detect_to_track_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    lda = K + (-K) % 4
    A = torch.randn(M, lda, generator=g).to(cuda)
    B = torch.randn(N, lda, generator=g).to(cuda)
    ldo = M if col else N + (-N) % 4
    out = torch.full((splits * (N if col else M) * ldo,), float("nan"), device=cuda)
    rc = lib.d2t_gemm_tf32x3_f32(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, lda, lda, ldo, col, splits, bn,
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    got = out.view(splits, -1).sum(0)
    got = got.view(N, M).t() if col else got.view(M, ldo)[:, :N]
    ref = A[:, :K].double() @ B[:, :K].double().t()
    mag = A[:, :K].double().abs() @ B[:, :K].double().abs().t()
    assert bool(((got.double() - ref).abs() <= 4e-6 * mag + 1e-30).all())


@pytest.mark.parametrize("case", range(24))
def test_track_head_shape_sweep(cuda, case):
    """Seeded random shapes of the fused head against the per-image composition: image counts whose position tiles do and do
    not fill whole waves (the spill path of the forward GEMM), r_hw 1 .. 7, 1 .. 5 outputs, channel counts that are not
    multiples of the GEMM's K block, maps down to 1x1."""
    rng = np.random.default_rng(7000 + case)
    k = int(rng.integers(1, 8))
    n_out = int(rng.integers(1, 6))
    N = int(rng.choice([1, 2, 3, 8, 9, 17]))
    C = int(rng.choice([1, 5, 31, 32, 33, 100, 257]))
    H, W = int(rng.integers(1, 40)), int(rng.integers(1, 64))
    if case % 8 == 7:
        N, H, W = 8, 38, 63            # 150 position tiles on 148 SMs
    R = int(rng.choice([1, 3, 40, 120]))
    g = torch.Generator(device="cpu").manual_seed(7100 + case)
    fm = torch.randn(N, C, H, W, generator=g).to(cuda)
    rois = torch.stack([torch.from_numpy(inside(cases.rois_random(R, 7200 + 20 * case + n))) for n in range(N)]).to(cuda)
    weight = (torch.randn(n_out, C * k * k, generator=g) / (C * k * k) ** 0.5).to(cuda)
    bias = torch.randn(n_out, generator=g).to(cuda)
    go = torch.randn(N, R, n_out, generator=g).to(cuda)
    out = th.track_head_forward(fm, rois, weight, bias, k)
    got = th.track_head_backward(go, fm, rois, weight, k)
    gw, gb = 0, 0
    for n in range(N):
        want = composition(fm[n], rois[n], weight, bias, k, go[n])
        close(out[n], want[0], f"t_hat[{n}]")
        close(got[0][n], want[1], f"grad_fm[{n}]")
        gw, gb = gw + want[2], gb + want[3]
    close(got[1], gw, "grad_weight")
    close(got[2], gb, "grad_bias")
