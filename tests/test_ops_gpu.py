"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the
C ABI, against (1) the C oracle, (2) the reference's own kernels (oracle/_ref, when built),
(3) the committed golden fixtures, (4) the reference's own tests (float64 gradcheck, OOB
known answer), and (5) size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): integer bin edges / live masks exact; float32 values
and gradients rtol 1e-4 (atol scaled to the data); float64 1e-10.
"""
from pathlib import Path

import numpy as np
import pytest
import torch
from torch.autograd import gradcheck

import cases
import oracle
from oracle import ref_cuda
import detect_to_track_b200 as d2t
from detect_to_track_b200 import roipool as rp_mod, ps_roipool as ps_mod, pointwise_correlation as pc_mod

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def dev(a, device):
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def close(got: torch.Tensor, want, dtype, scale=None, equal_nan=False):
    got = got.detach().cpu().numpy()
    want = np.asarray(want)
    rtol = 1e-4 if dtype == np.float32 else 1e-10
    s = float(np.nanmax(np.abs(want))) if scale is None else scale
    atol = (1e-5 if dtype == np.float32 else 1e-11) * max(s, 1e-30)
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, equal_nan=equal_nan)


def corr_live_mask(H, W, d, stride):
    """(H, W, 2d+1, 2d+1) bool: entry (i, j, ci, cj) is sampled by the reference's loops
    (pointwise_correlation_cuda.cu:92-100) iff ci < 2d, 0 <= i-d+ci < H and ((i-d+ci) - max(0, i-d)) % stride == 0,
    and the same along j."""
    def axis(n):
        m = np.zeros((n, 2 * d + 1), bool)
        for i in range(n):
            lo = max(0, i - d)
            for p in range(lo, min(i + d, n), stride):
                m[i, p - i + d] = True
        return m
    mi, mj = axis(H), axis(W)
    return mi[:, None, :, None] & mj[None, :, None, :]


# ------------------------------------------------------------------ correlation
CORR_CASES = [
    (1, 2, 10, 10, 3, 1), (2, 2, 11, 10, 3, 2), (2, 2, 10, 11, 3, 1), (1, 2, 11, 11, 3, 2),    # reference grid
    (1, 5, 7, 9, 4, 1), (2, 4, 9, 14, 2, 3), (1, 3, 5, 4, 8, 1), (1, 1, 1, 1, 1, 1),
    (2, 256, 32, 32, 4, 1),          # BASELINE config 1
    (1, 96, 38, 63, 8, 1),           # config-3 geometry, fewer channels
    (3, 40, 20, 21, 8, 1),
    (1, 16, 70, 66, 8, 1),           # wider than one tile
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("B,C,H,W,d,stride", CORR_CASES)
def test_corr_fwd_bwd_vs_oracle(cuda, B, C, H, W, d, stride, dtype):
    fm0, fm1, go = cases.corr_inputs(B, C, H, W, d, seed=11, dtype=dtype)
    out = pc_mod.pointwise_correlation_forward(dev(fm0, cuda), dev(fm1, cuda), d, stride)
    want = oracle.corr_fwd(fm0, fm1, d, stride)
    close(out, want, dtype)
    # structural live mask (SURVEY.md F4/F5 closed form): every dead entry is EXACTLY zero, in our output and in the oracle's
    live = torch.from_numpy(corr_live_mask(H, W, d, stride))
    assert torch.equal((out.cpu() != 0) & ~live, torch.zeros_like(live).expand_as(out))
    assert not bool(((torch.from_numpy(want) != 0) & ~live).any())
    assert float(out[..., 2 * d, :].abs().max()) == 0 and float(out[..., :, 2 * d].abs().max()) == 0
    g0, g1 = pc_mod.pointwise_correlation_backward(dev(go, cuda), dev(fm0, cuda), dev(fm1, cuda), d, stride)
    w0, w1 = oracle.corr_bwd(go, fm0, fm1, d, stride)
    close(g0, w0, dtype)
    close(g1, w1, dtype)


def test_corr_live_mask_is_exact_on_strictly_positive_inputs(cuda):
    """with strictly positive maps every sampled entry is > 0, so the zero pattern of the output IS the structural
    mask: torch.equal against the closed form, for both strides and both tuned / generic kernels."""
    for (B, C, H, W, d, stride) in [(2, 3, 11, 10, 3, 2), (1, 24, 38, 63, 8, 1), (2, 8, 12, 13, 4, 1), (1, 4, 9, 14, 2, 3)]:
        g = torch.Generator(device="cpu").manual_seed(3)
        fm0 = (torch.rand(B, C, H, W, generator=g) + 0.5).to(cuda)
        fm1 = (torch.rand(B, C, H, W, generator=g) + 0.5).to(cuda)
        out = pc_mod.pointwise_correlation_forward(fm0, fm1, d, stride)
        live = torch.from_numpy(corr_live_mask(H, W, d, stride)).to(cuda)
        assert torch.equal(out != 0, live.expand_as(out))


@pytest.mark.parametrize("name,C", [("c3", 512), ("c4", 1024), ("c5", 2048)])
def test_corr_bench_shapes_vs_reference_kernels(cuda, name, C):
    """BASELINE config 3 at FULL size with DEFAULT dispatch -- B = 8, 38x63, d = 8: the SIMT band forward and the
    tcgen05 (3xTF32) backward, the exact kernels and shapes bench.py times -- against the reference's own CUDA kernels
    (oracle/_ref) run on the same device.  Same tolerance as everywhere: rtol 1e-4, atol 1e-5 * max|ref|."""
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    B, H, W, d = 8, 38, 63, 8
    g = torch.Generator(device="cpu").manual_seed(1234)
    fm0 = (torch.randn(B, C, H, W, generator=g).relu_() / 16).to(cuda)
    fm1 = (torch.randn(B, C, H, W, generator=g).relu_() / 16).to(cuda)
    go = torch.randn(B, H, W, 17, 17, generator=g).to(cuda)
    out = pc_mod.pointwise_correlation_forward(fm0, fm1, d, 1)
    ref = ref_cuda.corr_fwd(fm0, fm1, d, 1)
    close(out, ref.cpu().numpy(), np.float32)
    live = torch.from_numpy(corr_live_mask(H, W, d, 1)).to(cuda)
    assert not bool(((out != 0) & ~live).any())
    g0, g1 = pc_mod.pointwise_correlation_backward(go, fm0, fm1, d, 1)
    r0, r1 = ref_cuda.corr_bwd(go, fm0, fm1, d, 1)
    close(g0, r0.cpu().numpy(), np.float32)
    close(g1, r1.cpu().numpy(), np.float32)


@pytest.mark.parametrize("B,C,H,W", [(1, 137, 38, 63), (3, 160, 20, 21), (1, 200, 70, 66), (2, 300, 9, 17), (2, 2048, 38, 63), (1, 128, 5, 70)])
def test_corr_fwd_tensor_core_kernel(cuda, B, C, H, W):
    """tcgen05 / 3xTF32 forward (default for d_max = 8, stride 1, C >= 128; MN-major operands).  Stated tolerance
    |err| <= (2e-6 + 1.5e-8 * C) * sum_c |fm0 * fm1| against the float64 generic kernel (magnitude sum from the same kernel
    on absolute values) -- the C term is the tensor core's truncating FP32 accumulator, reached by all-positive inputs,
    which are tested too; dead entries exactly zero; agrees with the FP32-pipe kernel at the op tolerance; ragged channel
    counts, maps smaller than a tile, partial tiles, chunks outside the image."""
    from detect_to_track_b200 import _lib
    lib = _lib.lib()
    d = 8
    g = torch.Generator(device="cpu").manual_seed(79)
    fm0 = torch.randn(B, C, H, W, generator=g).to(cuda)
    fm1 = torch.randn(B, C, H, W, generator=g).to(cuda)
    out = torch.full((B, H, W, 17, 17), float("nan"), device=cuda)
    rc = lib.d2t_corr_fwd_f32_tc(fm0.data_ptr(), fm1.data_ptr(), out.data_ptr(), B, C, H, W, d, 1, None, 0,
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    ref = pc_mod.pointwise_correlation_forward(fm0.double(), fm1.double(), d, 1)
    mag = pc_mod.pointwise_correlation_forward(fm0.double().abs(), fm1.double().abs(), d, 1)
    err = (out.double() - ref).abs()
    bound = 2e-6 + 1.5e-8 * C
    assert bool((err <= bound * mag + 1e-30).all()), float((err / (mag + 1e-30)).max())
    pos = torch.empty_like(out)     # same-signed terms: the accumulator bias adds up instead of cancelling
    a0, a1 = fm0.abs(), fm1.abs()
    rc = lib.d2t_corr_fwd_f32_tc(a0.data_ptr(), a1.data_ptr(), pos.data_ptr(), B, C, H, W, d, 1, None, 0,
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    assert bool(((pos.double() - mag).abs() <= bound * mag + 1e-30).all()), float(((pos.double() - mag).abs() / (mag + 1e-30)).max())
    live = torch.from_numpy(corr_live_mask(H, W, d, 1)).to(cuda)
    assert bool((out[~live.expand_as(out)] == 0).all())
    assert torch.equal(out, pc_mod.pointwise_correlation_forward(fm0, fm1, d, 1))     # the default dispatch runs this kernel
    n = lib.d2t_corr_fwd_simt_workspace_bytes(B, C, H, W, d, 1)
    ws = torch.empty(max(n, 1), dtype=torch.uint8, device=cuda)
    simt = torch.empty_like(out)
    rc = lib.d2t_corr_fwd_f32_simt(fm0.data_ptr(), fm1.data_ptr(), simt.data_ptr(), B, C, H, W, d, 1, ws.data_ptr(), n,
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    close(out, simt.cpu().numpy(), np.float32)


def test_corr_backward_is_batch_invariant(cuda):
    """the kernel family behind d2t_corr_bwd_f32 depends on (C, d_max, stride) only (include/d2t_b200.h), so the
    gradients of an image are bit-identical whether it is processed alone or inside a batch of 8."""
    for C in (64, 512):   # FP32-pipe family / tensor-core family
        fm0, fm1, go = (dev(a, cuda) for a in cases.corr_inputs(8, C, 38, 63, 8, seed=19, dtype=np.float32))
        g0, g1 = pc_mod.pointwise_correlation_backward(go, fm0, fm1, 8, 1)
        out = pc_mod.pointwise_correlation_forward(fm0, fm1, 8, 1)
        for b in (0, 5):
            h0, h1 = pc_mod.pointwise_correlation_backward(go[b:b + 1].contiguous(), fm0[b:b + 1].contiguous(),
                                                           fm1[b:b + 1].contiguous(), 8, 1)
            assert torch.equal(g0[b:b + 1], h0) and torch.equal(g1[b:b + 1], h1)
            if C >= 128:   # tensor-core forward: fixed channel order per item.  (The FP32-pipe forward cuts its stream-K
                # ranges by the total amount of work, so its summation order -- not its result within rtol -- follows B.)
                assert torch.equal(out[b:b + 1], pc_mod.pointwise_correlation_forward(fm0[b:b + 1].contiguous(),
                                                                                      fm1[b:b + 1].contiguous(), 8, 1))


def test_corr_nonfinite_inputs(cuda):
    """a NaN / Inf in one key pixel: the forward and the FP32-pipe backward poison exactly the windows that contain it,
    like the reference; the tensor-core backward is documented to spread it over the affected 8x16 tiles only
    (include/d2t_b200.h) -- positions of other tiles stay finite and correct."""
    B, C, H, W, d = 1, 128, 38, 63, 8
    fm0, fm1, go = (dev(a, cuda) for a in cases.corr_inputs(B, C, H, W, d, seed=23, dtype=np.float32))
    bad = fm1.clone()
    bad[0, 3, 20, 40] = float("inf")
    from detect_to_track_b200 import _lib
    lib = _lib.lib()

    def fwd_simt(a, b_):
        o = torch.empty((B, H, W, 17, 17), device=cuda)
        n = lib.d2t_corr_fwd_simt_workspace_bytes(B, C, H, W, d, 1)
        ws = torch.empty(max(n, 1), dtype=torch.uint8, device=cuda)
        rc = lib.d2t_corr_fwd_f32_simt(a.data_ptr(), b_.data_ptr(), o.data_ptr(), B, C, H, W, d, 1, ws.data_ptr(), n,
                                       torch.cuda.current_stream().cuda_stream)
        assert rc == 0, _lib.last_error()
        return o

    out = fwd_simt(fm0, bad)
    clean = fwd_simt(fm0, fm1)
    nonfinite = ~torch.isfinite(out)
    # exactly the (i, j, ci, cj) with i-d+ci == 20 and j-d+cj == 40
    want = torch.zeros_like(nonfinite)
    for i in range(max(0, 20 - d + 1), min(H, 20 + d + 1)):
        for j in range(max(0, 40 - d + 1), min(W, 40 + d + 1)):
            want[0, i, j, 20 - i + d, 40 - j + d] = True
    assert torch.equal(nonfinite, want)
    assert torch.equal(out[~want], clean[~want])
    # tensor-core forward (default at C >= 128): the windows containing the value are non-finite, and nothing outside the
    # 8 x 16 tiles whose halo patch holds pixel (20, 40) is touched (rows 8..31, columns 32..62)
    tc = pc_mod.pointwise_correlation_forward(fm0, bad, d, 1)
    tcc = pc_mod.pointwise_correlation_forward(fm0, fm1, d, 1)
    nf = ~torch.isfinite(tc)
    assert bool(nf[want].all())
    assert not bool(nf[0, :8].any()) and not bool(nf[0, 32:].any()) and not bool(nf[0, :, :32].any())
    assert torch.equal(tc[~nf], tcc[~nf])
    g0, g1 = torch.empty_like(fm0), torch.empty_like(fm1)
    rc = lib.d2t_corr_bwd_f32_simt(go.data_ptr(), fm0.data_ptr(), bad.data_ptr(), g0.data_ptr(), g1.data_ptr(), B, C, H, W, d, 1,
                                   None, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    nf0 = ~torch.isfinite(g0)
    assert bool(nf0[0, 3].any()) and not bool(nf0[0, :3].any()) and not bool(nf0[0, 4:].any())   # only channel 3 of grad_FM0
    assert bool(torch.isfinite(g1).all())                                                         # grad_FM1 does not read FM1
    # tensor-core family: rows 12..27 / cols 32..48 can see pixel (20, 40); tiles are 8 rows x 16 cols
    t0, t1 = pc_mod.pointwise_correlation_backward(go, fm0, bad, d, 1)
    nf = ~torch.isfinite(t0)
    assert not bool(nf[0, :, :8].any()) and not bool(nf[0, :, 32:].any()) and not bool(nf[0, :, :, :16].any())
    assert bool(torch.isfinite(t1).all())


@pytest.mark.parametrize("B,C,H,W,d,stride", [(2, 2, 11, 10, 3, 2), (2, 64, 32, 32, 4, 1), (1, 48, 38, 63, 8, 1)])
def test_corr_vs_reference_kernels(cuda, B, C, H, W, d, stride):
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    fm0, fm1, go = (dev(a, cuda) for a in cases.corr_inputs(B, C, H, W, d, seed=12, dtype=np.float32))
    close(pc_mod.pointwise_correlation_forward(fm0, fm1, d, stride), ref_cuda.corr_fwd(fm0, fm1, d, stride).cpu().numpy(), np.float32)
    g0, g1 = pc_mod.pointwise_correlation_backward(go, fm0, fm1, d, stride)
    r0, r1 = ref_cuda.corr_bwd(go, fm0, fm1, d, stride)
    close(g0, r0.cpu().numpy(), np.float32)
    close(g1, r1.cpu().numpy(), np.float32)


@pytest.mark.parametrize("d_max", [3])
@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("input_b", [1, 2])
@pytest.mark.parametrize("input_c", [2])
@pytest.mark.parametrize("input_hw", [10, 11])
def test_pointwise_correlation_gradients(cuda, d_max, stride, input_b, input_c, input_hw):
    """the reference's tests/test_pointwise_correlation.py:8-22, unchanged but for the import."""
    pc = d2t.PointwiseCorrelation(d_max, stride).cuda()
    fm_shape = (input_b, input_c, input_hw, input_hw)
    fm0 = torch.rand(*fm_shape).double().cuda().requires_grad_(True)
    fm1 = torch.rand(*fm_shape).double().cuda().requires_grad_(True)
    assert gradcheck(pc, (fm0, fm1))


@pytest.mark.parametrize("B,C,H,W", [(1, 96, 38, 63), (3, 40, 20, 21), (1, 16, 70, 66), (2, 300, 9, 17), (2, 2048, 38, 63)])
def test_corr_bwd_tensor_core_variant(cuda, B, C, H, W):
    """tcgen05 / 3xTF32 backward (d2t_corr_bwd_f32_tc).  Stated tolerance: |err| <= 2e-6 * sum |grad_out * fm| over the
    element's terms (measured 9e-7, tools/umma_sw128_test.cu); checked against the float64 generic kernel, with the
    magnitude sum obtained from the same kernel on absolute values.  Also bitwise reproducible."""
    from detect_to_track_b200 import _lib
    d = 8
    g = torch.Generator(device="cpu").manual_seed(78)
    fm0 = torch.randn(B, C, H, W, generator=g).to(cuda)
    fm1 = torch.randn(B, C, H, W, generator=g).to(cuda)
    go = torch.randn(B, H, W, 17, 17, generator=g).to(cuda)
    lib = _lib.lib()
    n = lib.d2t_corr_bwd_tc_workspace_bytes(B, C, H, W, d, 1)
    assert n == B * H * W * 256 * 4
    ws = torch.empty(n, dtype=torch.uint8, device=cuda)

    def run():
        g0, g1 = torch.full_like(fm0, float("nan")), torch.full_like(fm1, float("nan"))
        rc = lib.d2t_corr_bwd_f32_tc(go.data_ptr(), fm0.data_ptr(), fm1.data_ptr(), g0.data_ptr(), g1.data_ptr(), B, C, H, W,
                                     d, 1, ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, _lib.last_error()
        return g0, g1

    g0, g1 = run()
    r0, r1 = pc_mod.pointwise_correlation_backward(go.double(), fm0.double(), fm1.double(), d, 1)
    m0, m1 = pc_mod.pointwise_correlation_backward(go.double().abs(), fm0.double().abs(), fm1.double().abs(), d, 1)
    for got, ref, mag in ((g0, r0, m0), (g1, r1, m1)):
        err = (got.double() - ref).abs()
        assert bool((err <= 2e-6 * mag + 1e-30).all()), float((err / (mag + 1e-30)).max())
    # the op-level tolerance of the FP32 path (rtol 1e-4 at the result's scale) holds as well
    close(g0, r0.cpu().numpy(), np.float32)
    close(g1, r1.cpu().numpy(), np.float32)
    h0, h1 = run()
    assert torch.equal(g0, h0) and torch.equal(g1, h1)


@pytest.mark.parametrize("B,C,H,W", [(1, 37, 38, 63), (2, 300, 9, 17), (1, 520, 20, 21)])
def test_tensor_core_kernels_stay_inside_their_buffers(cuda, B, C, H, W):
    """compute-sanitizer is not available on the GPU pool, so the tcgen05 kernels are run with every output and the
    workspace embedded in larger buffers filled with a sentinel: the guard zones must come back untouched (ragged
    channel counts, maps smaller than a tile, partial tiles in both directions)."""
    from detect_to_track_b200 import _lib
    lib = _lib.lib()
    d, guard, sentinel = 8, 4096, 1234.5
    g = torch.Generator(device="cpu").manual_seed(5)
    fm0 = torch.randn(B, C, H, W, generator=g).to(cuda)
    fm1 = torch.randn(B, C, H, W, generator=g).to(cuda)
    go = torch.randn(B, H, W, 17, 17, generator=g).to(cuda)

    def guarded(n):
        buf = torch.full((n + 2 * guard,), sentinel, device=cuda)
        return buf, buf[guard:guard + n]

    def intact(buf, n):
        return bool((buf[:guard] == sentinel).all()) and bool((buf[guard + n:] == sentinel).all())

    stream = torch.cuda.current_stream().cuda_stream
    n_ws = lib.d2t_corr_bwd_tc_workspace_bytes(B, C, H, W, d, 1) // 4
    wsb, ws = guarded(n_ws)
    g0b, g0 = guarded(fm0.numel())
    g1b, g1 = guarded(fm1.numel())
    rc = lib.d2t_corr_bwd_f32_tc(go.data_ptr(), fm0.data_ptr(), fm1.data_ptr(), g0.data_ptr(), g1.data_ptr(), B, C, H, W, d, 1,
                                 ws.data_ptr(), n_ws * 4, stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert intact(wsb, n_ws) and intact(g0b, fm0.numel()) and intact(g1b, fm1.numel())
    assert bool(torch.isfinite(g0).all()) and bool(torch.isfinite(g1).all())

    # SIMT forward with its stream-K partial workspace
    n_fws = lib.d2t_corr_fwd_workspace_bytes(B, C, H, W, d, 1, 4) // 4
    fwb, fws = guarded(max(n_fws, 1))
    ob, o = guarded(B * H * W * 289)
    rc = lib.d2t_corr_fwd_f32(fm0.data_ptr(), fm1.data_ptr(), o.data_ptr(), B, C, H, W, d, 1, fws.data_ptr(), n_fws * 4, stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert intact(fwb, max(n_fws, 1)) and intact(ob, B * H * W * 289)
    assert not bool((o == sentinel).any())  # every output element is written (no memset needed)


@pytest.mark.parametrize("C,H,W,R", [(29, 38, 63, 97), (200, 17, 64, 300), (3, 38, 20, 5)])
def test_roipool_backward_stays_inside_its_buffers(cuda, C, H, W, R):
    """the float32 ROIPool backward with its output embedded in a sentinel-filled buffer and grad_out at the very end of
    its allocation: guard zones untouched, every output element written (ragged channel tiles, partial RoI groups)."""
    from detect_to_track_b200 import _lib
    lib = _lib.lib()
    k, guard, sentinel = 7, 4096, 1234.5
    rois_np = _roipool_rois(H, W, np.float32, R=R)
    R = rois_np.shape[0]
    g = torch.Generator(device="cpu").manual_seed(6)
    n_go = R * C * k * k
    gob = torch.full((guard + n_go,), sentinel, device=cuda)
    go = gob[guard:]
    go.copy_(torch.randn(n_go, generator=g))
    rois = dev(rois_np, cuda)
    n = C * H * W
    buf = torch.full((n + 2 * guard,), sentinel, device=cuda)
    gin = buf[guard:guard + n]
    rc = lib.d2t_roipool_bwd_f32(go.data_ptr(), rois.data_ptr(), gin.data_ptr(), R, C, H, W, k, None, 0,
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + n:] == sentinel).all())
    assert bool((gob[:guard] == sentinel).all())
    assert not bool((gin == sentinel).any()) and bool(torch.isfinite(gin).all())
    want = oracle.roipool_bwd(go.view(R, C, k, k).cpu().numpy(), rois_np, H, W)
    close(gin.view(C, H, W), want, np.float32)


@pytest.mark.parametrize("N,nT,H,W,k,R", [(1, 31, 38, 63, 7, 300), (3, 31, 38, 63, 7, 40), (2, 4, 17, 50, 5, 33), (1, 5, 9, 33, 8, 600),
                                           (2, 3, 20, 21, 9, 30)])
def test_psroipool_backward_stays_inside_its_buffers(cuda, N, nT, H, W, k, R):
    """the float32 PSROIPool backward (pool_ps3.cu and the fallback kernels) through the C ABI with its output, its gradient
    input and its WORKSPACE each embedded in sentinel-filled allocations: guard zones untouched, every output element written
    (the prep kernel zero-fills the channels nobody reads, the main kernel writes the rest), result equal to the oracle."""
    from detect_to_track_b200 import _lib
    lib = _lib.lib()
    guard, sentinel = 4096, 1234.5
    rois_np = np.stack([np.concatenate([cases.rois_edge_cases(H, W), cases.rois_random(R, 8100 + n), cases.ROIS_OOB.astype(np.float32)])
                        for n in range(N)]).astype(np.float32)
    Rt = rois_np.shape[1]
    g = torch.Generator(device="cpu").manual_seed(8)
    nCh = nT * k * k
    n_go, n_out = N * Rt * nCh, N * nCh * H * W
    gob = torch.full((guard + n_go,), sentinel, device=cuda)
    go = gob[guard:]
    go.copy_(torch.randn(n_go, generator=g))
    rois = dev(rois_np, cuda)
    buf = torch.full((n_out + 2 * guard,), sentinel, device=cuda)
    gin = buf[guard:guard + n_out]
    nws = lib.d2t_psroipool_bwd_batched_workspace_bytes(N, Rt, nT, H, W, k, 4)
    wsb = torch.full((max(nws, 1) + 2 * guard,), 0x5A, dtype=torch.uint8, device=cuda)
    rc = lib.d2t_psroipool_bwd_batched_f32(go.data_ptr(), rois.data_ptr(), gin.data_ptr(), N, Rt, nT, H, W, k, 0,
                                           wsb.data_ptr() + guard if nws else None, nws, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + n_out:] == sentinel).all())
    assert bool((gob[:guard] == sentinel).all())
    assert bool((wsb[:guard] == 0x5A).all()) and bool((wsb[guard + nws:] == 0x5A).all())
    assert not bool((gin == sentinel).any()) and bool(torch.isfinite(gin).all())
    gv = gin.view(N, nCh, H, W)
    for n in range(N):
        close(gv[n], oracle.psroipool_bwd(go.view(N, Rt, nT, k, k)[n].cpu().numpy(), rois_np[n], H, W), np.float32)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_pooling_with_no_rois(cuda, dtype):
    """R = 0: the forward returns an empty tensor and the backward a zero gradient of the map's shape (the reference
    allocates zeros and launches nothing: roipool_cuda.cu:141/172, ps_roipool_cuda.cu:155/188)."""
    C, H, W, k, nT = 10, 12, 13, 7, 2
    rois = torch.zeros(0, 4, dtype=dtype, device=cuda)
    fm = torch.randn(C, H, W, dtype=dtype, device=cuda)
    out = rp_mod.roipool_forward(fm, rois, k)
    assert tuple(out.shape) == (0, C, k, k)
    gin = rp_mod.roipool_backward(torch.zeros(0, C, k, k, dtype=dtype, device=cuda), rois, H, W)
    assert tuple(gin.shape) == (C, H, W) and not bool(gin.any())
    sfm = torch.randn(nT * k * k, H, W, dtype=dtype, device=cuda)
    pout = ps_mod.ps_roipool_forward(sfm, rois, nT, k)
    assert tuple(pout.shape) == (0, nT, k, k)
    pgin = ps_mod.ps_roipool_backward(torch.zeros(0, nT, k, k, dtype=dtype, device=cuda), rois, H, W)
    assert tuple(pgin.shape) == (nT * k * k, H, W) and not bool(pgin.any())


def test_pooling_on_maps_larger_than_the_packed_edge_range(cuda):
    """H or W above 255: the float32 fast paths pack bin edges into bytes and must hand such maps to the generic kernels."""
    C, H, W, k, nT = 3, 20, 300, 7, 1
    rois = _roipool_rois(H, W, np.float32, R=30)
    fm, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), 41, np.float32)
    want = oracle.roipool_fwd(fm, rois, k)
    out = rp_mod.roipool_forward(dev(fm, cuda), dev(rois, cuda), k)
    close(out, want, np.float32, scale=float(np.nanmax(np.abs(want))), equal_nan=True)
    close(rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W), oracle.roipool_bwd(go, rois, H, W), np.float32)
    sfm, sgo = cases.pool_inputs(nT * k * k, W, H, (rois.shape[0], nT, k, k), 42, np.float32)   # 300 rows x 20 columns
    np.testing.assert_array_equal(ps_mod.ps_roipool_forward(dev(sfm, cuda), dev(rois, cuda), nT, k).cpu().numpy(),
                                  oracle.psroipool_fwd(sfm, rois, nT, k))
    close(ps_mod.ps_roipool_backward(dev(sgo, cuda), dev(rois, cuda), W, H), oracle.psroipool_bwd(sgo, rois, W, H), np.float32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_psroipool_backward_on_wide_maps(cuda, dtype):
    """W above 255: the packed edge fields of every PSROIPool backward path are 16 bits wide now --
    float32 takes the one-launch channel-owner kernel while the plane fits shared memory and the per-pixel gather
    kernels beyond that, float64 always the latter."""
    k, nT = 7, 2
    for (H, W) in [(12, 300), (40, 700)]:
        rois = _roipool_rois(H, W, dtype, R=25)
        _, go = cases.pool_inputs(nT * k * k, H, W, (rois.shape[0], nT, k, k), 43, dtype)
        got = ps_mod.ps_roipool_backward(dev(go, cuda), dev(rois, cuda), H, W)
        close(got, oracle.psroipool_bwd(go, rois, H, W), dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_roipool_on_planes_larger_than_shared_memory(cuda, dtype):
    """a 250 x 300 plane (300 KB in float32) does not fit the shared-memory slab of any ROIPool kernel: the
    forward falls back to the per-output global-memory kernel (bit-identical to the reference order), the backward to
    row bands of the channel-owner kernel.  Also r_hw = 5 (not the tuned 7)."""
    C, H, W = 3, 250, 300
    for k in (7, 5):
        rois = _roipool_rois(H, W, dtype, R=20)
        fm, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), 44, dtype)
        want = oracle.roipool_fwd(fm, rois, k)
        out = rp_mod.roipool_forward(dev(fm, cuda), dev(rois, cuda), k, exact=True)
        if dtype == np.float32:
            np.testing.assert_array_equal(out.cpu().numpy(), want)
        else:
            close(out, want, dtype, scale=float(np.nanmax(np.abs(want))), equal_nan=True)
        close(rp_mod.roipool_forward(dev(fm, cuda), dev(rois, cuda), k), want, dtype, scale=float(np.nanmax(np.abs(want))), equal_nan=True)
        gin = rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W)
        close(gin, oracle.roipool_bwd(go, rois, H, W), dtype)
        assert torch.equal(gin, rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W))


def test_corr_bwd_explicit_families_agree(cuda):
    """d2t_corr_bwd_f32_simt (FP32 FMAs) and d2t_corr_bwd_f32_tc (tcgen05, 3xTF32) agree within the FP32 tolerance, and
    the default entry point picks the family the header documents (C >= 128 -> tensor cores, bit-identical to _tc)."""
    from detect_to_track_b200 import _lib
    lib = _lib.lib()
    for C in (72, 200):
        B, H, W, d = 2, 38, 63, 8
        fm0, fm1, go = (dev(a, cuda) for a in cases.corr_inputs(B, C, H, W, d, seed=15, dtype=np.float32))
        stream = torch.cuda.current_stream().cuda_stream
        n = lib.d2t_corr_bwd_tc_workspace_bytes(B, C, H, W, d, 1)
        ws = torch.empty(n, dtype=torch.uint8, device=cuda)
        res = {}
        for fam, args in (("simt", (None, 0)), ("tc", (ws.data_ptr(), n))):
            g0, g1 = torch.empty_like(fm0), torch.empty_like(fm1)
            rc = getattr(lib, f"d2t_corr_bwd_f32_{fam}")(go.data_ptr(), fm0.data_ptr(), fm1.data_ptr(), g0.data_ptr(), g1.data_ptr(),
                                                         B, C, H, W, d, 1, *args, stream)
            assert rc == 0, _lib.last_error()
            res[fam] = (g0, g1)
        close(res["tc"][0], res["simt"][0].cpu().numpy(), np.float32)
        close(res["tc"][1], res["simt"][1].cpu().numpy(), np.float32)
        dflt = pc_mod.pointwise_correlation_backward(go, fm0, fm1, d, 1)
        fam = "tc" if C >= 128 else "simt"
        assert torch.equal(dflt[0], res[fam][0]) and torch.equal(dflt[1], res[fam][1])
    # a missing workspace is an error, never a silent change of kernel family
    g0, g1 = torch.empty_like(fm0), torch.empty_like(fm1)
    rc = lib.d2t_corr_bwd_f32(go.data_ptr(), fm0.data_ptr(), fm1.data_ptr(), g0.data_ptr(), g1.data_ptr(), B, C, H, W, d, 1,
                              None, 0, stream)
    assert rc == 3


@pytest.mark.parametrize("C,H,W,d,stride,dtype", [(96, 38, 63, 8, 1, np.float32), (2048, 38, 63, 8, 1, np.float32),
                                                 (24, 12, 13, 4, 1, np.float32), (5, 9, 14, 2, 3, np.float32),
                                                 (3, 10, 11, 3, 2, np.float64)])
def test_corr_channel_major_output_is_bit_identical_to_the_permute(cuda, C, H, W, d, stride, dtype):
    """tracker glue fusion (correlation_tracker.py:64-80): d2t_corr_fwd_strided_* writing ((2d+1)^2, H, W) into a slice of a
    larger buffer == permute(2, 0, 1) of the reference-layout output, bit for bit; the rest of the buffer is untouched
    (whole tiles, stream-K split tiles at C = 2048, the generic kernel for other (d, stride), float64)."""
    fm0, fm1, _ = (dev(a, cuda) for a in cases.corr_inputs(1, C, H, W, d, seed=31, dtype=dtype))
    kk = (2 * d + 1) ** 2
    buf = torch.full((7 + kk + 5, H, W), 777.0, dtype=fm0.dtype, device=cuda)
    pc_mod.pointwise_correlation_forward_channel_major(fm0, fm1, d, stride, buf[7:7 + kk])
    want = pc_mod.pointwise_correlation_forward(fm0, fm1, d, stride).squeeze(0).view(H, W, -1).permute(2, 0, 1)
    assert torch.equal(buf[7:7 + kk], want)
    assert bool((buf[:7] == 777.0).all()) and bool((buf[7 + kk:] == 777.0).all())


def test_track_features_function_matches_cat_of_permutes(cuda):
    """TrackFeaturesFunction == torch.cat([reg0, reg1, corr maps permuted]) of the reference wiring: values bit-identical,
    gradients of all eight inputs equal to autograd through the reference composition."""
    H, W, d, Cr = 20, 21, 3, 6
    g = torch.Generator(device="cpu").manual_seed(9)
    mk = lambda *s: torch.randn(*s, generator=g).to(cuda).requires_grad_(True)
    ins = [mk(Cr, H, W), mk(Cr, H, W), mk(1, 8, H, W), mk(1, 8, H, W), mk(1, 12, H, W), mk(1, 12, H, W), mk(1, 16, H, W), mk(1, 16, H, W)]
    out = d2t.TrackFeaturesFunction.apply(*ins, d, 1)
    wt = torch.randn(out.shape, generator=g).to(cuda)
    (out * wt).sum().backward()
    ins2 = [t.detach().clone().requires_grad_(True) for t in ins]
    pc = d2t.PointwiseCorrelation(d, 1)
    feats = [pc(ins2[2 + 2 * n], ins2[3 + 2 * n]).squeeze(0).view(H, W, -1).permute(2, 0, 1) for n in range(3)]
    ref = torch.cat([ins2[0], ins2[1], *feats])
    (ref * wt).sum().backward()
    assert torch.equal(out, ref)
    for a, b in zip(ins, ins2):
        assert torch.equal(a.grad, b.grad)


def test_corr_backward_is_deterministic(cuda):
    fm0, fm1, go = (dev(a, cuda) for a in cases.corr_inputs(2, 64, 38, 63, 8, seed=13, dtype=np.float32))
    a = pc_mod.pointwise_correlation_backward(go, fm0, fm1, 8, 1)
    b = pc_mod.pointwise_correlation_backward(go, fm0, fm1, 8, 1)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_corr_full_size_adjoint_identity(cuda):
    """BASELINE config 3 (c5: C=2048, 38x63, d=8) at full size: the op is bilinear, so
    <fwd(x0,x1), g> == <x0, bwd0(g)> == <x1, bwd1(g)> -- checked in float64 accumulation."""
    g = torch.Generator(device="cpu").manual_seed(1234)
    fm0 = (torch.randn(1, 2048, 38, 63, generator=g).relu_() / 16).to(cuda)
    fm1 = (torch.randn(1, 2048, 38, 63, generator=g).relu_() / 16).to(cuda)
    go = torch.randn(1, 38, 63, 17, 17, generator=g).to(cuda)
    out = pc_mod.pointwise_correlation_forward(fm0, fm1, 8, 1)
    g0, g1 = pc_mod.pointwise_correlation_backward(go, fm0, fm1, 8, 1)
    lhs = (out.double() * go.double()).sum().item()
    # lhs is a sum of 5e5 terms of random sign (|lhs| ~ 300, sum |terms| ~ 5e5).  The tensor-core forward's FP32 accumulator
    # truncates, which biases every one of these all-positive 2048-channel sums by about -1.4e-5 relative
    # (tools/adjoint_probe.py), so the identity is asked to hold to 3e-5 |lhs| + 1e-8 sum|terms|
    tol = 3e-5 * abs(lhs) + 1e-8 * (out.double().abs() * go.double().abs()).sum().item()
    assert abs(lhs - (fm0.double() * g0.double()).sum().item()) <= tol
    assert abs(lhs - (fm1.double() * g1.double()).sum().item()) <= tol
    # dead rows / columns (F4)
    assert float(out[..., 16, :].abs().max()) == 0 and float(out[..., :, 16].abs().max()) == 0
    # spot-check 64 random outputs against a float64 dot product
    idx = torch.randint(0, 38 * 63 * 256, (64,), generator=g)
    for n in idx.tolist():
        p, t = divmod(n, 256)
        i, j = divmod(p, 63)
        ci, cj = divmod(t, 16)
        di, dj = i - 8 + ci, j - 8 + cj
        want = 0.0
        if 0 <= di < 38 and 0 <= dj < 63:
            want = (fm0[0, :, i, j].double() * fm1[0, :, di, dj].double()).sum().item()
        assert abs(out[0, i, j, ci, cj].item() - want) <= 1e-4 * abs(want) + 1e-6


# ------------------------------------------------------------------ bin edges (integer: exact)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("clamp_start", [True, False])
@pytest.mark.parametrize("H,W,k", [(38, 63, 7), (10, 11, 5), (11, 10, 6)])
def test_bin_edges_bit_exact(cuda, H, W, k, clamp_start, dtype):
    rois = np.concatenate([cases.rois_edge_cases(H, W, dtype), cases.rois_random(3000, 21, dtype), cases.ROIS_OOB.astype(dtype)])
    got = rp_mod.pool_bins(dev(rois, cuda), H, W, k, clamp_start).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.bins(rois, H, W, k, clamp_start))


# ------------------------------------------------------------------ ROIPool
def _roipool_rois(H, W, dtype, R=40):
    return np.concatenate([cases.rois_edge_cases(H, W, dtype), cases.rois_random(R, 31, dtype), cases.ROIS_OOB.astype(dtype)])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("C,H,W,k", [(2, 10, 10, 5), (2, 11, 10, 6), (37, 38, 63, 7), (300, 20, 30, 3)])
def test_roipool_vs_oracle(cuda, C, H, W, k, dtype):
    rois = _roipool_rois(H, W, dtype)
    fm, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), 32, dtype)
    want = oracle.roipool_fwd(fm, rois, k)
    out = rp_mod.roipool_forward(dev(fm, cuda), dev(rois, cuda), k)
    close(out, want, dtype, scale=float(np.nanmax(np.abs(want))), equal_nan=True)       # default (row-prefix) kernel
    assert np.array_equal(np.isnan(out.cpu().numpy()), np.isnan(want))                 # NaNs of empty bins (F7)
    if dtype == np.float32:
        # exact variant: same summation order as the reference => bit-identical
        ex = rp_mod.roipool_forward(dev(fm, cuda), dev(rois, cuda), k, exact=True)
        np.testing.assert_array_equal(ex.cpu().numpy(), want)
    gin = rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W)
    # empty bins contribute nothing to the gradient
    close(gin, oracle.roipool_bwd(go, rois, H, W), dtype)


@pytest.mark.parametrize("C,H,W,R", [(21, 12, 9, 700), (3, 38, 63, 5), (16, 40, 70, 33), (53, 7, 200, 64)])
def test_roipool_vec_kernels_shapes(cuda, C, H, W, R):
    """float32, r_hw = 7 runs the [pixel][16 channel] kernels (pool_vec.cu): channel counts that are not a multiple
    of 4, more RoIs than one edge-table chunk (512), maps narrower than a bin row, wide maps."""
    k = 7
    rois = _roipool_rois(H, W, np.float32, R=R)
    fm, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), 34, np.float32)
    want = oracle.roipool_fwd(fm, rois, k)
    out = rp_mod.roipool_forward(dev(fm, cuda), dev(rois, cuda), k)
    close(out, want, np.float32, scale=float(np.nanmax(np.abs(want))), equal_nan=True)
    assert np.array_equal(np.isnan(out.cpu().numpy()), np.isnan(want))
    gin = rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W)
    close(gin, oracle.roipool_bwd(go, rois, H, W), np.float32)
    assert torch.equal(gin, rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W))


@pytest.mark.parametrize("C,H,W,R", [(29, 38, 63, 300), (5, 38, 64, 1100), (18, 16, 20, 9), (1, 1, 1, 3), (7, 50, 70, 41)])
def test_roipool_backward_shapes(cuda, C, H, W, R):
    """float32, r_hw = 7 backward (pool_vec2.cu: raw cp.async staging, per-row RoI lists, update classes) against the
    oracle and bitwise reproducible; channel counts that are not a multiple of 4, RoI counts that are not a multiple
    of the group size and exceed every table, RoIs whose bins are thinner than a pixel (update classes 1 and 2)."""
    k = 7
    rois = _roipool_rois(H, W, np.float32, R=R)
    tiny = np.asarray([[0.4, 0.6, 3.0 / H, 4.5 / W], [0.7, 0.3, 5.0 / H, 2.0 / W], [0.2, 0.2, 0.5 / H, 0.4 / W]], np.float32)
    rois = np.concatenate([rois, tiny], 0)
    _, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), 35, np.float32)
    want = oracle.roipool_bwd(go, rois, H, W)
    a = rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W)
    close(a, want, np.float32)
    assert torch.equal(a, rp_mod.roipool_backward(dev(go, cuda), dev(rois, cuda), H, W))


def test_roipool_vs_reference_kernels(cuda):
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    C, H, W, k = 64, 38, 63, 7
    rois = _roipool_rois(H, W, np.float32, R=100)
    fm, go = cases.pool_inputs(C, H, W, (rois.shape[0], C, k, k), 33, np.float32)
    fm, go, rois = dev(fm, cuda), dev(go, cuda), dev(rois, cuda)
    ref = ref_cuda.roipool_fwd(fm, rois, k)
    out = rp_mod.roipool_forward(fm, rois, k, exact=True)
    assert torch.equal(torch.nan_to_num(out, nan=12345.0), torch.nan_to_num(ref, nan=12345.0))   # bit-exact
    close(rp_mod.roipool_forward(fm, rois, k), ref.cpu().numpy(), np.float32, equal_nan=True)   # default kernel
    close(rp_mod.roipool_backward(go, rois, H, W), ref_cuda.roipool_bwd(go, rois, H, W).cpu().numpy(), np.float32)


@pytest.mark.parametrize("r_hw", [5, 6])
@pytest.mark.parametrize("fm_c", [2])
@pytest.mark.parametrize("fm_h", [10, 11])
@pytest.mark.parametrize("fm_w", [10, 11])
def test_roipool_gradients(cuda, r_hw, fm_c, fm_h, fm_w):
    """the reference's tests/test_roipool.py:10-27."""
    rp = d2t.ROIPool(r_hw)
    fm = torch.rand(fm_c, fm_h, fm_w).double().cuda().requires_grad_(True)
    rois = torch.Tensor([[0.5, 0.5, 0.5, 0.5], [0.1, 0.1, 0.2, 0.3]]).double().cuda().requires_grad_(False)
    assert gradcheck(rp, (fm, rois))


def test_roipool_full_size_track_head(cuda):
    """BASELINE config 4 at full size (1891 channels, 300 RoIs): adjoint identity + determinism +
    a sampled comparison with the oracle."""
    C, H, W, k, R = 1891, 38, 63, 7, 300
    rois_np = cases.rois_random(R, 1238)
    g = torch.Generator(device="cpu").manual_seed(1238)
    fm = torch.randn(C, H, W, generator=g).to(cuda)
    go = torch.randn(R, C, k, k, generator=g).to(cuda)
    rois = dev(rois_np, cuda)
    out = rp_mod.roipool_forward(fm, rois, k)
    gin = rp_mod.roipool_backward(go, rois, H, W)
    # RoIs crossing the bottom/right border have EMPTY bins -> 0/0 = NaN in the reference (F7); those
    # bins receive no gradient, so they drop out of the adjoint identity.
    empty = torch.isnan(out)
    assert 0 < int(empty.sum()) < out.numel() // 4
    lhs = (torch.where(empty, torch.zeros_like(out), out).double() * go.double()).sum().item()
    rhs = (fm.double() * gin.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    assert torch.equal(gin, rp_mod.roipool_backward(go, rois, H, W))
    sel = [0, 5, 777, 1890]
    want = oracle.roipool_fwd(fm[sel].cpu().numpy(), rois_np, k)
    close(out[:, sel], want, np.float32, equal_nan=True)
    np.testing.assert_array_equal(rp_mod.roipool_forward(fm, rois, k, exact=True)[:, sel].cpu().numpy(), want)
    wg = oracle.roipool_bwd(go[:, sel].cpu().numpy().copy(), rois_np, H, W)
    close(gin[sel], wg, np.float32)


def test_roipool_full_size_track_head_vs_reference_kernels(cuda):
    """BASELINE config 4 at full size, the WHOLE tensors (not sampled channels) against the reference's own kernels
    (oracle/_ref) on the same device: 300 x 1891 x 7 x 7 forward (NaN pattern of empty bins included, exact variant
    bit-identical) and the 1891 x 38 x 63 backward."""
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    C, H, W, k, R = 1891, 38, 63, 7, 300
    g = torch.Generator(device="cpu").manual_seed(1238)
    fm = torch.randn(C, H, W, generator=g).to(cuda)
    go = torch.randn(R, C, k, k, generator=g).to(cuda)
    rois = dev(cases.rois_random(R, 1238), cuda)
    ref = ref_cuda.roipool_fwd(fm, rois, k)
    out = rp_mod.roipool_forward(fm, rois, k)
    assert torch.equal(torch.isnan(out), torch.isnan(ref))
    close(out, ref.cpu().numpy(), np.float32, equal_nan=True)
    ex = rp_mod.roipool_forward(fm, rois, k, exact=True)
    assert torch.equal(torch.nan_to_num(ex, nan=12345.0), torch.nan_to_num(ref, nan=12345.0))
    close(rp_mod.roipool_backward(go, rois, H, W), ref_cuda.roipool_bwd(go, rois, H, W).cpu().numpy(), np.float32)


# ------------------------------------------------------------------ PSROIPool
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("canonical", [False, True])
@pytest.mark.parametrize("nT,H,W,k", [(1, 10, 10, 6), (2, 11, 10, 7), (4, 38, 63, 7), (31, 38, 63, 7), (33, 12, 13, 3), (2, 20, 21, 9)])
def test_psroipool_vs_oracle(cuda, nT, H, W, k, canonical, dtype):
    rois = _roipool_rois(H, W, dtype, R=60)
    fm, go = cases.pool_inputs(nT * k * k, H, W, (rois.shape[0], nT, k, k), 34, dtype)
    out = ps_mod.ps_roipool_forward(dev(fm, cuda), dev(rois, cuda), nT, k, canonical)
    want = oracle.psroipool_fwd(fm, rois, nT, k, canonical)
    if dtype == np.float32:
        np.testing.assert_array_equal(out.cpu().numpy(), want)
    else:
        close(out, want, dtype, scale=1.0)
    gin = ps_mod.ps_roipool_backward(dev(go, cuda), dev(rois, cuda), H, W, canonical)
    close(gin, oracle.psroipool_bwd(go, rois, H, W, canonical), dtype)


def test_psroipool_vs_reference_kernels(cuda):
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    nT, H, W, k = 31, 38, 63, 7
    rois = _roipool_rois(H, W, np.float32, R=300)
    fm, go = cases.pool_inputs(nT * k * k, H, W, (rois.shape[0], nT, k, k), 35, np.float32)
    fm, go, rois = dev(fm, cuda), dev(go, cuda), dev(rois, cuda)
    assert torch.equal(ps_mod.ps_roipool_forward(fm, rois, nT, k), ref_cuda.psroipool_fwd(fm, rois, nT, k))
    close(ps_mod.ps_roipool_backward(go, rois, H, W), ref_cuda.psroipool_bwd(go, rois, H, W).cpu().numpy(), np.float32)


@pytest.mark.parametrize("n_targets", [1, 2])
@pytest.mark.parametrize("r_hw", [6, 7])
@pytest.mark.parametrize("fm_h", [10, 11])
@pytest.mark.parametrize("fm_w", [10, 11])
def test_ps_roipool_gradients(cuda, n_targets, r_hw, fm_h, fm_w):
    """the reference's tests/test_ps_roipool.py:8-30."""
    pr = d2t.PSROIPool(n_targets, r_hw)
    fm = torch.rand(n_targets * r_hw ** 2, fm_h, fm_w).double().cuda().requires_grad_(True)
    rois = (torch.Tensor([[0.5, 0.5, 0.1, 0.1], [0.1, 0.1, 0.2, 0.3], [1.5, 1.5, 0.2, 0.2]])
            .double().cuda().requires_grad_(False))
    assert gradcheck(pr, (fm, rois))


@pytest.mark.parametrize("n_targets", [1, 2])
@pytest.mark.parametrize("r_hw", [6, 7])
@pytest.mark.parametrize("fm_h", [10, 11])
@pytest.mark.parametrize("fm_w", [10, 11])
def test_ps_roipool_can_handle_oob(cuda, n_targets, r_hw, fm_h, fm_w):
    """the reference's tests/test_ps_roipool.py:33-44 (its only known-answer test)."""
    pr = d2t.PSROIPool(n_targets, r_hw).cuda()
    fm = torch.full((n_targets * r_hw ** 2, fm_h, fm_w), 10.0).cuda()
    rois = torch.as_tensor([[3.0, 3.0, 0.5, 0.5]]).cuda()
    ans = pr(fm, rois).cpu()
    assert torch.allclose(ans, torch.zeros(len(rois), n_targets, r_hw, r_hw))


def test_psroipool_full_size_cls_head(cuda):
    """BASELINE config 2 at full size: 300 RoIs, 31 targets: adjoint identity + determinism."""
    nT, H, W, k, R = 31, 38, 63, 7, 300
    rois = dev(cases.rois_random(R, 1237), cuda)
    g = torch.Generator(device="cpu").manual_seed(1237)
    fm = torch.randn(nT * k * k, H, W, generator=g).to(cuda)
    go = torch.randn(R, nT, k, k, generator=g).to(cuda)
    out = ps_mod.ps_roipool_forward(fm, rois, nT, k)
    gin = ps_mod.ps_roipool_backward(go, rois, H, W)
    lhs = (out.double() * go.double()).sum().item()
    assert abs(lhs - (fm.double() * gin.double()).sum().item()) <= 1e-5 * abs(lhs)
    assert torch.equal(gin, ps_mod.ps_roipool_backward(go, rois, H, W))
    assert int((gin.abs().sum(dim=(1, 2)) > 0).sum()) == 608


@pytest.mark.parametrize("canonical", [False, True])
@pytest.mark.parametrize("N,nT,H,W,k,R", [(3, 4, 38, 63, 7, 50), (2, 31, 38, 63, 7, 300), (4, 2, 11, 10, 6, 9), (1, 5, 20, 21, 3, 700),
                                           (2, 33, 12, 13, 3, 20), (2, 3, 20, 21, 9, 30)])
def test_psroipool_batched_equals_per_frame(cuda, N, nT, H, W, k, R, canonical):
    """the batched entry points (one set of launches for N frames) against N single-frame calls: the forward is
    bit-identical (both keep the reference's summation order); so is the backward wherever a batch and a single frame run
    the same third-generation kernel (pool_ps3.cu: one owner per element, ascending RoI order -- more than 8 targets; for
    fewer a batch keeps the row-list kernels and agrees within the FP32 tolerance); both match the oracle (R = 700 on a
    20x21 map: many RoIs per pixel)."""
    rng = np.random.default_rng(36)
    rois = np.stack([np.concatenate([cases.rois_edge_cases(H, W), cases.rois_random(R, 40 + n), cases.ROIS_OOB.astype(np.float32)])
                     for n in range(N)]).astype(np.float32)
    Rt = rois.shape[1]
    fm = rng.standard_normal((N, nT * k * k, H, W)).astype(np.float32)
    go = rng.standard_normal((N, Rt, nT, k, k)).astype(np.float32)
    out = ps_mod.ps_roipool_forward_batched(dev(fm, cuda), dev(rois, cuda), nT, k, canonical)
    gin = ps_mod.ps_roipool_backward_batched(dev(go, cuda), dev(rois, cuda), H, W, canonical)
    assert torch.equal(gin, ps_mod.ps_roipool_backward_batched(dev(go, cuda), dev(rois, cuda), H, W, canonical))
    for n in range(N):
        o1 = ps_mod.ps_roipool_forward(dev(fm[n], cuda), dev(rois[n], cuda), nT, k, canonical)
        g1 = ps_mod.ps_roipool_backward(dev(go[n], cuda), dev(rois[n], cuda), H, W, canonical)
        assert torch.equal(out[n], o1)
        if (8 < nT <= 32 and k <= 8) or N == 1:   # same kernel (pool_ps3.cu); 33 targets / r_hw = 9: the older kernels
            assert torch.equal(g1, gin[n])
        else:
            close(g1, gin[n].cpu().numpy(), np.float32)
        assert torch.equal(g1, ps_mod.ps_roipool_backward(dev(go[n], cuda), dev(rois[n], cuda), H, W, canonical))
        close(g1, oracle.psroipool_bwd(go[n], rois[n], H, W, canonical), np.float32)
        np.testing.assert_array_equal(out[n].cpu().numpy(), oracle.psroipool_fwd(fm[n], rois[n], nT, k, canonical))
        close(gin[n], oracle.psroipool_bwd(go[n], rois[n], H, W, canonical), np.float32)


@pytest.mark.parametrize("canonical", [False, True])
@pytest.mark.parametrize("nT,H,W,k,R", [(31, 38, 63, 7, 300), (4, 38, 63, 7, 300), (3, 17, 40, 5, 64), (1, 9, 9, 3, 5), (32, 12, 33, 2, 40)])
def test_psroipool_backward_zero_pattern_and_nonfinite_locality(cuda, nT, H, W, k, R, canonical):
    """properties of the reference's scatter that the float32 backward keeps (pool_ps3.cu: no difference arrays): a pixel
    that no cell covers is an exact 0 (same zero pattern as the oracle), and a non-finite gradient reaches exactly the pixels
    of its own cell (same finite pattern as the oracle).  Also every power-of-two lane split (nT = 1, 3, 4, 31, 32)."""
    rng = np.random.default_rng(77)
    rois = np.concatenate([cases.rois_edge_cases(H, W), cases.rois_random(R, 78), cases.ROIS_OOB.astype(np.float32)]).astype(np.float32)
    go = rng.standard_normal((rois.shape[0], nT, k, k)).astype(np.float32)
    go = np.where(go == 0, np.float32(1), go)
    want = oracle.psroipool_bwd(go, rois, H, W, canonical)
    got = ps_mod.ps_roipool_backward(dev(go, cuda), dev(rois, cuda), H, W, canonical)
    close(got, want, np.float32)
    # a pixel that receives contributions may still cancel to 0 in one summation order and not in another: compare the
    # zero pattern against the COVERAGE (the oracle's result for all-ones gradients), which no order can change
    cover = oracle.psroipool_bwd(np.ones_like(go), rois, H, W, canonical) != 0
    assert not bool((got.cpu().numpy()[~cover] != 0).any())
    bad = go.copy()
    idx = rng.integers(0, rois.shape[0], 4)
    bad[idx[0], 0, 0, 0] = np.inf
    bad[idx[1], nT - 1, k - 1, k - 1] = -np.inf
    bad[idx[2], nT // 2, k // 2, 0] = np.nan
    want_bad = oracle.psroipool_bwd(bad, rois, H, W, canonical)
    got_bad = ps_mod.ps_roipool_backward(dev(bad, cuda), dev(rois, cuda), H, W, canonical).cpu().numpy()
    np.testing.assert_array_equal(np.isfinite(got_bad), np.isfinite(want_bad))


def test_psroipool_batched_module_autograd(cuda):
    N, nT, H, W, k, R = 2, 4, 12, 13, 7, 20
    g = torch.Generator(device="cpu").manual_seed(5)
    fm = torch.randn(N, nT * k * k, H, W, generator=g).to(cuda).requires_grad_(True)
    rois = torch.stack([dev(cases.rois_random(R, 50 + n), cuda) for n in range(N)])
    out = d2t.PSROIPoolBatched(nT, k)(fm, rois)
    w = torch.randn(out.shape, generator=g).to(cuda)
    (out * w).sum().backward()
    fm2 = fm.detach().clone().requires_grad_(True)
    outs = torch.stack([d2t.PSROIPool(nT, k)(fm2[n], rois[n]) for n in range(N)])
    (outs * w).sum().backward()
    assert torch.equal(out, outs)
    close(fm.grad, fm2.grad.cpu().numpy(), np.float32)
    with pytest.raises(ValueError):
        d2t.PSROIPoolBatched(nT, k)(fm[:, :-1].contiguous(), rois)


@pytest.mark.parametrize("canonical", [False, True])
@pytest.mark.parametrize("N,nT,H,W,k,R", [(1, 31, 38, 63, 7, 300), (3, 4, 38, 63, 7, 50), (2, 2, 11, 10, 6, 9), (1, 5, 20, 21, 3, 700)])
def test_psroipool_vote_matches_pool_then_mean(cuda, N, nT, H, W, k, R, canonical):
    """PSROIPool + vote as one operator (rfcn.py:40-41) == PSROIPool -> mean(-1).mean(-1) of the API-parity op, values and
    gradients (FP32 rounding apart: the sum over bins is a shuffle tree, not two means), bitwise reproducible; single frame
    and batch, reference and canonical channel maps, edge-case and out-of-bounds RoIs."""
    rng = np.random.default_rng(46)
    rois = np.stack([np.concatenate([cases.rois_edge_cases(H, W), cases.rois_random(R, 60 + n), cases.ROIS_OOB.astype(np.float32)])
                     for n in range(N)]).astype(np.float32)
    fm = torch.from_numpy(rng.standard_normal((N, nT * k * k, H, W)).astype(np.float32)).to(cuda).requires_grad_(True)
    tr = dev(rois, cuda)
    w = torch.from_numpy(rng.standard_normal((N, rois.shape[1], nT)).astype(np.float32)).to(cuda)
    out = d2t.PSROIPoolVoteFunction.apply(fm, tr, nT, k, canonical)
    (out * w).sum().backward()
    fm2 = fm.detach().clone().requires_grad_(True)
    ref = d2t.PSROIPoolBatchedFunction.apply(fm2, tr, nT, k, canonical).mean(-1).mean(-1)
    (ref * w).sum().backward()
    close(out, ref.detach().cpu().numpy(), np.float32)
    close(fm.grad, fm2.grad.cpu().numpy(), np.float32)
    fm3 = fm.detach().clone().requires_grad_(True)
    out3 = d2t.PSROIPoolVoteFunction.apply(fm3, tr, nT, k, canonical)
    (out3 * w).sum().backward()
    assert torch.equal(out, out3) and torch.equal(fm.grad, fm3.grad)
    if N == 1:   # the (C, H, W) / (|R|, 4) form the R-FCN head calls
        o1 = d2t.PSROIPoolVoteFunction.apply(fm.detach()[0], tr[0], nT, k, canonical)
        assert torch.equal(o1, out.detach()[0])


def test_psroipool_single_frame_and_batched_kernels_agree(cuda):
    """the single-frame entry point runs the per-output kernel, the batched one the channel-owner kernels: bit-identical"""
    nT, H, W, k = 4, 38, 63, 7
    rois = _roipool_rois(H, W, np.float32, R=60)
    fm, _ = cases.pool_inputs(nT * k * k, H, W, (rois.shape[0], nT, k, k), 37, np.float32)
    tf, tr = dev(fm, cuda), dev(rois, cuda)
    assert torch.equal(ps_mod.ps_roipool_forward(tf, tr, nT, k), ps_mod.ps_roipool_forward_batched(tf[None], tr[None], nT, k)[0])


# ------------------------------------------------------------------ error behaviour + autograd wiring
def test_errors_match_reference(cuda):
    fm = torch.rand(1, 2, 8, 8, device=cuda)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        d2t.PointwiseCorrelation(2, 1)(fm.transpose(2, 3), fm)
    with pytest.raises(RuntimeError):
        d2t.ROIPool(3)(torch.rand(2, 8, 8, device=cuda), torch.rand(2, 4, device=cuda).double())   # dtype mismatch
    with pytest.raises(ValueError):
        d2t.PSROIPool(2, 3)(torch.rand(17, 8, 8, device=cuda), torch.rand(2, 4, device=cuda))


def test_autograd_none_grads_and_noncontiguous_grad_out(cuda):
    fm0 = torch.rand(1, 4, 9, 9, device=cuda, requires_grad=True)
    fm1 = torch.rand(1, 4, 9, 9, device=cuda, requires_grad=True)
    out = d2t.PointwiseCorrelationFunction.apply(fm0, fm1, 2, 1)
    out.permute(0, 1, 2, 4, 3).sum().backward()            # non-contiguous grad_out -> .contiguous() like the reference
    assert fm0.grad is not None and fm1.grad is not None
    fm = torch.rand(3, 9, 9, device=cuda, requires_grad=True)
    rois = torch.tensor([[0.5, 0.5, 0.4, 0.4]], device=cuda, requires_grad=True)
    d2t.ROIPool(3)(fm, rois).sum().backward()
    assert rois.grad is None and fm.grad is not None        # no gradient w.r.t. rois (roipool.py:57)


def test_runs_on_a_non_default_stream(cuda):
    fm0, fm1, _ = (dev(a, cuda) for a in cases.corr_inputs(1, 8, 12, 12, 3, seed=14, dtype=np.float32))
    want = pc_mod.pointwise_correlation_forward(fm0, fm1, 3, 1)
    s = torch.cuda.Stream(device=cuda)
    s.wait_stream(torch.cuda.current_stream(cuda))
    with torch.cuda.stream(s):
        got = pc_mod.pointwise_correlation_forward(fm0, fm1, 3, 1)
    s.synchronize()
    assert torch.equal(got, want)


# ------------------------------------------------------------------ golden fixtures (reference kernels' outputs)
def _golden(name):
    p = GOLDEN / f"{name}.npz"
    if not p.exists():
        pytest.skip(f"{p.name} not generated yet")
    return np.load(p)


@pytest.mark.parametrize("case", cases.GOLDEN_CORR, ids=lambda c: c[0])
def test_golden_corr(cuda, case):
    name, B, C, H, W, d, s, dt = case
    g = _golden(name)
    dtype = np.dtype(dt).type
    close(pc_mod.pointwise_correlation_forward(dev(g["fm0"], cuda), dev(g["fm1"], cuda), d, s), g["out"], dtype)
    g0, g1 = pc_mod.pointwise_correlation_backward(dev(g["go"], cuda), dev(g["fm0"], cuda), dev(g["fm1"], cuda), d, s)
    close(g0, g["g0"], dtype)
    close(g1, g["g1"], dtype)


@pytest.mark.parametrize("case", cases.GOLDEN_ROIPOOL, ids=lambda c: c[0])
def test_golden_roipool(cuda, case):
    name, C, H, W, k, R, dt = case
    g = _golden(name)
    dtype = np.dtype(dt).type
    out = rp_mod.roipool_forward(dev(g["fm"], cuda), dev(g["rois"], cuda), k)
    close(out, g["out"], dtype, scale=float(np.nanmax(np.abs(g["out"]))), equal_nan=True)
    if dt == "float32":
        ex = rp_mod.roipool_forward(dev(g["fm"], cuda), dev(g["rois"], cuda), k, exact=True)
        np.testing.assert_array_equal(ex.cpu().numpy(), g["out"])
    close(rp_mod.roipool_backward(dev(g["go"], cuda), dev(g["rois"], cuda), H, W), g["gin"], dtype, equal_nan=True)


@pytest.mark.parametrize("case", cases.GOLDEN_PSROIPOOL, ids=lambda c: c[0])
def test_golden_psroipool(cuda, case):
    name, nT, H, W, k, R, dt = case
    g = _golden(name)
    dtype = np.dtype(dt).type
    out = ps_mod.ps_roipool_forward(dev(g["fm"], cuda), dev(g["rois"], cuda), nT, k)
    if dt == "float32":
        np.testing.assert_array_equal(out.cpu().numpy(), g["out"])
    else:
        close(out, g["out"], dtype, scale=1.0)
    close(ps_mod.ps_roipool_backward(dev(g["go"], cuda), dev(g["rois"], cuda), H, W), g["gin"], dtype)


# ------------------------------------------------------------------ randomized shape sweep (dispatch boundaries)


@pytest.mark.parametrize("case", range(48))
def test_shape_sweep_pooling_vs_oracle(cuda, case):
    """Seeded random shapes for ROIPool / PSROIPool, single frame and batched, against the oracle: r_hw from 1 to 10, target
    counts on both sides of every lane split (1 .. 40), channel counts that are not multiples of the slab, maps from 1x1 to
    wider than 255, RoI counts from 1 to several hundred.  The point is the DISPATCH boundaries between kernel families
    (a batched r_hw = 9 backward used to reach a kernel built for r_hw <= 8)."""
    rng = np.random.default_rng(1000 + case)
    k = int(rng.integers(1, 11))
    nT = int(rng.choice([1, 2, 3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 40]))
    H, W = int(rng.integers(1, 48)), int(rng.integers(1, 80))
    if case % 6 == 5:
        W = int(rng.integers(256, 320))
    R = int(rng.choice([1, 2, 7, 33, 150, 400]))
    C = int(rng.integers(1, 40))
    N = int(rng.integers(1, 4))
    rois = np.stack([np.concatenate([cases.rois_random(R, 5000 + 10 * case + n), cases.rois_edge_cases(H, W)[:6]])
                     for n in range(N)]).astype(np.float32)
    Rt = rois.shape[1]
    # ROIPool (single frame)
    fm = rng.standard_normal((C, H, W)).astype(np.float32)
    go = rng.standard_normal((Rt, C, k, k)).astype(np.float32)
    want = oracle.roipool_fwd(fm, rois[0], k)
    out = rp_mod.roipool_forward(dev(fm, cuda), dev(rois[0], cuda), k)
    close(out, want, np.float32, scale=float(np.nanmax(np.abs(want))) if np.isfinite(want).any() else 1.0, equal_nan=True)
    close(rp_mod.roipool_backward(dev(go, cuda), dev(rois[0], cuda), H, W), oracle.roipool_bwd(go, rois[0], H, W), np.float32)
    # PSROIPool, both channel maps, single frame and batched
    canonical = bool(case & 1)
    sfm = rng.standard_normal((N, nT * k * k, H, W)).astype(np.float32)
    sgo = rng.standard_normal((N, Rt, nT, k, k)).astype(np.float32)
    outb = ps_mod.ps_roipool_forward_batched(dev(sfm, cuda), dev(rois, cuda), nT, k, canonical)
    ginb = ps_mod.ps_roipool_backward_batched(dev(sgo, cuda), dev(rois, cuda), H, W, canonical)
    for n in range(N):
        np.testing.assert_array_equal(outb[n].cpu().numpy(), oracle.psroipool_fwd(sfm[n], rois[n], nT, k, canonical))
        wantg = oracle.psroipool_bwd(sgo[n], rois[n], H, W, canonical)
        close(ginb[n], wantg, np.float32)
        close(ps_mod.ps_roipool_backward(dev(sgo[n], cuda), dev(rois[n], cuda), H, W, canonical), wantg, np.float32)
        np.testing.assert_array_equal(ps_mod.ps_roipool_forward(dev(sfm[n], cuda), dev(rois[n], cuda), nT, k, canonical).cpu().numpy(),
                                      oracle.psroipool_fwd(sfm[n], rois[n], nT, k, canonical))


@pytest.mark.parametrize("case", range(32))
def test_shape_sweep_correlation_vs_oracle(cuda, case):
    """Seeded random shapes for the correlation: displacements 1 .. 8, strides 1 .. 3, channel counts on both sides of the
    tensor-core threshold (128) and not multiples of the stage size, maps smaller than a tile and not multiples of it."""
    rng = np.random.default_rng(2000 + case)
    d = int(rng.choice([1, 2, 3, 4, 8, 8]))
    stride = int(rng.choice([1, 1, 1, 2, 3]))
    C = int(rng.choice([1, 3, 31, 64, 127, 128, 129, 160, 257]))
    B = int(rng.integers(1, 4))
    H, W = int(rng.integers(1, 30)), int(rng.integers(1, 40))
    fm0, fm1, go = cases.corr_inputs(B, C, H, W, d, 3000 + case)
    want = oracle.corr_fwd(fm0, fm1, d, stride)
    out = pc_mod.pointwise_correlation_forward(dev(fm0, cuda), dev(fm1, cuda), d, stride)
    close(out, want, np.float32, scale=float(np.abs(want).max()) if want.size else 1.0)
    g0, g1 = pc_mod.pointwise_correlation_backward(dev(go, cuda), dev(fm0, cuda), dev(fm1, cuda), d, stride)
    w0, w1 = oracle.corr_bwd(go, fm0, fm1, d, stride)
    close(g0, w0, np.float32)
    close(g1, w1, np.float32)
