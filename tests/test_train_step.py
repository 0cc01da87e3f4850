"""BASELINE config 5 (detect_to_track_b200/train_step.py): the restated model glue and losses against the reference's own
files where they can be imported (loss.py needs only torch), the data-parallel gradient averaging on CPU (gloo, 2 ranks),
and -- on the GPU -- one small training step through the CUDA ops, reference composition vs fused tracker."""
import importlib.util
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from detect_to_track_b200 import train_step as ts  # noqa: E402

REF_LOSS = Path("/root/reference/detect_to_track/loss.py")


@pytest.mark.skipif(not REF_LOSS.exists(), reason="reference not mounted")
def test_losses_match_the_reference_loss_module():
    spec = importlib.util.spec_from_file_location("ref_loss", REF_LOSS)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(3)
    c_hat = torch.softmax(torch.randn(2, 50, 5, generator=g), -1)
    c_star = torch.randint(0, 5, (2, 50), generator=g)
    b_hat, b_star = torch.randn(2, 50, 4, generator=g), torch.randn(2, 50, 4, generator=g)
    lw = torch.rand(2, 50, generator=g)
    torch.testing.assert_close(ts.focal_loss(c_hat, c_star), ref.FocalLoss(0.25, 2.0)(c_hat, c_star))
    torch.testing.assert_close(ts.bbox_loss(b_hat, b_star, c_star), ref.BBoxLoss()(b_hat, b_star, c_star))
    o_ref, b_ref = ref.RPNLoss(0.25, 2.0)(lw, c_hat, c_star, b_hat, b_star)
    torch.testing.assert_close((lw * ts.focal_loss(c_hat, c_star)).mean(), o_ref)
    torch.testing.assert_close(ts.bbox_loss(b_hat, b_star, c_star).mean(), b_ref)
    c_ref, bb_ref = ref.RCNNLoss(0.25, 2.0)(c_hat[0], c_star[0], b_hat[0], b_star[0])
    torch.testing.assert_close(ts.focal_loss(c_hat[:1], c_star[:1]).mean(), c_ref)
    torch.testing.assert_close(ts.bbox_loss(b_hat[:1], b_star[:1], c_star[:1]).mean(), bb_ref)
    t_ref = ref.TrackLoss()(b_hat[0], b_star[0])
    torch.testing.assert_close(torch.nn.functional.smooth_l1_loss(b_hat[0], b_star[0], reduction="none").mean(), t_ref)


def test_module_layout_follows_the_reference():
    """attribute names / state_dict keys of detect_track.py:52-55, rpn.py:19-21, rfcn.py, correlation_tracker.py; the
    backbone returns c3 / c4 / c5 at strides 8 / 16 / 16 (layer4 dilated, resnet.py:19-23); stages below 3 are frozen."""
    m = ts.DetectTrackModule("resnet50", 3, n_classes=30, k=7, d_max=8, r_hw=7)
    keys = set(m.state_dict())
    for k in ("rpn.conv.weight", "rpn.cls_fc.bias", "rcnn.channel_reduce.weight", "rcnn.cls_head.sm_conv.weight",
              "rcnn.reg_head.sm_conv.bias", "c_tracker.reg_fc.weight", "backbone.1.layer3.0.conv1.weight"):
        assert k in keys, k
    assert m.c_tracker.reg_fc.in_features == (3 * 289 + 2 * 512) * 49 == 92659
    assert m.rcnn.cls_head.sm_conv.out_channels == 31 * 49 and m.rpn.cls_fc.out_channels == 2 * ts.N_ANCHORS
    fm = m.backbone(torch.rand(1, 3, 64, 96))
    assert [tuple(fm[k].shape) for k in ("c3", "c4", "c5")] == [(1, 512, 8, 12), (1, 1024, 4, 6), (1, 2048, 4, 6)]
    frozen = [n for n, p in m.backbone.named_parameters() if not p.requires_grad]
    assert any("layer2" in n for n in frozen) and not any("layer3" in n or "layer4" in n for n in frozen)


def test_fast_backbone_is_the_same_function():
    """resnet_backbone(fast=True) -- conv + frozen BN folded into one convolution, channels_last stack -- has the same
    state_dict keys and computes the same pyramid and the same gradients (float64: to rounding of 1e-12)."""
    torch.manual_seed(0)
    a = ts.resnet_backbone("resnet50", 3, fast=False).double()
    b = ts.resnet_backbone("resnet50", 3, fast=True).double()
    assert list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())
    x = torch.rand(2, 3, 64, 96, dtype=torch.float64)
    fa, fb = a(x), b(x)
    for k in ("c3", "c4", "c5"):
        assert fb[k].is_contiguous()
        torch.testing.assert_close(fb[k], fa[k], rtol=1e-12, atol=1e-12)
    sum((v ** 2).sum() for v in fa.values()).backward()
    sum((v ** 2).sum() for v in fb.values()).backward()
    gb = dict(b.named_parameters())
    n_grads = 0
    for n, p in a.named_parameters():
        assert (p.grad is None) == (gb[n].grad is None), n
        if p.grad is not None:
            n_grads += 1
            assert float((p.grad - gb[n].grad).norm()) <= 1e-12 * float(p.grad.norm()) + 1e-300, n
    assert n_grads > 0
    # a changed frozen-BN buffer or frozen weight invalidates the cached scaled weights
    with torch.no_grad():
        for m in (a, b):
            m[1].layer1[0].bn1.weight.mul_(0.5)
            m[1].layer1[0].conv1.weight.mul_(1.5)
    torch.testing.assert_close(b(x)["c5"], a(x)["c5"], rtol=1e-12, atol=1e-12)


def test_synthetic_batch_shapes():
    b = ts.synthetic_batch(2, 64, 96, 11, 30, seed=5)
    assert len(b) == 2 and tuple(b[0]["x"].shape) == (2, 3, 64, 96)
    A = ts.N_ANCHORS * 4 * 6
    assert tuple(b[0]["c_star_rpn"].shape) == (2, A) and tuple(b[0]["rois"].shape) == (2, 11, 4)
    r = b[1]["track_rois"]
    assert bool(((r[:, :2] - r[:, 2:] / 2) >= 0).all()) and bool(((r[:, :2] + r[:, 2:] / 2) <= 1).all())


class _RpnOnlyStep(torch.nn.Module):
    """the RPN part of DetectTrackTrainStep.pair_losses (no custom CUDA op): what the gloo test can run on CPU"""

    def __init__(self):
        super().__init__()
        self.rpn = ts.RPN(16, ts.N_ANCHORS)

    def forward(self, item):
        o_hat, b_hat, _ = self.rpn(item["c4"])
        o_loss = (item["lw_rpn"] * ts.focal_loss(o_hat, item["c_star_rpn"])).mean()
        return o_loss + ts.bbox_loss(b_hat, item["b_star_rpn"], item["c_star_rpn"]).mean()


def _item(seed):
    g = torch.Generator().manual_seed(seed)
    A = ts.N_ANCHORS * 4 * 6
    return {"c4": torch.randn(2, 16, 4, 6, generator=g), "lw_rpn": torch.rand(2, A, generator=g),
            "c_star_rpn": (torch.rand(2, A, generator=g) < 0.2).long(), "b_star_rpn": torch.randn(2, A, 4, generator=g)}


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(7)
    net = _RpnOnlyStep()
    ddp = torch.nn.parallel.DistributedDataParallel(net, bucket_cap_mb=1)
    ddp(_item(100 + rank)).backward()
    q.put((rank, [p.grad.numpy().copy() for p in net.parameters()]))   # by value (tensors would travel as shared-memory handles)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradients_are_the_mean_over_pair_shards_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(7)
    net = _RpnOnlyStep()
    grads = []
    for r in range(2):
        net.zero_grad()
        net(_item(100 + r)).backward()
        grads.append([p.grad.clone() for p in net.parameters()])
    for g0, g1, a, b in zip(res[0], res[1], grads[0], grads[1]):
        torch.testing.assert_close(torch.from_numpy(g0), torch.from_numpy(g1))
        torch.testing.assert_close(torch.from_numpy(g0), (a + b) / 2)


@pytest.mark.gpu
def test_one_training_step_on_the_gpu_reference_composition_vs_fused(cuda):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    losses = {}
    for fused in (False, True):
        torch.manual_seed(11)
        model = ts.DetectTrackModule("resnet50", 3, fused_tracker=fused).to(cuda)
        stepm = ts.DetectTrackTrainStep(model)
        opt = ts.make_optimizer(stepm)
        batch = ts.synthetic_batch(2, 160, 192, 24, 30, seed=21, device=cuda)
        before = model.c_tracker.reg_fc.weight.detach().clone()
        loss, terms = stepm(batch)
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in stepm.named_parameters() if p.grad is not None}
        opt.step()
        assert bool(torch.isfinite(loss)) and bool(torch.isfinite(terms).all())
        assert not torch.equal(before, model.c_tracker.reg_fc.weight)
        assert "model.backbone.1.layer3.0.conv1.weight" in grads and "model.backbone.1.layer2.0.conv1.weight" not in grads
        losses[fused] = (loss.detach(), terms, grads)
    torch.testing.assert_close(losses[True][1], losses[False][1], rtol=1e-4, atol=1e-6)
    for n, gr in losses[False][2].items():
        ref_scale = float(gr.abs().max()) or 1.0
        torch.testing.assert_close(losses[True][2][n], gr, rtol=1e-3, atol=1e-4 * ref_scale, msg=n)


@pytest.mark.gpu
def test_fast_backbone_and_batched_step_match_the_reference_composition_on_the_gpu(cuda):
    """fast_backbone (folded frozen BN, channels_last stack and R-FCN conv) + batch_backbone (backbone / RPN once per minibatch)
    against the per-pair NCHW composition: same losses, same gradients (FP32 convolutions, different cuDNN kernels)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {}
    for fast in (False, True):
        torch.manual_seed(11)
        model = ts.DetectTrackModule("resnet50", 3, fused_tracker=False, fast_backbone=fast).to(cuda)
        stepm = ts.DetectTrackTrainStep(model, batch_backbone=fast)
        batch = ts.synthetic_batch(3, 160, 192, 24, 30, seed=21, device=cuda)
        loss, terms = stepm(batch)
        loss.backward()
        res[fast] = (terms, {n: p.grad.detach().clone() for n, p in stepm.named_parameters() if p.grad is not None})
    assert list(res[True][1]) == list(res[False][1])
    torch.testing.assert_close(res[True][0], res[False][0], rtol=1e-4, atol=1e-6)
    # every convolution runs a different cuDNN kernel (NHWC, other batch size), so single elements differ by more than an
    # elementwise FP32 tolerance wherever a ReLU input sits at zero; per parameter the gradients agree in norm
    worst = max((float((res[True][1][n] - gr).norm() / (gr.norm() + 1e-30)), n) for n, gr in res[False][1].items())
    assert worst[0] <= 2e-3, worst
