"""Device-side RoI pipeline (csrc/roi_pipeline.cu) against a plain numpy statement of the same (standard Faster R-CNN)
semantics.  The reference's own filter lives in ml_utils, whose source is absent: parity with it is unpinned."""
import numpy as np
import pytest
import torch

from detect_to_track_b200 import roi_pipeline as rp

pytestmark = pytest.mark.gpu


def numpy_pipeline(anchors, offsets, conf, conf_thresh, iou_thresh, max_rois, pre_nms):
    a, d = anchors.astype(np.float32), offsets.astype(np.float32)
    boxes = np.stack([a[:, 0] + d[:, 0] * a[:, 2], a[:, 1] + d[:, 1] * a[:, 3],
                      a[:, 2] * np.exp(d[:, 2]), a[:, 3] * np.exp(d[:, 3])], 1).astype(np.float32)
    idx = np.argsort(-conf, kind="stable")
    idx = [i for i in idx if conf[i] > conf_thresh][:pre_nms]

    def iou(p, q):
        pi0, pj0, pi1, pj1 = p[0] - p[2] / 2, p[1] - p[3] / 2, p[0] + p[2] / 2, p[1] + p[3] / 2
        qi0, qj0, qi1, qj1 = q[0] - q[2] / 2, q[1] - q[3] / 2, q[0] + q[2] / 2, q[1] + q[3] / 2
        inter = max(min(pi1, qi1) - max(pi0, qi0), 0) * max(min(pj1, qj1) - max(pj0, qj0), 0)
        uni = p[2] * p[3] + q[2] * q[3] - inter
        return inter / uni if uni > 0 else 0.0

    kept = []
    for i in idx:
        if len(kept) >= max_rois:
            break
        if all(iou(boxes[i], boxes[j]) <= iou_thresh for j in kept):
            kept.append(i)
    return boxes[kept]


@pytest.mark.parametrize("A,max_rois,pre_nms,conf_thresh", [(700, 300, 6000, 0.3), (2000, 50, 900, 0.1), (65, 3000, 6000, 0.9), (300, 10, 100, 2.0)])
def test_propose_regions_matches_numpy(cuda, A, max_rois, pre_nms, conf_thresh):
    rng = np.random.default_rng(A)
    anchors = np.concatenate([rng.uniform(0.1, 0.9, (A, 2)), rng.uniform(0.05, 0.5, (A, 2))], 1).astype(np.float32)
    offsets = (rng.standard_normal((A, 4)) * 0.15).astype(np.float32)
    conf = rng.uniform(0, 1, A).astype(np.float32)
    conf[::7] = conf[3]                       # ties: the lower anchor index goes first (stable sort)
    want = numpy_pipeline(anchors, offsets, conf, conf_thresh, 0.5, max_rois, pre_nms)
    rois, count = rp.propose_regions(torch.from_numpy(anchors).to(cuda), torch.from_numpy(offsets).to(cuda),
                                     torch.from_numpy(conf).to(cuda), conf_thresh, 0.5, max_rois, pre_nms)
    n = int(count)
    assert n == len(want)
    np.testing.assert_allclose(rois[:n].cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    assert not bool(rois[n:].any())
    again, c2 = rp.propose_regions(torch.from_numpy(anchors).to(cuda), torch.from_numpy(offsets).to(cuda),
                                   torch.from_numpy(conf).to(cuda), conf_thresh, 0.5, max_rois, pre_nms)
    assert torch.equal(rois, again) and int(c2) == n


def test_proposals_feed_the_heads_without_a_host_copy_of_the_anchors(cuda):
    """the surviving boxes go straight into PSROIPool / ROIPool on the device"""
    import detect_to_track_b200 as d2t
    g = torch.Generator(device="cpu").manual_seed(3)
    A, H, W = 15 * 8 * 10, 8, 10
    anchors = torch.cat([torch.rand(A, 2, generator=g) * 0.6 + 0.2, torch.rand(A, 2, generator=g) * 0.3 + 0.1], 1).to(cuda)
    rois, count = rp.propose_regions(anchors, (torch.randn(A, 4, generator=g) * 0.05).to(cuda), torch.rand(A, generator=g).to(cuda),
                                     max_rois=64)
    r = rois[:int(count)].contiguous()
    out = d2t.PSROIPool(2, 7)(torch.randn(98, H, W, generator=g).to(cuda), r)
    assert tuple(out.shape) == (int(count), 2, 7, 7) and bool(torch.isfinite(out).all())
