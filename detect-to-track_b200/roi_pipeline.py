"""Device-side region-proposal pipeline: anchor decode + confidence filter + NMS without leaving the GPU
(csrc/roi_pipeline.cu).

Extension.  The reference does this step on the host with `ml_utils`
(`frcnn_box_decode` + `region_filter`, trainer.py:178-190 / inference.py:78-84): two device->numpy copies of the RPN
outputs and one numpy->device copy of the surviving boxes per frame.  `ml_utils` is not available, so the semantics here
are the standard Faster R-CNN ones for fractional (centre_i, centre_j, height, width) boxes and parity with the
reference's filter is UNPINNED (stated in DESIGN.md).
"""
from typing import Tuple

import torch
from torch import Tensor

from . import _lib


@torch.no_grad()
def propose_regions(anchors: Tensor, offsets: Tensor, conf: Tensor, conf_thresh: float = 0.3, iou_thresh: float = 0.5,
                    max_rois: int = 3000, pre_nms: int = 6000) -> Tuple[Tensor, Tensor]:
    """anchors (|A|, 4) ijhw, offsets (|A|, 4) RPN regression output, conf (|A|) objectness ->
    rois (max_rois, 4) ijhw in descending score order, zero rows past `count`; count: 0-dim int32 tensor ON THE DEVICE
    (no host synchronisation happens here; `rois[:int(count)]` is the only sync a caller needs).
    Defaults: cfg/default.yaml:22-25 (TRAIN_ROI_CONF_THRESH 0.3, TRAIN_NMS_IOU_THRESH 0.5, TRAIN_MAX_ROIS 3000)."""
    for t, name in ((anchors, "anchors"), (offsets, "offsets"), (conf, "conf")):
        _lib.check_input(t, name)
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name} must be float32")
    A = anchors.size(0)
    if tuple(anchors.shape) != (A, 4) or tuple(offsets.shape) != (A, 4) or tuple(conf.shape) != (A,):
        raise RuntimeError("expected anchors (|A|, 4), offsets (|A|, 4), conf (|A|)")
    lib = _lib.lib()
    dev = anchors.device
    with torch.cuda.device(dev):
        boxes = torch.empty((A, 4), dtype=torch.float32, device=dev)
        scores = torch.empty((A,), dtype=torch.float32, device=dev)
        stream = _lib.stream_ptr(dev)
        _lib.check(lib.d2t_roi_decode_filter_f32(anchors.data_ptr(), offsets.data_ptr(), conf.data_ptr(), boxes.data_ptr(),
                                                 scores.data_ptr(), A, float(conf_thresh), stream), "roi_decode_filter")
        sorted_scores, order = torch.sort(scores, descending=True, stable=True)     # device sort: plumbing
        rois = torch.empty((max_rois, 4), dtype=torch.float32, device=dev)
        count = torch.empty((), dtype=torch.int32, device=dev)
        pre = min(pre_nms, 16384)
        nbytes = lib.d2t_roi_nms_workspace_bytes(A, pre)
        ws, ws_ptr, ws_n = _lib.workspace(nbytes, dev)
        _lib.check(lib.d2t_roi_nms_f32(boxes.data_ptr(), order.data_ptr(), sorted_scores.data_ptr(), rois.data_ptr(),
                                       count.data_ptr(), A, pre, max_rois, float(iou_thresh), ws_ptr, ws_n, stream), "roi_nms")
    return rois, count
