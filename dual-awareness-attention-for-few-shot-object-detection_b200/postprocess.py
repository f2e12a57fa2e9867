"""Detections after `DAnARCNN.forward`: the step that follows the hot path in the
reference's inference.py:108-142 -- de-normalise the box deltas, decode against the rois, clip, undo the image
scale, threshold the fg score and run NMS(cfg.TEST.NMS) -- as one device pipeline (decode kernel, radix sort,
greedy NMS), no host round trip per image."""
import ctypes

import torch

from . import _lib, ops
from .config import cfg


def detections(rois, cls_prob, bbox_pred, im_info, score_thresh=0.05, nms_thresh=None, stds=None, means=None):
    """rois [B,R,5], cls_prob [B*R,2], bbox_pred [B*R,4], im_info [B,3] (CUDA) ->
    (dets [B,R,5] = (x1,y1,x2,y2,score) in kept order, zero padded; counts [B] int32)."""
    if not rois.is_cuda:
        raise _lib.DanaError("dana_b200.postprocess needs CUDA tensors (there is no CPU fallback)")
    b, r, _ = rois.shape
    nms_thresh = cfg.TEST.NMS if nms_thresh is None else nms_thresh
    if stds is None:
        stds = cfg.TRAIN.BBOX_NORMALIZE_STDS if cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED else (1.0, 1.0, 1.0, 1.0)
    if means is None:
        means = cfg.TRAIN.BBOX_NORMALIZE_MEANS if cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED else (0.0, 0.0, 0.0, 0.0)
    c_stds = (ctypes.c_float * 4)(*[float(v) for v in stds])
    c_means = (ctypes.c_float * 4)(*[float(v) for v in means])
    lib = _lib.load()
    wsb = lib.dana_detections_workspace_bytes(b, r)
    dev = rois.device
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    dets = torch.empty((b, r, 5), dtype=torch.float32, device=dev)
    counts = torch.empty((b,), dtype=torch.int32, device=dev)
    ops._count(5)
    _lib.check(lib.dana_detections(ops._p(rois.contiguous().float()), ops._p(cls_prob.contiguous().float()),
                                   ops._p(bbox_pred.contiguous().float()), ops._p(im_info.contiguous().float()), b, r,
                                   ctypes.cast(c_stds, ctypes.c_void_p), ctypes.cast(c_means, ctypes.c_void_p),
                                   float(score_thresh), float(nms_thresh), ops._p(dets), ops._p(counts), ops._p(ws), wsb,
                                   ops._stream()), "dana_detections")
    return dets, counts
