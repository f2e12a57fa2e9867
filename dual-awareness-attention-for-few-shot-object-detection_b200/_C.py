"""Drop-in for the reference's pybind operator module `model._C`
(lib/model/csrc/vision.cpp:7-13): same function names, argument order and return conventions, backed
by the C ABI in libdana_b200.so.  CUDA tensors only: this build has no CPU path, and says so with the
reference's own error wording ("Not implemented on the CPU", csrc/ROIAlign.h:44)."""
import torch

from . import ops


def _cuda_only(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s: Not implemented on the CPU (dana_b200 is a CUDA-only build)" % what)


def nms(dets, scores, threshold, strict_gt=False):
    """nms(dets[N,4] f32, scores[N] f32, threshold) -> LongTensor[M] of kept input indices, ascending
    (csrc/nms.h:10-28).  Empty input returns an empty CPU long tensor, like the reference (:17-18).

    Comparison semantics: a box is suppressed when IoU >= threshold -- the reference's CPU operator
    (csrc/cpu/nms_cpu.cpp:60), which is the parity target (bit-exact keep indices).  The reference's CUDA operator
    compares with a strict `>` (csrc/cuda/nms.cu:60), so the two reference back ends disagree when an IoU equals the
    threshold exactly; `strict_gt=True` selects the CUDA operator's rule (IoU > t  <=>  IoU >= nextafter(t, +inf) in
    fp32, so the same kernel serves both)."""
    _cuda_only(dets, "nms")
    if dets.dtype != torch.float32:
        raise RuntimeError("nms: only float32 boxes are supported (csrc/cuda/nms.cu:71)")
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.long, device="cpu")
    thr = float(threshold)
    if strict_gt:
        import numpy as np
        thr = float(np.nextafter(np.float32(thr), np.float32(np.inf)))
    return ops.nms(dets, scores, thr)


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    """-> Tensor[R,C,ph,pw] (csrc/ROIAlign.h:11-27)."""
    _cuda_only(input, "roi_align_forward")
    return ops.roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio)


def roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels, height, width,
                       sampling_ratio):
    """-> Tensor[B,C,H,W] (csrc/ROIAlign.h:29-45)."""
    _cuda_only(grad, "roi_align_backward")
    return ops.roi_align_backward(grad, rois, spatial_scale, pooled_height, pooled_width, batch_size, channels,
                                  height, width, sampling_ratio)


def roi_pool_forward(input, rois, spatial_scale, pooled_height, pooled_width):
    raise RuntimeError("roi_pool_forward: POOLING_MODE 'pool' is not built (every shipped cfgs/*.yml uses 'align')")


def roi_pool_backward(grad, input, rois, argmax, spatial_scale, pooled_height, pooled_width, batch_size, channels,
                      height, width):
    raise RuntimeError("roi_pool_backward: POOLING_MODE 'pool' is not built (every shipped cfgs/*.yml uses 'align')")
