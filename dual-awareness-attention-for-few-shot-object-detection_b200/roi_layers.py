"""Mirror of the reference's `model.roi_layers` package (lib/model/roi_layers/{nms,roi_align}.py):
`nms`, `roi_align`, `ROIAlign` with the same call signatures, on top of the dana_b200 `_C` module."""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import _C

nms = _C.nms


class _ROIAlign(Function):
    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio):
        out_h, out_w = _pair(output_size)
        ctx.save_for_backward(roi)
        ctx.geom = (out_h, out_w, spatial_scale, sampling_ratio, tuple(input.shape))
        return _C.roi_align_forward(input, roi, spatial_scale, out_h, out_w, sampling_ratio)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (roi,) = ctx.saved_tensors
        out_h, out_w, spatial_scale, sampling_ratio, (b, c, h, w) = ctx.geom
        grad_in = _C.roi_align_backward(grad_output, roi, spatial_scale, out_h, out_w, b, c, h, w, sampling_ratio)
        return grad_in, None, None, None, None


roi_align = _ROIAlign.apply


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    def __repr__(self):
        return "ROIAlign(output_size=%s, spatial_scale=%s, sampling_ratio=%s)" % (
            self.output_size, self.spatial_scale, self.sampling_ratio)
