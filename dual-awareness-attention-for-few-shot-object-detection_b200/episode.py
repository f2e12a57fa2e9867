"""Episode construction on the GPU -- host-side mirror of the reference's loaders for the
one step that precedes the hot path: turning a decoded image into the `im_data` / `support_ims` tensors.

  prep_im_for_blob(im, pixel_means, target_size, max_size)      lib/model/utils/blob.py:35-52
  support_from_box(prepared_im, box, support_scale, 320)        lib/roi_data_layer/fs_loader.py:113-138
  support_from_image(im, pixel_means, 320)                      lib/roi_data_layer/inference_loader.py:95-109

Inputs are device tensors in the reference's host layout (HWC, BGR, uint8 or float32); outputs are CHW float32 device
tensors, i.e. what the loaders hand to `DAnARCNN.forward` after their `permute`.  All arithmetic runs in
`dana_episode_resize` (csrc/episode.cuh); there is no CPU path (CPU tensors raise)."""
import numpy as np
import torch

from . import _lib
from ._lib import check
from .ops import _count, _need_cuda, _p, _stream


def cv_round(v):
    """cvRound: round half to even (what cv2.resize uses for dsize when fx / fy are given)."""
    return int(np.rint(v))


def resize_into(src, crop, scales, dst_hw, out, means=(0.0, 0.0, 0.0)):
    """Low-level call: src [H,W,3] u8/f32 device tensor, crop (x, y, w, h), scales (sx, sy) source px per dst px,
    dst_hw (h, w) resized extent, out [3,OH,OW] fp32 (planes written in full: image + zero padding)."""
    _need_cuda(src, out)
    assert src.dim() == 3 and src.shape[2] == 3 and src.stride(2) == 1 and src.stride(1) == 3
    assert src.dtype in (torch.uint8, torch.float32) and out.dtype == torch.float32 and out.is_contiguous()
    _count(1)
    check(_lib.load().dana_episode_resize(_p(src), int(src.dtype == torch.float32), src.shape[0], src.shape[1],
                                          src.stride(0), int(crop[0]), int(crop[1]), int(crop[2]), int(crop[3]),
                                          float(scales[0]), float(scales[1]), int(dst_hw[1]), int(dst_hw[0]),
                                          float(means[0]), float(means[1]), float(means[2]), _p(out), out.shape[1],
                                          out.shape[2], _stream()), "dana_episode_resize")
    return out


def prep_im_for_blob(im, pixel_means, target_size, max_size=None, out=None):
    """blob.py:35-52.  im [H,W,3] (BGR, u8 or f32, device) -> (CHW fp32 [3,H',W'], im_scale).
    `max_size` is accepted and ignored exactly as in the reference (the cap is commented out there, :45-47).
    `out` may be a larger [3,OH,OW] canvas (im_list_to_blob's zero-padded blob, blob.py:17-33)."""
    h, w = int(im.shape[0]), int(im.shape[1])
    im_scale = float(target_size) / float(min(h, w))
    dst_w, dst_h = cv_round(w * im_scale), cv_round(h * im_scale)     # cv2.resize(dsize=None, fx, fy)
    if out is None:
        out = torch.empty((3, dst_h, dst_w), dtype=torch.float32, device=im.device)
    means = np.asarray(pixel_means, dtype=np.float32).reshape(-1)[:3]
    resize_into(im, (0, 0, w, h), (1.0 / im_scale, 1.0 / im_scale), (dst_h, dst_w), out, means)
    return out, im_scale


def _fit(box_h, box_w, target):
    """Long side -> target, the other side int(side * scale) (fs_loader.py:126-133)."""
    if box_h > box_w:
        scale = float(target) / float(box_h)
        return target, int(box_w * scale)          # (dst_h, dst_w)
    scale = float(target) / float(box_w)
    return int(box_h * scale), target


def support_from_box(prepared_im, box, support_scale, target_size=320, out=None):
    """fs_loader.py:117-138.  prepared_im [H,W,3] fp32 (already mean-subtracted and scaled, i.e. the `data` blob of
    the support image), box (x1,y1,x2,y2) in original pixels.  -> [3,target,target] fp32, zero padded."""
    b = (np.asarray(box, dtype=np.float64) * support_scale).astype(np.int16)
    x_min, y_min, x_max, y_max = int(b[0]), int(b[1]), int(b[2]), int(b[3])
    box_h, box_w = y_max - y_min, x_max - x_min
    # numpy slicing semantics of support_im[y_min:y_max+1, x_min:x_max+1]
    ih, iw = int(prepared_im.shape[0]), int(prepared_im.shape[1])
    cx0, cy0 = max(x_min, 0), max(y_min, 0)
    cw, ch = min(x_max + 1, iw) - cx0, min(y_max + 1, ih) - cy0
    dst_h, dst_w = _fit(box_h, box_w, target_size)
    if out is None:
        out = torch.empty((3, target_size, target_size), dtype=torch.float32, device=prepared_im.device)
    resize_into(prepared_im, (cx0, cy0, cw, ch), (cw / float(dst_w), ch / float(dst_h)), (dst_h, dst_w), out)
    return out


def support_from_image(im, pixel_means, target_size=320, out=None):
    """inference_loader.py:95-109: the whole (pre-cropped) support image, mean-subtracted at its own size
    (prep_im_for_blob with target = min side is the identity resize), long side -> target, zero padded."""
    h, w = int(im.shape[0]), int(im.shape[1])
    dst_h, dst_w = _fit(h, w, target_size)
    if out is None:
        out = torch.empty((3, target_size, target_size), dtype=torch.float32, device=im.device)
    means = np.asarray(pixel_means, dtype=np.float32).reshape(-1)[:3]
    resize_into(im, (0, 0, w, h), (w / float(dst_w), h / float(dst_h)), (dst_h, dst_w), out, means)
    return out
