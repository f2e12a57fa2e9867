"""ORACLE (test infrastructure, never imported by the product path): numpy restatement of the reference's episode
construction -- cv2.resize(INTER_LINEAR) on float32 images as OpenCV computes it (resizeGeneric_ with
HResizeLinear / VResizeLinear: row pass then column pass, coordinates (float)((d + .5) * scale - .5)), and the loader
logic around it: prep_im_for_blob (lib/model/utils/blob.py:35-52), the support crop of fs_loader.py:113-138 and
inference_loader.py:95-109.  Pinned against cv2 itself by tests/test_oracle_pins.py (golden vectors made by
oracle/make_golden_episode.py with the real cv2)."""
import numpy as np


def resize_linear_f32(src, dst_w, dst_h, scale_x=None, scale_y=None):
    """cv2.resize(src [H,W,C] float32, (dst_w, dst_h), INTER_LINEAR); scale_* = source px per dst px
    (cv2: 1/fx when fx is given, ssize/dsize otherwise)."""
    src = np.asarray(src, dtype=np.float32)
    h, w = src.shape[:2]
    sx_ = (w / float(dst_w)) if scale_x is None else scale_x
    sy_ = (h / float(dst_h)) if scale_y is None else scale_y
    dx = np.arange(dst_w, dtype=np.float64)
    fx = ((dx + 0.5) * sx_ - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = fx - sx.astype(np.float32)
    lo = sx < 0
    fx[lo], sx[lo] = 0.0, 0
    hi = sx >= w - 1
    fx[hi], sx[hi] = 0.0, w - 1
    sx1 = np.minimum(sx + 1, w - 1)
    dy = np.arange(dst_h, dtype=np.float64)
    fy = ((dy + 0.5) * sy_ - 0.5).astype(np.float32)
    sy = np.floor(fy).astype(np.int64)
    fy = fy - sy.astype(np.float32)
    y0 = np.clip(sy, 0, h - 1)
    y1 = np.clip(sy + 1, 0, h - 1)
    ax0 = (np.float32(1.0) - fx)[None, :, None]
    ax1 = fx[None, :, None]
    rows = src[:, sx, :] * ax0 + src[:, sx1, :] * ax1                      # row pass, float32
    by0 = (np.float32(1.0) - fy)[:, None, None]
    by1 = fy[:, None, None]
    return (rows[y0] * by0 + rows[y1] * by1).astype(np.float32)


def cv_round(v):
    return int(np.rint(v))


def prep_im_for_blob(im, pixel_means, target_size):
    """blob.py:35-52 (the MAX_SIZE cap is commented out in the reference)."""
    im = im.astype(np.float32, copy=True)
    im -= np.asarray(pixel_means, dtype=np.float32).reshape(1, 1, 3)
    h, w = im.shape[:2]
    im_scale = float(target_size) / float(min(h, w))
    dst_w, dst_h = cv_round(w * im_scale), cv_round(h * im_scale)
    return resize_linear_f32(im, dst_w, dst_h, 1.0 / im_scale, 1.0 / im_scale), im_scale


def _fit(box_h, box_w, target):
    if box_h > box_w:
        scale = float(target) / float(box_h)
        return target, int(box_w * scale)
    scale = float(target) / float(box_w)
    return int(box_h * scale), target


def support_from_box(prepared_im, box, support_scale, target_size=320):
    """fs_loader.py:117-138 -> [3, target, target] float32."""
    b = (np.asarray(box, dtype=np.float64) * support_scale).astype(np.int16)
    x_min, y_min, x_max, y_max = int(b[0]), int(b[1]), int(b[2]), int(b[3])
    box_h, box_w = y_max - y_min, x_max - x_min
    crop = prepared_im[y_min:y_max + 1, x_min:x_max + 1, :]
    dst_h, dst_w = _fit(box_h, box_w, target_size)
    r = resize_linear_f32(crop, dst_w, dst_h)
    out = np.zeros((3, target_size, target_size), dtype=np.float32)
    out[:, :r.shape[0], :r.shape[1]] = np.transpose(r, (2, 0, 1))
    return out


def support_from_image(im, pixel_means, target_size=320):
    """inference_loader.py:95-109 -> [3, target, target] float32."""
    p, _ = prep_im_for_blob(im, pixel_means, min(im.shape[0], im.shape[1]))
    dst_h, dst_w = _fit(p.shape[0], p.shape[1], target_size)
    r = resize_linear_f32(p, dst_w, dst_h)
    out = np.zeros((3, target_size, target_size), dtype=np.float32)
    out[:, :r.shape[0], :r.shape[1]] = np.transpose(r, (2, 0, 1))
    return out
