"""Base anchor windows of the RPN (host-side, computed once per engine).

Same enumeration as the reference (lib/model/rpn/generate_anchors.py:45-105): start from the
0-based window (0, 0, base-1, base-1), enumerate aspect ratios (rounded widths / heights), then
scales; the result is ratio-major.  float64, like the reference, and cast to fp32 by the caller."""
import numpy as np


def _window_to_whc(win):
    w = win[2] - win[0] + 1.0
    h = win[3] - win[1] + 1.0
    return w, h, win[0] + 0.5 * (w - 1.0), win[1] + 0.5 * (h - 1.0)


def _windows(ws, hs, cx, cy):
    ws = np.atleast_1d(np.asarray(ws, dtype=np.float64))[:, None]
    hs = np.atleast_1d(np.asarray(hs, dtype=np.float64))[:, None]
    half_w, half_h = 0.5 * (ws - 1.0), 0.5 * (hs - 1.0)
    return np.concatenate([cx - half_w, cy - half_h, cx + half_w, cy + half_h], axis=1)


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    w, h, cx, cy = _window_to_whc(np.array([0.0, 0.0, base_size - 1.0, base_size - 1.0]))
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    out = []
    for rw in _windows(ws, hs, cx, cy):
        w, h, cx2, cy2 = _window_to_whc(rw)
        out.append(_windows(w * scales, h * scales, cx2, cy2))
    return np.concatenate(out, axis=0)
