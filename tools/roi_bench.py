"""RoIAlign microbenchmark at the BASELINE.json shape (map [B,1024,38,63], 300 RoIs per image): the
reference-layout operator (NCHW fp32 -> [R,C,7,7] fp32, dana_roi_align_forward layout 0) and the
pipeline variant (NHWC -> bf16 pair).  Algorithmic bytes: 4*R*C*49 + 4*C*h*w*B + 20*R.
  ncu --set full --clock-control none --import-source on -k regex:roi_align7 -c 3 -o gpurun_out/roi python tools/roi_bench.py --iters 1 --only f32
L2 is flushed before every timed launch (256 MiB written, then 256 MiB read so that no dirty lines are left behind).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--sweep", action="store_true", help="also time the fp32 NHWC kernel per RoI size class")
ap.add_argument("--side", type=int, default=0, help="square RoIs of this side (pixels) instead of the mixture")
ap.add_argument("--no-flush", action="store_true", help="leave L2 warm between iterations (diagnostic)")
ap.add_argument("--only", default="", help="run only the named variant (nchw|pair|head|f32) -- for ncu captures")
a = ap.parse_args()
b, c, h, w = a.batch, 1024, 38, 63
rs = np.random.RandomState(0)
feat = torch.randn(b, c, h, w, device="cuda")
r = 300 * b
# proposal-like boxes: mixture of scales, clipped to the image
cx, cy = rs.uniform(0, 1000, r), rs.uniform(0, 600, r)
bw, bh = np.exp(rs.uniform(np.log(16), np.log(600), r)), np.exp(rs.uniform(np.log(16), np.log(500), r))
rois = np.stack([np.repeat(np.arange(b), 300), np.clip(cx - bw / 2, 0, 999), np.clip(cy - bh / 2, 0, 599),
                 np.clip(cx + bw / 2, 0, 999), np.clip(cy + bh / 2, 0, 599)], 1).astype(np.float32)
if a.side > 0:
    hh = min(a.side, 599)
    cx, cy = rs.uniform(a.side / 2, 1000 - a.side / 2, r), rs.uniform(hh / 2, 600 - hh / 2, r)
    rois = np.stack([np.repeat(np.arange(b), 300), cx - a.side / 2, cy - hh / 2, cx + a.side / 2, cy + hh / 2], 1).astype(np.float32)
rois = torch.from_numpy(rois).cuda()
nhwc = feat.permute(0, 2, 3, 1).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")


def l2_flush():
    """Write a 256 MiB buffer (evicts everything), then READ another 256 MiB one: the write alone leaves ~126 MB of
    dirty lines in L2 whose write-back would land inside the timed region of a write-bound kernel and be billed to it."""
    if a.no_flush:
        return
    flush.zero_()
    flush_rd.sum()


alg_bytes = 4.0 * r * c * 49 + 4.0 * c * h * w * b + 20.0 * r


def timeit(fn):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(a.iters):
        l2_flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / a.iters


peak = 6405.0
try:
    import json
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
if a.only:
    pe = torch.randn(49, c, device="cuda")
    fn = {"nchw": lambda: ops.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0),
          "pair": lambda: ops.roi_align_nhwc(nhwc, rois, 1.0 / 16, 7, 0, want_f32=False, want_pair=True),
          "head": lambda: ops.roi_align_head(nhwc, rois, 1.0 / 16, 0, pe=pe, want_f32=False, want_pair=True, want_qpe=True),
          "head16": lambda: ops.roi_align_head(nhwc, rois, 1.0 / 16, 0, want_f32=False, want_pair=True, want_f16=True),
          "f32": lambda: ops.roi_align_head(nhwc, rois, 1.0 / 16, 0, want_f32=True, want_pair=False, want_qpe=False)}[a.only]
    print("%s: %.4f ms" % (a.only, timeit(fn)))
    sys.exit(0)
ms = timeit(lambda: ops.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0))
print("roi_align NCHW fp32 (reference layout, incl. NCHW->NHWC staging): %.4f ms  %.0f GB/s algorithmic (%.1f MB)" %
      (ms, alg_bytes / ms / 1e6, alg_bytes / 1e6))
ms = timeit(lambda: ops.roi_align_nhwc(nhwc, rois, 1.0 / 16, 7, 0, want_f32=False, want_pair=True))
pb = 2 * 2.0 * r * c * 49 + 4.0 * c * h * w * b
print("roi_align NHWC -> bf16 hi/lo pair (pipeline):                      %.4f ms  %.0f GB/s algorithmic (%.1f MB)" %
      (ms, pb / ms / 1e6, pb / 1e6))
pe = torch.randn(49, c, device="cuda")
ms = timeit(lambda: ops.roi_align_head(nhwc, rois, 1.0 / 16, 0, pe=pe, want_f32=False, want_pair=True, want_qpe=True))
pb2 = 4 * 2.0 * r * c * 49 + 4.0 * c * h * w * b
print("roi_align_head NHWC -> pooled pair + (pooled+PE) pair:              %.4f ms  %.0f GB/s algorithmic (%.1f MB)" %
      (ms, pb2 / ms / 1e6, pb2 / 1e6))
ms = timeit(lambda: ops.roi_align_head(nhwc, rois, 1.0 / 16, 0, want_f32=False, want_pair=True, want_f16=True))
pb3 = (4 + 2) * 1.0 * r * c * 49 + 4.0 * c * h * w * b
print("roi_align_head NHWC -> pooled pair + fp16 plane (the mixed-mode step): %.4f ms  %.0f GB/s algorithmic (%.1f MB)  %.1f %% of HBM peak" %
      (ms, pb3 / ms / 1e6, pb3 / 1e6, 100 * pb3 / ms / 1e6 / peak))
g49 = torch.randn(r, 49, c, device="cuda")
ms = timeit(lambda: ops.roi_align_backward_nhwc(g49, rois, 1.0 / 16, b, h, w, 0))
print("roi_align backward NHWC (vector reductions, incl. the map memset):  %.4f ms  (reads %.1f MB of gradient)" % (ms, 4.0 * r * c * 49 / 1e6))
ms = timeit(lambda: ops.roi_align_head(nhwc, rois, 1.0 / 16, 0, want_f32=True, want_pair=False, want_qpe=False))
print("roi_align_head NHWC -> fp32 [R,7,7,C]:                              %.4f ms  %.0f GB/s algorithmic (%.1f MB)" %
      (ms, alg_bytes / ms / 1e6, alg_bytes / 1e6))
print("fraction of the measured HBM peak (%.0f GB/s): %.1f %%" % (peak, 100 * alg_bytes / ms / 1e6 / peak))
# reference points under the same protocol: what a pure write / a copy of the output-sized buffer achieves
_o = torch.empty((r, 7, 7, c), device="cuda")
_o2 = torch.empty((r, 7, 7, c), device="cuda")
ms = timeit(lambda: _o.zero_())
print("reference: memset of the %.0f MB output (pure write): %.4f ms  %.0f GB/s" % (_o.numel() * 4 / 1e6, ms, _o.numel() * 4 / ms / 1e6))
ms = timeit(lambda: _o.copy_(_o2))
print("reference: copy of the output (read + write, %.0f MB moved): %.4f ms  %.0f GB/s" % (_o.numel() * 8 / 1e6, ms, _o.numel() * 8 / ms / 1e6))
if a.sweep:
    # the gather cost of an RoI grows with its area (adaptive sampling visits every feature pixel it covers), the
    # output does not: per size class, where the kernel is write-bound and where it is gather-bound
    print("size sweep (square RoIs, 1200 per class, fp32 [R,7,7,C] out):  side px | ms | GB/s algorithmic | % HBM peak")
    for side in (32, 64, 112, 160, 224, 320, 448, 600):
        cx, cy = rs.uniform(side / 2, 1000 - side / 2, r), rs.uniform(min(side, 599) / 2, 600 - min(side, 599) / 2, r)
        hh = min(side, 599)
        rr = np.stack([np.repeat(np.arange(b), 300), cx - side / 2, cy - hh / 2, cx + side / 2, cy + hh / 2], 1).astype(np.float32)
        rr = torch.from_numpy(rr).cuda()
        ms = timeit(lambda: ops.roi_align_head(nhwc, rr, 1.0 / 16, 0, want_f32=True, want_pair=False, want_qpe=False))
        print("  %4d  %.4f ms  %6.0f GB/s  %5.1f %%" % (side, ms, alg_bytes / ms / 1e6, 100 * alg_bytes / ms / 1e6 / peak))
