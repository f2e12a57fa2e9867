"""Generates tests/golden/episode_cv2.npz with the REAL cv2 (the third-party library the reference's loaders call:
blob.py:48, fs_loader.py:128,132, inference_loader.py:102,106) on seeded inputs; run once in the build container:
    python oracle/make_golden_episode.py
The vectors pin oracle/episode_oracle.py (tests/test_oracle_pins.py)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rs = np.random.RandomState(1996)
out = {}
cases = [("down", 41, 57, 23, 31), ("up", 17, 23, 45, 66), ("same", 12, 16, 16, 12), ("thin", 3, 24, 50, 5),
         ("one", 1, 1, 7, 5)]
for name, h, w, dw, dh in cases:
    src = (rs.rand(h, w, 3) * 255 - 110).astype(np.float32)
    out[name + "_src"] = src
    out[name + "_dst"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    # the generic C++ path (no SIMD / IPP dispatch): the algorithm the oracle restates, bit for bit; the default
    # (optimised) build of cv2 4.13 differs from it by up to 4.3e-4 on the 8-bit pixel scale
    cv2.setUseOptimized(False)
    out[name + "_dst_generic"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    cv2.setUseOptimized(True)
# fx/fy form used by prep_im_for_blob
im = rs.randint(0, 256, size=(30, 47, 3)).astype(np.uint8)
means = np.array([[[102.9801, 115.9465, 122.7717]]], dtype=np.float32)
f = im.astype(np.float32) - means
scale = 48.0 / 30.0
out["prep_im"] = im
out["prep_dst"] = cv2.resize(f, None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
cv2.setUseOptimized(False)
out["prep_dst_generic"] = cv2.resize(f, None, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
cv2.setUseOptimized(True)
out["prep_scale"] = np.float64(scale)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "episode_cv2.npz"), **out)
print("wrote episode_cv2.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
