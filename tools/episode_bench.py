"""Episode construction microbenchmark: one 600x1000 query from a 375x500 uint8 image
(prep_im_for_blob) and six 320x320 support crops (fs_loader crop -> resize -> pad) per episode, on the device, next to
the same work with cv2 on one host core (what the reference's loaders do, num_workers = 0 in inference.py:85).
Algorithmic bytes: source window read once + canvas written once."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import episode  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
MEANS = [102.9801, 115.9465, 122.7717]
rs = np.random.RandomState(0)
im_h = rs.randint(0, 256, size=(375, 625, 3)).astype(np.uint8)       # -> 600 x 1000
sup_h = (rs.standard_normal((600, 1000, 3)) * 50).astype(np.float32)  # a prepared support image
boxes = [(50 + 40 * i, 30 + 20 * i, 250 + 60 * i, 200 + 50 * i) for i in range(6)]
im_d, sup_d = torch.from_numpy(im_h).cuda(), torch.from_numpy(sup_h).cuda()
q_out = torch.empty((3, 600, 1000), device="cuda")
s_out = torch.empty((6, 3, 320, 320), device="cuda")


def gpu_episode():
    episode.prep_im_for_blob(im_d, MEANS, 600, out=q_out)
    for i, b in enumerate(boxes):
        episode.support_from_box(sup_d, b, 1.0, 320, out=s_out[i])


for _ in range(3):
    gpu_episode()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        gpu_episode()
torch.cuda.current_stream().wait_stream(side)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    g.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
src_bytes = im_h.size + sum((b[2] - b[0] + 1) * (b[3] - b[1] + 1) * 12 for b in boxes)
dst_bytes = 4 * (3 * 600 * 1000 + 6 * 3 * 320 * 320)
print("GPU episode construction (1 query 375x625 u8 -> 600x1000, 6 support crops -> 320x320): %.4f ms/episode, "
      "%.0f GB/s algorithmic (%.1f MB), 7 launches" % (ms, (src_bytes + dst_bytes) / ms / 1e6, (src_bytes + dst_bytes) / 1e6))
try:
    import cv2
    cv2.setNumThreads(1)

    def cpu_episode():
        f = im_h.astype(np.float32) - np.array([[MEANS]], dtype=np.float32)
        q = cv2.resize(f, None, None, fx=1.6, fy=1.6, interpolation=cv2.INTER_LINEAR)
        outs = [np.ascontiguousarray(q.transpose(2, 0, 1))]
        for b in boxes:
            crop = sup_h[b[1]:b[3] + 1, b[0]:b[2] + 1]
            bh, bw = b[3] - b[1], b[2] - b[0]
            dsz = (320, int(bh * 320.0 / bw)) if bw >= bh else (int(bw * 320.0 / bh), 320)
            r = cv2.resize(crop, dsz, interpolation=cv2.INTER_LINEAR)
            canvas = np.zeros((3, 320, 320), dtype=np.float32)
            canvas[:, :r.shape[0], :r.shape[1]] = r.transpose(2, 0, 1)
            outs.append(canvas)
        return outs
    cpu_episode()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        cpu_episode()
    print("cv2 on one host core (the reference's loader path): %.3f ms/episode" % ((time.perf_counter() - t0) / n * 1e3))
except ImportError:
    print("cv2 not importable here: host timing skipped")
