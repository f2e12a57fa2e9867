"""Isolated timing of the tensor-core implicit-GEMM kernel on the layer shapes of the step
(CUDA events, L2 flushed between iterations).  Diagnostic; also the ncu target for the kernel:
  ncu --set full --clock-control none --import-source on -k regex:conv_gemm -o gpurun_out/gemm python tools/gemm_bench.py --only 3 --iters 1
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import ops  # noqa: E402
from dana_b200.ops import Pair  # noqa: E402

# (name, n, h, w, cin, cout, ksize, stride, residual)
SHAPES = [
    ("l1.conv1 1x1 256->64", 4, 150, 250, 256, 64, 1, 1, False),
    ("l1.conv2 3x3 64->64", 4, 150, 250, 64, 64, 3, 1, False),
    ("l1.conv3 1x1 64->256 +res", 4, 150, 250, 64, 256, 1, 1, True),
    ("l2.conv3 1x1 128->512 +res", 4, 75, 125, 128, 512, 1, 1, True),
    ("l3.conv1 1x1 1024->256", 4, 38, 63, 1024, 256, 1, 1, False),
    ("l3.conv2 3x3 256->256", 4, 38, 63, 256, 256, 3, 1, False),
    ("l3.conv3 1x1 256->1024 +res", 4, 38, 63, 256, 1024, 1, 1, True),
    ("rpn 3x3 2048->512", 4, 38, 63, 2048, 512, 3, 1, False),
    ("l4.conv2 3x3 512->512", 1200, 4, 4, 512, 512, 3, 1, False),
    ("l4.conv3 1x1 512->2048 +res", 1200, 4, 4, 512, 2048, 1, 1, True),
    ("l4.conv1 1x1 2048->512", 1200, 4, 4, 2048, 512, 1, 1, False),
    ("l2.conv2 3x3 128->128", 4, 75, 125, 128, 128, 3, 1, False),
    ("l2.conv1 1x1 512->128", 4, 75, 125, 512, 128, 1, 1, False),
    ("exp 1x1 64->256 no res", 4, 150, 250, 64, 256, 1, 1, False),
    ("exp 1x1 256->256 +res", 4, 150, 250, 256, 256, 1, 1, True),
    ("exp 1x1 64->64 +res", 4, 150, 250, 64, 64, 1, 1, True),
]

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", type=int, default=-1)
ap.add_argument("--precision", default="both")
a = ap.parse_args()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("%-30s %5s %9s %8s %9s %9s" % ("layer", "prec", "ms", "TFLOP/s", "GB/s(alg)", "MB(alg)"))
for idx, (name, n, h, w, cin, cout, ks, stride, has_res) in enumerate(SHAPES):
    if a.only >= 0 and idx != a.only:
        continue
    modes = {"both": ("x3", "x1"), "all": ("x3", "x1", "f16"), "bf16x3": ("x3",), "bf16": ("x1",), "f16": ("f16",)}[a.precision]
    for mode in modes:
        split = mode == "x3"
        mk = Pair.from_float_f16 if mode == "f16" else (lambda t: Pair.from_float(t, split))
        x = mk(torch.randn(n, h, w, cin, device="cuda"))
        wt = mk(torch.randn(cout, ks * ks * cin, device="cuda") * 0.02)
        sc = torch.rand(cout, device="cuda") + 0.5
        bi = torch.randn(cout, device="cuda")
        res = mk(torch.randn(n, h, w, cout, device="cuda")) if has_res else None
        out = Pair.empty_f16((n, h, w, cout), "cuda") if mode == "f16" else Pair.empty((n, h, w, cout), "cuda", split)
        run = lambda: ops.conv_nhwc(x, wt, cout, ksize=ks, stride=stride, scale=sc, bias=bi, res=res, relu=True, out=out, split=split)  # noqa: E731
        for _ in range(3):
            run()
        tot = 0.0
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / a.iters
        m = n * h * w
        planes = 2 if split else 1
        byt = planes * 2 * (m * cin + m * cout * (2 if has_res else 1) + cout * ks * ks * cin)
        print("%-30s %5s %9.4f %8.1f %9.0f %9.1f" % (name, mode, ms, 2.0 * m * cout * ks * ks * cin / ms / 1e9,
                                                    byt / ms / 1e6, byt / 1e6), flush=True)
