"""Backward-pass kernels in isolation at the training step's shapes (res101, one 800x1333 episode: 50x84 query map, ten
20x20 support maps): the two layout kernels against the HBM roofline (algorithmic bytes = every input element read once,
every output element written once) and the data- / weight-gradient GEMMs in TFLOP/s.  L2 is flushed between iterations.

  python tools/bwd_bench.py [--iters 20] [--only i]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import autograd_ops as A  # noqa: E402
from dana_b200 import ops  # noqa: E402

LAYERS = [  # name, cin, cout, ksize, stride, n, h, w
    ("layer3 conv1 1x1 1024->256 (query)", 1024, 256, 1, 1, 1, 50, 84),
    ("layer3 conv2 3x3 256->256 (query)", 256, 256, 3, 1, 1, 50, 84),
    ("layer3 conv3 1x1 256->1024 (query)", 256, 1024, 1, 1, 1, 50, 84),
    ("layer3 conv2 3x3 256->256 (supports)", 256, 256, 3, 1, 10, 20, 20),
    ("layer2 conv2 3x3 128->128 (query)", 128, 128, 3, 1, 1, 100, 167),
    ("layer3.0 conv1 1x1 s2 512->256 (query)", 512, 256, 1, 2, 1, 100, 167),
    ("layer4 conv2 3x3 512->512 (128 RoIs)", 512, 512, 3, 1, 128, 4, 4),
    ("RPN conv 3x3 2048->512", 2048, 512, 3, 1, 1, 50, 84),
]


def timed(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    print("%-42s %-14s %9s %10s %10s" % ("layer", "kernel", "ms", "GB/s(alg)", "TFLOP/s"))
    for li, (name, ci, co, ks, st, n, h, w) in enumerate(LAYERS):
        if args.only >= 0 and li != args.only:
            continue
        g = torch.Generator().manual_seed(li)
        oh, ow = (h, w) if ks == 3 else ((h - 1) // st + 1, (w - 1) // st + 1)
        px = n * oh * ow
        x = ops.split_f32(torch.randn(n, h, w, ci, generator=g).to(dev))
        wt = (torch.randn(co, ci, ks, ks, generator=g) / (ci * ks * ks) ** 0.5).to(dev)
        gy = torch.randn(n, oh, ow, co, generator=g).to(dev)
        y = torch.randn(n, oh, ow, co, generator=g).to(dev)
        fwd, dg = ops.pack_conv_weight(wt, None)
        taps = ks * ks
        flops = 2.0 * px * co * ci * taps
        res = {}
        # layout kernels
        t = timed(lambda: ops.grad_prepare(gy, y, want_f32=False, want_pair=True, want_t=True), args.iters, flush)
        res["grad_prepare"] = (t, (8.0 + 4.0 + 4.0) * px * co, None)          # read g, y; write NHWC pair + transposed pair
        t = timed(lambda: ops.im2col_t(x, ks, st), args.iters, flush)
        res["im2col_t"] = (t, 4.0 * px * ci * (1 + taps), None)               # read the pair once, write taps planes
        _, gp, gt = ops.grad_prepare(gy, y, want_f32=False, want_pair=True, want_t=True)
        xt = ops.im2col_t(x, ks, st)
        dx = torch.empty((n, oh, ow, ci), dtype=torch.float32, device=dev)
        t = timed(lambda: ops.conv_nhwc(gp, dg, ci, ksize=ks, stride=1, out_f32=dx), args.iters, flush)
        res["dgrad GEMM"] = (t, None, flops)
        dwk = torch.empty((co, taps * ci), dtype=torch.float32, device=dev)
        t = timed(lambda: ops.linear(gt, xt, taps * ci, out_f32=dwk), args.iters, flush)
        res["wgrad GEMM"] = (t, None, flops)
        t = timed(lambda: ops.unpack_conv_wgrad(dwk, None, co, ci, ks, ks), args.iters, flush)
        res["unpack"] = (t, 8.0 * co * ci * taps, None)
        t = timed(lambda: ops.pack_conv_weight(wt, None), args.iters, flush)
        res["pack (fwd+dgrad)"] = (t, (4.0 + 8.0) * co * ci * taps, None)
        y32 = torch.empty((n, oh, ow, co), dtype=torch.float32, device=dev)
        yp = ops.Pair.empty((n, oh, ow, co), dev)
        t = timed(lambda: ops.conv_nhwc(x, fwd, co, ksize=ks, stride=st, relu=True, out=yp, out_f32=y32), args.iters, flush)
        res["forward GEMM"] = (t, None, flops)
        t = timed(lambda: ops.conv_backward(gy, y, x, dg, None, ks, st, need_dx=True, need_dw=True, want_dres=False),
                  args.iters, flush)
        res["conv_backward (all)"] = (t, None, 2 * flops)
        for k, (ms, byts, fl) in res.items():
            print("%-42s %-20s %8.4f %10s %10s" % (name, k, ms, "%.0f" % (byts / ms / 1e6) if byts else "-",
                                                  "%.1f" % (fl / ms / 1e9) if fl else "-"))
    assert ops.device_error() == 0


if __name__ == "__main__":
    main()
