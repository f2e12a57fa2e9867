"""Per-launch table of the tensor-core GEMM kernel inside one forward step (CUDA events on the
launching stream): shape, ms, achieved TFLOP/s.  Diagnostic, not a benchmark."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import ops  # noqa: E402
from dana_b200.engine import DanaEngine  # noqa: E402
from dana_b200.synthetic import synthetic_episode, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--batch", type=int, default=4)
a = ap.parse_args()
eng = DanaEngine(synthetic_state_dict(1996), n_shot=3, precision=a.precision)
im, info, sup = [t.cuda() for t in synthetic_episode(3, a.batch)]
for _ in range(3):
    eng.forward(im, info, sup)
torch.cuda.synchronize()
ops.GEMM_TRACE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.forward(im, info, sup)
e1.record()
torch.cuda.synchronize()
tr, ops.GEMM_TRACE = ops.GEMM_TRACE, None
tot = 0.0
print("step %.3f ms (with events)" % e0.elapsed_time(e1))
print("%4s %9s %6s %7s %4s %9s %9s" % ("#", "M", "N", "K", "taps", "ms", "TFLOP/s"))
for i, (s, e, fl, (m, n, k, taps)) in enumerate(tr):
    ms = s.elapsed_time(e)
    tot += ms
    print("%4d %9d %6d %7d %4d %9.4f %9.1f" % (i, m, n, k, taps, ms, fl / ms / 1e9))
print("gemm total %.3f ms, %.1f GF, %.1f TFLOP/s" % (tot, sum(t[2] for t in tr) / 1e9, sum(t[2] for t in tr) / tot / 1e9))
