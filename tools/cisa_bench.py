"""BA + CISA block microbenchmark (BASELINE.json configs[4]): query 38x50x1024 vs support Ns x 1024,
units = way x shot support maps attended against one query feature.
Reports time, algorithmic TFLOP/s and GB/s against the measured peaks, per precision mode.
Algorithmic FLOPs / bytes: SURVEY.md section 8d formulas (C = 1024, d = 256)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200.engine import DanaEngine  # noqa: E402
from dana_b200.synthetic import synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--units", default="1,2,3,6,10,25,50,100")
ap.add_argument("--ns", default="196,400")
ap.add_argument("--precision", default="bf16x3,bf16")
ap.add_argument("--only-unit", type=int, default=0)
ap.add_argument("--eager", action="store_true", help="time eager launches (host launch overhead included) instead of a CUDA graph replay")
a = ap.parse_args()
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
sd = synthetic_state_dict(1996)
C, D, NQ = 1024, 256, 1900
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("%-7s %4s %5s %9s %9s %8s %9s %8s" % ("prec", "Ns", "units", "ms", "TFLOP/s", "%peakTF", "GB/s", "%peakBW"))
for prec in a.precision.split(","):
    for ns_side in [int(v) for v in a.ns.split(",")]:
        hs = int(round(ns_side ** 0.5))
        ns = hs * hs
        for units in [int(v) for v in a.units.split(",")]:
            if a.only_unit and units != a.only_unit:
                continue
            eng = DanaEngine(sd, n_shot=units, precision=prec)
            base = torch.relu(torch.randn(1, C, 38, 50, device="cuda"))
            sup = torch.relu(torch.randn(1, units, C, hs, hs, device="cuda"))
            from dana_b200 import ops
            from dana_b200.ops import Pair
            split = prec == "bf16x3"
            bp = ops.split_f32(base.permute(0, 2, 3, 1).contiguous(), split).view(NQ, C)
            dense = Pair.empty((NQ, C), "cuda", split)
            supp = ops.split_f32(sup.reshape(units, C, hs, hs).permute(0, 2, 3, 1).contiguous(), split)
            run = lambda: eng.rpn_attention(bp, dense, 1, NQ, supp, 1)  # noqa: E731
            for _ in range(3):
                run()
            if not a.eager:   # one graph replay per iteration: device time of the block, not python launch overhead
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    run()
                    with torch.cuda.graph(graph, stream=side):
                        run()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                run = graph.replay
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
            flops = 2.0 * NQ * C * D + units * (2.0 * ns * C * D + 2.0 * NQ * ns * D + 2.0 * NQ * ns * C + 8.0 * ns * C)
            byts = 2.0 * C * (NQ + units * ns) + 4.0 * NQ * C + 1.05e6
            tf = flops / ms / 1e9
            gbs = byts / ms / 1e6
            print("%-7s %4d %5d %9.4f %9.1f %8.1f %9.0f %8.1f" % (prec, ns, units, ms, tf, 100 * tf / peaks["bf16_tflops"],
                                                                   gbs, 100 * gbs / peaks["hbm_gbs"]), flush=True)
