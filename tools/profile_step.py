"""One profiled forward step for ncu (never a benchmark number):
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [--precision bf16x3] [--batch 4]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200.engine import DanaEngine  # noqa: E402
from dana_b200.synthetic import synthetic_episode, synthetic_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--height", type=int, default=600)
ap.add_argument("--width", type=int, default=1000)
ap.add_argument("--shots", type=int, default=3)
ap.add_argument("--sets", type=int, default=2)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
p = synthetic_state_dict(1996)
im, info, sup = synthetic_episode(3, a.batch, a.height, a.width, a.shots * a.sets)
im, info, sup = im.cuda(), info.cuda(), sup.cuda()
eng = DanaEngine(p, n_shot=a.shots, precision=a.precision)
for _ in range(a.warmup):
    eng.forward(im, info, sup)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(im, info, sup)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
