"""Kernel-level profile of one training step (torch.profiler, CUPTI): total GPU time, launch count and the top
kernels.  Diagnostic for tools/train_bench.py; numbers under a profiler are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]] + sys.argv[1:]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200.config import cfg_from_file, cfg_from_list, reset_cfg  # noqa: E402
from dana_b200.dana import DAnARCNN  # noqa: E402
from dana_b200.synthetic import synthetic_episode, synthetic_state_dict  # noqa: E402
from dana_b200.train_step import SGDTrainer  # noqa: E402

layers = int(os.environ.get("LAYERS", "101"))
h, w, shots = int(os.environ.get("H", "800")), int(os.environ.get("W", "1333")), int(os.environ.get("SHOTS", "5"))
reset_cfg()
cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_layers=layers, num_way=2,
               num_shot=shots, precision="bf16x3")
net.create_architecture()
net.load_state_dict(synthetic_state_dict(1996, num_layers=layers), strict=False)
net.cuda().train()
tr = SGDTrainer(net, cuda_graph=os.environ.get('EAGER', '0') != '1')
im, info, sup = synthetic_episode(100, 1, h, w, 2 * shots)
gt = torch.zeros(1, 50, 5)
gt[0, 0] = torch.tensor([100.0, 120.0, 400.0, 380.0, 1.0])
gt[0, 1] = torch.tensor([500.0, 200.0, 760.0, 520.0, 1.0])
args = [t.cuda() for t in (im, info, gt, torch.tensor([2]), sup)]
np.random.seed(0)
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.step(*args)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type is not None]
kern = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time_total if hasattr(e, "device_time_total") else 0 for e in kern)
by = {}
for e in kern:
    d = by.setdefault(e.name[:90], [0, 0.0])
    d[0] += 1
    d[1] += e.device_time_total
print("GPU kernels: %d launches, %.2f ms total device time" % (len(kern), tot / 1e3))
own = sum(v[1] for k, v in by.items() if "dana::" in k)
print("own kernels (dana::): %.2f ms = %.1f %% of device time" % (own / 1e3, 100.0 * own / max(tot, 1)))
for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:32]:
    print("%8.3f ms  n=%5d  %s" % (v[1] / 1e3, v[0], k))
