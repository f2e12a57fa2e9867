"""Robustness sweep of the training step (captured and eager) over shapes: odd image sizes, 1 / 3 / 10 shots, res50 / res101,
batch 1 / 2 / 4.  Checks: capture did not fall back, losses finite and decreasing-ish over 4 steps, captured == eager on the
first step, no device error.  Diagnostic (parity proper is in tests/test_gpu_train.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import ops  # noqa: E402
from dana_b200.config import cfg_from_file, cfg_from_list, reset_cfg  # noqa: E402
from dana_b200.dana import DAnARCNN  # noqa: E402
from dana_b200.synthetic import synthetic_episode, synthetic_state_dict  # noqa: E402
from dana_b200.train_step import SGDTrainer  # noqa: E402

CASES = [(50, 2, 600, 1000, 3, 320), (50, 1, 375, 500, 1, 320), (50, 3, 333, 517, 2, 224), (101, 1, 608, 1008, 5, 256),
         (50, 1, 480, 640, 10, 320)]
reset_cfg()
cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
for layers, b, h, w, k, ss in CASES:
    sd = synthetic_state_dict(1996, num_layers=layers)
    im, info, sup = synthetic_episode(3, b, h, w, 2 * k, support_size=ss)
    rs = np.random.RandomState(b + h)
    gt = torch.zeros(b, 50, 5)
    for i in range(b):
        n = rs.randint(1, 5)
        bw, bh = rs.uniform(40, w / 2, n), rs.uniform(40, h / 2, n)
        x1, y1 = rs.uniform(0, w - bw - 1), rs.uniform(0, h - bh - 1)
        gt[i, :n] = torch.from_numpy(np.stack([x1, y1, x1 + bw, y1 + bh, np.ones(n)], 1).astype(np.float32))
    args = [t.cuda() for t in (im, info, gt, torch.ones(b), sup)]
    res = {}
    for mode in ("eager", "captured"):
        net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_layers=layers, num_way=2,
                       num_shot=k, precision="bf16x3")
        net.create_architecture()
        net.load_state_dict(sd, strict=False)
        net.cuda().train()
        tr = SGDTrainer(net, cuda_graph=(mode == "captured"))
        np.random.seed(5)
        losses = [float(tr.step(*args)[0]) for _ in range(4)]
        fell_back = mode == "captured" and any(c is False for c in tr._captured.values())
        res[mode] = (losses, fell_back)
        del tr, net
        torch.cuda.empty_cache()
    le, lc = res["eager"][0], res["captured"][0]
    ok = (all(np.isfinite(le + lc)) and not res["captured"][1] and abs(le[0] - lc[0]) <= 1e-4 * abs(le[0]) and
          ops.device_error() == 0)
    print("res%d bs%d %dx%d %d-shot support %d: eager %s captured %s  %s" % (
        layers, b, h, w, k, ss, ["%.4f" % v for v in le], ["%.4f" % v for v in lc], "ok" if ok else "FAILED"), flush=True)
