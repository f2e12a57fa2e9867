"""Robustness sweep: the eval forward on the other BASELINE.json configurations' shapes (res101 5-way 5-shot at
800x1333, 1-shot, 10-shot, odd image sizes), checking shapes / finiteness / determinism and printing the step time.
Diagnostic only (parity is asserted by the test-suite at sizes the oracle can reach)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import ops  # noqa: E402
from dana_b200.engine import DanaEngine  # noqa: E402
from dana_b200.synthetic import synthetic_episode, synthetic_state_dict  # noqa: E402

CASES = [  # (layers, batch, H, W, sets, shots, support_size)
    (101, 2, 800, 1333, 5, 5, 320),
    (50, 1, 600, 800, 1, 1, 320),
    (50, 3, 480, 640, 1, 10, 320),
    (50, 2, 375, 500, 2, 3, 224),
    (50, 1, 1000, 600, 3, 2, 320),
    (101, 1, 608, 1008, 2, 5, 256),
    (50, 1, 333, 517, 2, 3, 200),      # nothing divides anything
    (50, 2, 600, 1000, 1, 3, 288),     # 18x18 supports: Ns = 324 (wide softmax tile, odd half width)
]
for layers, b, h, w, sets, k, ss in CASES:
    sd = synthetic_state_dict(1996, num_layers=layers)
    for prec in ("mixed", "bf16x3", "bf16"):
        eng = DanaEngine(sd, num_layers=layers, n_shot=k, precision=prec)
        im, info, sup = synthetic_episode(5, b, h, w, sets * k, support_size=ss)
        im, info, sup = im.cuda(), info.cuda(), sup.cuda()
        out = eng.forward(im, info, sup)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out2 = eng.forward(im, info, sup)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        rois, cls_prob, bbox = out
        ok = (tuple(rois.shape) == (b, 300, 5) and tuple(cls_prob.shape) == (sets * b * 300, 2) and
              tuple(bbox.shape) == (b * 300, 4) and all(torch.isfinite(t).all().item() for t in out) and
              all(torch.equal(x, y) for x, y in zip(out, out2)) and ops.device_error() == 0)
        print("res%d bs%d %dx%d %d-set %d-shot support %d %-7s %8.2f ms  %s" % (layers, b, h, w, sets, k, ss, prec, ms,
                                                                              "ok" if ok else "FAILED"), flush=True)
        del eng
        torch.cuda.empty_cache()
