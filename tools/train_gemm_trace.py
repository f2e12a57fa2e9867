"""Per-shape GEMM times of one EAGER training step (ops.GEMM_TRACE: CUDA events around every dana_conv_gemm launch issued
from Python; the launches inside dana_conv_backward are not individually visible here).  Diagnostic."""
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
os.environ["DANA_TRAIN_SIDE_STREAM"] = "0"
reset_cfg()
cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_layers=101, num_way=2,
               num_shot=5, precision="bf16x3")
net.create_architecture()
net.load_state_dict(synthetic_state_dict(1996, num_layers=101), strict=False)
net.cuda().train()
tr = SGDTrainer(net)
im, info, sup = synthetic_episode(100, 1, 800, 1333, 10)
gt = torch.zeros(1, 50, 5)
gt[0, 0] = torch.tensor([100.0, 120.0, 400.0, 380.0, 1.0])
args = [t.cuda() for t in (im, info, gt, torch.tensor([1]), sup)]
np.random.seed(0)
for _ in range(2):
    tr.step(*args)
torch.cuda.synchronize()
ops.GEMM_TRACE = []
tr.step(*args)
torch.cuda.synchronize()
rows = {}
for e0, e1, flops, shape, meta in ops.GEMM_TRACE:
    d = rows.setdefault(shape, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += e0.elapsed_time(e1)
    d[2] += flops
ops.GEMM_TRACE = None
tot = sum(v[1] for v in rows.values())
print("forward-side GEMM launches (fwd convs, linears, bmm fwd+bwd): %.2f ms" % tot)
for shape, v in sorted(rows.items(), key=lambda kv: -kv[1][1])[:25]:
    print("M=%6d N=%5d K=%6d taps=%d  n=%3d  %7.3f ms  %6.1f TFLOP/s" % (shape[0], shape[1], shape[2], shape[3], v[0], v[1], v[2] / v[1] / 1e9))
