"""cProfile of the host side of one training step (diagnostic)."""
import cProfile
import os
import pstats
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import dana_b200  # noqa: E402,F401
from dana_b200.config import cfg_from_file, cfg_from_list, reset_cfg  # noqa: E402
from dana_b200.dana import DAnARCNN  # noqa: E402
from dana_b200.synthetic import synthetic_episode, synthetic_state_dict  # noqa: E402
from dana_b200.train_step import SGDTrainer  # noqa: E402
layers = int(os.environ.get("LAYERS", "101"))
reset_cfg()
cfg_from_file(os.path.join(ROOT, "cfgs", "res50.yml"))
cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_layers=layers, num_way=2,
               num_shot=5, precision="bf16x3")
net.create_architecture()
net.load_state_dict(synthetic_state_dict(1996, num_layers=layers), strict=False)
net.cuda().train()
tr = SGDTrainer(net)
im, info, sup = synthetic_episode(100, 1, 800, 1333, 10)
gt = torch.zeros(1, 50, 5)
gt[0, 0] = torch.tensor([100.0, 120.0, 400.0, 380.0, 1.0])
args = [t.cuda() for t in (im, info, gt, torch.tensor([1]), sup)]
np.random.seed(0)
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    tr.step(*args)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
