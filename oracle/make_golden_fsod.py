"""Generates tests/golden/fsod_attention.npz by running the UNMODIFIED reference FSOD module (eval mode, CPU) and
capturing what it feeds to its RPN (correlation_feat, lib/model/framework/fsod.py:112-116).  Build container only:
  python oracle/make_golden_fsod.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

import dana_oracle as O  # noqa: E402
import ref_loader  # noqa: E402

FSOD_CASE = dict(seed=1996, input_seed=13, height=160, width=224, n_shot=2)


def main():
    model = ref_loader.load()  # noqa: F841
    from model.utils.config import cfg_from_file, cfg_from_list
    cfg_from_file("/root/reference/cfgs/res50.yml")
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    from model.framework.fsod import FSOD
    fc = FSOD_CASE
    torch.manual_seed(0)
    net = FSOD(["bg", "fg"], 50, pretrained=False, num_way=2, num_shot=fc["n_shot"])
    net.create_architecture()
    params = O.make_params(fc["seed"])
    trunk = {k: v for k, v in params.items() if k.startswith(("RCNN_base.", "RCNN_top."))}
    missing, unexpected = net.load_state_dict(trunk, strict=False)
    assert not unexpected
    net.eval()
    im, info, sup = O.synth_inputs(fc["input_seed"], 1, fc["height"], fc["width"], fc["n_shot"])
    cap = {}
    net.RCNN_rpn.register_forward_hook(lambda m, i, o: cap.__setitem__("corr", i[0].detach()))
    with torch.no_grad():
        net(im, info, torch.zeros(1, 1, 5), torch.zeros(1), sup)
    corr = cap["corr"]
    np.savez_compressed(os.path.join(GOLD, "fsod_attention.npz"), corr_sample=corr.reshape(-1)[::5].numpy().copy(),
                        corr_shape=np.array(corr.shape), corr_abs_sum=float(corr.abs().sum()))
    print("fsod correlation_feat", tuple(corr.shape), float(corr.abs().max()))


if __name__ == "__main__":
    main()
