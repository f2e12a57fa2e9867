"""Generates tests/golden/forward_train_small.npz by running the UNMODIFIED reference DAnARCNN in TRAIN mode on CPU
(python from /root/reference/lib, its CPU operators compiled in place) on seeded inputs, numpy RNG seeded right before
the forward.  Run in the build container only:   python oracle/make_golden_train.py
Pins oracle/train_oracle.py (tests/test_oracle_pins.py::test_train_forward_vs_reference)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

import dana_oracle as O  # noqa: E402
import ref_loader  # noqa: E402

TRAIN_CASE = dict(seed=1996, attn_std=0.05, input_seed=5, batch=2, height=256, width=384, n_shot=2, np_seed=7)


def train_inputs():
    tc = TRAIN_CASE
    im, info, sup = O.synth_inputs(tc["input_seed"], tc["batch"], tc["height"], tc["width"], 2 * tc["n_shot"])
    gt = torch.zeros(tc["batch"], 50, 5)
    gt[0, 0] = torch.tensor([20.0, 30.0, 120.0, 140.0, 1.0])
    gt[0, 1] = torch.tensor([150.0, 10.0, 370.0, 230.0, 1.0])
    gt[1, 0] = torch.tensor([50.0, 40.0, 180.0, 150.0, 1.0])
    return im, info, gt, torch.tensor([2, 1]), sup


def main():
    model = ref_loader.load()  # noqa: F841
    from model.utils.config import cfg_from_file, cfg_from_list
    cfg_from_file("/root/reference/cfgs/res50.yml")
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    from model.framework.dana import DAnARCNN
    tc = TRAIN_CASE
    net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_way=2,
                   num_shot=tc["n_shot"])
    net.create_architecture()
    missing, unexpected = net.load_state_dict(O.make_params(tc["seed"], attn_std=tc["attn_std"]), strict=False)
    assert not unexpected
    net.train()
    im, info, gt, nb, sup = train_inputs()
    cap = {}
    net.RCNN_rpn.RPN_anchor_target.register_forward_hook(lambda m, i, o: cap.__setitem__("rpn_targets", [t.detach().clone() for t in o]))
    np.random.seed(tc["np_seed"])
    with torch.no_grad():
        rois, cls_prob, bbox_pred, l1, l2, l3, l4, rois_label = net(im, info, gt, nb, sup)
    lab = cap["rpn_targets"][0]
    np.savez_compressed(os.path.join(GOLD, "forward_train_small.npz"), rois=rois.numpy(), cls_prob=cls_prob.numpy(),
                        bbox_pred=bbox_pred.numpy(), rois_label=rois_label.numpy(),
                        losses=np.array([float(l1), float(l2), float(l3), float(l4)], dtype=np.float64),
                        rpn_labels=lab.numpy().astype(np.int8), rpn_bbox_targets_abs_sum=float(cap["rpn_targets"][1].abs().sum()),
                        rpn_outside_sum=float(cap["rpn_targets"][3].sum()))
    print("losses", float(l1), float(l2), float(l3), float(l4), "fg rois", int(rois_label.sum()),
          "rpn fg/bg", int((lab == 1).sum()), int((lab == 0).sum()))


if __name__ == "__main__":
    main()
