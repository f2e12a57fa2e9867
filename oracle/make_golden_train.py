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
    # the training step's gradients (train.py:129-138: sum of the four losses, backward): per-parameter L2 norm and 16
    # evenly spaced elements of every trainable tensor -- pins the oracle's autograd and, through it, the CUDA backward
    # The reference has no CPU RoIAlign backward (csrc/ROIAlign.h:44 raises "Not implemented on the CPU"): that ONE
    # operator is substituted by the exact adjoint of its (linear) forward, taken from torchvision's roi_align with the
    # same sampling rules (aligned=False); everything else in the backward is the reference's own autograd graph.
    import torchvision

    def roi_align_backward_cpu(grad, rois, scale, ph, pw, bs, ch, h, w, ratio):
        with torch.enable_grad():          # the caller is @once_differentiable
            x = torch.zeros(bs, ch, h, w, dtype=grad.dtype, requires_grad=True)
            o = torchvision.ops.roi_align(x, rois, (ph, pw), scale, ratio, False)
            return torch.autograd.grad(o, x, grad)[0].detach()
    model._C.roi_align_backward = roi_align_backward_cpu
    np.random.seed(tc["np_seed"])
    net.zero_grad()
    out = net(im, info, gt, nb, sup)
    (out[3].mean() + out[4].mean() + out[5].mean() + out[6].mean()).backward()
    g_names, g_norms, g_samples = [], [], []
    for name, prm in net.named_parameters():
        if prm.requires_grad:
            assert prm.grad is not None, name
            g = prm.grad.detach().double().reshape(-1)
            g_names.append(name)
            g_norms.append(float(g.norm()))
            g_samples.append(g[torch.linspace(0, g.numel() - 1, 16).long()].numpy())
    lab = cap["rpn_targets"][0]
    np.savez_compressed(os.path.join(GOLD, "forward_train_small.npz"), rois=rois.numpy(), cls_prob=cls_prob.numpy(),
                        bbox_pred=bbox_pred.numpy(), rois_label=rois_label.numpy(),
                        losses=np.array([float(l1), float(l2), float(l3), float(l4)], dtype=np.float64),
                        rpn_labels=lab.numpy().astype(np.int8), rpn_bbox_targets_abs_sum=float(cap["rpn_targets"][1].abs().sum()),
                        rpn_outside_sum=float(cap["rpn_targets"][3].sum()), grad_names=np.array(g_names),
                        grad_norms=np.array(g_norms), grad_samples=np.stack(g_samples))
    print("losses", float(l1), float(l2), float(l3), float(l4), "fg rois", int(rois_label.sum()),
          "rpn fg/bg", int((lab == 1).sum()), int((lab == 0).sum()))


if __name__ == "__main__":
    main()
