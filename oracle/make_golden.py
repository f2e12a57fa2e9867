"""Generates tests/golden/*.npz by running the UNMODIFIED reference (python from /root/reference/lib,
CPU operators compiled in place from /root/reference/lib/model/csrc) on seeded synthetic inputs.

Run in the build container only:   python oracle/make_golden.py
Inputs are never stored: every test regenerates them from the numpy seeds recorded here, so the
fixtures hold outputs (or strided samples + checksums of large tensors) only.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

import dana_oracle as O  # noqa: E402
import ref_loader  # noqa: E402


# ---- seeded input builders shared with the tests (tests import them from here) ----------------
def nms_case(name):
    rs = np.random.RandomState({"random300": 11, "clustered1000": 12, "ties": 13, "big6000": 14}[name])
    if name == "ties":
        boxes = np.array([[0, 0, 9, 9], [0, 0, 9, 4], [1, 1, 9, 9], [20, 20, 30, 30], [20, 20, 30, 30.5],
                          [0, 0, 9, 9]], dtype=np.float32)
        scores = np.array([0.9, 0.8, 0.7, 0.6, 0.5, 0.4], dtype=np.float32)
        return boxes, scores, 0.5
    n = {"random300": 300, "clustered1000": 1000, "big6000": 6000}[name]
    x1 = rs.uniform(0, 900, n)
    y1 = rs.uniform(0, 500, n)
    w = rs.uniform(1, 200, n)
    h = rs.uniform(1, 200, n)
    boxes = np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    if name == "clustered1000":
        boxes[n // 2:] = boxes[: n - n // 2] + rs.normal(0, 3, (n - n // 2, 4)).astype(np.float32)
    scores = rs.permutation(n).astype(np.float32) / n  # distinct
    return boxes, scores.astype(np.float32), 0.7 if name != "random300" else 0.3


def roi_align_case():
    rs = np.random.RandomState(21)
    feat = rs.standard_normal((2, 8, 12, 15)).astype(np.float32)
    r = 40
    x1 = rs.uniform(-20, 220, r)
    y1 = rs.uniform(-20, 170, r)
    w = rs.uniform(0.5, 200, r)
    h = rs.uniform(0.5, 150, r)
    rois = np.stack([rs.randint(0, 2, r).astype(np.float64), x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    rois[0, 1:] = [0, 0, 239, 191]          # full map
    rois[1, 1:] = [100, 100, 90, 95]        # x2 < x1
    rois[2, 1:] = [-300, -300, -200, -250]  # fully outside
    rois[3, 1:] = [17.3, 33.9, 17.9, 34.2]  # sub-pixel
    return feat, rois


def proposal_case():
    rs = np.random.RandomState(31)
    b, a, fh, fw = 2, 12, 6, 8
    cls = rs.standard_normal((b, 2 * a, fh, fw)).astype(np.float32)
    prob = torch.softmax(torch.from_numpy(cls).view(b, 2, a * fh, fw), 1).view(b, 2 * a, fh, fw).numpy()
    bbox = (rs.standard_normal((b, 4 * a, fh, fw)) * 0.4).astype(np.float32)
    im_info = np.array([[96.0, 128.0, 1.0], [90.0, 120.0, 1.0]], dtype=np.float32)
    return prob, bbox, im_info


FORWARD_CASE = dict(seed=1996, height=128, width=192, n_shot=2, attn_std=0.05)
# the headline shape itself (BASELINE.json configs[1], one episode): 600x1000 query, 3 shots (the reference's eval
# forward uses the positive set only)
FORWARD_FULL_CASE = dict(seed=2024, height=600, width=1000, n_shot=3, attn_std=0.05)


def postprocess_case():
    """Detections input for the post-processing pin (inference.py:108-142 / utils.py:312-317): 300 boxes with
    clustered overlaps and scores around the 0.05 threshold."""
    rs = np.random.RandomState(31)
    n = 300
    cx, cy = rs.uniform(50, 950, n), rs.uniform(50, 550, n)
    cx[100:] = cx[:200] + rs.normal(0, 6, 200)
    cy[100:] = cy[:200] + rs.normal(0, 6, 200)
    w, h = rs.uniform(20, 300, n), rs.uniform(20, 300, n)
    boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
    scores = rs.uniform(0, 1, n).astype(np.float32) ** 2
    return boxes, scores


def reference_NMS_function(ref_c, nms_thresh):
    """utils.py cannot be imported (it pulls the dataset stack), so its NMS() is exec'd from the source text,
    lines 312-317, with the names it uses bound to the reference's own objects."""
    import easydict  # noqa: F401  (shim installed by ref_loader)
    src = open("/root/reference/utils.py").read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def NMS("))
    body = "\n".join(src[start:start + 6])
    from model.roi_layers import nms as ref_nms
    from model.utils.config import cfg
    cfg.TEST.NMS = nms_thresh
    ns = {"torch": torch, "nms": ref_nms, "cfg": cfg}
    exec(compile(body, "utils.py:NMS", "exec"), ns)
    return ns["NMS"]


def sample(t, step=7):
    return t.detach().reshape(-1)[::step].numpy().copy()


def main():
    os.makedirs(GOLD, exist_ok=True)
    model = ref_loader.load()
    from model.utils.config import cfg, cfg_from_file, cfg_from_list
    cfg_from_file("/root/reference/cfgs/res50.yml")
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    ref_c = model._C

    # anchors: the reference function (generate_anchors.py:45) at its default and at the CLI default scales
    from model.rpn.generate_anchors import generate_anchors
    np.savez(os.path.join(GOLD, "anchors.npz"), default9=generate_anchors(),
             scales4_32=generate_anchors(scales=np.array([4, 8, 16, 32]), ratios=np.array([0.5, 1, 2])))

    # nms: the reference's compiled CPU kernel (csrc/cpu/nms_cpu.cpp)
    out = {}
    for name in ("random300", "clustered1000", "ties", "big6000"):
        boxes, scores, thr = nms_case(name)
        out[name] = ref_c.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
    np.savez(os.path.join(GOLD, "nms.npz"), **out)

    # roi_align: the reference's compiled CPU kernel (csrc/cpu/ROIAlign_cpu.cpp)
    feat, rois = roi_align_case()
    ra = ref_c.roi_align_forward(torch.from_numpy(feat), torch.from_numpy(rois), 1.0 / 16, 7, 7, 0).numpy()
    ra2 = ref_c.roi_align_forward(torch.from_numpy(feat), torch.from_numpy(rois), 1.0 / 16, 7, 7, 2).numpy()
    np.savez(os.path.join(GOLD, "roi_align.npz"), adaptive=ra, ratio2=ra2)

    # proposal layer: the reference module (rpn/proposal_layer.py)
    from model.rpn.proposal_layer import _ProposalLayer
    prob, bbox, im_info = proposal_case()
    layer = _ProposalLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
    rois_test = layer((torch.from_numpy(prob), torch.from_numpy(bbox), torch.from_numpy(im_info), "TEST")).numpy()
    np.savez(os.path.join(GOLD, "proposals.npz"), rois_test=rois_test)

    # full eval forward of the reference DAnARCNN with deterministic weights
    fc = FORWARD_CASE
    from model.framework.dana import DAnARCNN
    torch.manual_seed(0)
    net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_way=2,
                   num_shot=fc["n_shot"])
    net.create_architecture()
    params = O.make_params(fc["seed"], attn_std=fc["attn_std"])
    missing, unexpected = net.load_state_dict(params, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    net.eval()
    im, info, sup = O.synth_inputs(fc["seed"], 1, fc["height"], fc["width"], fc["n_shot"])
    cap = {}
    net.RCNN_base.register_forward_hook(lambda m, i, o: cap.setdefault("base", []).append(o.detach()))
    net.RCNN_rpn.register_forward_hook(lambda m, i, o: cap.__setitem__("corr", i[0].detach()))
    net.RCNN_roi_align.register_forward_hook(lambda m, i, o: cap.__setitem__("pooled", o.detach()))
    with torch.no_grad():
        rois, cls_prob, bbox_pred, *_ = net(im, info, torch.zeros(1, 1, 5), torch.zeros(1), sup)
    base_feat, sup_feat = cap["base"][0], cap["base"][1]
    dense = cap["corr"][:, 1024:]
    np.savez(os.path.join(GOLD, "forward_small.npz"),
             rois=rois.numpy(), cls_prob=cls_prob.numpy(), bbox_pred=bbox_pred.numpy(),
             base_feat_sample=sample(base_feat), base_feat_abs_sum=float(base_feat.abs().sum()),
             support_feat_sample=sample(sup_feat, 97), dense_sample=sample(dense),
             dense_abs_sum=float(dense.abs().sum()), pooled_sample=sample(cap["pooled"], 101),
             base_shape=np.array(base_feat.shape), dense_shape=np.array(dense.shape))
    # the same at the headline shape (one 600x1000 query, 3 shots): sampled outputs of the UNMODIFIED reference
    ff = FORWARD_FULL_CASE
    net3 = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_way=2,
                    num_shot=ff["n_shot"])
    net3.create_architecture()
    missing, unexpected = net3.load_state_dict(O.make_params(ff["seed"], attn_std=ff["attn_std"]), strict=False)
    assert not unexpected
    net3.eval()
    im, info, sup = O.synth_inputs(ff["seed"], 1, ff["height"], ff["width"], ff["n_shot"])
    cap = {}
    net3.RCNN_base.register_forward_hook(lambda m, i, o: cap.setdefault("base", []).append(o.detach()))
    net3.RCNN_rpn.register_forward_hook(lambda m, i, o: cap.__setitem__("corr", i[0].detach()))
    net3.RCNN_roi_align.register_forward_hook(lambda m, i, o: cap.__setitem__("pooled", o.detach()))
    with torch.no_grad():
        rois, cls_prob, bbox_pred, *_ = net3(im, info, torch.zeros(1, 1, 5), torch.zeros(1), sup)
    base_feat = cap["base"][0]
    dense = cap["corr"][:, 1024:]
    np.savez(os.path.join(GOLD, "forward_full.npz"), rois=rois.numpy(), cls_prob=cls_prob.numpy(),
             bbox_pred=bbox_pred.numpy(), base_feat_sample=sample(base_feat, 53), dense_sample=sample(dense, 53),
             pooled_sample=sample(cap["pooled"], 1009), base_shape=np.array(base_feat.shape))

    # detections post-processing: the reference's NMS() (utils.py:312-317) exec'd from its source
    boxes, scores = postprocess_case()
    NMS = reference_NMS_function(ref_c, 0.3)
    dets = NMS(torch.from_numpy(boxes), torch.from_numpy(scores))
    np.savez(os.path.join(GOLD, "postprocess.npz"), dets=dets.numpy())
    print("golden vectors written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print("  %-22s %8d bytes" % (f, os.path.getsize(os.path.join(GOLD, f))))


if __name__ == "__main__":
    main()
