"""Pins the oracle (oracle/dana_oracle.py + oracle/c/dana_oracle.c) against golden vectors produced by
the unmodified reference (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import dana_oracle as O
import make_golden as MG


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_anchor_known_answer():
    # the table the reference carries in a comment (lib/model/rpn/generate_anchors.py:17-37)
    want = np.array([[-83, -39, 100, 56], [-175, -87, 192, 104], [-359, -183, 376, 200], [-55, -55, 72, 72],
                     [-119, -119, 136, 136], [-247, -247, 264, 264], [-35, -79, 52, 96], [-79, -167, 96, 184],
                     [-167, -343, 184, 360]], dtype=np.float64)
    # that table is the 1-based MATLAB output; the python function works on the 0-based window
    # (0, 0, 15, 15) (generate_anchors.py:52), hence the uniform -1
    np.testing.assert_array_equal(O.generate_anchors(), want - 1)


def test_anchors_vs_reference(golden_dir):
    g = _g(golden_dir, "anchors.npz")
    np.testing.assert_array_equal(O.generate_anchors(), g["default9"])
    np.testing.assert_array_equal(O.generate_anchors(scales=(4, 8, 16, 32)), g["scales4_32"])


@pytest.mark.parametrize("case", ["random300", "clustered1000", "ties", "big6000"])
def test_nms_vs_reference(golden_dir, case):
    boxes, scores, thr = MG.nms_case(case)
    keep = O.nms(boxes, scores, thr).numpy()
    np.testing.assert_array_equal(keep, _g(golden_dir, "nms.npz")[case])


def test_nms_edge_cases():
    assert O.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).numel() == 0
    # IoU exactly 0.5 at thr 0.5 is suppressed (>=), SURVEY.md section 7 hard part 3
    k = O.nms(np.array([[0, 0, 9, 9], [0, 0, 9, 4]], np.float32), np.array([.9, .8], np.float32), 0.5)
    assert k.tolist() == [0]
    k = O.nms(np.array([[0, 0, 9, 9], [50, 50, 60, 60], [0, 0, 9, 9], [51, 51, 60, 60]], np.float32),
              np.array([.1, .5, .9, .7], np.float32), 0.5)
    assert k.tolist() == [2, 3]


def test_roi_align_vs_reference(golden_dir):
    feat, rois = MG.roi_align_case()
    g = _g(golden_dir, "roi_align.npz")
    np.testing.assert_array_equal(O.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0).numpy(), g["adaptive"])
    np.testing.assert_array_equal(O.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 2).numpy(), g["ratio2"])


def test_proposal_layer_vs_reference(golden_dir):
    prob, bbox, im_info = MG.proposal_case()
    rois = O.proposal_layer(torch.from_numpy(prob), torch.from_numpy(bbox), torch.from_numpy(im_info),
                            O.generate_anchors(scales=(4, 8, 16, 32)), 16, 6000, 300, 0.7)
    np.testing.assert_array_equal(rois.numpy(), _g(golden_dir, "proposals.npz")["rois_test"])


@pytest.fixture(scope="module")
def forward_small():
    fc = MG.FORWARD_CASE
    p = O.make_params(fc["seed"], attn_std=fc["attn_std"])
    im, info, sup = O.synth_inputs(fc["seed"], 1, fc["height"], fc["width"], fc["n_shot"])
    with torch.no_grad():
        return O.dana_forward_eval(p, im, info, sup, fc["n_shot"])


def test_forward_vs_reference(golden_dir, forward_small):
    g = _g(golden_dir, "forward_small.npz")
    out = forward_small
    assert tuple(out["base_feat"].shape) == tuple(g["base_shape"])
    # identical torch ops in the identical order -> bit-exact on the same machine; allow 1e-6 rel
    # so the pin also holds across CPU kernels (different BLAS blocking on the GPU box's host)
    def close(a, b, tol=2e-5):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), np.abs(a - b).max()
    close(MG.sample(out["base_feat"]), g["base_feat_sample"])
    close(MG.sample(out["support_feat"].reshape(-1, *out["support_feat"].shape[3:]), 97), g["support_feat_sample"])
    close(MG.sample(out["dense"]), g["dense_sample"])
    close(float(out["dense"].abs().sum()), g["dense_abs_sum"], 1e-5)
    close(MG.sample(out["pooled"], 101), g["pooled_sample"])
    close(out["rois"].numpy(), g["rois"], 1e-5)
    close(out["bbox_pred"].numpy(), g["bbox_pred"], 1e-4)
    close(out["cls_prob"].numpy(), g["cls_prob"], 1e-4)


def test_forward_full_size_vs_reference(golden_dir):
    """The oracle against the UNMODIFIED reference at the headline shape itself (one 600x1000 query, 3 shots;
    oracle/make_golden.py FORWARD_FULL_CASE): sampled base / dense / pooled features, rois, cls_prob, bbox_pred."""
    g = _g(golden_dir, "forward_full.npz")
    ff = MG.FORWARD_FULL_CASE
    p = O.make_params(ff["seed"], attn_std=ff["attn_std"])
    im, info, sup = O.synth_inputs(ff["seed"], 1, ff["height"], ff["width"], ff["n_shot"])
    with torch.no_grad():
        out = O.dana_forward_eval(p, im, info, sup, ff["n_shot"])

    def close(a, b, tol=2e-5):
        a = np.asarray(a, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), np.abs(a - b).max()
    assert tuple(out["base_feat"].shape) == tuple(g["base_shape"]) == (1, 1024, 38, 63)
    close(MG.sample(out["base_feat"], 53), g["base_feat_sample"])
    close(MG.sample(out["dense"], 53), g["dense_sample"])
    close(MG.sample(out["pooled"], 1009), g["pooled_sample"])
    close(out["rois"].numpy(), g["rois"], 1e-5)
    close(out["bbox_pred"].numpy(), g["bbox_pred"], 1e-4)
    close(out["cls_prob"].numpy(), g["cls_prob"], 1e-4)


def test_postprocess_vs_reference_NMS(golden_dir):
    """oracle/postprocess_oracle.py's sort + NMS composition against the reference's own NMS() (utils.py:312-317,
    exec'd from its source by oracle/make_golden.py): same detections in the same order."""
    boxes, scores = MG.postprocess_case()
    g = _g(golden_dir, "postprocess.npz")["dets"]
    b, s = torch.from_numpy(boxes), torch.from_numpy(scores)
    order = torch.sort(s, dim=0, descending=True, stable=True)[1]
    dets = torch.cat((b, s.unsqueeze(1)), 1)[order]
    keep = O.nms(b[order].numpy(), s[order].numpy(), 0.3)
    np.testing.assert_array_equal(dets[keep.view(-1).long()].numpy(), g)


def test_fsod_attention_feature_vs_reference(golden_dir):
    """oracle/fsod_oracle.py (sibling model FSOD's attention-RPN feature, fsod.py:90-112) against the UNMODIFIED
    reference FSOD module (oracle/make_golden_fsod.py hooks the input of its RPN)."""
    import fsod_oracle as FO
    import make_golden_fsod as MF
    g = _g(golden_dir, "fsod_attention.npz")
    fc = MF.FSOD_CASE
    p = O.make_params(fc["seed"])
    im, info, sup = O.synth_inputs(fc["input_seed"], 1, fc["height"], fc["width"], fc["n_shot"])
    with torch.no_grad():
        corr = FO.fsod_attention_feature(p, im, sup, fc["n_shot"])
    assert tuple(corr.shape) == tuple(g["corr_shape"])
    a, b = corr.reshape(-1)[::5].double().numpy(), g["corr_sample"].astype(np.float64)
    assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()


def test_ref_extension_matches_oracle_when_present():
    """The reference's own compiled CPU operators (oracle/_ref) travel with the repo; when present they
    must agree with the C restatement on a fresh random case (bit-exact)."""
    import build_ref
    ref_c = build_ref.load()
    if ref_c is None:
        pytest.skip("oracle/_ref not built")
    rs = np.random.RandomState(5)
    n = 2000
    x1, y1 = rs.uniform(0, 600, n), rs.uniform(0, 400, n)
    boxes = np.stack([x1, y1, x1 + rs.uniform(1, 150, n), y1 + rs.uniform(1, 150, n)], 1).astype(np.float32)
    scores = (rs.permutation(n) / n).astype(np.float32)
    np.testing.assert_array_equal(O.nms(boxes, scores, 0.6).numpy(),
                                  ref_c.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.6).numpy())
    feat = rs.standard_normal((1, 5, 20, 31)).astype(np.float32)
    # NB the reference CPU kernel reads rois.data<T>() without .contiguous() (ROIAlign_cpu.cpp:254):
    # hand it C-contiguous rois
    rois = np.ascontiguousarray(
        np.concatenate([np.zeros((30, 1)), np.sort(rs.uniform(-10, 500, (30, 4)), 1)], 1).astype(np.float32))
    a = O.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0).numpy()
    b = ref_c.roi_align_forward(torch.from_numpy(feat), torch.from_numpy(rois), 1.0 / 16, 7, 7, 0).numpy()
    np.testing.assert_array_equal(a, b)


# ------------------------------------------------------------------ episode construction
@pytest.mark.parametrize("case", ["down", "up", "same", "thin", "one"])
def test_episode_resize_vs_cv2(golden_dir, case):
    """oracle/episode_oracle.py against the real cv2 (oracle/make_golden_episode.py): bit-exact with cv2's generic
    C++ path, within 1e-3 (8-bit pixel scale) of the default SIMD build the reference's loaders actually run."""
    import episode_oracle as E
    g = _g(golden_dir, "episode_cv2.npz")
    src, dst = g[case + "_src"], g[case + "_dst"]
    got = E.resize_linear_f32(src, dst.shape[1], dst.shape[0])
    np.testing.assert_array_equal(got, g[case + "_dst_generic"])
    np.testing.assert_allclose(got, dst, rtol=0, atol=1e-3)


def test_episode_prep_im_vs_cv2(golden_dir):
    import episode_oracle as E
    g = _g(golden_dir, "episode_cv2.npz")
    got, scale = E.prep_im_for_blob(g["prep_im"], [102.9801, 115.9465, 122.7717], 48)
    assert scale == float(g["prep_scale"])
    np.testing.assert_array_equal(got, g["prep_dst_generic"])
    np.testing.assert_allclose(got, g["prep_dst"], rtol=0, atol=1e-3)


def test_episode_support_crop_layout():
    """fs_loader.py:117-138 on a synthetic image: long side -> 320, zero padding, CHW."""
    import episode_oracle as E
    rs = np.random.RandomState(3)
    im = rs.standard_normal((90, 140, 3)).astype(np.float32) * 50
    out = E.support_from_box(im, (10.2, 20.7, 60.9, 50.1), 1.5, 320)
    assert out.shape == (3, 320, 320)
    # box*1.5 -> int16 (15, 31, 91, 75): w 76 > h 44 -> width 320, height int(44 * 320 / 76) = 185
    assert np.all(out[:, 185:, :] == 0) and np.any(out[:, 184, :] != 0)
    tall = E.support_from_image((rs.rand(50, 20, 3) * 255).astype(np.uint8), [1.0, 2.0, 3.0], 320)
    assert np.all(tall[:, :, 128:] == 0) and np.any(tall[:, :, 127] != 0)


# ------------------------------------------------------------------ training branch: oracle only
def test_train_forward_vs_reference(golden_dir):
    """oracle/train_oracle.py (anchor targets, RPN losses, proposal targets, both head passes, hard-negative-mined
    classification loss) against the UNMODIFIED reference run in train mode on CPU (oracle/make_golden_train.py), same
    numpy RNG seed: sampled labels identical, losses to 1e-6 relative."""
    import make_golden_train as MT
    import train_oracle as T
    g = _g(golden_dir, "forward_train_small.npz")
    tc = MT.TRAIN_CASE
    p = O.make_params(tc["seed"], attn_std=tc["attn_std"])
    im, info, gt, nb, sup = MT.train_inputs()
    np.random.seed(tc["np_seed"])
    with torch.no_grad():
        out = T.dana_forward_train(p, im, info, gt, nb, sup, tc["n_shot"])
    got = np.array([float(out[k]) for k in ("rpn_loss_cls", "rpn_loss_box", "RCNN_loss_cls", "RCNN_loss_bbox")])
    np.testing.assert_allclose(got, g["losses"], rtol=1e-6)
    np.testing.assert_array_equal(out["rois_label"].numpy(), g["rois_label"])
    np.testing.assert_array_equal(out["rpn_targets"][0].numpy().astype(np.int8), g["rpn_labels"])
    np.testing.assert_allclose(out["rois"].numpy(), g["rois"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(out["cls_prob"].numpy(), g["cls_prob"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["bbox_pred"].numpy(), g["bbox_pred"], rtol=0, atol=1e-6)
    assert abs(float(out["rpn_targets"][1].abs().sum()) - float(g["rpn_bbox_targets_abs_sum"])) <= 1e-3
    assert abs(float(out["rpn_targets"][3].sum()) - float(g["rpn_outside_sum"])) <= 1e-5
    assert tuple(out["rois"].shape) == (2, 128, 5) and tuple(out["cls_prob"].shape) == (512, 2)


def test_train_backward_vs_reference(golden_dir):
    """Autograd of the oracle's training forward against the gradients of the UNMODIFIED reference's own
    loss.backward() (oracle/make_golden_train.py: per-parameter norms and 16 sampled elements; the reference's missing
    CPU RoIAlign backward is the one substituted operator, documented there)."""
    import make_golden_train as MT
    import torchvision
    import train_oracle as T
    g = _g(golden_dir, "forward_train_small.npz")
    tc = MT.TRAIN_CASE
    p = O.make_params(tc["seed"], attn_std=tc["attn_std"])
    names = [str(s) for s in g["grad_names"]]
    for n in names:
        p[n].requires_grad_(True)
    im, info, gt, nb, sup = MT.train_inputs()
    np.random.seed(tc["np_seed"])
    out = T.dana_forward_train(p, im, info, gt, nb, sup, tc["n_shot"],
                               roi_align_fn=lambda f, r, s, ph, pw, q: torchvision.ops.roi_align(f, r, (ph, pw), s, q, False))
    (out["rpn_loss_cls"] + out["rpn_loss_box"] + out["RCNN_loss_cls"] + out["RCNN_loss_bbox"]).backward()
    assert len(names) == 70 and not any(n.startswith(("RCNN_base.0.", "RCNN_base.1.", "RCNN_base.4.")) or ".bn" in n for n in names)
    # biases that cancel analytically (softmax over positions / mean-centering: unary, channel and q / k projection
    # biases) have gradients of pure rounding noise: compared on the absolute scale of the largest gradient
    tiny = 1e-7 * float(g["grad_norms"].max())
    for i, n in enumerate(names):
        gr = p[n].grad.double().reshape(-1)
        assert abs(float(gr.norm()) - g["grad_norms"][i]) <= 1e-4 * g["grad_norms"][i] + tiny, n
        smp = gr[torch.linspace(0, gr.numel() - 1, 16).long()].numpy()
        np.testing.assert_allclose(smp, g["grad_samples"][i], rtol=1e-3, atol=1e-4 * g["grad_norms"][i] + tiny, err_msg=n)


def test_train_target_layers_edge_cases():
    """Target layers on hand-made inputs: zero-padded gt columns never match, an anchor that is the best for a gt is
    positive even below the threshold, images without positives contribute no regression weight."""
    import train_oracle as T
    base = torch.from_numpy(O.generate_anchors(scales=(4, 8, 16, 32))).float()
    gt = torch.zeros(1, 5, 5)
    gt[0, 0] = torch.tensor([10.0, 12.0, 70.0, 60.0, 1.0])
    info = torch.tensor([[160.0, 224.0, 1.0]])
    np.random.seed(0)
    labels, tgt, iw, ow = T.anchor_target_layer(10, 14, gt, info, base)
    assert tuple(labels.shape) == (1, 1, 12 * 10, 14) and tuple(tgt.shape) == (1, 48, 10, 14)
    assert int((labels == 1).sum()) >= 1 and set(np.unique(labels.numpy())) <= {-1.0, 0.0, 1.0}
    assert float(iw.sum()) == 4.0 * int((labels == 1).sum())
    n_ex = int((labels >= 0).sum())
    assert abs(float(ow.sum()) - 4.0) <= 1e-5 and n_ex <= 256
    ov = T.bbox_overlaps_batch(torch.tensor([[10.0, 12.0, 70.0, 60.0], [0.0, 0.0, 0.0, 0.0]]), gt)
    assert float(ov[0, 0, 0]) == 1.0 and (ov[0, 0, 1:] == 0).all() and (ov[0, 1] == -1).all()
    rois = torch.zeros(1, 6, 5)
    rois[0, :, 1:] = torch.tensor([[10.0, 12.0, 70.0, 60.0], [12.0, 14.0, 68.0, 58.0], [100.0, 100.0, 150.0, 150.0],
                                   [0.0, 0.0, 30.0, 30.0], [150.0, 20.0, 220.0, 90.0], [5.0, 5.0, 200.0, 150.0]])
    np.random.seed(1)
    r, lab, t, iw2, ow2 = T.proposal_target_layer(rois, gt)
    assert tuple(r.shape) == (1, 128, 5) and int(lab.sum()) >= 1 and int(lab.sum()) <= 32
    assert torch.equal(iw2 > 0, ow2 > 0) and float(iw2.sum()) == 4.0 * int(lab.sum())
    loss = T.smooth_l1_loss(torch.zeros(2, 4), torch.tensor([[0.5, 2.0, 0.0, -3.0], [0.0, 0.0, 0.0, 0.0]]),
                            torch.ones(2, 4), torch.ones(2, 4))
    assert abs(float(loss) - (0.125 + 1.5 + 2.5) / 2) <= 1e-6


def _ref_available():
    import ref_loader
    return ref_loader.available()


@pytest.mark.skipif(not _ref_available(), reason="/root/reference is mounted in the build container only")
@pytest.mark.parametrize("seed,n_gt", [(0, 3), (1, 40), (2, 1)])
def test_train_target_layers_vs_reference_live(seed, n_gt):
    """The reference's own _AnchorTargetLayer / _ProposalTargetLayer (imported live) against the oracle restatement on
    random inputs, same numpy seed: every sampling branch (more than 128 positive anchors, background subsampling,
    fg + bg / bg-only proposal sampling) must produce identical labels, targets and weights."""
    import ref_loader
    import train_oracle as T
    ref_loader.load()
    from model.utils.config import cfg_from_file, cfg_from_list
    cfg_from_file("/root/reference/cfgs/res50.yml")
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    from model.rpn.anchor_target_layer import _AnchorTargetLayer
    from model.rpn.proposal_target_layer_cascade import _ProposalTargetLayer
    rs = np.random.RandomState(seed)
    b, fh, fw = 2, 24, 36
    info = torch.tensor([[fh * 16.0, fw * 16.0, 1.0]] * b)
    gt = torch.zeros(b, 50, 5)
    for i in range(b):
        for j in range(n_gt):
            w, h = rs.uniform(30, 300), rs.uniform(30, 250)
            x1, y1 = rs.uniform(0, fw * 16 - w - 1), rs.uniform(0, fh * 16 - h - 1)
            gt[i, j] = torch.tensor([x1, y1, x1 + w, y1 + h, 1.0])
    nb = torch.tensor([n_gt] * b)
    score = torch.zeros(b, 24, fh, fw)
    ref_at = _AnchorTargetLayer(16, [4, 8, 16, 32], [0.5, 1, 2])
    np.random.seed(100 + seed)
    want = ref_at((score, gt, info, nb))
    base = torch.from_numpy(O.generate_anchors(scales=(4, 8, 16, 32))).float()
    np.random.seed(100 + seed)
    got = T.anchor_target_layer(fh, fw, gt, info, base)
    for w_, g_ in zip(want, got):
        assert torch.equal(w_, g_)
    if n_gt == 40:
        assert int((got[0] == 1).sum()) == 2 * 128          # the fg subsampling branch was taken on both images
    # proposals: jittered copies of the gt boxes (fg) + random boxes (bg), or only far-away boxes (bg-only branch)
    r = 300
    rois = torch.zeros(b, r, 5)
    for i in range(b):
        x1, y1 = rs.uniform(0, fw * 16 - 60, r), rs.uniform(0, fh * 16 - 60, r)
        boxes = np.stack([x1, y1, x1 + rs.uniform(10, 59, r), y1 + rs.uniform(10, 59, r)], 1)
        if seed != 2:
            k = min(n_gt, 20)
            boxes[:k] = gt[i, :k, :4].numpy() + rs.normal(0, 3, (k, 4))
        rois[i, :, 0] = i
        rois[i, :, 1:] = torch.from_numpy(boxes.astype(np.float32))
    ref_pt = _ProposalTargetLayer(2)
    np.random.seed(200 + seed)
    want = ref_pt(rois, gt, nb)
    np.random.seed(200 + seed)
    got = T.proposal_target_layer(rois, gt)
    for w_, g_ in zip(want, got):
        assert torch.equal(w_, g_.to(w_.dtype))
