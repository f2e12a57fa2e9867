"""GPU parity of the composed path (engine.DanaEngine / DAnARCNN.forward) against
  (a) the golden vectors produced by the UNMODIFIED reference (tests/golden/forward_small.npz) and
  (b) the oracle (oracle/dana_oracle.py) on seeded inputs, stage by stage.

Tolerance (north star: 1e-3 relative): every stage is asserted at 1e-3 max-norm relative in the
bf16x3 parity mode.  Proposals are discontinuous in the scores (hard part 2), so
the per-RoI stages are compared with the oracle's rois teacher-forced, and the free-running rois are
compared as a set.  The plain bf16 mode is reported against a looser, stated bound (5e-2)."""
import os

import numpy as np
import pytest
import torch

import dana_oracle as O
import make_golden as MG

pytestmark = pytest.mark.gpu
TOL = 1e-3


def relerr(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def roi_set_match(got, want, tol=0.05):
    """Fraction of reference rois (non-padded) that have a counterpart within `tol` px in `got`."""
    got, want = got.cpu(), want.cpu()
    hits = total = 0
    for i in range(want.shape[0]):
        w = want[i][(want[i, :, 1:].abs().sum(1) > 0)][:, 1:]
        g = got[i][:, 1:]
        if w.numel() == 0:
            continue
        d = (w[:, None, :] - g[None, :, :]).abs().max(2)[0].min(1)[0]
        hits += int((d <= tol).sum())
        total += w.shape[0]
    return hits / max(total, 1)


@pytest.fixture(scope="module")
def small_case():
    import dana_b200  # noqa: F401
    from dana_b200.engine import DanaEngine
    fc = MG.FORWARD_CASE
    p = O.make_params(fc["seed"], attn_std=fc["attn_std"])
    im, info, sup = O.synth_inputs(fc["seed"], 1, fc["height"], fc["width"], fc["n_shot"])
    eng = DanaEngine(p, n_shot=fc["n_shot"], precision="bf16x3")
    return p, im, info, sup, eng


def test_forward_vs_reference_golden(small_case, golden_dir):
    """Engine output vs the unmodified reference DAnARCNN.forward (eval) on the same weights / inputs."""
    p, im, info, sup, eng = small_case
    g = np.load(os.path.join(golden_dir, "forward_small.npz"))
    want = ("base_feat", "support_feat", "dense", "pooled")
    rois, cls_prob, bbox, ex = eng.forward(im.cuda(), info.cuda(), sup.cuda(), want=want)
    assert relerr(MG.sample(ex["base_feat"].cpu()), g["base_feat_sample"]) <= TOL
    assert relerr(MG.sample(ex["support_feat"].cpu(), 97), g["support_feat_sample"]) <= TOL
    assert relerr(MG.sample(ex["dense"].cpu().contiguous()), g["dense_sample"]) <= TOL
    assert roi_set_match(rois, torch.from_numpy(g["rois"])) >= 0.97
    # per-RoI outputs with the reference's rois teacher-forced
    rois2, cls_prob, bbox, ex = eng.forward(im.cuda(), info.cuda(), sup.cuda(), want=want,
                                            teacher={"rois": torch.from_numpy(g["rois"]).cuda()})
    assert relerr(MG.sample(ex["pooled"].cpu().contiguous(), 101), g["pooled_sample"]) <= TOL
    assert relerr(bbox, g["bbox_pred"]) <= TOL
    assert relerr(cls_prob, g["cls_prob"]) <= TOL


@pytest.mark.parametrize("precision,tol", [("mixed", TOL), ("bf16x3", TOL), ("bf16", 5e-2)])
def test_forward_stages_vs_oracle(precision, tol):
    """2 support sets x 2 shots, every stage against the oracle (free-running trunk, teacher-forced rois)."""
    import dana_b200  # noqa: F401
    from dana_b200.engine import DanaEngine
    k, sets = 2, 2
    p = O.make_params(1996, attn_std=0.05)
    im, info, sup = O.synth_inputs(7, 2, 112, 176, k * sets)
    with torch.no_grad():
        ref = O.dana_forward_eval(p, im, info, sup, k)
    eng = DanaEngine(p, n_shot=k, precision=precision)
    want = ("base_feat", "support_feat", "dense", "pooled", "fc7", "cls_score", "support_pooled", "rpn_fg")
    rois, cls_prob, bbox, ex = eng.forward(im.cuda(), info.cuda(), sup.cuda(), want=want)
    assert relerr(ex["base_feat"], ref["base_feat"]) <= tol
    assert relerr(ex["support_feat"], ref["support_feat"].reshape(-1, 1024, 20, 20)) <= tol
    assert relerr(ex["support_pooled"], ref["support_pooled"].reshape(-1, 1024, 7, 7)) <= tol
    assert relerr(ex["dense"], ref["dense"]) <= tol
    assert relerr(ex["rpn_fg"], ref["rpn_cls_prob"][:, 12:].permute(0, 2, 3, 1).reshape(2, -1)) <= tol
    assert relerr(ex["rpn_deltas"], ref["rpn_bbox_pred"].permute(0, 2, 3, 1).reshape(2, -1, 4)) <= tol
    if precision == "bf16x3":
        assert roi_set_match(rois, ref["rois"]) >= 0.97
    if precision == "mixed":    # fp16 RPN deltas (4e-4 relative) move the boxes of the large anchors by up to ~0.2 px
        assert roi_set_match(rois, ref["rois"], tol=0.5) >= 0.97
    rois, cls_prob, bbox, ex = eng.forward(im.cuda(), info.cuda(), sup.cuda(), want=want,
                                           teacher={"rois": ref["rois"].cuda()})
    assert relerr(ex["pooled"], ref["pooled"]) <= tol
    assert relerr(ex["fc7"], ref["fc7"]) <= tol
    assert relerr(bbox, ref["bbox_pred"]) <= tol
    assert relerr(ex["cls_score"], ref["cls_score"]) <= tol * 2   # 2-logit scores are O(0.1): looser in max-norm
    assert relerr(cls_prob, ref["cls_prob"]) <= tol
    assert tuple(cls_prob.shape) == (sets * 2 * 300, 2) and tuple(bbox.shape) == (2 * 300, 4)


@pytest.mark.parametrize("nq_hw,ns_hw,shots", [((38, 50), (20, 20), 1), ((38, 50), (20, 20), 3), ((38, 50), (14, 14), 1),
                                               ((38, 50), (14, 14), 3), ((38, 63), (20, 20), 6), ((38, 50), (14, 14), 6)])
@pytest.mark.parametrize("attn_std,feat_scale", [(0.01, 1.0), (0.05, 1.0), (0.05, 4.0)])
def test_ba_cisa_block_vs_oracle(nq_hw, ns_hw, shots, attn_std, feat_scale):
    """The lifted BA+CISA block (BASELINE.json configs 1/5): (Nq, Ns) in {(1900,400),(1900,196),(2394,400)} x
    units in {1,3,6}; post-ReLU N(0,1) features.  Logit range, stated with the tolerance:
    |logit| <= 0.3 at attn_std 0.01, <= ~6 at 0.05, and ~75 with the features scaled x4 (trained-like scale, where
    plain bf16 operands are off by 3e-2 and only the split operands hold 1e-3)."""
    import dana_b200  # noqa: F401
    from dana_b200.engine import DanaEngine
    p = O.make_params(11, attn_std=attn_std)
    rs = np.random.RandomState(shots * 100 + ns_hw[0])
    base = torch.from_numpy(np.maximum(rs.standard_normal((1, 1024) + nq_hw), 0).astype(np.float32)) * feat_scale
    sup = torch.from_numpy(np.maximum(rs.standard_normal((1, shots, 1024) + ns_hw), 0).astype(np.float32)) * feat_scale
    with torch.no_grad():
        want = O.ba_cisa_rpn(base, sup, p, True)
    eng = DanaEngine(p, n_shot=shots, precision="bf16x3")
    got = eng.ba_cisa_block(base.cuda(), sup.cuda())
    assert relerr(got, want) <= TOL
    if feat_scale == 1.0:
        eng16 = DanaEngine(p, n_shot=shots, precision="bf16")
        assert relerr(eng16.ba_cisa_block(base.cuda(), sup.cuda()), want) <= 2e-2


def test_module_boundary_eval_forward():
    """DAnARCNN mirror: reference ctor / create_architecture / load_state_dict / 8-tuple eval forward."""
    import dana_b200  # noqa: F401
    from dana_b200.config import cfg, cfg_from_file, cfg_from_list, reset_cfg
    from dana_b200.dana import DAnARCNN
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    reset_cfg()
    cfg_from_file(os.path.join(root, "cfgs", "res50.yml"))
    cfg_from_list(["ANCHOR_SCALES", "[4,8,16,32]", "ANCHOR_RATIOS", "[0.5,1,2]", "MAX_NUM_GT_BOXES", "50"])
    fc = MG.FORWARD_CASE
    net = DAnARCNN(["bg", "fg"], "concat", 256, 256, pretrained=False, semantic_enhance=True, num_way=2,
                   num_shot=fc["n_shot"])
    net.create_architecture()
    missing, unexpected = net.load_state_dict(O.make_params(fc["seed"], attn_std=fc["attn_std"]), strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing)
    net.cuda().eval()
    im, info, sup = O.synth_inputs(fc["seed"], 1, fc["height"], fc["width"], fc["n_shot"])
    out = net(im.cuda(), info.cuda(), torch.zeros(1, 1, 5).cuda(), torch.zeros(1).cuda(), sup.cuda())
    assert len(out) == 8 and out[3:7] == (0, 0, 0, 0) and out[7] is None
    g = np.load(os.path.join(root, "tests", "golden", "forward_small.npz"))
    assert tuple(out[0].shape) == tuple(g["rois"].shape)
    assert roi_set_match(out[0], torch.from_numpy(g["rois"]), tol=0.5) >= 0.97     # default precision: mixed
    with pytest.raises(NotImplementedError):
        net.train()(im.cuda(), info.cuda(), torch.zeros(1, 1, 5).cuda(), torch.zeros(1).cuda(), sup.cuda())


def test_support_feature_cache_matches_full_forward(small_case):
    """encode_supports() once + forward(support_feats=...) == forward on the raw crops (bitwise: same kernels)."""
    p, im, info, sup, eng = small_case
    a = eng.forward(im.cuda(), info.cuda(), sup.cuda())
    cache = eng.encode_supports(sup.cuda())
    b = eng.forward(im.cuda(), info.cuda(), None, support_feats=cache)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_res101_five_sets_vs_oracle():
    """BASELINE.json configs[3] in miniature: the res101 trunk (layer3 = 23 blocks) with 5 support sets x 1 shot,
    every stage against the oracle.  (the reference builds resnet50 for every
    num_layers and ignores n_way in eval; both generalisations are stated in DESIGN.md.)"""
    import dana_b200  # noqa: F401
    from dana_b200.engine import DanaEngine
    k, sets = 1, 5
    p = O.make_params(101, num_layers=101, attn_std=0.05)
    im, info, sup = O.synth_inputs(11, 1, 96, 160, k * sets)
    with torch.no_grad():
        ref = O.dana_forward_eval(p, im, info, sup, k, cfg={"num_layers": 101})
    eng = DanaEngine(p, num_layers=101, n_shot=k, precision="bf16x3")
    want = ("base_feat", "dense", "pooled", "fc7")
    rois, cls_prob, bbox, ex = eng.forward(im.cuda(), info.cuda(), sup.cuda(), want=want,
                                           teacher={"rois": ref["rois"].cuda()})
    assert relerr(ex["base_feat"], ref["base_feat"]) <= TOL
    assert relerr(ex["dense"], ref["dense"]) <= TOL
    assert relerr(ex["pooled"], ref["pooled"]) <= TOL
    assert relerr(ex["fc7"], ref["fc7"]) <= TOL
    assert relerr(bbox, ref["bbox_pred"]) <= TOL
    assert relerr(cls_prob, ref["cls_prob"]) <= TOL
    assert tuple(cls_prob.shape) == (sets * 300, 2)


@pytest.fixture(scope="module")
def full_size_ref():
    k, sets = 3, 2
    p = O.make_params(1996, attn_std=0.05)
    im, info, sup = O.synth_inputs(21, 1, 600, 1000, k * sets)
    with torch.no_grad():
        ref = O.dana_forward_eval(p, im, info, sup, k)
    return p, im, info, sup, ref


@pytest.mark.parametrize("precision", ["mixed", "bf16x3"])
def test_full_size_query_vs_oracle(full_size_ref, precision):
    """The headline configuration itself (BASELINE.json configs[1], one episode of it): a 600x1000 query with 2 support
    sets x 3 shots of 320x320 crops, every stage against the oracle at the north-star tolerance (1e-3, max-norm), in
    both parity modes (the benchmarked mixed mode and the all-bf16x3 mode)."""
    import dana_b200  # noqa: F401
    from dana_b200.engine import DanaEngine
    k, sets = 3, 2
    p, im, info, sup, ref = full_size_ref
    eng = DanaEngine(p, n_shot=k, precision=precision)
    want = ("base_feat", "dense", "pooled", "fc7", "rpn_fg")
    rois, cls_prob, bbox, ex = eng.forward(im.cuda(), info.cuda(), sup.cuda(), want=want)
    assert tuple(ex["base_feat"].shape) == (1, 1024, 38, 63)
    assert relerr(ex["base_feat"], ref["base_feat"]) <= TOL
    assert relerr(ex["dense"], ref["dense"]) <= TOL
    assert relerr(ex["rpn_fg"], ref["rpn_cls_prob"][:, 12:].permute(0, 2, 3, 1).reshape(1, -1)) <= TOL
    assert roi_set_match(rois, ref["rois"], tol=0.05 if precision == "bf16x3" else 0.5) >= 0.97
    rois, cls_prob, bbox, ex = eng.forward(im.cuda(), info.cuda(), sup.cuda(), want=want,
                                           teacher={"rois": ref["rois"].cuda()})
    assert relerr(ex["pooled"], ref["pooled"]) <= TOL
    assert relerr(ex["fc7"], ref["fc7"]) <= TOL
    assert relerr(bbox, ref["bbox_pred"]) <= TOL
    assert relerr(cls_prob, ref["cls_prob"]) <= TOL
    assert tuple(cls_prob.shape) == (sets * 300, 2)
