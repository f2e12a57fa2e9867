"""Size-independent properties of the hot path at BASELINE.json's FULL sizes (configs[1]: bs 4, 600x1000 queries,
2-way 3-shot, 300 RoIs per image on a [4,1024,38,63] map; NMS at the TRAIN pre-NMS size 12000), where the oracle
would take minutes: determinism, batch-order equivariance, NMS idempotence and the no-overlap invariant, RoIAlign
linearity and partition of unity, CUDA-graph replay == eager.  The per-stage numerics against the oracle are
asserted at smaller sizes in test_gpu_forward.py / test_gpu_ops.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _iou_matrix(b):
    area = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    xx1 = torch.maximum(b[:, None, 0], b[None, :, 0])
    yy1 = torch.maximum(b[:, None, 1], b[None, :, 1])
    xx2 = torch.minimum(b[:, None, 2], b[None, :, 2])
    yy2 = torch.minimum(b[:, None, 3], b[None, :, 3])
    inter = (xx2 - xx1 + 1).clamp_min(0) * (yy2 - yy1 + 1).clamp_min(0)
    return inter / (area[:, None] + area[None, :] - inter)


@pytest.fixture(scope="module", params=["mixed", "bf16x3"])
def full_case(request):
    import dana_b200  # noqa: F401
    from dana_b200.engine import DanaEngine
    from dana_b200.synthetic import synthetic_episode, synthetic_state_dict
    eng = DanaEngine(synthetic_state_dict(1996), n_shot=3, precision=request.param)
    im, info, sup = synthetic_episode(77, 4, 600, 1000, 6)
    return eng, im.cuda(), info.cuda(), sup.cuda()


def test_full_size_forward_is_deterministic(full_case):
    """Two runs of the full benchmark step give identical bits (stream-K partial sums are combined in a fixed
    order; the NMS is a deterministic greedy scan), and every output is finite."""
    eng, im, info, sup = full_case
    a = eng.forward(im, info, sup)
    b = eng.forward(im, info, sup)
    assert tuple(a[0].shape) == (4, 300, 5) and tuple(a[1].shape) == (2 * 4 * 300, 2) and tuple(a[2].shape) == (4 * 300, 4)
    for x, y in zip(a, b):
        assert torch.isfinite(x).all()
        assert torch.equal(x, y)
    assert torch.allclose(a[1].sum(1), torch.ones_like(a[1][:, 0]), atol=1e-5)      # cls_prob rows are distributions


def test_full_size_forward_batch_equivariance(full_case):
    """Episodes are independent (the premise of the multi-GPU sharding): permuting the batch permutes
    the outputs.  Tile boundaries move with the position in the batch, so equality is to fp32 rounding, and the
    proposals (discontinuous in the scores) are compared as sets."""
    eng, im, info, sup = full_case
    # mixed mode: an fp32 accumulation-order difference can flip the fp16 rounding of an RPN activation (2^-11 relative),
    # which moves the decoded boxes of the large anchors by ~0.1 px
    tol = 0.5 if eng.precision == "mixed" else 0.05
    perm = [2, 0, 3, 1]
    rois, cls_prob, bbox = eng.forward(im, info, sup)
    rois_p, cls_p, bbox_p = eng.forward(im[perm].contiguous(), info[perm].contiguous(), sup[perm].contiguous())
    for j, i in enumerate(perm):
        a, b = rois[i, :, 1:], rois_p[j, :, 1:]
        d = (a[:, None, :] - b[None, :, :]).abs().max(2)[0].min(1)[0]
        assert (d <= tol).float().mean().item() >= 0.97
        assert (rois_p[j, :, 0] == j).all()


def test_graph_replay_equals_eager_full_size(full_case):
    from dana_b200.engine import GraphedForward
    eng, im, info, sup = full_case
    eager = eng.forward(im, info, sup)
    g = GraphedForward(eng, im, info, sup)
    for _ in range(2):
        out = g(im, info, sup)
        for x, y in zip(eager, out):
            assert torch.equal(x, y)


@pytest.mark.parametrize("n,thr", [(12000, 0.7), (6000, 0.3)])
def test_nms_idempotent_and_non_overlapping(n, thr):
    """keep = nms(boxes): (1) no two kept boxes overlap by >= thr, (2) every suppressed box overlaps a kept box of
    higher score by >= thr, (3) nms(kept boxes) keeps all of them (idempotence)."""
    import dana_b200  # noqa: F401
    from dana_b200 import ops
    rs = np.random.RandomState(n)
    x1, y1 = rs.uniform(0, 900, n), rs.uniform(0, 500, n)
    boxes = np.stack([x1, y1, x1 + rs.uniform(1, 200, n), y1 + rs.uniform(1, 200, n)], 1).astype(np.float32)
    scores = (rs.permutation(n) / n).astype(np.float32)
    b, s = torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda()
    keep = ops.nms(b, s, thr)
    assert (keep[1:] > keep[:-1]).all()                              # ascending input indices (nms_cpu.cpp:64)
    kb, ks = b[keep], s[keep]
    iou = _iou_matrix(kb.double())
    iou.fill_diagonal_(0)
    assert iou.max().item() < thr
    mask = torch.ones(n, dtype=torch.bool, device="cuda")
    mask[keep] = False
    sb, ss = b[mask].double(), s[mask]
    best = torch.zeros(sb.shape[0], dtype=torch.bool, device="cuda")
    for c0 in range(0, sb.shape[0], 2048):                           # chunked [suppressed x kept] IoU
        blk = sb[c0:c0 + 2048]
        area_k = (kb[:, 2] - kb[:, 0] + 1) * (kb[:, 3] - kb[:, 1] + 1)
        area_s = (blk[:, 2] - blk[:, 0] + 1) * (blk[:, 3] - blk[:, 1] + 1)
        iw = (torch.minimum(blk[:, None, 2], kb[None, :, 2].double()) - torch.maximum(blk[:, None, 0], kb[None, :, 0].double()) + 1).clamp_min(0)
        ih = (torch.minimum(blk[:, None, 3], kb[None, :, 3].double()) - torch.maximum(blk[:, None, 1], kb[None, :, 1].double()) + 1).clamp_min(0)
        inter = iw * ih
        ov = inter / (area_s[:, None] + area_k[None, :].double() - inter)
        higher = ks[None, :] > ss[c0:c0 + 2048, None]
        best[c0:c0 + 2048] = ((ov >= thr - 1e-6) & higher).any(1)
    assert best.all()
    again = ops.nms(kb.contiguous(), ks.contiguous(), thr)
    assert again.numel() == keep.numel()


def test_roi_align_linearity_and_partition_of_unity_full_size():
    """RoIAlign is linear in the map, and a constant map pools to the same constant wherever every sample of a bin
    lies inside the map (weights of a bin sum to one) -- on the benchmark's shapes, all three layouts."""
    import dana_b200  # noqa: F401
    from dana_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    b, c, h, w = 4, 1024, 38, 63
    f1 = torch.randn(b, h, w, c, device="cuda", generator=g)
    f2 = torch.randn(b, h, w, c, device="cuda", generator=g)
    rs = np.random.RandomState(5)
    r = 300 * b
    cx, cy = rs.uniform(0, 1000, r), rs.uniform(0, 600, r)
    bw, bh = np.exp(rs.uniform(np.log(16), np.log(600), r)), np.exp(rs.uniform(np.log(16), np.log(500), r))
    rois = np.stack([np.repeat(np.arange(b), 300), np.clip(cx - bw / 2, 0, 999), np.clip(cy - bh / 2, 0, 599),
                     np.clip(cx + bw / 2, 0, 999), np.clip(cy + bh / 2, 0, 599)], 1).astype(np.float32)
    rois = torch.from_numpy(rois).cuda()

    def pool(f):
        return ops.roi_align_head(f, rois, 1.0 / 16, 0, want_f32=True, want_pair=False, want_qpe=False)[0]
    o1, o2 = pool(f1), pool(f2)
    o12 = pool((0.5 * f1 - 2.0 * f2).contiguous())
    ref = 0.5 * o1 - 2.0 * o2
    assert ((o12 - ref).abs().max() / ref.abs().max()).item() <= 1e-5
    ones = pool(torch.full((b, h, w, c), 3.25, device="cuda"))
    # RoIs clipped to the image: all samples valid except beyond the last map row / column (y > 37 or x > 62 is
    # still inside [-1, H] x [-1, W], clamped) -> every bin is exactly the constant
    assert (ones - 3.25).abs().max().item() <= 1e-5
    # the reference-layout operator agrees with the pipeline layout on the same inputs
    nchw = f1.permute(0, 3, 1, 2).contiguous()
    o_ref_layout = ops.roi_align_forward(nchw, rois, 1.0 / 16, 7, 7, 0)
    assert torch.equal(o_ref_layout, o1.permute(0, 3, 1, 2).contiguous())
