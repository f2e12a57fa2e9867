"""GPU parity tests of the operator-level C ABI (dana_nms, dana_proposals, dana_roi_align_*,
dana_conv_gemm, stem) against the oracle and the golden vectors of the unmodified reference.

Bars: bit-exact for NMS keep indices and proposal selection on identical inputs; 1e-5 relative
(max-norm) for fp32 gathers; the tensor-core GEMM is checked operand-exact (against the same
bf16-rounded operands in fp64) and, in bf16x3 mode, against unrounded fp32 operands at 2e-5."""
import os

import numpy as np
import pytest
import torch

import dana_oracle as O
import make_golden as MG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import dana_b200  # noqa: F401
    from dana_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------ NMS
@pytest.mark.parametrize("case", ["random300", "clustered1000", "ties", "big6000"])
def test_nms_golden(ops, golden_dir, case):
    boxes, scores, thr = MG.nms_case(case)
    keep = ops.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr)
    assert keep.dtype == torch.int64 and keep.is_cuda
    want = np.load(os.path.join(golden_dir, "nms.npz"))[case]
    np.testing.assert_array_equal(keep.cpu().numpy(), want)          # reference's own CPU kernel
    np.testing.assert_array_equal(keep.cpu().numpy(), O.nms(boxes, scores, thr).numpy())


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 129, 1000, 12000])
@pytest.mark.parametrize("thr", [0.3, 0.7])
def test_nms_random_vs_oracle(ops, n, thr):
    rs = np.random.RandomState(100 + n)
    x1, y1 = rs.uniform(0, 900, n), rs.uniform(0, 500, n)
    boxes = np.stack([x1, y1, x1 + rs.uniform(1, 200, n), y1 + rs.uniform(1, 200, n)], 1).astype(np.float32)
    if n >= 129:  # jittered copies: more than half get suppressed
        boxes[n // 2:] = boxes[: n - n // 2] + rs.normal(0, 2, (n - n // 2, 4)).astype(np.float32)
    scores = (rs.permutation(n) / max(n, 1)).astype(np.float32)
    keep = ops.nms(torch.from_numpy(boxes).cuda(), torch.from_numpy(scores).cuda(), thr)
    np.testing.assert_array_equal(keep.cpu().numpy(), O.nms(boxes, scores, thr).numpy())


def test_nms_edge_cases(ops):
    e = ops.nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), 0.5)
    assert e.numel() == 0 and e.dtype == torch.int64 and e.device.type == "cpu"   # csrc/nms.h:17-18
    # IoU == thr suppresses (>=): the CUDA kernel of the reference used >
    k = ops.nms(torch.tensor([[0., 0, 9, 9], [0, 0, 9, 4]]).cuda(), torch.tensor([.9, .8]).cuda(), 0.5)
    assert k.tolist() == [0]
    # tied scores: lower input index wins on both sides
    b = torch.tensor([[0., 0, 9, 9], [0, 0, 9, 9], [50, 50, 60, 60], [50, 50, 60, 60]])
    s = torch.tensor([.5, .5, .7, .7])
    assert ops.nms(b.cuda(), s.cuda(), 0.5).tolist() == O.nms(b.numpy(), s.numpy(), 0.5).tolist() == [0, 2]
    # saturated scores (softmax exactly 1.0f): many ties
    rs = np.random.RandomState(3)
    n = 500
    x1, y1 = rs.uniform(0, 300, n), rs.uniform(0, 300, n)
    bb = np.stack([x1, y1, x1 + rs.uniform(5, 80, n), y1 + rs.uniform(5, 80, n)], 1).astype(np.float32)
    ss = np.where(rs.uniform(size=n) < 0.5, 1.0, rs.uniform(size=n)).astype(np.float32)
    np.testing.assert_array_equal(ops.nms(torch.from_numpy(bb).cuda(), torch.from_numpy(ss).cuda(), 0.7).cpu().numpy(),
                                  O.nms(bb, ss, 0.7).numpy())


# ------------------------------------------------------------------------------------ RoIAlign
def test_roi_align_golden(ops, golden_dir):
    feat, rois = MG.roi_align_case()
    g = np.load(os.path.join(golden_dir, "roi_align.npz"))
    f, r = torch.from_numpy(feat).cuda(), torch.from_numpy(rois).cuda()
    for ratio, key in ((0, "adaptive"), (2, "ratio2")):
        out = ops.roi_align_forward(f, r, 1.0 / 16, 7, 7, ratio)
        assert relerr(out, torch.from_numpy(g[key])) <= 1e-5
    out_f32, pair = ops.roi_align_nhwc(f.permute(0, 2, 3, 1).contiguous(), r, 1.0 / 16, 7, 0)
    assert relerr(out_f32.permute(0, 3, 1, 2), torch.from_numpy(g["adaptive"])) <= 1e-5
    assert relerr(pair.float().permute(0, 3, 1, 2), torch.from_numpy(g["adaptive"])) <= 2e-5


def test_roi_align_full_size_vs_oracle(ops):
    """SURVEY.md 8d KAT: map [2,1024,38,63], 300 rois incl. sub-pixel, full-image, out-of-bounds, degenerate."""
    rs = np.random.RandomState(42)
    feat = rs.standard_normal((2, 1024, 38, 63)).astype(np.float32)
    r = 300
    x1, y1 = rs.uniform(-30, 980, r), rs.uniform(-30, 590, r)
    rois = np.stack([rs.randint(0, 2, r).astype(np.float64), x1, y1, x1 + rs.uniform(0.3, 600, r),
                     y1 + rs.uniform(0.3, 400, r)], 1).astype(np.float32)
    rois[0, 1:] = [0, 0, 999, 599]
    rois[1, 1:] = [500, 300, 400, 200]
    rois[2, 1:] = [1500, 900, 1600, 950]
    rois[3, 1:] = [10.2, 10.7, 10.9, 11.1]
    want = O.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0)
    got = ops.roi_align_forward(torch.from_numpy(feat).cuda(), torch.from_numpy(rois).cuda(), 1.0 / 16, 7, 7, 0)
    assert tuple(got.shape) == (300, 1024, 7, 7)
    assert relerr(got, want) <= 1e-5
    # empty roi list
    e = ops.roi_align_forward(torch.from_numpy(feat).cuda(), torch.zeros(0, 5).cuda(), 1.0 / 16, 7, 7, 0)
    assert tuple(e.shape) == (0, 1024, 7, 7)


def test_roi_align_backward_vs_autograd(ops):
    import torchvision
    rs = np.random.RandomState(9)
    feat = torch.from_numpy(rs.standard_normal((2, 16, 20, 31))).cuda()
    rois = torch.from_numpy(np.concatenate([rs.randint(0, 2, (25, 1)).astype(np.float64),
                                            np.sort(rs.uniform(0, 300, (25, 2)), 1)[:, [0]],
                                            rs.uniform(0, 150, (25, 1)), rs.uniform(300, 480, (25, 1)),
                                            rs.uniform(150, 310, (25, 1))], 1)).cuda()
    f = feat.clone().requires_grad_(True)
    o = torchvision.ops.roi_align(f, rois, (7, 7), 1.0 / 16, 0, False)    # same sampling rules, fp64 autograd
    go = torch.randn_like(o)
    o.backward(go)
    got = ops.roi_align_backward(go.float(), rois.float(), 1.0 / 16, 7, 7, 2, 16, 20, 31, 0)
    assert relerr(got, f.grad) <= 1e-4


@pytest.mark.parametrize("b,c,h,w,ratio", [(2, 64, 38, 63, 0), (1, 1024, 20, 31, 0), (2, 8, 9, 11, 2), (1, 516, 5, 3, 0)])
def test_roi_align_backward_layouts_vs_autograd(ops, b, c, h, w, ratio):
    """Backward of the 7x7 RoIAlign as the transpose of the separable forward (NHWC, 16-byte vector reductions), both
    layouts, RoIs from sub-pixel to larger than the map (per-sample path), against fp64 autograd of torchvision's
    roi_align, which implements the same sampling rules (aligned=False).  The reference has NO CPU backward
    (csrc/ROIAlign.h:44 raises), so there is no reference-derived golden for this operator: its CUDA kernel
    (ROIAlign_cuda.cu:178-254) is the adjoint of the forward, which is what autograd differentiates."""
    import torchvision
    rs = np.random.RandomState(b * 7 + c)
    feat = torch.from_numpy(rs.standard_normal((b, c, h, w))).cuda()
    r = 60
    iw, ih = w * 16.0, h * 16.0
    x1, y1 = rs.uniform(-0.1 * iw, iw, r), rs.uniform(-0.1 * ih, ih, r)
    ww = np.exp(rs.uniform(np.log(0.5), np.log(2.5 * iw), r))
    hh = np.exp(rs.uniform(np.log(0.5), np.log(2.5 * ih), r))
    rois = torch.from_numpy(np.stack([rs.randint(0, b, r).astype(np.float64), x1, y1, x1 + ww, y1 + hh], 1)).cuda()
    f = feat.clone().requires_grad_(True)
    o = torchvision.ops.roi_align(f, rois, (7, 7), 1.0 / 16, ratio, False)
    go = torch.randn_like(o)
    o.backward(go)
    got = ops.roi_align_backward(go.float(), rois.float(), 1.0 / 16, 7, 7, b, c, h, w, ratio)
    assert relerr(got, f.grad) <= 1e-4
    if c % 4 == 0:
        g_nhwc = ops.roi_align_backward_nhwc(go.float().reshape(r, c, 49).transpose(1, 2).contiguous(), rois.float(),
                                             1.0 / 16, b, h, w, ratio)
        assert relerr(g_nhwc.permute(0, 3, 1, 2), f.grad) <= 1e-4


# ------------------------------------------------------------------------------------ proposals
def test_proposals_golden(ops, golden_dir):
    prob, bbox, im_info = MG.proposal_case()
    a = 12
    b, _, fh, fw = bbox.shape
    fg = torch.from_numpy(prob[:, a:]).permute(0, 2, 3, 1).reshape(b, -1).contiguous().cuda()
    deltas = torch.from_numpy(bbox).permute(0, 2, 3, 1).reshape(b, -1, 4).contiguous().cuda()
    base = torch.from_numpy(O.generate_anchors(scales=(4, 8, 16, 32))).float().cuda()
    rois = ops.proposals(fg, deltas, base, torch.from_numpy(im_info).cuda(), fh, fw, 16, 6000, 300, 0.7)
    want = torch.from_numpy(np.load(os.path.join(golden_dir, "proposals.npz"))["rois_test"])
    # boxes go through expf on the device vs. torch's CPU exp: <= 2 ulp on coordinates of O(100)
    assert (rois.cpu() - want).abs().max().item() <= 1e-3
    assert torch.equal(rois.cpu()[:, :, 0], want[:, :, 0])


@pytest.mark.parametrize("b,fh,fw,pre,post", [(4, 38, 63, 6000, 300), (1, 38, 50, 12000, 2000), (2, 50, 84, 6000, 1000)])
def test_proposals_full_size_vs_oracle(ops, b, fh, fw, pre, post):
    rs = np.random.RandomState(fh * fw + b)
    a = 12
    prob = torch.from_numpy(rs.uniform(0, 1, (b, 2 * a, fh, fw)).astype(np.float32))
    bbox = torch.from_numpy((rs.standard_normal((b, 4 * a, fh, fw)) * 0.3).astype(np.float32))
    im_info = torch.tensor([[fh * 16.0 - 8, fw * 16.0 - 5, 1.0]] * b)
    base = O.generate_anchors(scales=(4, 8, 16, 32))
    want, want_sc = O.proposal_layer(prob, bbox, im_info, base, 16, pre, post, 0.7, return_scores=True)
    fg = prob[:, a:].permute(0, 2, 3, 1).reshape(b, -1).contiguous().cuda()
    deltas = bbox.permute(0, 2, 3, 1).reshape(b, -1, 4).contiguous().cuda()
    rois, sc, cnt = ops.proposals(fg, deltas, torch.from_numpy(base).float().cuda(), im_info.cuda(), fh, fw, 16, pre,
                                  post, 0.7, want_scores=True)
    # selection is bit-exact: the kept scores (inputs, untouched) must be identical, row by row
    assert torch.equal(sc.cpu(), want_sc)
    assert (rois.cpu() - want).abs().max().item() <= 1e-3
    assert cnt.cpu().tolist() == [int((want_sc[i] > 0).sum()) for i in range(b)]


# ------------------------------------------------------------------------------------ tensor-core GEMM / conv
@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("m,k,n", [(128, 64, 64), (300, 192, 72), (2394, 1024, 256), (1000, 1200, 1024), (77, 3136, 1024),
                                   (1200, 2048, 4), (14700, 147, 1024), (300, 96, 512), (500, 416, 640)])
def test_linear(ops, split, m, k, n):
    torch.manual_seed(m + k + n)
    kp = (k + 7) // 8 * 8                                                   # row pitch (TMA wants 16-byte strides)
    xf = torch.randn(m, kp, device="cuda")
    wf = torch.randn(n, kp, device="cuda") * 0.05
    bias = torch.randn(n, device="cuda")
    P = ops.Pair
    xh, wh = xf.to(torch.bfloat16), wf.to(torch.bfloat16)
    xl, wl = (xf - xh.float()).to(torch.bfloat16), (wf - wh.float()).to(torch.bfloat16)
    x, w = xf[:, :k], wf[:, :k]
    xp = P(xh[:, :k], xl[:, :k] if split else None)                         # strided (pitch >= k) operands
    wp = P(wh[:, :k], wl[:, :k] if split else None)
    out = torch.empty(m, n, device="cuda")
    ops.linear(xp, wp, n, bias=bias, out_f32=out, relu=True)
    ref = torch.relu(xp.float().double() @ wp.float().double().t() + bias.double())
    assert relerr(out, ref) <= 2e-5
    if split:  # fp32-equivalent: also close to the unrounded operands
        ref32 = torch.relu(x.double() @ w.double().t() + bias.double())
        assert relerr(out, ref32) <= 3e-5


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("n,h,w,cin,cout,ks,stride", [(2, 38, 63, 256, 256, 3, 1), (3, 20, 20, 128, 128, 3, 1),
                                                      (16, 4, 4, 512, 512, 3, 1), (2, 75, 125, 256, 128, 1, 2),
                                                      (8, 7, 7, 1024, 512, 1, 2), (1, 150, 250, 64, 256, 1, 1),
                                                      (1, 38, 63, 2048, 512, 3, 1), (5, 9, 11, 64, 72, 1, 1),
                                                      # window-per-column-shift schedule (DX3): 64 / 128 / 256 channels,
                                                      # maps that do and do not divide the 8x16 / 16x8 tiles
                                                      (2, 150, 250, 64, 64, 3, 1), (1, 37, 29, 64, 64, 3, 1),
                                                      (2, 75, 125, 128, 128, 3, 1), (3, 40, 40, 128, 128, 3, 1),
                                                      (1, 5, 3, 256, 256, 3, 1), (4, 80, 80, 64, 64, 3, 1)])
def test_conv(ops, split, n, h, w, cin, cout, ks, stride):
    torch.manual_seed(h * w + cin)
    F = torch.nn.functional
    x = torch.randn(n, h, w, cin, device="cuda")
    wt = torch.randn(cout, cin, ks, ks, device="cuda") / (cin * ks * ks) ** 0.5
    scale = torch.rand(cout, device="cuda") + 0.5
    bias = torch.randn(cout, device="cuda") * 0.1
    xp = ops.Pair.from_float(x, split)
    wp = ops.Pair.from_float(wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous(), split)
    oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1
    rp = ops.Pair.from_float(torch.randn(n, oh, ow, cout, device="cuda"), split)
    out = ops.conv_nhwc(xp, wp, cout, ksize=ks, stride=stride, scale=scale, bias=bias, res=rp, relu=True, split=split)
    xr = xp.float().double().permute(0, 3, 1, 2)
    wr = wp.float().double().reshape(cout, ks, ks, cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xr, wr, stride=stride, padding=ks // 2) * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1)
    ref = torch.relu(ref + rp.float().double().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    tol = (2e-5 if cin * ks * ks < 8192 else 1e-4) if split else 6e-3   # bf16 output rounding dominates when !split
    assert relerr(out.float(), ref) <= tol


def test_stem_vs_oracle(ops):
    p = O.make_params(5)
    for (b, h, w) in [(1, 97, 131), (2, 320, 320), (1, 600, 1000)]:
        im = torch.from_numpy((np.random.RandomState(h).standard_normal((b, 3, h, w)) * 50).astype(np.float32))
        want = O.stem(im, p).permute(0, 2, 3, 1)
        g, bb, m, v = [p["RCNN_base.1." + n].cuda() for n in ("weight", "bias", "running_mean", "running_var")]
        sc = g / torch.sqrt(v + 1e-5)
        got = ops.stem(im.cuda(), ops.pack_stem_weight(p["RCNN_base.0.weight"].cuda()), sc.contiguous(), (bb - m * sc).contiguous())
        assert tuple(got.hi.shape) == tuple(want.shape)
        assert relerr(got.float(), want) <= 2e-5


def test_roi_align_head_fused_outputs(ops):
    """dana_roi_align_head: pooled fp32 / bf16 pair / pair of pooled + positional encoding, vs the oracle."""
    rs = np.random.RandomState(77)
    feat = rs.standard_normal((2, 64, 38, 63)).astype(np.float32)
    r = 200
    x1, y1 = rs.uniform(-30, 980, r), rs.uniform(-30, 590, r)
    rois = np.stack([rs.randint(0, 2, r).astype(np.float64), x1, y1, x1 + rs.uniform(0.3, 700, r),
                     y1 + rs.uniform(0.3, 500, r)], 1).astype(np.float32)
    rois[0, 1:] = [0, 0, 999, 599]
    rois[1, 1:] = [500, 300, 400, 200]
    rois[2, 1:] = [1500, 900, 1600, 950]
    rois[3, 1:] = [10.2, 10.7, 10.9, 11.1]
    want = O.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0)                       # [R,C,7,7]
    pe = O.positional_encoding(49, 64)
    nhwc = torch.from_numpy(feat).cuda().permute(0, 2, 3, 1).contiguous()
    f32, pair, qpe = ops.roi_align_head(nhwc, torch.from_numpy(rois).cuda(), 1.0 / 16, 0, pe=pe.cuda().contiguous(),
                                        want_f32=True, want_pair=True, want_qpe=True)
    assert relerr(f32.permute(0, 3, 1, 2), want) <= 1e-5
    assert relerr(pair.float().permute(0, 3, 1, 2), want) <= 2e-5
    want_q = want.reshape(r, 64, 49).transpose(1, 2) + pe.unsqueeze(0)             # [R,49,C]
    assert relerr(qpe.float().reshape(r, 49, 64), want_q) <= 2e-5


@pytest.mark.parametrize("b,c,h,w", [(1, 8, 5, 3), (2, 16, 9, 11), (1, 1024, 3, 40), (2, 512, 38, 63)])
def test_roi_align_orders_and_tiny_maps(ops, b, c, h, w):
    """Every gather order of the 7x7 kernel (two-row cache for RoIs with <= 2 row taps per bin, rolling otherwise, the
    direct path for RoIs larger than the map) and the dynamic tap loop taken when the map is narrower than the compiled tap count, against
    the oracle: RoI sizes from a fraction of a pixel to several times the map, all three layouts."""
    rs = np.random.RandomState(b * 100 + c + h + w)
    feat = rs.standard_normal((b, c, h, w)).astype(np.float32)
    r = 160
    iw, ih = w * 16.0, h * 16.0
    x1, y1 = rs.uniform(-0.2 * iw, iw, r), rs.uniform(-0.2 * ih, ih, r)
    ww = np.exp(rs.uniform(np.log(0.5), np.log(2.5 * iw), r))
    hh = np.exp(rs.uniform(np.log(0.5), np.log(2.5 * ih), r))
    rois = np.stack([rs.randint(0, b, r).astype(np.float64), x1, y1, x1 + ww, y1 + hh], 1).astype(np.float32)
    want = O.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0)
    f, rr = torch.from_numpy(feat).cuda(), torch.from_numpy(rois).cuda()
    assert relerr(ops.roi_align_forward(f, rr, 1.0 / 16, 7, 7, 0), want) <= 1e-5
    nhwc = f.permute(0, 2, 3, 1).contiguous()
    if c % 4 == 0:
        out, _, _ = ops.roi_align_head(nhwc, rr, 1.0 / 16, 0, want_f32=True, want_pair=False, want_qpe=False)
        assert relerr(out.permute(0, 3, 1, 2), want) <= 1e-5
    # fixed sampling ratio 2 (grid_h = 2 everywhere: three row taps per bin on small RoIs)
    want2 = O.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 2)
    assert relerr(ops.roi_align_forward(f, rr, 1.0 / 16, 7, 7, 2), want2) <= 1e-5


@pytest.mark.parametrize("maps,shots,ns,c,seg", [(6, 3, 49, 64, 56), (4, 1, 196, 32, 200), (10, 5, 7, 8, 8)])
def test_transpose_segments(ops, maps, shots, ns, c, seg):
    """Key-major relayout used by the head ((V W^T)^T as the B operand of the P contraction): exact data movement
    plus the bf16 pair split, pad columns zero."""
    g = torch.Generator().manual_seed(maps * ns)
    x = torch.randn(maps, ns, c, generator=g)
    pitch = (shots * seg + 7) // 8 * 8 + 8
    out = ops.transpose_segments(x.cuda(), shots, seg, pitch)
    got = out.float().cpu()                                     # hi + lo
    want = torch.zeros(maps // shots, c, pitch)
    for m in range(maps):
        want[m // shots, :, (m % shots) * seg:(m % shots) * seg + ns] = x[m].t()
    assert tuple(got.shape) == tuple(want.shape)
    assert (got - want).abs().max().item() <= 2e-5 * x.abs().max().item()
    mask = want == 0
    assert (got[mask] == 0).all()


# ------------------------------------------------------------------------------------ fp16 planes (mixed-precision mode)
@pytest.mark.parametrize("m,k,n", [(128, 64, 64), (1000, 1200, 1024), (19200, 4608, 512), (300, 512, 96)])
def test_linear_f16_operands(ops, m, k, n):
    """kind::f16 MMAs with IEEE fp16 operand planes and an fp16 single-plane output, operand-exact in fp64
    (tolerance = fp16 output rounding, 2^-11)."""
    torch.manual_seed(m + k)
    x = ops.Pair.from_float_f16(torch.randn(m, k, device="cuda"))
    w = ops.Pair.from_float_f16(torch.randn(n, k, device="cuda") * 0.05)
    bias = torch.randn(n, device="cuda")
    ref = torch.relu(x.hi.double() @ w.hi.double().t() + bias.double())
    out32 = torch.empty(m, n, device="cuda")
    ops.linear(x, w, n, bias=bias, out_f32=out32, relu=True)
    assert relerr(out32, ref) <= 2e-5
    if n % 32 == 0:
        out16 = ops.Pair.empty_f16((m, n), "cuda")
        ops.linear(x, w, n, bias=bias, out=out16, relu=True)
        assert relerr(out16.hi.float(), ref) <= 6e-4


@pytest.mark.parametrize("n,h,w,cin,cout,ks,stride", [(16, 4, 4, 512, 512, 3, 1), (8, 7, 7, 1024, 512, 1, 2),
                                                      (1, 38, 63, 2048, 512, 3, 1), (9, 4, 4, 512, 2048, 1, 1)])
def test_conv_f16_planes(ops, n, h, w, cin, cout, ks, stride):
    """layer4 / RPN convolutions of the mixed mode: fp16 activation, weight, residual and output planes."""
    torch.manual_seed(h * w + cin)
    F = torch.nn.functional
    xp = ops.Pair.from_float_f16(torch.randn(n, h, w, cin, device="cuda"))
    wt = torch.randn(cout, cin, ks, ks, device="cuda") / (cin * ks * ks) ** 0.5
    wp = ops.Pair.from_float_f16(wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous())
    scale = torch.rand(cout, device="cuda") + 0.5
    bias = torch.randn(cout, device="cuda") * 0.1
    oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1
    rp = ops.Pair.from_float_f16(torch.randn(n, oh, ow, cout, device="cuda"))
    out = ops.conv_nhwc(xp, wp, cout, ksize=ks, stride=stride, scale=scale, bias=bias, res=rp, relu=True, out_f16=True)
    assert out.is_f16 and out.lo is None
    xr = xp.hi.double().permute(0, 3, 1, 2)
    wr = wp.hi.double().reshape(cout, ks, ks, cin).permute(0, 3, 1, 2)
    ref = F.conv2d(xr, wr, stride=stride, padding=ks // 2) * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1)
    ref = torch.relu(ref + rp.hi.double().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert relerr(out.hi.float(), ref) <= 6e-4


def test_split_operands_to_f16_plane(ops):
    """The P.V contraction of the mixed mode: bf16x3 operands, per-image weights and bias, fp16 single-plane output
    written into a channel slice of a wider buffer (the dense half of the RPN input)."""
    torch.manual_seed(3)
    b, rows, k, n = 2, 700, 1200, 1024
    x = ops.Pair.from_float(torch.rand(b * rows, k, device="cuda") / k)
    w = ops.Pair.from_float(torch.randn(b, n, k, device="cuda"))
    bias = torch.randn(b, n, device="cuda") * 0.1
    buf = torch.zeros(b * rows, 2 * n, dtype=torch.float16, device="cuda")
    ops.linear(x, ops.Pair(w.hi.view(b * n, k), w.lo.view(b * n, k)), n, alpha=0.5, bias=bias, bias_sn=n,
               out=ops.Pair(buf[:, n:]), batch=b, b_batch_stride=n * k)
    ref = torch.stack([0.5 * (x.float()[i * rows:(i + 1) * rows].double() @ w.float()[i].double().t()) + bias[i].double()
                       for i in range(b)]).reshape(b * rows, n)
    assert relerr(buf[:, n:].float(), ref) <= 6e-4
    assert (buf[:, :n] == 0).all()


def test_linear_row_bias(ops):
    """(x + PE) W^T = x W^T + PE W^T: the per-bin bias table enters as a residual with a zero stride over RoIs."""
    torch.manual_seed(5)
    r, bins, k, n = 37, 49, 1024, 256
    x = torch.randn(r * bins, k, device="cuda")
    w = torch.randn(n, k, device="cuda") * 0.03
    pe = torch.randn(bins, k, device="cuda")
    table = (pe.double() @ w.double().t()).float().contiguous()
    out = torch.empty(r * bins, n, device="cuda")
    ops.linear(ops.Pair.from_float(x), ops.Pair.from_float(w), n, out_f32=out, row_bias=table)
    ref = (x.view(r, bins, k) + pe).double().view(-1, k) @ w.double().t()
    assert relerr(out, ref) <= 3e-5


def test_f16_plane_producers(ops):
    """merge_pair (+ fp16 plane into a strided slice), spatial_mean on an fp16 plane, RoIAlign fp16 output."""
    torch.manual_seed(8)
    x = torch.randn(2, 9, 11, 64, device="cuda") * 3
    xp = ops.Pair.from_float(x)
    buf = torch.zeros(2, 9, 11, 128, dtype=torch.float16, device="cuda")
    f32 = ops.merge_pair(xp, out16=buf[..., :64])
    assert relerr(f32, x) <= 2e-5
    assert torch.equal(buf[..., :64], f32.half()) and (buf[..., 64:] == 0).all()
    t = torch.randn(50, 16, 2048, device="cuda")
    m32, mp = ops.spatial_mean(ops.Pair.from_float_f16(t))
    assert relerr(m32, t.half().double().mean(1)) <= 2e-6
    assert relerr(mp.float(), m32) <= 2e-5
    m32b, _ = ops.spatial_mean(ops.Pair.from_float(t))
    assert relerr(m32b, t.double().mean(1)) <= 2e-5
    rs = np.random.RandomState(4)
    feat = torch.from_numpy(rs.standard_normal((2, 20, 31, 64)).astype(np.float32)).cuda()
    rois = torch.from_numpy(np.stack([rs.randint(0, 2, 40).astype(np.float64), rs.uniform(0, 200, 40), rs.uniform(0, 100, 40),
                                      rs.uniform(250, 480, 40), rs.uniform(150, 310, 40)], 1).astype(np.float32)).cuda()
    f32, pair, _, h16 = ops.roi_align_head(feat, rois, 1.0 / 16, 0, want_f32=True, want_pair=True, want_f16=True)
    assert torch.equal(h16.hi, f32.half())
    assert relerr(pair.float(), f32) <= 2e-5


def test_nms_strict_gt_matches_reference_cuda_rule(ops):
    """The reference's CUDA operator suppresses on IoU > thr (cuda/nms.cu:60), its CPU operator on >= (nms_cpu.cpp:60):
    _C.nms(strict_gt=True) selects the former; they differ exactly when an IoU equals the threshold."""
    from dana_b200 import _C
    b = torch.tensor([[0., 0, 9, 9], [0, 0, 9, 4]]).cuda()      # IoU = 50 / 100 = 0.5 exactly
    s = torch.tensor([.9, .8]).cuda()
    assert _C.nms(b, s, 0.5).tolist() == [0]
    assert _C.nms(b, s, 0.5, strict_gt=True).tolist() == [0, 1]
    rs = np.random.RandomState(1)
    n = 2000
    x1, y1 = rs.uniform(0, 500, n), rs.uniform(0, 300, n)
    bb = np.stack([x1, y1, x1 + rs.uniform(5, 120, n), y1 + rs.uniform(5, 120, n)], 1).astype(np.float32)
    ss = rs.permutation(n).astype(np.float32) / n
    thr = np.float32(0.7)
    want = O.nms(bb, ss, float(np.nextafter(thr, np.float32(np.inf)))).numpy()
    got = _C.nms(torch.from_numpy(bb).cuda(), torch.from_numpy(ss).cuda(), float(thr), strict_gt=True).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("m,k,n", [(300, 512, 96), (1000, 192, 160), (4000, 64, 320)])
def test_linear_pair_output_partial_last_tile(ops, m, k, n):
    """Row-contiguous (FAST) epilogue with n_out a multiple of 32 but not of the tile width: the chunks of the last
    N-tile that lie past n_out must not be written (they would land in the next row)."""
    torch.manual_seed(n)
    x = ops.Pair.from_float(torch.randn(m, k, device="cuda"))
    w = ops.Pair.from_float(torch.randn(n, k, device="cuda") * 0.05)
    res = ops.Pair.from_float(torch.randn(m, n, device="cuda"))
    buf = ops.Pair.zeros((m + 1, n), "cuda")                      # one guard row behind the output
    out = ops.Pair(buf.hi[:m], buf.lo[:m])
    rows_total = m
    ops.conv_gemm(x, (k, rows_total, 1, 1), (k, k * m, k * m), w, n, (rows_total, 1, 1), (n, n * m, n * m), out=out,
                  res=res, r_strides=(n, n * m, n * m), relu=True)
    ref = torch.relu(x.float().double() @ w.float().double().t() + res.float().double())
    assert relerr(out.float(), ref) <= 2e-5
    assert (buf.hi[m] == 0).all() and (buf.lo[m] == 0).all()


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("rows,ns,shots,batch", [(1900, 400, 3, 1), (300, 400, 1, 2), (777, 324, 2, 1), (130, 512, 2, 1),
                                                 (500, 260, 5, 1)])
def test_fused_softmax_wide_segments(ops, split, rows, ns, shots, batch):
    """Attention logits + softmax in one launch for segments of 257..512 keys (the reference's 20x20 supports: 400):
    single 512-column TMEM accumulator, three-pass epilogue.  Against softmax_j(alpha * q k^T) of the same rounded
    operands in fp64, per shot segment; pad columns zero."""
    torch.manual_seed(ns + rows)
    d = 256
    q = ops.Pair.from_float(torch.randn(batch * rows, d, device="cuda"), split)
    k = ops.Pair.from_float(torch.randn(batch * shots * ns, d, device="cuda") * 1.5, split)
    sp = (ns + 7) // 8 * 8
    pitch = (shots * sp + 7) // 8 * 8 + 8
    p = ops.Pair.zeros((batch * rows, pitch), "cuda", split)
    ops.linear(q, k, shots * ns, alpha=1.0 / 16.0, out=p, batch=batch, b_batch_stride=shots * ns * d, softmax_ns=ns,
               softmax_pitch=sp)
    qf, kf = q.float().double().view(batch, rows, d), k.float().double().view(batch, shots, ns, d)
    got = p.float().view(batch, rows, pitch)
    for s in range(shots):
        ref = torch.softmax(torch.einsum("brd,bnd->brn", qf, kf[:, s]) / 16.0, dim=2)
        seg = got[:, :, s * sp:s * sp + ns]
        assert relerr(seg, ref) <= (3e-5 if split else 8e-3)
        assert (got[:, :, s * sp + ns:(s + 1) * sp] == 0).all()
    assert (got[:, :, shots * sp:] == 0).all()


# ------------------------------------------------------------------------------------ sibling model FSOD (8f rank 4)
@pytest.mark.parametrize("b,h,w,c,split", [(2, 20, 31, 64, True), (1, 38, 63, 1024, True), (3, 7, 9, 8, False)])
def test_depthwise_xcorr_vs_torch(ops, b, h, w, c, split):
    """F.conv2d(feat, kernel, groups=C) with a per-image 7x7 kernel (fsod.py:106-112), NHWC, against torch fp64."""
    torch.manual_seed(h * w)
    x = torch.randn(b, h, w, c, device="cuda")
    k = torch.randn(b, 7, 7, c, device="cuda") * 0.1
    xp = ops.Pair.from_float(x, split)
    out, pair = ops.depthwise_xcorr(xp, k, want_f32=True, want_pair=True, split=split)
    xr = xp.float().double().permute(0, 3, 1, 2)
    ref = torch.stack([torch.nn.functional.conv2d(xr[i:i + 1], k[i].double().permute(2, 0, 1).unsqueeze(1), groups=c)[0]
                       for i in range(b)]).permute(0, 2, 3, 1)
    assert tuple(out.shape) == (b, h - 6, w - 6, c)
    assert relerr(out, ref) <= 2e-6
    assert relerr(pair.float(), ref) <= (2e-5 if split else 6e-3)
    m = ops.group_mean(torch.arange(24, dtype=torch.float32, device="cuda").view(6, 2, 2), 3)
    assert torch.equal(m.cpu(), torch.arange(24, dtype=torch.float32).view(2, 3, 2, 2).mean(1))
