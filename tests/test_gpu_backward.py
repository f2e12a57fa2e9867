"""Backward pass building blocks: the autograd bindings of the tensor-core GEMM
(dana_b200/autograd_ops.py: data- and weight-gradients on dana_conv_gemm, operands laid out by dana_grad_prepare /
dana_im2col_t) and the fused SGD update, against torch autograd in float64 on the same inputs.

The reference has no backward code of its own for these layers (it differentiates nn.Conv2d / nn.Linear / torch.bmm
through torch.autograd, train.py:138), so the float64 autograd of the same torch functionals IS the reference here.
Tolerance 1e-3 normwise (north_star), measured ~1e-5 with split-bf16 operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("cin,cout,ksize,stride,h,w,n", [
    (64, 128, 1, 1, 19, 23, 2),       # 1x1
    (128, 64, 3, 1, 17, 21, 2),       # 3x3, pad 1
    (256, 128, 1, 2, 19, 23, 3),      # strided 1x1 (first conv / downsample of a stage), odd extent
    (512, 72, 1, 1, 9, 13, 1),        # RPN heads: 24 + 48 outputs in one GEMM, not a multiple of 8 / 64
    (64, 64, 3, 1, 40, 40, 5),        # pixel count over several 64-pixel tiles and stream-K'd weight gradient
])
@pytest.mark.parametrize("relu,res,scale,bias", [(True, True, True, True), (False, False, False, True),
                                                  (True, False, True, False)])
def test_conv_autograd_vs_float64(cin, cout, ksize, stride, h, w, n, relu, res, scale, bias):
    import dana_b200  # noqa: F401
    from dana_b200 import autograd_ops as A
    g = torch.Generator(device="cpu").manual_seed(cin * 7 + cout + ksize + stride)
    dev = "cuda"
    x = torch.randn(n, h, w, cin, generator=g).to(dev).requires_grad_(True)
    wt = (torch.randn(cout, cin, ksize, ksize, generator=g) / (cin * ksize * ksize) ** 0.5).to(dev).requires_grad_(True)
    sc = (0.5 + torch.rand(cout, generator=g)).to(dev) if scale else None
    bs = torch.randn(cout, generator=g).to(dev).requires_grad_(True) if bias else None
    oh, ow = (h, w) if ksize == 3 else ((h - 1) // stride + 1, (w - 1) // stride + 1)
    rs = torch.randn(n, oh, ow, cout, generator=g).to(dev).requires_grad_(True) if res else None
    gy = torch.randn(n, oh, ow, cout, generator=g).to(dev)

    A.begin_step()
    y = A.conv(x, wt, bias=bs, res=rs, scale=sc, relu=relu, ksize=ksize, stride=stride)
    y.backward(gy)
    got = dict(y=y.detach(), dx=x.grad, dw=wt.grad, db=None if bs is None else bs.grad, dr=None if rs is None else rs.grad)

    xd = x.detach().double().permute(0, 3, 1, 2).requires_grad_(True)
    wd = wt.detach().double().requires_grad_(True)
    bd = None if bs is None else bs.detach().double().requires_grad_(True)
    rd = None if rs is None else rs.detach().double().permute(0, 3, 1, 2).requires_grad_(True)
    weff = wd if sc is None else wd * sc.double().view(-1, 1, 1, 1)
    yr = F.conv2d(xd, weff, bd, stride=stride, padding=1 if ksize == 3 else 0)
    if rd is not None:
        yr = yr + rd
    y_pre = yr
    if relu:
        # the mask is taken from the product's own output: a pre-activation within rounding of zero may fall on either
        # side, and a flipped mask element is a discontinuity of the function, not an error of the gradient kernels
        yr = yr * (y.detach().permute(0, 3, 1, 2) > 0).double()
    yr.backward(gy.double().permute(0, 3, 1, 2))
    ref = dict(y=(F.relu(y_pre) if relu else y_pre).detach().permute(0, 2, 3, 1), dx=xd.grad.permute(0, 2, 3, 1), dw=wd.grad,
               db=None if bd is None else bd.grad, dr=None if rd is None else rd.grad.permute(0, 2, 3, 1))
    for k in ref:
        if ref[k] is None:
            assert got[k] is None
            continue
        assert got[k].shape == ref[k].shape, k
        assert _rel(got[k], ref[k]) < 2e-4, (k, _rel(got[k], ref[k]))
    from dana_b200 import ops
    assert ops.device_error() == 0


def test_linear_autograd_vs_float64():
    import dana_b200  # noqa: F401
    from dana_b200 import autograd_ops as A
    torch.manual_seed(5)
    x = torch.randn(3, 49, 1024, device="cuda", requires_grad=True)
    for n_out in (256, 1):                  # the projections and the C -> 1 unary / channel-attention layers
        w = (torch.randn(n_out, 1024, device="cuda") / 32).requires_grad_(True)
        b = torch.randn(n_out, device="cuda", requires_grad=True)
        A.begin_step()
        x.grad = None
        y = A.linear(x, w, b)
        gy = torch.randn_like(y)
        y.backward(gy)
        xd, wd, bd = [t.detach().double().requires_grad_(True) for t in (x, w, b)]
        yr = F.linear(xd, wd, bd)
        yr.backward(gy.double())
        assert _rel(y, yr) < 2e-4
        assert _rel(x.grad, xd.grad) < 2e-4
        assert _rel(w.grad, wd.grad) < 2e-4
        assert _rel(b.grad, bd.grad) < 2e-4


@pytest.mark.parametrize("b,m,n,k", [(2, 300, 196, 256), (3, 2394, 400, 256), (2, 49, 49, 256), (2, 130, 1024, 400),
                                     (1, 77, 50, 36)])
def test_bmm_nt_autograd_vs_float64(b, m, n, k):
    import dana_b200  # noqa: F401
    from dana_b200 import autograd_ops as A
    torch.manual_seed(b * 100 + m)
    a = torch.randn(b, m, k, device="cuda", requires_grad=True)
    bb = torch.randn(b, n, k, device="cuda", requires_grad=True)
    c = A.bmm_nt(a, bb)
    gc = torch.randn_like(c)
    c.backward(gc)
    ad, bd = a.detach().double().requires_grad_(True), bb.detach().double().requires_grad_(True)
    cr = torch.bmm(ad, bd.transpose(1, 2))
    cr.backward(gc.double())
    assert _rel(c, cr) < 2e-4
    assert _rel(a.grad, ad.grad) < 2e-4
    assert _rel(bb.grad, bd.grad) < 2e-4


def test_sgd_momentum_matches_torch_optim():
    import dana_b200  # noqa: F401
    from dana_b200 import ops
    torch.manual_seed(3)
    n = 100003
    p0 = torch.randn(n, device="cuda")
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([ref], lr=0.01, momentum=0.9, weight_decay=1e-4)
    p, m = p0.clone(), torch.zeros(n, device="cuda")
    for step in range(3):
        g = torch.randn(n, device="cuda")
        ref.grad = g.clone()
        opt.step()
        ops.sgd_momentum(p, g, m, 0.01, 0.9, 1e-4)
    assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-6)
