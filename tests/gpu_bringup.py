"""GPU bring-up diagnostics (not a pytest file): runs each kernel family against a torch fp64/fp32
reference and prints max errors.  Usage: python tests/gpu_bringup.py [group ...]
Each group should be run in its own process (a trapped kernel poisons the CUDA context)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dana_b200  # noqa: E402,F401
from dana_b200 import ops  # noqa: E402
from dana_b200.ops import Pair  # noqa: E402

DEV = "cuda"


def relerr(a, b):
    a = a.double()
    b = b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def report(name, err, tol):
    print("%-58s err=%.3e tol=%.1e %s" % (name, err, tol, "OK" if err <= tol else "FAIL"), flush=True)
    return err <= tol


def g_gemm():
    torch.manual_seed(0)
    ok = True
    for split in (False, True):
        for (m, k, n) in [(128, 64, 64), (128, 128, 128), (256, 256, 256), (300, 192, 72), (2394, 1024, 256),
                          (1000, 1200, 1024), (4096, 512, 2048), (77, 3136, 1024), (1200, 2048, 4)]:
            x = torch.randn(m, k, device=DEV)
            w = torch.randn(n, k, device=DEV) * 0.05
            bias = torch.randn(n, device=DEV)
            xp, wp = Pair.from_float(x, split), Pair.from_float(w, split)
            out = torch.empty(m, n, device=DEV)
            ops.linear(xp, wp, n, bias=bias, out_f32=out)
            torch.cuda.synchronize()
            ref = xp.float().double() @ wp.float().double().t() + bias.double()
            ref32 = x.double() @ w.double().t() + bias.double()
            e_op = relerr(out, ref)      # vs the same rounded operands: tests the kernel
            e_fp = relerr(out, ref32)    # vs fp32 operands: tests the precision mode
            ok &= report("linear split=%d m=%d k=%d n=%d (operand-exact)" % (split, m, k, n), e_op,
                         2e-5 if split else 2e-5)
            print("    vs unrounded fp32 operands: %.3e" % e_fp)
    return ok


def g_conv():
    torch.manual_seed(1)
    ok = True
    F = torch.nn.functional
    for split in (True, False):
        for (n, h, w, cin, cout, ks, stride) in [(2, 38, 63, 256, 256, 3, 1), (3, 20, 20, 128, 128, 3, 1),
                                                 (16, 4, 4, 512, 512, 3, 1), (2, 75, 125, 256, 128, 1, 2),
                                                 (8, 7, 7, 1024, 512, 1, 2), (2, 150, 250, 64, 256, 1, 1),
                                                 (1, 38, 63, 2048, 512, 3, 1)]:
            x = torch.randn(n, h, w, cin, device=DEV)
            wt = torch.randn(cout, cin, ks, ks, device=DEV) * (1.0 / (cin * ks * ks) ** 0.5)
            scale = torch.rand(cout, device=DEV) + 0.5
            bias = torch.randn(cout, device=DEV) * 0.1
            xp = Pair.from_float(x, split)
            wp = Pair.from_float(wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous(), split)
            oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1
            res = torch.randn(n, oh, ow, cout, device=DEV)
            rp = Pair.from_float(res, split)
            out = ops.conv_nhwc(xp, wp, cout, ksize=ks, stride=stride, scale=scale, bias=bias, res=rp, relu=True,
                                split=split)
            torch.cuda.synchronize()
            xr = xp.float().double().permute(0, 3, 1, 2)
            wr = wp.float().double().reshape(cout, ks, ks, cin).permute(0, 3, 1, 2)
            ref = F.conv2d(xr, wr, stride=stride, padding=ks // 2)
            ref = ref * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1)
            ref = torch.relu(ref + rp.float().double().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
            ok &= report("conv split=%d n=%d %dx%d cin=%d cout=%d k=%d s=%d" % (split, n, h, w, cin, cout, ks, stride),
                         relerr(out.float(), ref), (2e-5 if cin * ks * ks < 8192 else 1e-4) if split else 6e-3)
    return ok


def g_gemm_batched():
    torch.manual_seed(2)
    ok = True
    b, rows, k, n = 3, 500, 1200, 1024
    x = torch.rand(b * rows, k, device=DEV)
    w = torch.randn(b, n, k, device=DEV) * 0.05
    bias = torch.randn(b, n, device=DEV)
    xp = Pair.from_float(x, True)
    wp = Pair.from_float(w, True)
    out = torch.empty(b * rows, n, device=DEV)
    ops.linear(xp, Pair(wp.hi.view(b * n, k), wp.lo.view(b * n, k)), n, bias=bias, out_f32=out, batch=b,
               b_batch_stride=n * k, bias_sn=n, alpha=0.5)
    torch.cuda.synchronize()
    ref = 0.5 * torch.bmm(xp.float().double().view(b, rows, k), wp.float().double().transpose(1, 2)) \
        + bias.double().view(b, 1, n)
    ok &= report("batched linear + per-batch bias", relerr(out.view(b, rows, n), ref), 2e-5)
    return ok


def nms_ref(boxes, scores, thr):
    import numpy as np
    b = boxes.cpu().numpy().astype(np.float32)
    s = scores.cpu().numpy()
    order = np.argsort(-s, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1 + np.float32(1)) * (y2 - y1 + np.float32(1))
    sup = np.zeros(len(b), dtype=bool)
    for _i in range(len(b)):
        i = order[_i]
        if sup[i]:
            continue
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest])
        yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest])
        yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1 + np.float32(1))
        h = np.maximum(np.float32(0), yy2 - yy1 + np.float32(1))
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        sup[rest[ovr >= np.float32(thr)]] = True
    return np.nonzero(~sup)[0]


def g_nms():
    import numpy as np
    torch.manual_seed(3)
    ok = True
    for n in (1, 63, 64, 65, 300, 2000, 6000, 12000):
        for thr in (0.3, 0.7):
            x1 = torch.rand(n) * 900
            y1 = torch.rand(n) * 500
            w = torch.rand(n) * 199 + 1
            h = torch.rand(n) * 199 + 1
            boxes = torch.stack([x1, y1, x1 + w, y1 + h], 1)
            if n >= 300:  # clustered variant
                boxes[n // 2:] = boxes[: n - n // 2] + torch.randn(n - n // 2, 4) * 3
            scores = torch.rand(n)
            t0 = time.time()
            keep = ops.nms(boxes.to(DEV), scores.to(DEV), thr)
            torch.cuda.synchronize()
            dt = time.time() - t0
            ref = nms_ref(boxes, scores, thr)
            same = keep.cpu().numpy().tolist() == ref.tolist()
            print("nms n=%d thr=%.1f kept=%d ref=%d %s (%.1f ms)" % (n, thr, keep.numel(), len(ref),
                                                                      "OK" if same else "FAIL", dt * 1e3), flush=True)
            ok &= same
    return ok


def g_roialign():
    import torchvision
    torch.manual_seed(4)
    ok = True
    b, c, h, w = 2, 256, 38, 63
    feat = torch.randn(b, c, h, w, device=DEV)
    r = 200
    x1 = torch.rand(r) * 900
    y1 = torch.rand(r) * 550
    bw = torch.rand(r) * 400 + 0.5
    bh = torch.rand(r) * 300 + 0.5
    rois = torch.stack([torch.randint(0, b, (r,)).float(), x1, y1, x1 + bw, y1 + bh], 1)
    rois[0, 1:] = torch.tensor([0., 0., 1007., 607.])    # full image, past the border
    rois[1, 1:] = torch.tensor([100., 100., 90., 95.])   # degenerate
    rois[2, 1:] = torch.tensor([-50., -40., 30., 20.])   # negative start
    rois[3, 1:] = torch.tensor([990., 590., 1200., 700.])
    rois = rois.to(DEV)
    out = ops.roi_align_forward(feat, rois, 1.0 / 16, 7, 7, 0)
    ref = torchvision.ops.roi_align(feat.double(), rois.double(), (7, 7), 1.0 / 16, 0, False)
    ok &= report("roi_align NCHW vs torchvision", relerr(out, ref), 1e-5)
    out2, pair = ops.roi_align_nhwc(feat.permute(0, 2, 3, 1).contiguous(), rois, 1.0 / 16, 7, 0)
    ok &= report("roi_align NHWC fp32", relerr(out2.permute(0, 3, 1, 2), ref), 1e-5)
    ok &= report("roi_align NHWC pair", relerr(pair.float().permute(0, 3, 1, 2), ref), 1e-5)
    # backward vs autograd of torchvision
    f2 = feat.double().clone().requires_grad_(True)
    o = torchvision.ops.roi_align(f2, rois.double(), (7, 7), 1.0 / 16, 0, False)
    go = torch.randn_like(o)
    o.backward(go)
    gin = ops.roi_align_backward(go.float(), rois, 1.0 / 16, 7, 7, b, c, h, w, 0)
    ok &= report("roi_align backward", relerr(gin, f2.grad), 1e-4)
    return ok


def g_stem():
    torch.manual_seed(5)
    F = torch.nn.functional
    ok = True
    for (b, h, w) in [(1, 64, 96), (2, 320, 320), (1, 600, 1000), (1, 97, 131)]:
        im = torch.randn(b, 3, h, w, device=DEV) * 50
        wt = torch.randn(64, 3, 7, 7, device=DEV) * 0.05
        scale = torch.rand(64, device=DEV) + 0.5
        bias = torch.randn(64, device=DEV)
        out = ops.stem(im, ops.pack_stem_weight(wt), scale, bias)
        ref = F.conv2d(im.double(), wt.double(), stride=2, padding=3)
        ref = torch.relu(ref * scale.double().view(1, -1, 1, 1) + bias.double().view(1, -1, 1, 1))
        ref = F.max_pool2d(ref, 3, 2, 0, ceil_mode=True).permute(0, 2, 3, 1)
        if tuple(out.hi.shape) != tuple(ref.shape):
            print("stem shape mismatch", out.hi.shape, ref.shape)
            ok = False
            continue
        ok &= report("stem %dx%dx%d" % (b, h, w), relerr(out.float(), ref), 2e-5)
    return ok


def g_misc():
    torch.manual_seed(6)
    ok = True
    # support prepare
    maps, shots, ns, c = 6, 3, 400, 1024
    s = torch.relu(torch.randn(maps, ns, c, device=DEV))
    pe = torch.randn(ns, c, device=DEV)
    ba_w = torch.randn(c, device=DEV) * 0.05
    ba_b = torch.randn(1, device=DEV)
    un_w = torch.randn(c, device=DEV) * 0.05
    un_b = torch.randn(1, device=DEV)
    pitch = 1216
    vc, vt, rbar = ops.support_prepare(Pair.from_float(s), pe, shots, ba_w=ba_w, ba_b=ba_b, gamma=0.1, un_w=un_w,
                                       un_b=un_b, unary_gamma=0.1, vt_pitch=pitch)
    sd = Pair.from_float(s).float().double()
    v = sd + pe.double()
    wsp = torch.softmax(v @ ba_w.double() + ba_b.double(), 1)
    g = torch.einsum("mn,mnc->mc", wsp, v)
    v = v + 0.1 * torch.nn.functional.leaky_relu(g).unsqueeze(1)
    u = torch.softmax(v @ un_w.double() + un_b.double(), 1)
    r = torch.einsum("mn,mnc->mc", u, v)
    ok &= report("support vc", relerr(vc.float().view(maps, ns, c), v - v.mean(1, keepdim=True)), 1e-5)
    vt_ref = v.view(maps // shots, shots, ns, c).permute(0, 3, 1, 2).reshape(maps // shots, c, shots * ns)
    ok &= report("support vt", relerr(vt.float()[:, :, : shots * ns], vt_ref), 1e-5)
    ok &= report("support vt pad zero", vt.float()[:, :, shots * ns:].abs().max().item(), 0.0)
    ok &= report("support rbar", relerr(rbar, 0.1 * r.view(maps // shots, shots, c).mean(1)), 1e-5)
    # center rows (small + large groups)
    for (gr, rows) in [(50, 49), (3, 2394)]:
        x = torch.randn(gr * rows, 256, device=DEV) + 3
        o = ops.center_rows(x, gr, rows)
        xr = x.double().view(gr, rows, 256)
        ok &= report("center_rows %dx%d" % (gr, rows), relerr(o.float().view(gr, rows, 256), xr - xr.mean(1, keepdim=True)),
                     1e-5)
    # attention softmax
    lg = torch.randn(777, 1216, device=DEV) * 3
    p = ops.attn_softmax(lg, 3, 400)
    ref = torch.softmax(lg[:, :1200].double().view(777, 3, 400), 2).reshape(777, 1200)
    ok &= report("attn_softmax", relerr(p.float()[:, :1200], ref), 1e-5)
    ok &= report("attn_softmax pad", p.float()[:, 1200:].abs().max().item(), 0.0)
    # rpn fg prob
    x = torch.randn(2, 5, 7, 72, device=DEV)
    fg, d = ops.rpn_fg_prob(x, 12)
    sc = x[..., :24].double()
    ref = torch.softmax(torch.stack([sc[..., :12], sc[..., 12:]], 0), 0)[1].reshape(2, -1)
    ok &= report("rpn_fg_prob", relerr(fg, ref), 1e-6)
    ok &= report("rpn deltas", relerr(d, x[..., 24:].reshape(2, -1, 4)), 0.0)
    # avgpool / spatial mean / softmax2 / converts
    a = torch.randn(4, 20, 20, 128, device=DEV)
    o = ops.avgpool(Pair.from_float(a), 14)
    ref = torch.nn.functional.avg_pool2d(Pair.from_float(a).float().double().permute(0, 3, 1, 2), 14, 1).permute(0, 2, 3, 1)
    ok &= report("avgpool", relerr(o, ref), 1e-5)
    a = torch.randn(30, 16, 256, device=DEV)
    o, op = ops.spatial_mean(Pair.from_float(a))
    ok &= report("spatial_mean", relerr(o, Pair.from_float(a).float().double().mean(1)), 1e-5)
    a = torch.randn(100, 2, device=DEV)
    ok &= report("softmax2", relerr(ops.softmax2(a), torch.softmax(a.double(), 1)), 1e-6)
    a = torch.randn(2, 5, 6, 64, device=DEV)
    ok &= report("nhwc->nchw", relerr(ops.nhwc_pair_to_nchw(Pair.from_float(a)), a.permute(0, 3, 1, 2)), 1e-5)
    return ok


def g_proposals():
    import numpy as np
    torch.manual_seed(7)
    ok = True
    for (b, fh, fw, pre, post) in [(2, 10, 14, 6000, 300), (4, 38, 63, 6000, 300), (1, 38, 50, 12000, 2000)]:
        A = 12
        hwa = fh * fw * A
        fg = torch.rand(b, hwa)
        deltas = torch.randn(b, hwa, 4) * 0.3
        base = torch.tensor([[-38, -16, 53, 31], [-84, -40, 99, 55], [-176, -88, 191, 103], [-360, -184, 375, 199],
                             [-24, -24, 39, 39], [-56, -56, 71, 71], [-120, -120, 135, 135], [-248, -248, 263, 263],
                             [-14, -36, 29, 51], [-36, -80, 51, 95], [-80, -168, 95, 183], [-168, -344, 183, 359]],
                            dtype=torch.float32)
        im_info = torch.tensor([[fh * 16.0 - 8, fw * 16.0 - 8, 1.0]] * b)
        t0 = time.time()
        rois, sc, cnt = ops.proposals(fg.to(DEV), deltas.to(DEV), base.to(DEV), im_info.to(DEV), fh, fw, 16, pre, post,
                                      0.7, want_scores=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        # reference on CPU (torch, following proposal_layer.py)
        sx = torch.arange(fw) * 16.0
        sy = torch.arange(fh) * 16.0
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack([xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)], 1)
        anchors = (base.view(1, A, 4) + shifts.view(-1, 1, 4)).view(1, hwa, 4).expand(b, hwa, 4)
        wid = anchors[..., 2] - anchors[..., 0] + 1.0
        hei = anchors[..., 3] - anchors[..., 1] + 1.0
        cx = anchors[..., 0] + 0.5 * wid
        cy = anchors[..., 1] + 0.5 * hei
        pcx = deltas[..., 0] * wid + cx
        pcy = deltas[..., 1] * hei + cy
        pw = torch.exp(deltas[..., 2]) * wid
        ph = torch.exp(deltas[..., 3]) * hei
        props = torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], 2)
        allok = True
        for i in range(b):
            props[i, :, 0::2].clamp_(0, im_info[i, 1] - 1)
            props[i, :, 1::2].clamp_(0, im_info[i, 0] - 1)
            order = torch.sort(fg[i], descending=True, stable=True)[1]
            if pre > 0 and pre < fg.numel():
                order = order[:pre]
            pi = props[i][order]
            keep = nms_ref(pi, torch.arange(len(order), 0, -1).float(), 0.7)[:post]
            ref = torch.zeros(post, 5)
            ref[:, 0] = i
            ref[: len(keep), 1:] = pi[keep]
            got = rois[i].cpu()
            close = torch.allclose(got, ref, atol=1e-3, rtol=1e-5)
            nk = int(cnt[i].item())
            if not close or nk != len(keep):
                # count rows that match exactly anyway
                match = (got - ref).abs().max(1)[0] < 1e-3
                print("  image %d: kept %d ref %d, rows matching %d/%d" % (i, nk, len(keep), int(match.sum()), post))
                allok = False
        print("proposals b=%d %dx%d pre=%d post=%d %s (%.1f ms)" % (b, fh, fw, pre, post, "OK" if allok else "FAIL",
                                                                   dt * 1e3), flush=True)
        ok &= allok
    return ok


def g_engine():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import dana_oracle as O
    from dana_b200.engine import DanaEngine
    ok = True
    k, sets = 2, 2
    p = O.make_params(1996, attn_std=0.05)
    im, info, sup = O.synth_inputs(7, 1, 128, 192, k * sets)
    with torch.no_grad():
        ref = O.dana_forward_eval(p, im, info, sup, k)
    want = ("base_feat", "support_feat", "dense", "pooled", "fc7", "cls_score", "support_pooled", "rpn_fg")
    for prec in ("bf16x3", "bf16"):
        eng = DanaEngine(p, n_shot=k, precision=prec)
        t0 = time.time()
        rois, cls_prob, bbox, ex = eng.forward(im.to(DEV), info.to(DEV), sup.to(DEV), want=want)
        torch.cuda.synchronize()
        print("engine %s forward %.1f ms" % (prec, (time.time() - t0) * 1e3))
        tol = 1e-3 if prec == "bf16x3" else 5e-2
        ok &= report(prec + " base_feat", relerr(ex["base_feat"].cpu(), ref["base_feat"]), tol)
        ok &= report(prec + " support_feat", relerr(ex["support_feat"].cpu(), ref["support_feat"].reshape(-1, 1024, 20, 20)), tol)
        ok &= report(prec + " support_pooled", relerr(ex["support_pooled"].cpu(), ref["support_pooled"].reshape(-1, 1024, 7, 7)), tol)
        ok &= report(prec + " dense", relerr(ex["dense"].cpu(), ref["dense"]), tol)
        fg_ref = ref["rpn_cls_prob"][:, 12:].permute(0, 2, 3, 1).reshape(1, -1)
        ok &= report(prec + " rpn_fg", relerr(ex["rpn_fg"].cpu(), fg_ref), tol)
        dl_ref = ref["rpn_bbox_pred"].permute(0, 2, 3, 1).reshape(1, -1, 4)
        ok &= report(prec + " rpn_deltas", relerr(ex["rpn_deltas"].cpu(), dl_ref), tol)
        same_rois = (rois.cpu() - ref["rois"]).abs().max().item()
        print("    rois max abs diff (free running) %.3e" % same_rois)
        # teacher-forced rois so that the per-RoI stages are comparable row by row
        rois, cls_prob, bbox, ex = eng.forward(im.to(DEV), info.to(DEV), sup.to(DEV), want=want,
                                               teacher={"rois": ref["rois"].to(DEV)})
        ok &= report(prec + " pooled (teacher rois)", relerr(ex["pooled"].cpu(), ref["pooled"]), tol)
        ok &= report(prec + " fc7", relerr(ex["fc7"].cpu(), ref["fc7"]), tol)
        ok &= report(prec + " bbox_pred", relerr(bbox.cpu(), ref["bbox_pred"]), tol)
        ok &= report(prec + " cls_score", relerr(ex["cls_score"].cpu(), ref["cls_score"]), tol)
        ok &= report(prec + " cls_prob", relerr(cls_prob.cpu(), ref["cls_prob"]), tol)
    return ok


def g_bench():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import dana_oracle as O
    from dana_b200.engine import DanaEngine
    p = O.make_params(1996)
    im, info, sup = O.synth_inputs(3, 4, 600, 1000, 6)
    im, info, sup = im.to(DEV), info.to(DEV), sup.to(DEV)
    for prec in ("bf16x3", "bf16"):
        eng = DanaEngine(p, n_shot=3, precision=prec)
        for _ in range(2):
            eng.forward(im, info, sup)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            eng.forward(im, info, sup)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("bench %s: %.2f ms/step (bs 4) -> %.1f img/s ; peak mem %.1f GB" %
              (prec, ms, 4e3 / ms, torch.cuda.max_memory_allocated() / 2**30), flush=True)
    return True


GROUPS = {"gemm": g_gemm, "conv": g_conv, "batched": g_gemm_batched, "nms": g_nms, "roialign": g_roialign,
          "stem": g_stem, "misc": g_misc, "proposals": g_proposals, "engine": g_engine, "bench": g_bench}

if __name__ == "__main__":
    names = sys.argv[1:] or list(GROUPS)
    allok = True
    for nme in names:
        print("==== %s" % nme, flush=True)
        try:
            r = GROUPS[nme]()
            err = ops.device_error()
            if err:
                print("DEVICE ERROR CODE %d" % err)
                r = False
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            r = False
        print("==== %s: %s" % (nme, "PASS" if r else "FAIL"), flush=True)
        allok &= bool(r)
    sys.exit(0 if allok else 1)
