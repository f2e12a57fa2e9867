"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + torch-CPU functional ops, no nn.Module state) of the reference algorithm
for the DAnA forward hot path.  Nothing under the product package imports this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker
or as the timed CPU baseline, never as a shipped compute path.

Every function cites the reference lines it follows (paths relative to the reference repo root).
Pinning: tests/test_oracle_pins.py checks these functions against golden vectors produced by the
UNMODIFIED reference python (oracle/make_golden.py, run in the build container where
/root/reference is mounted) and against the reference's own compiled CPU operators (oracle/_ref).
The reference itself ships no tests or golden vectors, so those generated
vectors are the pin.

Third-party arithmetic: conv / linear / bmm / softmax / sort are torch (the reference pins
pytorch 1.2, env.yml:61; this container has 2.11) -- the oracle is "torch CPU fp32 executing the
reference's algorithm", which is what BASELINE.json calls the reference's own PyTorch/CPU path.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
C_SRC = os.path.join(HERE, "c", "dana_oracle.c")
C_LIB = os.path.join(HERE, "liboracle_c.so")

# --------------------------------------------------------------------------------------------
# C restatements (nms, roi_align) -- built with gcc, loaded with ctypes
# --------------------------------------------------------------------------------------------
_clib = None


def build_c(force=False):
    if not force and os.path.exists(C_LIB) and os.path.getmtime(C_LIB) >= os.path.getmtime(C_SRC):
        return C_LIB
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", C_LIB, C_SRC, "-lm"])
    return C_LIB


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(build_c())
        lib.oracle_nms.restype = ctypes.c_int64
        lib.oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p]
        lib.oracle_roi_align_fwd.restype = None
        lib.oracle_roi_align_fwd.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_int, ctypes.c_void_p]
        _clib = lib
    return _clib


def nms(boxes, scores, thresh):
    """lib/model/csrc/cpu/nms_cpu.cpp:6-67.  boxes [N,4], scores [N] (torch or numpy) -> int64 kept
    input indices ascending.  Ties in score are broken by lower input index."""
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float32))
    s = np.ascontiguousarray(np.asarray(scores, dtype=np.float32))
    n = b.shape[0]
    keep = np.empty((max(n, 1),), dtype=np.int64)
    m = _c().oracle_nms(b.ctypes.data, s.ctypes.data, n, float(thresh), keep.ctypes.data)
    return torch.from_numpy(keep[:m].copy())


def roi_align_forward(inp, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio):
    """lib/model/csrc/cpu/ROIAlign_cpu.cpp:18-257.  inp [B,C,H,W] fp32, rois [R,5] -> [R,C,ph,pw]."""
    x = np.ascontiguousarray(np.asarray(inp, dtype=np.float32))
    r = np.ascontiguousarray(np.asarray(rois, dtype=np.float32))
    _, c, h, w = x.shape
    out = np.zeros((r.shape[0], c, pooled_h, pooled_w), dtype=np.float32)
    if r.shape[0]:
        _c().oracle_roi_align_fwd(x.ctypes.data, r.ctypes.data, r.shape[0], c, h, w, pooled_h, pooled_w,
                                  float(spatial_scale), int(sampling_ratio), out.ctypes.data)
    return torch.from_numpy(out)


# --------------------------------------------------------------------------------------------
# anchors / box decoding / proposal layer
# --------------------------------------------------------------------------------------------
def _anchor_whc(a):
    w = a[2] - a[0] + 1
    h = a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def _anchors_from(ws, hs, xc, yc):
    ws = np.asarray(ws, dtype=np.float64).reshape(-1, 1)
    hs = np.asarray(hs, dtype=np.float64).reshape(-1, 1)
    return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    """lib/model/rpn/generate_anchors.py:45-105: ratio-major, then scale; float64 [A,4]."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    base = np.array([0, 0, base_size - 1, base_size - 1], dtype=np.float64)
    w, h, xc, yc = _anchor_whc(base)
    size_ratios = (w * h) / ratios
    ws = np.round(np.sqrt(size_ratios))
    hs = np.round(ws * ratios)
    ratio_anchors = _anchors_from(ws, hs, xc, yc)
    rows = []
    for ra in ratio_anchors:
        w, h, xc, yc = _anchor_whc(ra)
        rows.append(_anchors_from(w * scales, h * scales, xc, yc))
    return np.vstack(rows)


def anchor_grid(base_anchors, feat_h, feat_w, feat_stride):
    """proposal_layer.py:79-93: anchors[(y*W + x)*A + a] = base[a] + (x, y, x, y)*stride, fp32 [H*W*A, 4]."""
    base = torch.from_numpy(np.asarray(base_anchors)).float()
    sx = np.arange(0, feat_w) * feat_stride
    sy = np.arange(0, feat_h) * feat_stride
    gx, gy = np.meshgrid(sx, sy)
    shifts = torch.from_numpy(np.vstack((gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel())).transpose()).float()
    return (base.view(1, -1, 4) + shifts.view(-1, 1, 4)).reshape(-1, 4)


def bbox_transform_inv(boxes, deltas):
    """lib/model/rpn/bbox_transform.py:77-103 for one delta group per box. boxes/deltas [B,N,4]."""
    widths = boxes[:, :, 2] - boxes[:, :, 0] + 1.0
    heights = boxes[:, :, 3] - boxes[:, :, 1] + 1.0
    ctr_x = boxes[:, :, 0] + 0.5 * widths
    ctr_y = boxes[:, :, 1] + 0.5 * heights
    pcx = deltas[:, :, 0] * widths + ctr_x
    pcy = deltas[:, :, 1] * heights + ctr_y
    pw = torch.exp(deltas[:, :, 2]) * widths
    ph = torch.exp(deltas[:, :, 3]) * heights
    return torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), 2)


def clip_boxes(boxes, im_info):
    """lib/model/rpn/bbox_transform.py:125-133: x to [0, W-1], y to [0, H-1]; im_info rows (H, W, scale)."""
    out = boxes.clone()
    for i in range(out.shape[0]):
        out[i, :, 0].clamp_(0, float(im_info[i, 1]) - 1)
        out[i, :, 1].clamp_(0, float(im_info[i, 0]) - 1)
        out[i, :, 2].clamp_(0, float(im_info[i, 1]) - 1)
        out[i, :, 3].clamp_(0, float(im_info[i, 0]) - 1)
    return out


def proposal_layer(rpn_cls_prob, rpn_bbox_pred, im_info, base_anchors, feat_stride, pre_nms_top_n, post_nms_top_n,
                   nms_thresh, nms_fn=None, return_scores=False):
    """lib/model/rpn/proposal_layer.py:49-190.  rpn_cls_prob [B,2A,H,W] (first A = bg), rpn_bbox_pred
    [B,4A,H,W].  Returns rois [B, post, 5] zero padded.  The RPN_MIN_SIZE filter is disabled in the
    reference (:123) and stays disabled; the pre-NMS truncation compares against numel of the whole
    batch (:148).  The descending sort is made stable (ties -> lower index), see nms()."""
    nms_fn = nms_fn or nms
    num_a = base_anchors.shape[0]
    b, _, fh, fw = rpn_bbox_pred.shape
    scores = rpn_cls_prob[:, num_a:, :, :].permute(0, 2, 3, 1).reshape(b, -1)
    deltas = rpn_bbox_pred.permute(0, 2, 3, 1).reshape(b, -1, 4)
    anchors = anchor_grid(base_anchors, fh, fw, feat_stride).to(scores.dtype)
    anchors = anchors.view(1, -1, 4).expand(b, -1, 4)
    props = clip_boxes(bbox_transform_inv(anchors, deltas), im_info)
    out = scores.new_zeros(b, post_nms_top_n, 5)
    out_scores = scores.new_zeros(b, post_nms_top_n)
    for i in range(b):
        order = torch.sort(scores[i], descending=True, stable=True)[1]
        if 0 < pre_nms_top_n < scores.numel():
            order = order[:pre_nms_top_n]
        p_i = props[i][order]
        s_i = scores[i][order]
        keep = nms_fn(p_i, s_i, nms_thresh).long().view(-1)
        if post_nms_top_n > 0:
            keep = keep[:post_nms_top_n]
        out[i, :, 0] = i
        out[i, : keep.numel(), 1:] = p_i[keep]
        out_scores[i, : keep.numel()] = s_i[keep]
    return (out, out_scores) if return_scores else out


# --------------------------------------------------------------------------------------------
# attention blocks
# --------------------------------------------------------------------------------------------
def positional_encoding(max_len, d_model=1024, dtype=torch.float32):
    """lib/model/framework/dana.py:311-320: 1-D sinusoid over the flattened position index, [max_len, d]."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0., max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / float(d_model)))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def _lin(x, p, name):
    return F.linear(x, p[name + ".weight"].to(x.dtype), p[name + ".bias"].to(x.dtype))


def _attend(q_centered, supports, p, prefix, d, unary_gamma=0.1, enhance_prefix=None, channel_gamma=0.1):
    """Shared core of dana.py:126-150 (RPN level) and :268-281 (head).  q_centered [G, Nq, d];
    supports: list over shots of [G, Ns, C] (positional encoding already applied).  Returns the
    shot-mean attended feature [G, Nq, C]."""
    outs = []
    for s in supports:
        if enhance_prefix is not None:                                   # BA block, dana.py:133-137
            w = F.softmax(_lin(s, p, enhance_prefix), 1)
            g = torch.bmm(w.transpose(1, 2), s)
            s = s + channel_gamma * F.leaky_relu(g)
        k = _lin(s, p, prefix + "_adapt_k_layer")                        # :140 / :271
        k = k - k.mean(1, keepdim=True)                                  # :141 / :272
        att = torch.bmm(q_centered, k.transpose(1, 2)) / math.sqrt(d)    # :142 / :273
        att = F.softmax(att, dim=2)                                      # :143 / :274
        un = F.softmax(_lin(s, p, prefix + "_unary_layer"), dim=1)       # :144-145 / :275-276
        att = att + unary_gamma * un.transpose(1, 2)                     # :146 / :277
        outs.append(torch.bmm(att, s))                                   # :147 / :278
    return torch.stack(outs, 0).mean(0)                                  # :150 / :281


def ba_cisa_rpn(base_feat, support_feat, p, semantic_enhance=True, channel_gamma=0.1, unary_gamma=0.1,
                reduce_dim=256):
    """RPN-level BA + CISA, dana.py:117-151.  base_feat [B,C,h,w]; support_feat [B,K,C,hs,ws]
    (positive set).  The positional encoding uses max_len = hs*ws (400 in the reference, which
    hard-codes 20x20 supports; BASELINE.json's 14x14 case is the same formula with max_len 196).
    Returns the dense support feature [B,C,h,w]."""
    b, c, h, w = base_feat.shape
    k = support_feat.shape[1]
    ns = support_feat.shape[3] * support_feat.shape[4]
    pe = positional_encoding(ns, c, base_feat.dtype)
    query = base_feat.reshape(b, c, -1).transpose(1, 2)                    # [B, hw, C]   (:118)
    q = _lin(query, p, "rpn_adapt_q_layer")                               # :124
    q = q - q.mean(1, keepdim=True)                                       # :125
    shots = [support_feat[:, i].reshape(b, c, ns).transpose(1, 2) + pe.unsqueeze(0) for i in range(k)]  # :117,128
    dense = _attend(q, shots, p, "rpn", reduce_dim, unary_gamma,
                    "rpn_channel_k_layer" if semantic_enhance else None, channel_gamma)
    return dense.transpose(1, 2).reshape(b, c, h, w)                      # :151


def rcnn_head_attention(pooled, support_pooled, p, unary_gamma=0.1, reduce_dim=256):
    """Classification branch of rcnn_head, dana.py:247-290 ('concat').  pooled [R,C,7,7] (R = B*n_roi,
    image-major), support_pooled [B,K,C,7,7].  Returns (cls_score [R,2], cls_prob [R,2])."""
    r, c = pooled.shape[0], pooled.shape[1]
    b, k = support_pooled.shape[0], support_pooled.shape[1]
    per = r // b
    pe = positional_encoding(49, c, pooled.dtype).unsqueeze(0)
    q_mat = pooled.reshape(r, c, 49).transpose(1, 2) + pe                  # :257,259
    q = _lin(q_mat, p, "rcnn_adapt_q_layer")                              # :266
    q = q - q.mean(1, keepdim=True)                                       # :267
    shots = []
    for i in range(k):                                                    # :254-258 (repeat per RoI)
        s = support_pooled[:, i].reshape(b, c, 49).transpose(1, 2) + pe    # [B,49,C]
        shots.append(s.repeat_interleave(per, dim=0))                     # [R,49,C]
    dense = _attend(q, shots, p, "rcnn", reduce_dim, unary_gamma)
    corr = torch.cat([q_mat, dense], 2)                                   # :284
    corr = _lin(corr, p, "rcnn_transform_layer")                          # :288
    hid = F.relu(_lin(corr.reshape(r, -1), p, "output_score_layer.linear1"))   # FFN :302-305
    score = _lin(hid, p, "output_score_layer.linear2")
    return score, F.softmax(score, 1)                                     # :290


# --------------------------------------------------------------------------------------------
# ResNet trunk (Caffe-style bottleneck: stride on the first 1x1) with frozen BN
# --------------------------------------------------------------------------------------------
BN_EPS = 1e-5


def _bn(x, p, name):
    return F.batch_norm(x, p[name + ".running_mean"].to(x.dtype), p[name + ".running_var"].to(x.dtype),
                        p[name + ".weight"].to(x.dtype), p[name + ".bias"].to(x.dtype), False, 0.0, BN_EPS)


def _bottleneck(x, p, name, stride):
    """lib/model/framework/resnet.py:66-102."""
    out = F.relu(_bn(F.conv2d(x, p[name + ".conv1.weight"].to(x.dtype), stride=stride), p, name + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, p[name + ".conv2.weight"].to(x.dtype), padding=1), p, name + ".bn2"))
    out = _bn(F.conv2d(out, p[name + ".conv3.weight"].to(x.dtype)), p, name + ".bn3")
    if (name + ".downsample.0.weight") in p:
        x = _bn(F.conv2d(x, p[name + ".downsample.0.weight"].to(x.dtype), stride=stride), p, name + ".downsample.1")
    return F.relu(out + x)


RES_LAYERS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


def _stage(x, p, prefix, blocks, stride):
    for i in range(blocks):
        x = _bottleneck(x, p, "%s.%d" % (prefix, i), stride if i == 0 else 1)
    return x


def stem(x, p):
    """conv1 / bn1 / relu / maxpool(3, 2, pad 0, ceil_mode) -- resnet.py:109-113."""
    x = F.relu(_bn(F.conv2d(x, p["RCNN_base.0.weight"].to(x.dtype), stride=2, padding=3), p, "RCNN_base.1"))
    return F.max_pool2d(x, 3, 2, 0, ceil_mode=True)


def rcnn_base(x, p, num_layers=50):
    """RCNN_base = conv1, bn1, relu, maxpool, layer1..3 (dana.py:344-345)."""
    l = RES_LAYERS[num_layers]
    x = stem(x, p)
    x = _stage(x, p, "RCNN_base.4", l[0], 1)
    x = _stage(x, p, "RCNN_base.5", l[1], 2)
    return _stage(x, p, "RCNN_base.6", l[2], 2)


def head_to_tail(pooled, p, num_layers=50):
    """RCNN_top (layer4) + spatial mean (dana.py:346,387-389)."""
    return _stage(pooled, p, "RCNN_top.0", RES_LAYERS[num_layers][3], 2).mean(3).mean(2)


def rpn_head(feat, p, num_a):
    """_RPN.forward up to the proposal layer (lib/model/rpn/rpn.py:58-72)."""
    x = F.relu(F.conv2d(feat, p["RCNN_rpn.RPN_Conv.weight"].to(feat.dtype), p["RCNN_rpn.RPN_Conv.bias"].to(feat.dtype),
                        padding=1))
    cls = F.conv2d(x, p["RCNN_rpn.RPN_cls_score.weight"].to(feat.dtype), p["RCNN_rpn.RPN_cls_score.bias"].to(feat.dtype))
    b, _, h, w = cls.shape
    prob = F.softmax(cls.view(b, 2, num_a * h, w), 1).view(b, 2 * num_a, h, w)     # rpn.py:47-56,67-69
    bbox = F.conv2d(x, p["RCNN_rpn.RPN_bbox_pred.weight"].to(feat.dtype), p["RCNN_rpn.RPN_bbox_pred.bias"].to(feat.dtype))
    return prob, bbox


DEFAULT_CFG = dict(anchor_scales=(4, 8, 16, 32), anchor_ratios=(0.5, 1, 2), feat_stride=16, pre_nms_top_n=6000,
                   post_nms_top_n=300, nms_thresh=0.7, pooling_size=7, num_layers=50, semantic_enhance=True,
                   channel_gamma=0.1, unary_gamma=0.1)


def dana_forward_eval(p, im_data, im_info, support_ims, n_shot, cfg=None, roi_align_fn=None, nms_fn=None,
                      teacher=None):
    """_DAnARCNN.forward, eval branch (dana.py:87-220).  support_ims [B, sets*K, 3, Hs, Ws]; set 0 is the
    positive set that drives the RPN-level attention and the first head pass; every further set gets
    one more head pass (the reference does this for the negative set in training, dana.py:189-194, and
    ignores n_way in eval, :110-115 -- with sets == 1 this is exactly the reference's eval forward).
    Returns a dict of every stage output (used for teacher-forced per-stage parity)."""
    c = dict(DEFAULT_CFG)
    c.update(cfg or {})
    roi_align_fn = roi_align_fn or roi_align_forward
    out = {}
    b = im_data.shape[0]
    base_anchors = generate_anchors(ratios=c["anchor_ratios"], scales=c["anchor_scales"])
    num_a = base_anchors.shape[0]
    base_feat = rcnn_base(im_data, p, c["num_layers"])                               # :98
    out["base_feat"] = base_feat
    s = support_ims.reshape(-1, *support_ims.shape[2:])
    s_feat = rcnn_base(s, p, c["num_layers"])                                        # :111
    sets = support_ims.shape[1] // n_shot
    s_feat = s_feat.view(b, sets, n_shot, *s_feat.shape[1:])
    out["support_feat"] = s_feat
    k = s_feat.shape[-1] - 6                                                         # AvgPool2d(14,1): 20 -> 7
    s_pooled = F.avg_pool2d(s_feat.reshape(-1, *s_feat.shape[3:]), k, 1).view(b, sets, n_shot, s_feat.shape[3], 7, 7)
    out["support_pooled"] = s_pooled
    dense = ba_cisa_rpn(base_feat, s_feat[:, 0], p, c["semantic_enhance"], c["channel_gamma"], c["unary_gamma"])
    out["dense"] = dense
    corr = torch.cat([base_feat, dense], 1)                                          # :154
    prob, bbox = rpn_head(corr, p, num_a)
    out["rpn_cls_prob"], out["rpn_bbox_pred"] = prob, bbox
    rois = proposal_layer(prob, bbox, im_info, base_anchors, c["feat_stride"], c["pre_nms_top_n"],
                          c["post_nms_top_n"], c["nms_thresh"], nms_fn)
    if teacher is not None and "rois" in teacher:
        rois = teacher["rois"]
    out["rois"] = rois
    pooled = roi_align_fn(base_feat, rois.view(-1, 5), 1.0 / 16.0, c["pooling_size"], c["pooling_size"], 0)  # :183
    pooled = pooled.to(base_feat.dtype)
    out["pooled"] = pooled
    fc7 = head_to_tail(pooled, p, c["num_layers"])
    out["fc7"] = fc7
    out["bbox_pred"] = _lin(fc7, p, "RCNN_bbox_pred")                                # :246
    scores, probs = [], []
    for si in range(sets):                                                           # :196 (+ :190 per extra set)
        sc, pr = rcnn_head_attention(pooled, s_pooled[:, si], p, c["unary_gamma"])
        scores.append(sc)
        probs.append(pr)
    out["cls_score"] = torch.cat(scores, 0)
    out["cls_prob"] = torch.cat(probs, 0)
    return out


# --------------------------------------------------------------------------------------------
# deterministic parameters (numpy RandomState -> identical in the fixture generator, the tests,
# the oracle and the CUDA engine)
# --------------------------------------------------------------------------------------------
def param_shapes(num_layers=50, num_anchors=12, semantic_enhance=True):
    """State-dict keys/shapes of DAnARCNN('concat') (dana.py:45-84,344-348)."""
    shapes = {}

    def bn(name, ch):
        shapes[name + ".weight"] = (ch,)
        shapes[name + ".bias"] = (ch,)
        shapes[name + ".running_mean"] = (ch,)
        shapes[name + ".running_var"] = (ch,)

    def stage(prefix, inplanes, planes, blocks):
        for i in range(blocks):
            n = "%s.%d" % (prefix, i)
            cin = inplanes if i == 0 else planes * 4
            shapes[n + ".conv1.weight"] = (planes, cin, 1, 1)
            bn(n + ".bn1", planes)
            shapes[n + ".conv2.weight"] = (planes, planes, 3, 3)
            bn(n + ".bn2", planes)
            shapes[n + ".conv3.weight"] = (planes * 4, planes, 1, 1)
            bn(n + ".bn3", planes * 4)
            if i == 0:
                shapes[n + ".downsample.0.weight"] = (planes * 4, cin, 1, 1)
                bn(n + ".downsample.1", planes * 4)

    l = RES_LAYERS[num_layers]
    shapes["RCNN_base.0.weight"] = (64, 3, 7, 7)
    bn("RCNN_base.1", 64)
    stage("RCNN_base.4", 64, 64, l[0])
    stage("RCNN_base.5", 256, 128, l[1])
    stage("RCNN_base.6", 512, 256, l[2])
    stage("RCNN_top.0", 1024, 512, l[3])
    for pre in ("rpn", "rcnn"):
        shapes[pre + "_unary_layer.weight"] = (1, 1024)
        shapes[pre + "_unary_layer.bias"] = (1,)
        shapes[pre + "_adapt_q_layer.weight"] = (256, 1024)
        shapes[pre + "_adapt_q_layer.bias"] = (256,)
        shapes[pre + "_adapt_k_layer.weight"] = (256, 1024)
        shapes[pre + "_adapt_k_layer.bias"] = (256,)
    if semantic_enhance:
        shapes["rpn_channel_k_layer.weight"] = (1, 1024)
        shapes["rpn_channel_k_layer.bias"] = (1,)
    shapes["RCNN_rpn.RPN_Conv.weight"] = (512, 2048, 3, 3)
    shapes["RCNN_rpn.RPN_Conv.bias"] = (512,)
    shapes["RCNN_rpn.RPN_cls_score.weight"] = (2 * num_anchors, 512, 1, 1)
    shapes["RCNN_rpn.RPN_cls_score.bias"] = (2 * num_anchors,)
    shapes["RCNN_rpn.RPN_bbox_pred.weight"] = (4 * num_anchors, 512, 1, 1)
    shapes["RCNN_rpn.RPN_bbox_pred.bias"] = (4 * num_anchors,)
    shapes["rcnn_transform_layer.weight"] = (64, 2048)
    shapes["rcnn_transform_layer.bias"] = (64,)
    shapes["output_score_layer.linear1.weight"] = (1024, 3136)
    shapes["output_score_layer.linear1.bias"] = (1024,)
    shapes["output_score_layer.linear2.weight"] = (2, 1024)
    shapes["output_score_layer.linear2.bias"] = (2,)
    shapes["RCNN_bbox_pred.weight"] = (4, 2048)
    shapes["RCNN_bbox_pred.bias"] = (4,)
    return shapes


def make_params(seed=1996, num_layers=50, num_anchors=12, semantic_enhance=True, attn_std=0.01, head_std=0.01,
                random_bias=True):
    """Deterministic synthetic weights.  Distributions follow the reference init (conv N(0, sqrt(2/n)),
    resnet.py:123-129; linear / RPN N(0, .01), bbox N(0, .001), dana.py:46-69,234-238) but BN statistics
    and affine terms are randomised (so BN folding is exercised) and biases are non-zero when
    random_bias.  attn_std scales the attention projections (the reference's 0.01 gives |logit| < 0.3;
    tests also use 0.05).  Returns dict name -> torch fp32 tensor; key order is deterministic."""
    rs = np.random.RandomState(seed)
    p = {}
    for name, shape in param_shapes(num_layers, num_anchors, semantic_enhance).items():
        if name.endswith("running_mean"):
            v = rs.normal(0, 0.1, shape)
        elif name.endswith("running_var"):
            v = rs.uniform(0.5, 1.5, shape)
            if name.startswith("RCNN_base.1."):
                v = v * 230.0  # conv1 sees pixel-scale inputs (std 50): bring the stem output to O(1)
        elif ".bn" in name or ".downsample.1." in name or name.startswith("RCNN_base.1."):
            if name.endswith(".bn3.weight"):
                v = rs.uniform(0.3, 0.5, shape)  # damp the residual branch so 16 blocks stay O(1)
            elif name.endswith("weight"):
                v = rs.uniform(0.8, 1.2, shape)
            else:
                v = rs.normal(0, 0.05, shape)
        elif len(shape) == 4 and name.startswith(("RCNN_base", "RCNN_top")):
            v = rs.normal(0, math.sqrt(2.0 / (shape[2] * shape[3] * shape[0])), shape)
        elif name.endswith(".bias"):
            v = rs.normal(0, 0.02, shape) if random_bias else np.zeros(shape)
        elif name.startswith(("rpn_", "rcnn_adapt", "rcnn_unary")):
            v = rs.normal(0, attn_std, shape)
        elif name.startswith("RCNN_bbox_pred"):
            v = rs.normal(0, 0.001, shape)
        else:
            v = rs.normal(0, head_std, shape)
        p[name] = torch.from_numpy(np.asarray(v, dtype=np.float32))
    return p


def synth_inputs(seed, batch, height, width, n_support, support_size=320):
    """Synthetic episode: pixel-scale gaussian images, im_info = (H, W, 1)."""
    rs = np.random.RandomState(seed)
    im = torch.from_numpy((rs.standard_normal((batch, 3, height, width)) * 50).astype(np.float32))
    sup = torch.from_numpy((rs.standard_normal((batch, n_support, 3, support_size, support_size)) * 50)
                           .astype(np.float32))
    info = torch.tensor([[float(height), float(width), 1.0]] * batch)
    return im, info, sup
