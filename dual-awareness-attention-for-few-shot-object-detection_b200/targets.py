"""Target layers of the training branch, host side.

The reference computes its RPN / R-CNN training targets with torch index ops driven by numpy's GLOBAL random state
(lib/model/rpn/anchor_target_layer.py:48-193, lib/model/rpn/proposal_target_layer_cascade.py:33-213): label
assignment is a few hundred kFLOP of IoU arithmetic followed by `np.random.permutation` / `np.random.rand` draws whose
sizes depend on the data, i.e. host work with a device round trip by construction.  This module does the same work
in numpy float32 (same expression order, so the threshold decisions agree bit for bit with torch CPU fp32) and makes
the SAME numpy RNG calls in the SAME order, so that under one `np.random.seed` the reference and this path draw the
same samples.  The anchor targets depend on the ground truth only and are computed while the GPU runs the trunk; the
proposal targets need the proposal layer's rois (one D2H of [B, 2000, 5] floats, like the reference's own syncs).

Outputs are laid out for the device kernels of csrc/train_ops.cuh: anchors in (y, x, a) order -- the order of the
NHWC RPN output -- instead of the reference's [B, 1, A*H, W] / [B, 4A, H, W] views of the same values."""
import numpy as np

from .anchors import generate_anchors

F32 = np.float32


def anchor_grid(base_anchors, feat_h, feat_w, feat_stride):
    """All anchors, index (y*W + x)*A + a (anchor_target_layer.py:59-76, proposal_layer.py:79-93)."""
    sx = (np.arange(feat_w) * feat_stride).astype(F32)
    sy = (np.arange(feat_h) * feat_stride).astype(F32)
    gx, gy = np.meshgrid(sx, sy)
    shifts = np.stack([gx.ravel(), gy.ravel(), gx.ravel(), gy.ravel()], 1)
    return (np.asarray(base_anchors, dtype=F32)[None, :, :] + shifts[:, None, :]).reshape(-1, 4)


def overlaps_batch(boxes, gt_boxes):
    """IoU of boxes [N,4] (shared) or [B,N,4] with gt_boxes [B,G,5] -> [B,N,G] (bbox_transform.py:168-260): "+1"
    extents; a zero-area ground-truth column (padding) scores 0, a zero-area box row scores -1."""
    gt = np.asarray(gt_boxes, dtype=F32)
    b = gt.shape[0]
    bx = np.asarray(boxes, dtype=F32)
    if bx.ndim == 2:
        bx = np.broadcast_to(bx[None], (b,) + bx.shape)
    gw = gt[:, :, 2] - gt[:, :, 0] + F32(1)
    gh = gt[:, :, 3] - gt[:, :, 1] + F32(1)
    aw = bx[:, :, 2] - bx[:, :, 0] + F32(1)
    ah = bx[:, :, 3] - bx[:, :, 1] + F32(1)
    g_area = (gw * gh)[:, None, :]
    a_area = (aw * ah)[:, :, None]
    iw = np.minimum(bx[:, :, None, 2], gt[:, None, :, 2]) - np.maximum(bx[:, :, None, 0], gt[:, None, :, 0]) + F32(1)
    ih = np.minimum(bx[:, :, None, 3], gt[:, None, :, 3]) - np.maximum(bx[:, :, None, 1], gt[:, None, :, 1]) + F32(1)
    iw = np.maximum(iw, F32(0))
    ih = np.maximum(ih, F32(0))
    inter = iw * ih
    ov = inter / (a_area + g_area - inter)
    ov = np.where(((gw == 1) & (gh == 1))[:, None, :], F32(0), ov)
    ov = np.where(((aw == 1) & (ah == 1))[:, :, None], F32(-1), ov)
    return ov.astype(F32)


def regression_targets(ex, gt):
    """(dx, dy, dw, dh) of gt boxes w.r.t. example boxes, [...,4] each (bbox_transform.py:37-75)."""
    ex = np.asarray(ex, dtype=F32)
    gt = np.asarray(gt, dtype=F32)
    ew = ex[..., 2] - ex[..., 0] + F32(1)
    eh = ex[..., 3] - ex[..., 1] + F32(1)
    ecx = ex[..., 0] + F32(0.5) * ew
    ecy = ex[..., 1] + F32(0.5) * eh
    gw = gt[..., 2] - gt[..., 0] + F32(1)
    gh = gt[..., 3] - gt[..., 1] + F32(1)
    gcx = gt[..., 0] + F32(0.5) * gw
    gcy = gt[..., 1] + F32(0.5) * gh
    return np.stack([(gcx - ecx) / ew, (gcy - ecy) / eh, np.log(gw / ew), np.log(gh / eh)], -1).astype(F32)


def _used_columns(gt):
    """Number of leading ground-truth columns that hold a box in at least one image.  The trailing zero padding
    (gt_boxes is padded to MAX_NUM_GT_BOXES = 50) scores 0 against everything (-1 against zero-area rows), so
    leaving it out changes neither the row maxima / argmaxima nor the per-column tests -- it only spares the host
    50 columns of IoU arithmetic for the 1-8 boxes an image has."""
    real = ((gt[:, :, 2] - gt[:, :, 0] != 0) | (gt[:, :, 3] - gt[:, :, 1] != 0)).any(0)
    idx = np.nonzero(real)[0]
    return int(idx[-1]) + 1 if idx.size else 1


def anchor_targets(feat_h, feat_w, gt_boxes, im_info, base_anchors, feat_stride=16, *, negative_overlap=0.3,
                   positive_overlap=0.7, clobber_positives=False, fg_fraction=0.5, batchsize=256, inside_weight=1.0,
                   positive_weight=-1.0, allowed_border=0):
    """anchor_target_layer.py:48-193.  gt_boxes [B,G,5] (zero padded), im_info [B,3].
    -> labels int8 [B, H*W*A] (1 fg, 0 bg, -1 ignored), bbox_targets f32 [B, H*W*A, 4], inside_w f32 [B, H*W*A],
       outside_w f32 [B, H*W*A], all in (y, x, a) anchor order.
    Reference quirks kept: the image-border test uses image 0's size for the whole batch (:92-95), the RPN loss
    normaliser is the LAST image's example count (:163-165, python loop variable)."""
    gt = np.asarray(gt_boxes, dtype=F32)
    b = gt.shape[0]
    all_anchors = anchor_grid(base_anchors, feat_h, feat_w, feat_stride)
    total = all_anchors.shape[0]
    im_h, im_w = int(im_info[0][0]), int(im_info[0][1])
    inside = np.nonzero((all_anchors[:, 0] >= -allowed_border) & (all_anchors[:, 1] >= -allowed_border) &
                        (all_anchors[:, 2] < im_w + allowed_border) & (all_anchors[:, 3] < im_h + allowed_border))[0]
    anchors = all_anchors[inside]
    n_in = inside.size
    labels = np.full((b, n_in), -1, dtype=np.int8)
    ov = overlaps_batch(anchors, gt[:, :_used_columns(gt)])            # [B, n_in, G']
    max_ov = ov.max(2)
    argmax_ov = ov.argmax(2)
    gt_max = ov.max(1)                                                 # [B, G]
    if not clobber_positives:
        labels[max_ov < F32(negative_overlap)] = 0
    gt_max = np.where(gt_max == 0, F32(1e-5), gt_max)
    best = (ov == gt_max[:, None, :]).sum(2)
    if best.sum() > 0:
        labels[best > 0] = 1
    labels[max_ov >= F32(positive_overlap)] = 1
    if clobber_positives:
        labels[max_ov < F32(negative_overlap)] = 0
    num_fg = int(fg_fraction * batchsize)
    sum_fg = (labels == 1).sum(1)
    sum_bg = (labels == 0).sum(1)
    i = 0
    for i in range(b):
        if sum_fg[i] > num_fg:
            fg = np.nonzero(labels[i] == 1)[0]
            perm = np.random.permutation(fg.size)                      # :131
            labels[i, fg[perm[:fg.size - num_fg]]] = -1
        num_bg = batchsize - int((labels[i] == 1).sum())
        if sum_bg[i] > num_bg:
            bg = np.nonzero(labels[i] == 0)[0]
            perm = np.random.permutation(bg.size)                      # :143
            labels[i, bg[perm[:bg.size - num_bg]]] = -1
    gt_sel = np.take_along_axis(gt[:, :, :4], argmax_ov[:, :, None].repeat(4, 2), 1)
    tgt = regression_targets(np.broadcast_to(anchors[None], (b, n_in, 4)), gt_sel)
    if positive_weight >= 0:
        raise NotImplementedError("TRAIN.RPN_POSITIVE_WEIGHT >= 0 is not used by any shipped configuration")
    num_examples = int((labels[i] >= 0).sum())
    w_out = F32(1.0 / num_examples)          # python double division, stored as fp32 (:165-166)
    labels_all = np.full((b, total), -1, dtype=np.int8)
    labels_all[:, inside] = labels
    tgt_all = np.zeros((b, total, 4), dtype=F32)
    tgt_all[:, inside] = tgt
    in_all = np.zeros((b, total), dtype=F32)
    in_all[:, inside] = np.where(labels == 1, F32(inside_weight), F32(0))
    out_all = np.zeros((b, total), dtype=F32)
    out_all[:, inside] = np.where(labels >= 0, w_out, F32(0))
    return labels_all, tgt_all, in_all, out_all


def proposal_targets(all_rois, gt_boxes, *, rois_per_image=128, fg_fraction=0.25, fg_thresh=0.5, bg_thresh_hi=0.5,
                     bg_thresh_lo=0.0, normalize_means=(0.0, 0.0, 0.0, 0.0), normalize_stds=(0.1, 0.1, 0.2, 0.2),
                     inside_weights=(1.0, 1.0, 1.0, 1.0), normalize_targets=True):
    """proposal_target_layer_cascade.py:33-213.  all_rois [B,N,5] (batch index, box), gt_boxes [B,G,5] (box, class).
    -> rois f32 [B,R,5], labels f32 [B,R], bbox_targets [B,R,4], inside_w [B,R,4], outside_w [B,R,4]."""
    rois_in = np.asarray(all_rois, dtype=F32)
    gt = np.asarray(gt_boxes, dtype=F32)
    b, g = gt.shape[0], gt.shape[1]
    app = np.zeros((b, g, 5), dtype=F32)
    app[:, :, 1:5] = gt[:, :, :4]
    rois_all = np.concatenate([rois_in, app], 1)                       # the ground-truth boxes join the candidates (:44)
    rpi = int(rois_per_image)
    fg_rpi = int(np.round(fg_fraction * rpi)) or 1
    ov = overlaps_batch(rois_all[:, :, 1:5], gt[:, :_used_columns(gt)])
    max_ov = ov.max(2)
    assign = ov.argmax(2)
    labels = np.take_along_axis(gt[:, :, 4], assign, 1)
    labels_b = np.zeros((b, rpi), dtype=F32)
    rois_b = np.zeros((b, rpi, 5), dtype=F32)
    gt_b = np.zeros((b, rpi, 5), dtype=F32)
    for i in range(b):
        fg = np.nonzero(max_ov[i] >= F32(fg_thresh))[0]
        bg = np.nonzero((max_ov[i] < F32(bg_thresh_hi)) & (max_ov[i] >= F32(bg_thresh_lo)))[0]
        nf, nb = fg.size, bg.size
        if nf > 0 and nb > 0:
            fg_this = min(fg_rpi, nf)
            fg = fg[np.random.permutation(nf)[:fg_this]]               # :159
            bg_this = rpi - fg_this
            bg = bg[np.floor(np.random.rand(bg_this) * nb).astype(np.int64)]     # :168 (with replacement)
        elif nf > 0:
            fg = fg[np.floor(np.random.rand(rpi) * nf).astype(np.int64)]         # :175
            fg_this, bg = rpi, bg[:0]
        elif nb > 0:
            bg = bg[np.floor(np.random.rand(rpi) * nb).astype(np.int64)]         # :183
            fg_this, fg = 0, fg[:0]
        else:
            raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
        keep = np.concatenate([fg, bg])
        labels_b[i] = labels[i, keep]
        if fg_this < rpi:
            labels_b[i, fg_this:] = 0
        rois_b[i] = rois_all[i, keep]
        rois_b[i, :, 0] = i
        gt_b[i] = gt[i, assign[i, keep]]
    t = regression_targets(rois_b[:, :, 1:5], gt_b[:, :, :4])
    if normalize_targets:
        t = (t - np.asarray(normalize_means, dtype=F32)) / np.asarray(normalize_stds, dtype=F32)
    pos = (labels_b > 0).astype(F32)[:, :, None]
    pos = pos * (labels_b.sum(1) != 0).astype(F32)[:, None, None]      # an image without positives contributes nothing (:91)
    tgt = (t * pos).astype(F32)
    in_w = (pos * np.asarray(inside_weights, dtype=F32)).astype(F32)
    out_w = (in_w > 0).astype(F32)
    return rois_b, labels_b, tgt, in_w, out_w


__all__ = ["anchor_grid", "overlaps_batch", "regression_targets", "anchor_targets", "proposal_targets",
           "generate_anchors"]
