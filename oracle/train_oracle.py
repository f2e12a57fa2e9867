"""ORACLE (test infrastructure, never imported by the product path): torch-CPU / numpy restatement of the TRAINING
branch of _DAnARCNN.forward: anchor targets, RPN losses, proposal targets, the two head
passes (positive / negative support set) and the hard-negative-mined classification loss.  Groundwork for the training
step: pinned against the unmodified reference run in train mode on CPU (oracle/make_golden_train.py ->
tests/golden/forward_train_small.npz, tests/test_oracle_pins.py).

The reference subsamples with numpy's global RNG; this restatement makes the same calls in the same order
(anchor_target_layer.py:131,143, proposal_target_layer_cascade.py:159,168,175,183), so under the same
np.random.seed both produce the same samples.

Training configuration constants (lib/model/utils/config.py:90-200, cfgs/res50.yml)."""
import numpy as np
import torch
import torch.nn.functional as F

import dana_oracle as O

TRAIN_CFG = dict(rpn_pre_nms_top_n=12000, rpn_post_nms_top_n=2000, rpn_nms_thresh=0.7, rpn_negative_overlap=0.3,
                 rpn_positive_overlap=0.7, rpn_clobber_positives=False, rpn_fg_fraction=0.5, rpn_batchsize=256,
                 rpn_bbox_inside_weight=1.0, rpn_positive_weight=-1.0, batch_size=128, fg_fraction=0.25, fg_thresh=0.5,
                 bg_thresh_hi=0.5, bg_thresh_lo=0.0, bbox_normalize_means=(0.0, 0.0, 0.0, 0.0),
                 bbox_normalize_stds=(0.1, 0.1, 0.2, 0.2), bbox_inside_weights=(1.0, 1.0, 1.0, 1.0),
                 bbox_normalize_targets_precomputed=True)


def bbox_overlaps_batch(anchors, gt_boxes):
    """lib/model/rpn/bbox_transform.py:168-260.  anchors [N,4] or [B,N,4|5], gt_boxes [B,K,5] -> [B,N,K]; zero-area gt
    columns (padding) are 0, zero-area anchors are -1."""
    b = gt_boxes.shape[0]
    if anchors.dim() == 2:
        anchors = anchors.view(1, -1, 4).expand(b, -1, 4)
    elif anchors.shape[2] == 5:
        anchors = anchors[:, :, 1:5]
    gt = gt_boxes[:, :, :4]
    gx = gt[:, :, 2] - gt[:, :, 0] + 1
    gy = gt[:, :, 3] - gt[:, :, 1] + 1
    ax = anchors[:, :, 2] - anchors[:, :, 0] + 1
    ay = anchors[:, :, 3] - anchors[:, :, 1] + 1
    g_area = (gx * gy).unsqueeze(1)
    a_area = (ax * ay).unsqueeze(2)
    bx, qx = anchors.unsqueeze(2), gt.unsqueeze(1)
    iw = (torch.min(bx[..., 2], qx[..., 2]) - torch.max(bx[..., 0], qx[..., 0]) + 1).clamp_min(0)
    ih = (torch.min(bx[..., 3], qx[..., 3]) - torch.max(bx[..., 1], qx[..., 1]) + 1).clamp_min(0)
    ua = a_area + g_area - iw * ih
    ov = iw * ih / ua
    ov = ov.masked_fill(((gx == 1) & (gy == 1)).unsqueeze(1).expand_as(ov), 0)
    ov = ov.masked_fill(((ax == 1) & (ay == 1)).unsqueeze(2).expand_as(ov), -1)
    return ov


def bbox_transform_batch(ex, gt):
    """bbox_transform.py:37-75: regression targets of `gt` boxes w.r.t. `ex` boxes ([N,4] or [B,N,4])."""
    if ex.dim() == 2:
        ex = ex.unsqueeze(0)
    ew = ex[..., 2] - ex[..., 0] + 1.0
    eh = ex[..., 3] - ex[..., 1] + 1.0
    ecx = ex[..., 0] + 0.5 * ew
    ecy = ex[..., 1] + 0.5 * eh
    gw = gt[..., 2] - gt[..., 0] + 1.0
    gh = gt[..., 3] - gt[..., 1] + 1.0
    gcx = gt[..., 0] + 0.5 * gw
    gcy = gt[..., 1] + 0.5 * gh
    return torch.stack(((gcx - ecx) / ew, (gcy - ecy) / eh, torch.log(gw / ew), torch.log(gh / eh)), 2)


def smooth_l1_loss(pred, targets, inside_w, outside_w, sigma=1.0, dim=(1,)):
    """lib/model/utils/net_utils.py:71-85."""
    s2 = sigma ** 2
    d = inside_w * (pred - targets)
    a = d.abs()
    sign = (a < 1.0 / s2).float()
    loss = outside_w * (d.pow(2) * (s2 / 2.0) * sign + (a - 0.5 / s2) * (1.0 - sign))
    for i in sorted(dim, reverse=True):
        loss = loss.sum(i)
    return loss.mean()


def anchor_target_layer(feat_h, feat_w, gt_boxes, im_info, base_anchors, feat_stride=16, cfg=None):
    """lib/model/rpn/anchor_target_layer.py:48-193 -> (labels [B,1,A*H,W], bbox_targets [B,4A,H,W],
    inside weights, outside weights).  Consumes np.random.permutation like the reference (:131,:143)."""
    c = dict(TRAIN_CFG)
    c.update(cfg or {})
    b = gt_boxes.shape[0]
    num_a = base_anchors.shape[0]
    all_anchors = O.anchor_grid(base_anchors, feat_h, feat_w, feat_stride).to(gt_boxes.dtype)     # [(y*w+x)*A + a, 4]
    total = all_anchors.shape[0]
    keep = ((all_anchors[:, 0] >= 0) & (all_anchors[:, 1] >= 0) & (all_anchors[:, 2] < int(im_info[0][1])) &
            (all_anchors[:, 3] < int(im_info[0][0])))                                             # :92-95 (image 0's size)
    inds_inside = torch.nonzero(keep).view(-1)
    anchors = all_anchors[inds_inside]
    n_in = inds_inside.numel()
    labels = gt_boxes.new_full((b, n_in), -1)
    inside_w = gt_boxes.new_zeros((b, n_in))
    outside_w = gt_boxes.new_zeros((b, n_in))
    overlaps = bbox_overlaps_batch(anchors, gt_boxes)
    max_ov, argmax_ov = overlaps.max(2)
    gt_max, _ = overlaps.max(1)
    if not c["rpn_clobber_positives"]:
        labels[max_ov < c["rpn_negative_overlap"]] = 0
    gt_max[gt_max == 0] = 1e-5
    hit = overlaps.eq(gt_max.view(b, 1, -1).expand_as(overlaps)).sum(2)
    if hit.sum() > 0:
        labels[hit > 0] = 1
    labels[max_ov >= c["rpn_positive_overlap"]] = 1
    if c["rpn_clobber_positives"]:
        labels[max_ov < c["rpn_negative_overlap"]] = 0
    num_fg = int(c["rpn_fg_fraction"] * c["rpn_batchsize"])
    sum_fg = (labels == 1).int().sum(1)
    sum_bg = (labels == 0).int().sum(1)
    i = 0
    for i in range(b):
        if sum_fg[i] > num_fg:
            fg = torch.nonzero(labels[i] == 1).view(-1)
            perm = torch.from_numpy(np.random.permutation(fg.numel())).long()
            labels[i][fg[perm[:fg.numel() - num_fg]]] = -1
        num_bg = c["rpn_batchsize"] - int((labels[i] == 1).sum())
        if sum_bg[i] > num_bg:
            bg = torch.nonzero(labels[i] == 0).view(-1)
            perm = torch.from_numpy(np.random.permutation(bg.numel())).long()
            labels[i][bg[perm[:bg.numel() - num_bg]]] = -1
    gt_sel = torch.gather(gt_boxes[:, :, :4], 1, argmax_ov.unsqueeze(2).expand(-1, -1, 4))
    bbox_targets = bbox_transform_batch(anchors, gt_sel)
    inside_w[labels == 1] = c["rpn_bbox_inside_weight"]
    assert c["rpn_positive_weight"] < 0
    num_examples = int((labels[i] >= 0).sum())          # the reference uses the LAST image's count (:165, loop variable i)
    outside_w[labels == 1] = 1.0 / num_examples
    outside_w[labels == 0] = 1.0 / num_examples

    def unmap(data, fill):
        if data.dim() == 2:
            ret = data.new_full((b, total), fill)
            ret[:, inds_inside] = data
        else:
            ret = data.new_full((b, total, data.shape[2]), fill)
            ret[:, inds_inside, :] = data
        return ret
    labels = unmap(labels, -1).view(b, feat_h, feat_w, num_a).permute(0, 3, 1, 2).reshape(b, 1, num_a * feat_h, feat_w)
    bbox_targets = unmap(bbox_targets, 0).view(b, feat_h, feat_w, num_a * 4).permute(0, 3, 1, 2).contiguous()
    inside_w = unmap(inside_w, 0).view(b, total, 1).expand(b, total, 4).reshape(b, feat_h, feat_w, 4 * num_a)
    outside_w = unmap(outside_w, 0).view(b, total, 1).expand(b, total, 4).reshape(b, feat_h, feat_w, 4 * num_a)
    return labels, bbox_targets, inside_w.permute(0, 3, 1, 2).contiguous(), outside_w.permute(0, 3, 1, 2).contiguous()


def rpn_losses(rpn_cls_score, rpn_bbox_pred, targets, num_a):
    """lib/model/rpn/rpn.py:96-116.  rpn_cls_score [B,2A,H,W] (pre-softmax)."""
    labels, bbox_targets, inside_w, outside_w = targets
    b, _, h, w = rpn_cls_score.shape
    score = rpn_cls_score.view(b, 2, num_a * h, w).permute(0, 2, 3, 1).reshape(-1, 2)
    lab = labels.view(-1)
    keep = torch.nonzero(lab.ne(-1)).view(-1)
    loss_cls = F.cross_entropy(score[keep], lab[keep].long())
    loss_box = smooth_l1_loss(rpn_bbox_pred, bbox_targets, inside_w, outside_w, sigma=3, dim=(1, 2, 3))
    return loss_cls, loss_box


def proposal_target_layer(all_rois, gt_boxes, cfg=None):
    """lib/model/rpn/proposal_target_layer_cascade.py:33-213 -> (rois [B,R,5], labels [B,R], bbox_targets [B,R,4],
    inside weights, outside weights), R = TRAIN.BATCH_SIZE.  numpy RNG calls as in :159,:168,:175,:183."""
    c = dict(TRAIN_CFG)
    c.update(cfg or {})
    b = gt_boxes.shape[0]
    app = gt_boxes.new_zeros(gt_boxes.shape)
    app[:, :, 1:5] = gt_boxes[:, :, :4]
    all_rois = torch.cat([all_rois, app], 1)                                                    # :44
    rpi = int(c["batch_size"])
    fg_rpi = int(np.round(c["fg_fraction"] * rpi)) or 1
    overlaps = bbox_overlaps_batch(all_rois, gt_boxes)
    max_ov, gt_assign = overlaps.max(2)
    labels = torch.gather(gt_boxes[:, :, 4], 1, gt_assign)
    labels_batch = labels.new_zeros((b, rpi))
    rois_batch = all_rois.new_zeros((b, rpi, 5))
    gt_rois_batch = all_rois.new_zeros((b, rpi, 5))
    for i in range(b):
        fg = torch.nonzero(max_ov[i] >= c["fg_thresh"]).view(-1)
        bg = torch.nonzero((max_ov[i] < c["bg_thresh_hi"]) & (max_ov[i] >= c["bg_thresh_lo"])).view(-1)
        nf, nb = fg.numel(), bg.numel()
        if nf > 0 and nb > 0:
            fg_this = min(fg_rpi, nf)
            fg = fg[torch.from_numpy(np.random.permutation(nf)).long()[:fg_this]]
            bg_this = rpi - fg_this
            bg = bg[torch.from_numpy(np.floor(np.random.rand(bg_this) * nb)).long()]
        elif nf > 0:
            fg = fg[torch.from_numpy(np.floor(np.random.rand(rpi) * nf)).long()]
            fg_this, bg = rpi, bg[:0]
        elif nb > 0:
            bg = bg[torch.from_numpy(np.floor(np.random.rand(rpi) * nb)).long()]
            fg_this, fg = 0, fg[:0]
        else:
            raise ValueError("bg_num_rois = 0 and fg_num_rois = 0, this should not happen!")
        keep = torch.cat([fg, bg], 0)
        labels_batch[i] = labels[i][keep]
        if fg_this < rpi:
            labels_batch[i][fg_this:] = 0
        rois_batch[i] = all_rois[i][keep]
        rois_batch[i, :, 0] = i
        gt_rois_batch[i] = gt_boxes[i][gt_assign[i][keep]]
    t = bbox_transform_batch(rois_batch[:, :, 1:5], gt_rois_batch[:, :, :4])
    if c["bbox_normalize_targets_precomputed"]:
        t = (t - t.new_tensor(c["bbox_normalize_means"])) / t.new_tensor(c["bbox_normalize_stds"])
    pos = (labels_batch > 0).unsqueeze(2).to(t.dtype)
    pos = pos * (labels_batch.sum(1) != 0).view(b, 1, 1).to(t.dtype)      # images with no positive label are skipped (:91)
    bbox_targets = t * pos
    inside_w = pos * t.new_tensor(c["bbox_inside_weights"])
    outside_w = (inside_w > 0).to(t.dtype)
    return rois_batch, labels_batch, bbox_targets, inside_w, outside_w


def rcnn_cls_loss(cls_score_all, rois_label):
    """dana.py:203-214: 2-way cross entropy over all fg, the hardest bg of the positive-support half (2 x fg, at most
    a quarter of the rows) and the hardest bg of the negative-support half (<= fg)."""
    n = rois_label.shape[0]
    fg = torch.nonzero(rois_label == 1).view(-1)
    bg = torch.nonzero(rois_label == 0).view(-1)
    sm = F.softmax(cls_score_all, dim=1)
    bg_num_0 = max(1, min(fg.numel() * 2, int(n * 0.25)))
    bg_num_1 = max(1, min(fg.numel(), bg_num_0))
    order = torch.sort(sm[bg, 1], descending=True)[1]
    real_bg = bg[order]
    top0 = real_bg[real_bg < int(n * 0.5)][:bg_num_0]
    top1 = real_bg[real_bg >= int(n * 0.5)][:bg_num_1]
    idx = torch.cat([fg, top0, top1], 0)
    return F.cross_entropy(cls_score_all[idx], rois_label[idx])


def dana_forward_train(p, im_data, im_info, gt_boxes, num_boxes, support_ims, n_shot, cfg=None, roi_align_fn=None,
                       nms_fn=None):
    """_DAnARCNN.forward, training branch (dana.py:87-220, n_way = 2: positive set then negative set).
    Returns a dict with rois, cls_prob, bbox_pred, the four losses and rois_label."""
    c = dict(O.DEFAULT_CFG)
    c.update(cfg or {})
    roi_align_fn = roi_align_fn or O.roi_align_forward
    b = im_data.shape[0]
    base_anchors = O.generate_anchors(ratios=c["anchor_ratios"], scales=c["anchor_scales"])
    num_a = base_anchors.shape[0]
    base_feat = O.rcnn_base(im_data, p, c["num_layers"])
    s_feat = O.rcnn_base(support_ims.reshape(-1, *support_ims.shape[2:]), p, c["num_layers"])
    s_feat = s_feat.view(b, 2, n_shot, *s_feat.shape[1:])
    k = s_feat.shape[-1] - 6
    s_pooled = F.avg_pool2d(s_feat.reshape(-1, *s_feat.shape[3:]), k, 1).view(b, 2, n_shot, s_feat.shape[3], 7, 7)
    dense = O.ba_cisa_rpn(base_feat, s_feat[:, 0], p, c["semantic_enhance"], c["channel_gamma"], c["unary_gamma"])
    corr = torch.cat([base_feat, dense], 1)
    x = F.relu(F.conv2d(corr, p["RCNN_rpn.RPN_Conv.weight"], p["RCNN_rpn.RPN_Conv.bias"], padding=1))
    cls_score = F.conv2d(x, p["RCNN_rpn.RPN_cls_score.weight"], p["RCNN_rpn.RPN_cls_score.bias"])
    _, _, fh, fw = cls_score.shape
    prob = F.softmax(cls_score.view(b, 2, num_a * fh, fw), 1).view(b, 2 * num_a, fh, fw)
    bbox = F.conv2d(x, p["RCNN_rpn.RPN_bbox_pred.weight"], p["RCNN_rpn.RPN_bbox_pred.bias"])
    t = TRAIN_CFG
    # the proposal layer works on .data (proposal_layer.py:61-62): no gradient through the boxes
    rois = O.proposal_layer(prob.detach(), bbox.detach(), im_info, base_anchors, c["feat_stride"], t["rpn_pre_nms_top_n"],
                            t["rpn_post_nms_top_n"], t["rpn_nms_thresh"], nms_fn)
    all_rois = rois
    targets = anchor_target_layer(fh, fw, gt_boxes, im_info, torch.from_numpy(base_anchors).float(), c["feat_stride"])
    rpn_loss_cls, rpn_loss_box = rpn_losses(cls_score, bbox, targets, num_a)
    rois, rois_label, rois_target, in_w, out_w = proposal_target_layer(rois, gt_boxes)
    rois_label = rois_label.view(-1).long()
    pooled = roi_align_fn(base_feat, rois.view(-1, 5), 1.0 / 16.0, c["pooling_size"], c["pooling_size"], 0).to(base_feat.dtype)
    fc7 = O.head_to_tail(pooled, p, c["num_layers"])
    bbox_pred = O._lin(fc7, p, "RCNN_bbox_pred")
    sc_pos, pr_pos = O.rcnn_head_attention(pooled, s_pooled[:, 0], p, c["unary_gamma"])
    sc_neg, pr_neg = O.rcnn_head_attention(pooled, s_pooled[:, 1], p, c["unary_gamma"])
    cls_prob = torch.cat([pr_pos, pr_neg], 0)
    cls_score_all = torch.cat([sc_pos, sc_neg], 0)
    label_all = torch.cat([rois_label, torch.zeros_like(rois_label)], 0)
    loss_bbox = smooth_l1_loss(bbox_pred, rois_target.view(-1, 4), in_w.view(-1, 4), out_w.view(-1, 4))
    loss_cls = rcnn_cls_loss(cls_score_all, label_all)
    return dict(rois=rois, cls_prob=cls_prob, bbox_pred=bbox_pred, rpn_loss_cls=rpn_loss_cls, rpn_loss_box=rpn_loss_box,
                RCNN_loss_cls=loss_cls, RCNN_loss_bbox=loss_bbox, rois_label=label_all, rpn_targets=targets,
                all_rois=all_rois, rpn_cls_score=cls_score, rpn_bbox_pred=bbox, cls_score=cls_score_all)
