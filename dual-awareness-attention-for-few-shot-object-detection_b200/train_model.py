"""Training branch of _DAnARCNN.forward WITH its autograd graph (BASELINE configs[3]):
`loss.backward()` of train.py:138 runs every convolution / projection / attention product backward on the tcgen05
GEMM through dana_b200.autograd_ops, the RoIAlign backward on dana_roi_align_backward.

What runs where:
  frozen stem + layer1 (dana.py:350-368, FIXED_BLOCKS = 1)   the eval kernels, no graph
  layer2, layer3 (query and support crops), layer4           autograd_ops.conv: fwd + data-grad + weight-grad GEMMs
  RPN 3x3 conv + heads, q / k projections, transform, FFN    autograd_ops.conv / linear
  attention logits and P.V products (dana.py:142,147,273,278) autograd_ops.bmm_nt
  RoIAlign (dana.py:183)                                     dana_roi_align_forward / dana_roi_align_backward (NHWC)
  proposals + NMS (no gradient, rpn.py:74-78)                dana_proposals
  anchor / proposal targets                                  host numpy, the reference's RNG call order (targets.py)
  softmax, mean-centering, leaky-ReLU gate, PE add, the C -> 1 weighted sums, average pool, losses
                                                             torch element-wise ops (ATen) -- glue, < 2 % of the FLOPs

Activations are fp32 NHWC; every GEMM operand is a split-bf16 pair (fp32-equivalent products)."""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import autograd_ops as A
from . import ops, targets
from .config import cfg
from .engine import BN_EPS, RES_LAYERS, _Block, positional_encoding


# Module switch: independent branches of the step (the two trunks, the three head branches) on side streams.  Pays only when
# the GPU, not the host, is the limit -- i.e. under CUDA-graph capture (train_step._CapturedStep turns it on); the eager
# step is launch-bound and only gains stream-switch overhead from it (84 -> 106 ms).
SIDE_STREAM = False


class _RoIAlignNHWC(torch.autograd.Function):
    """ROIAlign(7, 7, 1/16, 0) on an NHWC map (roi_layers/roi_align.py:12-45): forward and backward kernels."""

    @staticmethod
    def forward(ctx, feat, rois):
        out, _ = ops.roi_align_nhwc(feat.contiguous(), rois, 1.0 / 16.0, 7, 0, want_f32=True, want_pair=False)
        ctx.save_for_backward(rois)
        ctx.shape = tuple(feat.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        (rois,) = ctx.saved_tensors
        b, h, w, c = ctx.shape
        return ops.roi_align_backward_nhwc(g.reshape(g.shape[0], 49, c), rois, 1.0 / 16.0, b, h, w, 0), None


class FrozenStem:
    """conv1 / bn1 / relu / maxpool / layer1 -- frozen in training (dana.py:350-359) -- packed once for the eval kernels."""

    def __init__(self, sd, device):
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        self.stem_w = ops.pack_stem_weight(f32(sd["RCNN_base.0.weight"]), True)
        g, b, m, v = [f32(sd["RCNN_base.1." + n]) for n in ("weight", "bias", "running_mean", "running_var")]
        self.scale = (g / torch.sqrt(v + BN_EPS)).contiguous()
        self.bias = (b - m * self.scale).contiguous()
        n_blocks = len([k for k in sd if k.startswith("RCNN_base.4.") and k.endswith(".conv1.weight")])
        sdf = {k: v.detach().float() for k, v in sd.items() if k.startswith("RCNN_base.4.")}
        self.blocks = [_Block(sdf, "RCNN_base.4.%d" % i, device, True) for i in range(n_blocks)]

    def __call__(self, im_nchw):
        x = ops.stem(im_nchw.float().contiguous(), self.stem_w, self.scale, self.bias, split=True)
        for blk in self.blocks:
            y = ops.conv_nhwc(x, blk.c1.w, blk.c1.n_out, ksize=1, scale=blk.c1.scale, bias=blk.c1.bias, relu=True)
            y = ops.conv_nhwc(y, blk.c2.w, blk.c2.n_out, ksize=3, scale=blk.c2.scale, bias=blk.c2.bias, relu=True)
            res = x if blk.down is None else ops.conv_nhwc(x, blk.down.w, blk.down.n_out, ksize=1, scale=blk.down.scale,
                                                           bias=blk.down.bias)
            x = ops.conv_nhwc(y, blk.c3.w, blk.c3.n_out, ksize=1, scale=blk.c3.scale, bias=blk.c3.bias, res=res, relu=True)
        out = ops.merge_pair(x)
        out._dana_pair = x
        return out


class TrainGraph:
    """Functional training forward over a name -> tensor dict `p` (the module's parameters and buffers)."""

    def __init__(self, p, num_layers=50, n_shot=3, semantic_enhance=True, channel_gamma=0.1, unary_gamma=0.1):
        self.p = p
        self.layers = RES_LAYERS[num_layers]
        self.n_shot = n_shot
        self.semantic_enhance = semantic_enhance
        self.channel_gamma, self.unary_gamma = channel_gamma, unary_gamma
        self._bn = {}
        self._pe = {}
        self._side = None
        self._side2 = None

    # ------------------------------------------------------------------ trunk
    def bn(self, name):
        """Frozen BatchNorm as an affine fold (dana.py:362-368): (scale, shift)."""
        v = self._bn.get(name)
        if v is None:
            p = self.p
            with torch.no_grad():
                s = p[name + ".weight"].float() / torch.sqrt(p[name + ".running_var"].float() + BN_EPS)
                t = p[name + ".bias"].float() - p[name + ".running_mean"].float() * s
            v = self._bn[name] = (s.contiguous(), t.contiguous())
        return v

    def bottleneck(self, x, name, stride):
        """resnet.py:66-102 (stride on the first 1x1)."""
        p = self.p
        s1, t1 = self.bn(name + ".bn1")
        s2, t2 = self.bn(name + ".bn2")
        s3, t3 = self.bn(name + ".bn3")
        y = A.conv(x, p[name + ".conv1.weight"], bias=t1, scale=s1, relu=True, ksize=1, stride=stride)
        y = A.conv(y, p[name + ".conv2.weight"], bias=t2, scale=s2, relu=True, ksize=3)
        if (name + ".downsample.0.weight") in p:
            sd, td = self.bn(name + ".downsample.1")
            res = A.conv(x, p[name + ".downsample.0.weight"], bias=td, scale=sd, ksize=1, stride=stride)
        else:
            res = x
        return A.conv(y, p[name + ".conv3.weight"], bias=t3, scale=s3, res=res, relu=True)

    def stage(self, x, prefix, blocks, stride):
        for i in range(blocks):
            x = self.bottleneck(x, "%s.%d" % (prefix, i), stride if i == 0 else 1)
        return x

    def trunk_tail(self, x):
        """layer2 + layer3 on the frozen layer1 output (NHWC fp32)."""
        x = self.stage(x, "RCNN_base.5", self.layers[1], 2)
        return self.stage(x, "RCNN_base.6", self.layers[2], 2)

    def head_to_tail(self, pooled):
        """RCNN_top + spatial mean (dana.py:387-389): [R,7,7,1024] -> [R,2048]."""
        return self.stage(pooled, "RCNN_top.0", self.layers[3], 2).mean(dim=(1, 2))

    # ------------------------------------------------------------------ attention
    def pe(self, n, c, device):
        key = (n, c, device)
        if key not in self._pe:
            self._pe[key] = positional_encoding(n, c).to(device)
        return self._pe[key]

    def lin(self, x, name, relu=False):
        return A.linear(x, self.p[name + ".weight"], self.p[name + ".bias"], relu=relu)

    def attend(self, qc, shots, prefix, d, enhance=None, repeat=1):
        """dana.py:126-150 / 268-281 for every shot: BA gate, k-projection, centred logits, softmax + unary term, P.V;
        the shot mean.  shots: [G0, Ns, C] each (PE applied); with repeat > 1 each support row serves `repeat`
        consecutive query rows (the head's per-RoI copy, dana.py:254-258): the support-only quantities are computed
        once per image and repeated, which is the same function."""
        outs = []
        for s in shots:
            if enhance is not None:                                           # BA block, dana.py:133-137
                w = F.softmax(self.lin(s, enhance), 1)                        # [G0, Ns, 1]
                g = (w * s).sum(1, keepdim=True)                              # bmm(w^T, s): a weighted sum over positions
                s = s + self.channel_gamma * F.leaky_relu(g)
            k = self.lin(s, prefix + "_adapt_k_layer")
            k = k - k.mean(1, keepdim=True)
            un = F.softmax(self.lin(s, prefix + "_unary_layer"), 1)           # [G0, Ns, 1]
            if repeat > 1:
                k, un, s = [t.repeat_interleave(repeat, dim=0) for t in (k, un, s)]
            att = F.softmax(A.bmm_nt(qc, k) / math.sqrt(d), dim=2)
            att = att + self.unary_gamma * un.transpose(1, 2)
            outs.append(A.bmm_nt(att, s.transpose(1, 2)))                     # att @ s
        return torch.stack(outs, 0).mean(0)

    def rpn_attention(self, base, sup_pos):
        """dana.py:117-151.  base [B,h,w,C]; sup_pos [B,K,hs,ws,C] -> dense [B,h,w,C]."""
        b, h, w, c = base.shape
        k, ns = sup_pos.shape[1], sup_pos.shape[2] * sup_pos.shape[3]
        q = self.lin(base.view(b, h * w, c), "rpn_adapt_q_layer")
        q = q - q.mean(1, keepdim=True)
        pe = self.pe(ns, c, base.device)
        shots = [sup_pos[:, i].reshape(b, ns, c) + pe for i in range(k)]
        dense = self.attend(q, shots, "rpn", 256, "rpn_channel_k_layer" if self.semantic_enhance else None)
        return dense.view(b, h, w, c)

    def rcnn_head(self, pooled, sup_pooled, per):
        """dana.py:247-290.  pooled [R,7,7,C]; sup_pooled [B,K,7,7,C] -> cls_score [R,2]."""
        r, c = pooled.shape[0], pooled.shape[-1]
        b, k = sup_pooled.shape[0], sup_pooled.shape[1]
        pe = self.pe(49, c, pooled.device)
        q_mat = pooled.reshape(r, 49, c) + pe
        q = self.lin(q_mat, "rcnn_adapt_q_layer")
        q = q - q.mean(1, keepdim=True)
        shots = [sup_pooled[:, i].reshape(b, 49, c) + pe for i in range(k)]
        dense = self.attend(q, shots, "rcnn", 256, None, repeat=per)
        corr = self.lin(torch.cat([q_mat, dense], 2), "rcnn_transform_layer")          # [R,49,64]
        hid = self.lin(corr.reshape(r, -1), "output_score_layer.linear1", relu=True)
        return self.lin(hid, "output_score_layer.linear2")

    # ------------------------------------------------------------------ losses (rpn.py:96-116, dana.py:199-215)
    @staticmethod
    def smooth_l1(pred, tgt, in_w, out_w, sigma, dims):
        s2 = sigma * sigma
        d = in_w * (pred - tgt)
        a = d.abs()
        small = (a < 1.0 / s2).float()
        loss = out_w * (d * d * (s2 / 2.0) * small + (a - 0.5 / s2) * (1.0 - small))
        return loss.sum(dim=dims).mean()

    def rpn_losses(self, rpn_raw, anchor_t, num_a):
        """rpn.py:96-116 with static shapes (no index_select of the kept anchors: a 0 / 1 weight per anchor instead, so
        that the step can be captured into a CUDA graph): mean cross entropy over the anchors labelled 0 / 1."""
        labels, tgt, in_w, out_w = anchor_t                 # (y, x, a) anchor order, flat per image
        b = rpn_raw.shape[0]
        raw = rpn_raw.view(b, -1, 6 * num_a)
        score = torch.stack([raw[..., :num_a].reshape(-1), raw[..., num_a:2 * num_a].reshape(-1)], 1)
        lab = labels.view(-1).long()
        keep = (lab != -1).float()
        ce = F.cross_entropy(score, lab.clamp_min(0), reduction="none")
        loss_cls = (ce * keep).sum() / keep.sum().clamp_min(1.0)
        deltas = raw[..., 2 * num_a:].reshape(b, -1, 4)
        loss_box = self.smooth_l1(deltas, tgt, in_w.unsqueeze(2), out_w.unsqueeze(2), 3.0, (1, 2))
        return loss_cls, loss_box

    @staticmethod
    def rcnn_cls_loss(scores, labels):
        """dana.py:203-214: all fg, the hardest bg of the positive-support half (2 x fg, at most a quarter of the rows)
        and of the negative-support half (<= fg) -- as a 0 / 1 weight per row computed on the device (ranks from two
        sorts), static shapes, no host read-back."""
        n = labels.shape[0]
        half = int(n * 0.5)
        idx = torch.arange(n, device=scores.device)
        fg = labels == 1
        bg = labels == 0
        nfg = fg.sum()
        bg0 = torch.clamp(torch.minimum(nfg * 2, torch.full_like(nfg, int(n * 0.25))), min=1)
        bg1 = torch.clamp(torch.minimum(nfg, bg0), min=1)
        p1 = F.softmax(scores.detach(), dim=1)[:, 1]
        neg_inf = torch.full_like(p1, float("-inf"))
        sel = fg
        for in_half, quota in ((idx < half, bg0), (idx >= half, bg1)):
            cand = bg & in_half
            order = torch.sort(torch.where(cand, p1, neg_inf), descending=True)[1]
            rank = torch.empty_like(order)
            rank[order] = idx
            sel = sel | (cand & (rank < quota))
        w = sel.float()
        ce = F.cross_entropy(scores, labels, reduction="none")
        return (ce * w).sum() / w.sum()

    # ------------------------------------------------------------------ forward
    def part1(self, stem, im_data, im_info, support_ims, base_anchors, feat_stride=16):
        """Trunks, RPN-level attention, RPN head and the proposal layer: everything up to the point where the host draws
        the training targets.  No host synchronisation (capturable into a CUDA graph).  -> state dict."""
        p, k = self.p, self.n_shot
        dev = im_data.device
        b = im_data.shape[0]
        num_a = base_anchors.shape[0]
        A.begin_step()
        # The support trunk runs on a side stream next to the query trunk: one 800x1333 query or ten 320x320 crops make
        # 33 pixel tiles per layer -- a launch fills a quarter to half of the 148 SMs -- so the two independent trunks
        # share the GPU.  autograd runs each node's backward on its forward's stream, so the two backward passes overlap
        # as well, and under capture they become parallel branches of the CUDA graph.
        # Shared by both trunks, so made BEFORE the fork (the cache would otherwise hand one stream planes that the other
        # stream's pack kernel is still writing): operand planes of every layer2 / layer3 weight, the BN folds.
        for name, w in p.items():
            if name.startswith(("RCNN_base.5.", "RCNN_base.6.")) and name.endswith(("conv1.weight", "conv2.weight",
                                                                                       "conv3.weight", "downsample.0.weight")):
                bn_name = name[:-len("conv1.weight")] + "bn" + name[-len("1.weight")] if "conv" in name else \
                    name[:-len("0.weight")] + "1"
                A._packed(w, self.bn(bn_name)[0], need_dgrad=True)
        cur = torch.cuda.current_stream()
        use_side = SIDE_STREAM and os.environ.get("DANA_TRAIN_SIDE_STREAM", "1") != "0"
        if use_side and self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side if use_side else cur
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            with torch.no_grad():
                s1 = stem(support_ims.reshape(-1, *support_ims.shape[2:]))
            sup = self.trunk_tail(s1)                                         # [B*2K,hs,ws,1024]
            hs, ws, c = sup.shape[1], sup.shape[2], sup.shape[3]
            if hs != ws:
                raise ValueError("support feature maps must be square (AvgPool2d of dana.py:42)")
            sup = sup.view(b, 2, k, hs, ws, c)                                # positive set, negative set (dana.py:100-108)
            sup_pooled = F.avg_pool2d(sup.reshape(-1, hs, ws, c).permute(0, 3, 1, 2), hs - 6, 1).permute(0, 2, 3, 1)
            sup_pooled = sup_pooled.reshape(b, 2, k, 7, 7, c)
        with torch.no_grad():
            q1 = stem(im_data)
        base = self.trunk_tail(q1)                                            # [B,h,w,1024]
        _, qh, qw, _ = base.shape
        cur.wait_stream(side)
        if use_side:
            sup.record_stream(cur)
            sup_pooled.record_stream(cur)

        dense = self.rpn_attention(base, sup[:, 0])
        corr = torch.cat([base, dense], 3)                                    # dana.py:154
        x = A.conv(corr, p["RCNN_rpn.RPN_Conv.weight"], bias=p["RCNN_rpn.RPN_Conv.bias"], relu=True, ksize=3)
        w_out = torch.cat([p["RCNN_rpn.RPN_cls_score.weight"], p["RCNN_rpn.RPN_bbox_pred.weight"]], 0)
        b_out = torch.cat([p["RCNN_rpn.RPN_cls_score.bias"], p["RCNN_rpn.RPN_bbox_pred.bias"]], 0)
        rpn_raw = A.conv(x, w_out, bias=b_out, ksize=1)                       # [B,h,w,6A]: bg | fg | deltas
        # ---- proposal layer: no gradient (rpn.py:74-78)
        with torch.no_grad():
            fg, deltas = ops.rpn_fg_prob(rpn_raw.detach().contiguous(), num_a)
            rois_all = ops.proposals(fg, deltas, base_anchors, im_info.to(dev).float(), qh, qw, feat_stride,
                                     cfg.TRAIN.RPN_PRE_NMS_TOP_N, cfg.TRAIN.RPN_POST_NMS_TOP_N, cfg.TRAIN.RPN_NMS_THRESH)
        return dict(base=base, sup_pooled=sup_pooled, rpn_raw=rpn_raw, rois_all=rois_all, qh=qh, qw=qw, num_a=num_a, b=b)

    @staticmethod
    def anchor_targets_host(qh, qw, gt_host, info_host, anchors_host, feat_stride=16):
        """anchor_target_layer.py:48-193 on the host (numpy RNG draws :131,:143)."""
        return targets.anchor_targets(
            qh, qw, gt_host, info_host, anchors_host, feat_stride,
            negative_overlap=cfg.TRAIN.RPN_NEGATIVE_OVERLAP, positive_overlap=cfg.TRAIN.RPN_POSITIVE_OVERLAP,
            clobber_positives=cfg.TRAIN.RPN_CLOBBER_POSITIVES, fg_fraction=cfg.TRAIN.RPN_FG_FRACTION,
            batchsize=cfg.TRAIN.RPN_BATCHSIZE, inside_weight=cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS[0],
            positive_weight=cfg.TRAIN.RPN_POSITIVE_WEIGHT)

    @staticmethod
    def proposal_targets_host(rois_all_host, gt_host):
        """proposal_target_layer_cascade.py:33-213 on the host (numpy RNG draws :159,:168,:175,:183)."""
        return targets.proposal_targets(
            rois_all_host, gt_host, rois_per_image=cfg.TRAIN.BATCH_SIZE, fg_fraction=cfg.TRAIN.FG_FRACTION,
            fg_thresh=cfg.TRAIN.FG_THRESH, bg_thresh_hi=cfg.TRAIN.BG_THRESH_HI, bg_thresh_lo=cfg.TRAIN.BG_THRESH_LO,
            normalize_means=cfg.TRAIN.BBOX_NORMALIZE_MEANS, normalize_stds=cfg.TRAIN.BBOX_NORMALIZE_STDS,
            inside_weights=cfg.TRAIN.BBOX_INSIDE_WEIGHTS, normalize_targets=cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED)

    def part2(self, st, anchor_t, sample):
        """Losses of the RPN, RoIAlign on the sampled RoIs, layer4 + box regressor, both head passes and the R-CNN
        losses.  anchor_t: (labels int8, bbox targets, inside w, outside w) device tensors; sample: (rois [B,R,5],
        labels [B,R], bbox targets, inside w, outside w) fp32 device tensors.  No host synchronisation."""
        b, num_a = st["b"], st["num_a"]
        rois, lab_s, tgt_s, inw_s, outw_s = sample
        rpn_loss_cls, rpn_loss_box = self.rpn_losses(st["rpn_raw"], anchor_t, num_a)
        per = rois.shape[1]
        r = b * per
        pooled = _RoIAlignNHWC.apply(st["base"], rois.view(-1, 5).contiguous())     # [R,7,7,1024]
        # Three independent branches on the pooled RoIs -- layer4 + box regressor, the positive-set head pass, the
        # negative-set head pass (dana.py:186-194) -- each a chain of small launches (128 RoIs x 49 bins): side by side on
        # three streams, forward and (autograd follows the forward's streams) backward.  The weights the two head passes
        # share are packed before the fork.
        cur = torch.cuda.current_stream()
        use_side = SIDE_STREAM and os.environ.get("DANA_TRAIN_SIDE_STREAM", "1") != "0"
        if use_side:
            for name in ("rcnn_adapt_q_layer", "rcnn_adapt_k_layer", "rcnn_transform_layer", "output_score_layer.linear1"):
                w = self.p[name + ".weight"]
                A._packed(w.view(w.shape[0], w.shape[1], 1, 1), None, need_dgrad=True)
            if self._side is None:
                self._side = torch.cuda.Stream(device=pooled.device)
            if self._side2 is None:
                self._side2 = torch.cuda.Stream(device=pooled.device)
            heads = []
            for stream, set_index in ((self._side, 0), (self._side2, 1)):
                stream.wait_stream(cur)
                with torch.cuda.stream(stream):
                    heads.append(self.rcnn_head(pooled, st["sup_pooled"][:, set_index], per))   # dana.py:189-194
            fc7 = self.head_to_tail(pooled)
            bbox_pred = self.lin(fc7, "RCNN_bbox_pred")
            for stream, t in ((self._side, heads[0]), (self._side2, heads[1])):
                cur.wait_stream(stream)
                t.record_stream(cur)
            sc_pos, sc_neg = heads
        else:
            fc7 = self.head_to_tail(pooled)
            bbox_pred = self.lin(fc7, "RCNN_bbox_pred")
            sc_pos = self.rcnn_head(pooled, st["sup_pooled"][:, 0], per)            # dana.py:189-194
            sc_neg = self.rcnn_head(pooled, st["sup_pooled"][:, 1], per)
        cls_score = torch.cat([sc_pos, sc_neg], 0)
        cls_prob = F.softmax(cls_score.detach(), 1)
        labels = lab_s.view(-1).long()
        rois_label = torch.cat([labels, torch.zeros_like(labels)], 0)         # dana.py:195-196
        loss_bbox = self.smooth_l1(bbox_pred, tgt_s.view(r, 4), inw_s.view(r, 4), outw_s.view(r, 4), 1.0, (1,))
        loss_cls = self.rcnn_cls_loss(cls_score, rois_label)
        return rois, cls_prob, bbox_pred.detach(), rpn_loss_cls, rpn_loss_box, loss_cls, loss_bbox, rois_label

    def forward(self, stem, im_data, im_info, gt_boxes, num_boxes, support_ims, base_anchors, feat_stride=16,
                teacher=None):
        """-> (rois [B,R,5], cls_prob [2BR,2], bbox_pred [BR,4], rpn_loss_cls, rpn_loss_box, RCNN_loss_cls,
        RCNN_loss_bbox, rois_label [2BR]) -- the reference's 8-tuple; the four losses carry the autograd graph."""
        dev = im_data.device
        self._bn.clear()
        st = self.part1(stem, im_data, im_info, support_ims, base_anchors, feat_stride)
        gt_host = gt_boxes.detach().float().cpu().numpy()
        info_host = im_info.detach().float().cpu().numpy()
        # anchor targets first, then proposal targets: the order of the reference's numpy RNG draws.  The anchor targets
        # do not depend on the network: they are drawn while the GPU is still working on part 1.
        anchor_t = self.anchor_targets_host(st["qh"], st["qw"], gt_host, info_host, base_anchors.cpu().numpy(), feat_stride)
        rois_all = st["rois_all"]
        if teacher and "rois" in teacher:
            rois_all = teacher["rois"].to(dev).float().contiguous()
        sample = self.proposal_targets_host(rois_all.cpu().numpy(), gt_host)
        anchor_t = [torch.from_numpy(np.ascontiguousarray(t)).to(dev) for t in anchor_t]
        sample = [torch.from_numpy(np.ascontiguousarray(t)).to(dev).float() for t in sample]
        return self.part2(st, anchor_t, sample)
