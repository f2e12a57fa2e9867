"""`DAnARCNN` -- the module boundary of the reference (lib/model/framework/dana.py:327-389): same
constructor, `create_architecture()`, parameter names / shapes (346-key state-dict, so reference
checkpoints load with `load_state_dict`), `train()` semantics (BatchNorm always in eval) and
`forward(im_data, im_info, gt_boxes, num_boxes, support_ims, all_cls_gt_boxes=None)` 8-tuple.

The forward runs on `engine.DanaEngine` (hand-written sm_100a kernels behind the C ABI); the nn.Module only owns
the parameters.  Train mode returns the reference's 8-tuple with the four losses (target layers on the host with the
reference's numpy RNG call order, loss kernels on the device); the BACKWARD pass of this path is not built
(SURVEY.md section 8 row a15): the losses carry no autograd graph."""
import math

import torch
import torch.nn as nn

from .config import cfg
from .engine import RES_LAYERS, DanaEngine


class _Bottleneck(nn.Module):
    """Caffe-style bottleneck: the stride sits on the first 1x1 (resnet.py:66-102)."""

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample


def _stage(inplanes, planes, blocks, stride):
    down = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(planes * 4))
    layers = [_Bottleneck(inplanes, planes, stride, down)]
    layers += [_Bottleneck(planes * 4, planes) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


class _FFN(nn.Module):
    def __init__(self, in_channel, hidden):
        super().__init__()
        self.linear1 = nn.Linear(in_channel, hidden)
        self.linear2 = nn.Linear(hidden, 2)
        self.relu = nn.ReLU()


class _RPNParams(nn.Module):
    """Parameter holder with the names of `_RPN` (lib/model/rpn/rpn.py:28-36)."""

    def __init__(self, din, num_anchors):
        super().__init__()
        self.RPN_Conv = nn.Conv2d(din, 512, 3, 1, 1, bias=True)
        self.RPN_cls_score = nn.Conv2d(512, 2 * num_anchors, 1, 1, 0)
        self.RPN_bbox_pred = nn.Conv2d(512, 4 * num_anchors, 1, 1, 0)


class DAnARCNN(nn.Module):
    def __init__(self, classes, attention_type, rpn_reduce_dim=256, rcnn_reduce_dim=256, gamma=0.1,
                 semantic_enhance=False, num_layers=50, pretrained=False, num_way=2, num_shot=5, pos_encoding=True,
                 precision="mixed", use_cuda_graph=False):
        super().__init__()
        if attention_type != "concat":
            raise NotImplementedError("only attention_type='concat' (the shipped configuration, utils.py:119) is built")
        if not pos_encoding:
            raise NotImplementedError("pos_encoding=False hits a typo in the reference (dana.py:130) and is not built")
        if rpn_reduce_dim != 256 or rcnn_reduce_dim != 256:
            raise NotImplementedError("reduce dims other than 256 are not built")
        self.model_path = "data/pretrained_model/resnet50_caffe.pth"
        self.dout_base_model = 1024
        self.pretrained = pretrained
        self.classes = classes
        self.n_classes = len(classes)
        self.n_way, self.n_shot = num_way, num_shot
        self.attention_type = attention_type
        self.channel_gamma, self.unary_gamma = gamma, 0.1
        self.semantic_enhance = semantic_enhance
        self.rpn_reduce_dim, self.rcnn_reduce_dim = rpn_reduce_dim, rcnn_reduce_dim
        # the reference ignores num_layers and always builds resnet50 (dana.py:337); 101 is the stated extension
        self.num_layers = num_layers if num_layers in RES_LAYERS else 50
        self.precision = precision
        # capture the eval forward into a CUDA graph per input shape (falls back to eager launches if capture fails)
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}
        dim_in = 1024

        def lin(i, o, std=0.01):
            m = nn.Linear(i, o)
            nn.init.normal_(m.weight, std=std)
            nn.init.constant_(m.bias, 0)
            return m
        self.rpn_unary_layer = lin(dim_in, 1)
        self.rcnn_unary_layer = lin(dim_in, 1)
        self.rpn_adapt_q_layer = lin(dim_in, rpn_reduce_dim)
        self.rpn_adapt_k_layer = lin(dim_in, rpn_reduce_dim)
        self.rcnn_adapt_q_layer = lin(dim_in, rcnn_reduce_dim)
        self.rcnn_adapt_k_layer = lin(dim_in, rcnn_reduce_dim)
        if semantic_enhance:
            self.rpn_channel_k_layer = lin(dim_in, 1)
        num_anchors = len(cfg.ANCHOR_SCALES) * len(cfg.ANCHOR_RATIOS)
        self.RCNN_rpn = _RPNParams(2048, num_anchors)
        self.rcnn_transform_layer = nn.Linear(2048, 64)
        self.output_score_layer = _FFN(64 * 49, dim_in)
        self._engine = None
        self._engine_key = None

    # ---- reference API --------------------------------------------------------------------
    def create_architecture(self):
        self._init_modules()
        self._init_weights()

    def _init_modules(self):
        l = RES_LAYERS[self.num_layers]
        conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.RCNN_base = nn.Sequential(conv1, nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                                       nn.MaxPool2d(3, 2, 0, ceil_mode=True), _stage(64, 64, l[0], 1),
                                       _stage(256, 128, l[1], 2), _stage(512, 256, l[2], 2))
        self.RCNN_top = nn.Sequential(_stage(1024, 512, l[3], 2))
        self.RCNN_bbox_pred = nn.Linear(2048, 4)
        for m in list(self.RCNN_base.modules()) + list(self.RCNN_top.modules()):
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        if self.pretrained:
            self._load_pretrained_resnet(self.model_path)
        # frozen parts (dana.py:350-368): conv1, bn1, the first FIXED_BLOCKS stages and every BatchNorm
        for p in self.RCNN_base[0].parameters():
            p.requires_grad = False
        for p in self.RCNN_base[1].parameters():
            p.requires_grad = False
        assert 0 <= cfg.RESNET.FIXED_BLOCKS < 4
        for idx, need in ((6, 3), (5, 2), (4, 1)):
            if cfg.RESNET.FIXED_BLOCKS >= need:
                for p in self.RCNN_base[idx].parameters():
                    p.requires_grad = False
        for m in list(self.RCNN_base.modules()) + list(self.RCNN_top.modules()):
            if isinstance(m, nn.BatchNorm2d):
                for p in m.parameters():
                    p.requires_grad = False

    # torchvision-style resnet key prefix -> position in RCNN_base / RCNN_top (dana.py:344-346)
    _RESNET_KEY_MAP = (("conv1.", "RCNN_base.0."), ("bn1.", "RCNN_base.1."), ("layer1.", "RCNN_base.4."),
                       ("layer2.", "RCNN_base.5."), ("layer3.", "RCNN_base.6."), ("layer4.", "RCNN_top.0."))

    def _load_pretrained_resnet(self, path):
        """dana.py:337-341: `resnet.load_state_dict({k: v for k, v in torch.load(model_path).items() if k in
        resnet.state_dict()})` -- the caffe-converted resnet checkpoint (conv1 / bn1 / layer1..4 keys; fc is dropped
        because the trunk has none).  Like the reference, keys the trunk does not have are ignored and a key of the
        trunk that the file lacks is an error (strict load of the filtered dict)."""
        print("Loading pretrained weights from %s" % path)
        state = torch.load(path, map_location="cpu")
        own = dict(self.RCNN_base.state_dict(prefix="RCNN_base."))
        own.update(self.RCNN_top.state_dict(prefix="RCNN_top."))
        mapped = {}
        for k, v in state.items():
            for src, dst in self._RESNET_KEY_MAP:
                if k.startswith(src):
                    name = dst + k[len(src):]
                    if name in own:
                        mapped[name] = v
                    break
        missing = [k for k in own if k not in mapped and not k.endswith("num_batches_tracked")]
        if missing:
            raise RuntimeError("pretrained resnet checkpoint %s lacks %d trunk tensors (first: %s)"
                               % (path, len(missing), missing[0]))
        for name, v in mapped.items():
            if tuple(own[name].shape) != tuple(v.shape):
                raise RuntimeError("size mismatch for %s: checkpoint %s vs model %s"
                                   % (name, tuple(v.shape), tuple(own[name].shape)))
            own[name].copy_(v)

    def _init_weights(self):
        def normal_init(m, mean, stddev, truncated=False):
            if truncated:
                m.weight.data.normal_().fmod_(2).mul_(stddev).add_(mean)
            else:
                m.weight.data.normal_(mean, stddev)
                m.bias.data.zero_()
        normal_init(self.RCNN_rpn.RPN_Conv, 0, 0.01, cfg.TRAIN.TRUNCATED)
        normal_init(self.RCNN_rpn.RPN_cls_score, 0, 0.01, cfg.TRAIN.TRUNCATED)
        normal_init(self.RCNN_rpn.RPN_bbox_pred, 0, 0.01, cfg.TRAIN.TRUNCATED)
        normal_init(self.RCNN_bbox_pred, 0, 0.001, cfg.TRAIN.TRUNCATED)

    def train(self, mode=True):
        nn.Module.train(self, mode)
        if mode and hasattr(self, "RCNN_base"):
            self.RCNN_base.eval()
            self.RCNN_base[5].train()
            self.RCNN_base[6].train()
            for m in list(self.RCNN_base.modules()) + list(self.RCNN_top.modules()):
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    # ---- engine cache ---------------------------------------------------------------------
    def _param_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in self.buffers())

    def engine(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("DAnARCNN (dana_b200) runs on CUDA only: call .cuda() first (there is no CPU fallback)")
        key = (dev, self._param_version(), tuple(cfg.ANCHOR_SCALES), tuple(cfg.ANCHOR_RATIOS), cfg.FEAT_STRIDE[0])
        if self._engine is None or key != self._engine_key:
            self._engine = DanaEngine(self.state_dict(), device=dev, num_layers=self.num_layers, n_shot=self.n_shot,
                                      semantic_enhance=self.semantic_enhance, channel_gamma=self.channel_gamma,
                                      unary_gamma=self.unary_gamma, precision=self.precision,
                                      anchor_scales=tuple(cfg.ANCHOR_SCALES), anchor_ratios=tuple(cfg.ANCHOR_RATIOS),
                                      feat_stride=cfg.FEAT_STRIDE[0])
            self._engine_key = key
            # graphs captured against the previous engine replay kernels that read ITS weight buffers: drop them
            # (also releases their private memory pools)
            self._graphs.clear()
        return self._engine

    def _forward_train(self, im_data, im_info, gt_boxes, num_boxes, support_ims, teacher=None):
        """Training branch of _DAnARCNN.forward (dana.py:100-108,158-215): the forward and the four losses.
        support_ims [B, 2*K, 3, H, W]: K positive then K negative crops per image (n_way = 2).  Anchor / proposal
        targets are drawn on the host with numpy's global RNG exactly like the reference (dana_b200.targets); the
        losses are device kernels.  Returns the reference's 8-tuple; the losses are 0-dim CUDA tensors WITHOUT a
        graph (the torch.no_grad() variant: validation losses); with gradients enabled forward() takes
        _forward_train_graph instead."""
        import numpy as np

        from . import ops, targets
        if self.n_way != 2:
            raise NotImplementedError("the training branch is built for n_way = 2 (positive + negative support set), "
                                      "like the reference's (dana.py:100-108)")
        if support_ims.shape[1] != 2 * self.n_shot:
            raise ValueError("train mode expects %d support crops per image (n_way * n_shot), got %d"
                             % (2 * self.n_shot, support_ims.shape[1]))
        eng = self.engine()
        dev = im_data.device
        b = im_data.shape[0]
        qh, qw = eng._trunk_hw(im_data.shape[2], im_data.shape[3])
        gt_host = gt_boxes.detach().float().cpu().numpy()
        info_host = im_info.detach().float().cpu().numpy()
        state = {}

        def hook(rois):
            # anchor targets first, then proposal targets: the order of the reference's numpy RNG draws
            # (RCNN_rpn.forward -> anchor_target_layer, then RCNN_proposal_target; dana.py:158-170)
            state["anchor"] = targets.anchor_targets(
                qh, qw, gt_host, info_host, eng.base_anchors.cpu().numpy(), eng.feat_stride,
                negative_overlap=cfg.TRAIN.RPN_NEGATIVE_OVERLAP, positive_overlap=cfg.TRAIN.RPN_POSITIVE_OVERLAP,
                clobber_positives=cfg.TRAIN.RPN_CLOBBER_POSITIVES, fg_fraction=cfg.TRAIN.RPN_FG_FRACTION,
                batchsize=cfg.TRAIN.RPN_BATCHSIZE, inside_weight=cfg.TRAIN.RPN_BBOX_INSIDE_WEIGHTS[0],
                positive_weight=cfg.TRAIN.RPN_POSITIVE_WEIGHT)
            state["all_rois"] = rois
            sample = targets.proposal_targets(
                rois.detach().cpu().numpy(), gt_host, rois_per_image=cfg.TRAIN.BATCH_SIZE,
                fg_fraction=cfg.TRAIN.FG_FRACTION, fg_thresh=cfg.TRAIN.FG_THRESH, bg_thresh_hi=cfg.TRAIN.BG_THRESH_HI,
                bg_thresh_lo=cfg.TRAIN.BG_THRESH_LO, normalize_means=cfg.TRAIN.BBOX_NORMALIZE_MEANS,
                normalize_stds=cfg.TRAIN.BBOX_NORMALIZE_STDS, inside_weights=cfg.TRAIN.BBOX_INSIDE_WEIGHTS,
                normalize_targets=cfg.TRAIN.BBOX_NORMALIZE_TARGETS_PRECOMPUTED)
            state["sample"] = sample
            return torch.from_numpy(sample[0])

        rois, cls_prob, bbox_pred, ex = eng.forward(
            im_data, im_info.data, support_ims, pre_nms_top_n=cfg.TRAIN.RPN_PRE_NMS_TOP_N,
            post_nms_top_n=cfg.TRAIN.RPN_POST_NMS_TOP_N, nms_thresh=cfg.TRAIN.RPN_NMS_THRESH,
            pooling_size=cfg.POOLING_SIZE, want=("rpn_raw", "cls_score"), rois_hook=hook, teacher=teacher)
        labels, tgt, in_w, out_w = [torch.from_numpy(np.ascontiguousarray(t)).to(dev) for t in state["anchor"]]
        rpn = ops.rpn_losses(ex["rpn_raw"], labels, tgt, in_w, out_w, eng.num_a)
        _, lab_s, tgt_s, inw_s, outw_s = [torch.from_numpy(np.ascontiguousarray(t)).to(dev) for t in state["sample"]]
        r = b * lab_s.shape[1]
        rcnn = ops.rcnn_losses(ex["cls_score"], lab_s.view(-1), bbox_pred, tgt_s.view(r, 4), inw_s.view(r, 4),
                               outw_s.view(r, 4))
        rois_label = torch.cat([lab_s.view(-1), torch.zeros_like(lab_s.view(-1))]).long()      # dana.py:195-196
        return rois, cls_prob, bbox_pred, rpn[0], rpn[1], rcnn[0], rcnn[1], rois_label

    def _forward_train_graph(self, im_data, im_info, gt_boxes, num_boxes, support_ims, teacher=None):
        """Training branch with the autograd graph (train.py:129-139: the four losses are summed and `.backward()`ed):
        dana_b200.train_model.TrainGraph over this module's own parameters -- forward, data- and weight-gradients of
        every trainable convolution / projection / attention product on the tcgen05 GEMM."""
        from .anchors import generate_anchors
        from .train_model import FrozenStem, TrainGraph
        if self.n_way != 2:
            raise NotImplementedError("the training branch is built for n_way = 2 (positive + negative support set), "
                                      "like the reference's (dana.py:100-108)")
        if support_ims.shape[1] != 2 * self.n_shot:
            raise ValueError("train mode expects %d support crops per image (n_way * n_shot), got %d"
                             % (2 * self.n_shot, support_ims.shape[1]))
        if cfg.RESNET.FIXED_BLOCKS != 1:
            raise NotImplementedError("the training graph is built for RESNET.FIXED_BLOCKS = 1 (config.py:223)")
        dev = im_data.device
        if dev.type != "cuda":
            raise RuntimeError("DAnARCNN (dana_b200) runs on CUDA only: call .cuda() first (there is no CPU fallback)")
        named = dict(self.named_parameters())
        named.update(dict(self.named_buffers()))
        frozen = [t for n, t in named.items() if n.startswith(("RCNN_base.0.", "RCNN_base.1.", "RCNN_base.4."))]
        key = tuple((t.data_ptr(), t._version) for t in frozen)
        if getattr(self, "_stem_key", None) != key:
            self._stem, self._stem_key = FrozenStem(named, dev), key
        anchors = torch.from_numpy(generate_anchors(ratios=tuple(cfg.ANCHOR_RATIOS), scales=tuple(cfg.ANCHOR_SCALES)))
        graph = TrainGraph(named, num_layers=self.num_layers, n_shot=self.n_shot, semantic_enhance=self.semantic_enhance,
                           channel_gamma=self.channel_gamma, unary_gamma=self.unary_gamma)
        return graph.forward(self._stem, im_data, im_info.data, gt_boxes, num_boxes, support_ims,
                             anchors.float().to(dev), cfg.FEAT_STRIDE[0], teacher=teacher)

    def forward(self, im_data, im_info, gt_boxes, num_boxes, support_ims, all_cls_gt_boxes=None):
        if cfg.POOLING_MODE != "align":
            raise NotImplementedError("POOLING_MODE %r is not built (shipped configs use 'align')" % cfg.POOLING_MODE)
        if self.training:
            if torch.is_grad_enabled():
                return self._forward_train_graph(im_data, im_info, gt_boxes, num_boxes, support_ims)
            return self._forward_train(im_data, im_info, gt_boxes, num_boxes, support_ims)
        if cfg.POOLING_MODE != "align":
            raise NotImplementedError("POOLING_MODE %r is not built (shipped configs use 'align')" % cfg.POOLING_MODE)
        eng = self.engine()
        kw = dict(pre_nms_top_n=cfg.TEST.RPN_PRE_NMS_TOP_N, post_nms_top_n=cfg.TEST.RPN_POST_NMS_TOP_N,
                  nms_thresh=cfg.TEST.RPN_NMS_THRESH, pooling_size=cfg.POOLING_SIZE)
        rois = None
        if self.use_cuda_graph:
            key = (tuple(im_data.shape), tuple(support_ims.shape), tuple(sorted(kw.items())))
            g = self._graphs.get(key)
            if g is not None and g is not False and g.engine is not eng:   # belt and braces: never replay a stale capture
                g = None
            if g is None:
                try:
                    from .engine import GraphedForward
                    g = GraphedForward(eng, im_data, im_info.data, support_ims, **kw)
                except Exception as e:  # noqa: BLE001  -- capture is an optimisation, never a requirement
                    import warnings
                    warnings.warn("dana_b200: CUDA graph capture failed (%r); running eagerly" % (e,))
                    g = False
                self._graphs[key] = g
            if g:
                rois, cls_prob, bbox_pred = g(im_data, im_info.data, support_ims)
        if rois is None:
            rois, cls_prob, bbox_pred = eng.forward(im_data, im_info.data, support_ims, **kw)
        # eval: losses are python 0 and rois_label is None (dana.py:173-179,216-220)
        return rois, cls_prob, bbox_pred, 0, 0, 0, 0, None
