"""Synthetic weights / episodes for benchmarks and smoke runs (there is no network for checkpoints or
datasets).  Weights follow the reference's init distributions (conv N(0, sqrt(2/n)), resnet.py:123-129;
linear / RPN N(0, .01), bbox N(0, .001), dana.py:46-69,234-238) with randomised frozen-BN statistics,
damped so that activations stay O(1) through the 16 residual blocks on pixel-scale inputs -- the RPN
scores then spread over (0, 1) instead of saturating, which keeps the proposal/NMS stage realistic."""
import math

import torch

from .engine import RES_LAYERS


def _bn(sd, g, name, ch, gamma=(0.8, 1.2), var_scale=1.0):
    sd[name + ".weight"] = torch.empty(ch).uniform_(gamma[0], gamma[1], generator=g)
    sd[name + ".bias"] = torch.randn(ch, generator=g) * 0.05
    sd[name + ".running_mean"] = torch.randn(ch, generator=g) * 0.1
    sd[name + ".running_var"] = torch.empty(ch).uniform_(0.5, 1.5, generator=g) * var_scale


def _conv(g, co, ci, k):
    return torch.randn(co, ci, k, k, generator=g) * math.sqrt(2.0 / (k * k * co))


def _stage(sd, g, prefix, inplanes, planes, blocks):
    for i in range(blocks):
        n = "%s.%d" % (prefix, i)
        cin = inplanes if i == 0 else planes * 4
        sd[n + ".conv1.weight"] = _conv(g, planes, cin, 1)
        _bn(sd, g, n + ".bn1", planes)
        sd[n + ".conv2.weight"] = _conv(g, planes, planes, 3)
        _bn(sd, g, n + ".bn2", planes)
        sd[n + ".conv3.weight"] = _conv(g, planes * 4, planes, 1)
        _bn(sd, g, n + ".bn3", planes * 4, gamma=(0.3, 0.5))
        if i == 0:
            sd[n + ".downsample.0.weight"] = _conv(g, planes * 4, cin, 1)
            _bn(sd, g, n + ".downsample.1", planes * 4)


def synthetic_state_dict(seed=1996, num_layers=50, num_anchors=12, semantic_enhance=True):
    """State-dict with the reference's key names and shapes."""
    g = torch.Generator().manual_seed(seed)
    l = RES_LAYERS[num_layers]
    sd = {"RCNN_base.0.weight": _conv(g, 64, 3, 7)}
    _bn(sd, g, "RCNN_base.1", 64, var_scale=230.0)          # pixel-scale input (std 50) -> O(1) stem output
    _stage(sd, g, "RCNN_base.4", 64, 64, l[0])
    _stage(sd, g, "RCNN_base.5", 256, 128, l[1])
    _stage(sd, g, "RCNN_base.6", 512, 256, l[2])
    _stage(sd, g, "RCNN_top.0", 1024, 512, l[3])

    def lin(name, o, i, std=0.01):
        sd[name + ".weight"] = torch.randn(o, i, generator=g) * std
        sd[name + ".bias"] = torch.zeros(o)
    for pre in ("rpn", "rcnn"):
        lin(pre + "_unary_layer", 1, 1024)
        lin(pre + "_adapt_q_layer", 256, 1024)
        lin(pre + "_adapt_k_layer", 256, 1024)
    if semantic_enhance:
        lin("rpn_channel_k_layer", 1, 1024)
    sd["RCNN_rpn.RPN_Conv.weight"] = torch.randn(512, 2048, 3, 3, generator=g) * 0.01
    sd["RCNN_rpn.RPN_Conv.bias"] = torch.zeros(512)
    sd["RCNN_rpn.RPN_cls_score.weight"] = torch.randn(2 * num_anchors, 512, 1, 1, generator=g) * 0.01
    sd["RCNN_rpn.RPN_cls_score.bias"] = torch.zeros(2 * num_anchors)
    sd["RCNN_rpn.RPN_bbox_pred.weight"] = torch.randn(4 * num_anchors, 512, 1, 1, generator=g) * 0.01
    sd["RCNN_rpn.RPN_bbox_pred.bias"] = torch.zeros(4 * num_anchors)
    lin("rcnn_transform_layer", 64, 2048)
    lin("output_score_layer.linear1", 1024, 3136)
    lin("output_score_layer.linear2", 2, 1024)
    lin("RCNN_bbox_pred", 4, 2048, std=0.001)
    return sd


def synthetic_episode(seed, batch, height=600, width=1000, n_support=6, support_size=320, pin=False):
    """(im_data [B,3,H,W], im_info [B,3], support_ims [B,n_support,3,S,S]) on the host: N(0,1)*50 pixels
    (mean-subtracted BGR scale, config.py:258), im_info = (H, W, 1)."""
    g = torch.Generator().manual_seed(seed)
    im = torch.randn(batch, 3, height, width, generator=g) * 50
    sup = torch.randn(batch, n_support, 3, support_size, support_size, generator=g) * 50
    info = torch.tensor([[float(height), float(width), 1.0]] * batch)
    if pin:
        im, sup, info = im.pin_memory(), sup.pin_memory(), info.pin_memory()
    return im, info, sup
