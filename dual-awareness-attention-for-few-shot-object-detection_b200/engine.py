"""Forward engine of the DAnA hot path on B200: packs a reference-compatible state-dict into
tensor-core friendly device buffers (NHWC, bf16 hi/lo planes, folded BN) and sequences the C-ABI
kernels for `_DAnARCNN.forward` in eval mode (lib/model/framework/dana.py:87-220).

Python here is host orchestration only: shapes, buffer allocation (torch caching allocator) and
kernel order.  Every FLOP runs in libdana_b200.so."""
import math
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .anchors import generate_anchors
from .ops import Pair

BN_EPS = 1e-5
RES_LAYERS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}


def positional_encoding(max_len, d_model=1024):
    """Sinusoid over the flattened position index (dana.py:311-320); host-side constant table."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0., max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0., d_model, 2) * -(math.log(10000.0) / float(d_model)))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


class _Conv:
    """One conv (+ frozen BN) packed for the implicit-GEMM kernel: weight [co, kh*kw*ci] (tap-major)."""

    def __init__(self, w, bn, device, split, conv_bias=None, f16=False):
        co, ci, kh, kw = w.shape
        self.n_out, self.ksize = co, kh
        wk = w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous().to(device)
        self.w = Pair.from_float_f16(wk) if f16 else Pair.from_float(wk, split)
        if bn is not None:
            g, b, m, v = [t.to(device=device, dtype=torch.float32) for t in bn]
            self.scale = (g / torch.sqrt(v + BN_EPS)).contiguous()
            self.bias = (b - m * self.scale).contiguous()
        else:
            self.scale = None
            self.bias = None if conv_bias is None else conv_bias.to(device=device, dtype=torch.float32).contiguous()


class _Block:
    def __init__(self, sd, name, device, split, f16=False):
        def bn(n):
            return (sd[n + ".weight"], sd[n + ".bias"], sd[n + ".running_mean"], sd[n + ".running_var"])
        self.c1 = _Conv(sd[name + ".conv1.weight"], bn(name + ".bn1"), device, split, f16=f16)
        self.c2 = _Conv(sd[name + ".conv2.weight"], bn(name + ".bn2"), device, split, f16=f16)
        self.c3 = _Conv(sd[name + ".conv3.weight"], bn(name + ".bn3"), device, split, f16=f16)
        self.down = None
        if (name + ".downsample.0.weight") in sd:
            self.down = _Conv(sd[name + ".downsample.0.weight"], bn(name + ".downsample.1"), device, split, f16=f16)


class GraphedForward:
    """One eval forward of a DanaEngine captured into a CUDA graph (streams and graphs instead of a tracing
    compiler): ~150 kernel launches become one graph launch, which removes the host launch pacing that leaves
    the GPU idle between the many small kernels of the support / head stages.  Inputs are copied into static
    buffers, outputs are cloned out of the graph's private pool."""

    def __init__(self, engine, im_data, im_info, support_ims, **kw):
        self.engine = engine          # the capture is only valid for this engine's buffers (keeps them alive, too)
        self.static_in = [im_data.detach().clone(), im_info.detach().float().clone(), support_ims.detach().clone()]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):          # warm-up outside capture: lazy workspaces, kernel attributes
            for _ in range(2):
                engine.forward(*self.static_in, **kw)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = engine.forward(*self.static_in, **kw)

    def __call__(self, im_data, im_info, support_ims):
        for dst, src in zip(self.static_in, (im_data, im_info, support_ims)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return tuple(t.clone() for t in self.static_out)


class DanaEngine:
    """precision:
      'mixed'   bf16x3 everywhere except the tensor-bound tail -- the RPN 3x3 conv + its 1x1 heads and layer4 run on
                single fp16 planes (11 significant bits per operand, one MMA per product, fp32 accumulation); every
                stage stays within the 1e-3 north-star tolerance (measured 2e-4 .. 6e-4 on those stages)
      'bf16x3'  hi/lo bf16 operands everywhere: fp32-equivalent products (2e-5 .. 8e-5 on every stage)
      'bf16'    plain bf16 operands: throughput mode, does not meet 1e-3."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", num_layers=50, n_shot=3,
                 semantic_enhance=True, channel_gamma=0.1, unary_gamma=0.1, precision="bf16x3",
                 anchor_scales=(4, 8, 16, 32), anchor_ratios=(0.5, 1, 2), feat_stride=16):
        assert precision in ("mixed", "bf16x3", "bf16")
        self.device = torch.device(device)
        self.split = precision in ("bf16x3", "mixed")
        self.f16 = precision == "mixed"
        self.precision = precision
        self.num_layers = num_layers
        self.n_shot = n_shot
        self.semantic_enhance = semantic_enhance
        self.channel_gamma, self.unary_gamma = channel_gamma, unary_gamma
        self.feat_stride = feat_stride
        self.base_anchors = torch.from_numpy(
            generate_anchors(ratios=anchor_ratios, scales=anchor_scales)).float().to(self.device)
        self.num_a = self.base_anchors.shape[0]
        self._pe = {}
        self._prop_ws = None
        self.stage_events = None      # bench.py: list receiving (stage name, CUDA event) at the stage boundaries
        self._side = None             # side stream of forward()'s support-side branch
        self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        dev, split = self.device, self.split
        sd = {k: v.detach().float() for k, v in sd.items()}
        f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self.stem_w = ops.pack_stem_weight(f32(sd["RCNN_base.0.weight"]), split)
        g, b, m, v = [f32(sd["RCNN_base.1." + n]) for n in ("weight", "bias", "running_mean", "running_var")]
        self.stem_scale = (g / torch.sqrt(v + BN_EPS)).contiguous()
        self.stem_bias = (b - m * self.stem_scale).contiguous()
        layers = RES_LAYERS[self.num_layers]
        self.stages = []
        for prefix, blocks in (("RCNN_base.4", layers[0]), ("RCNN_base.5", layers[1]), ("RCNN_base.6", layers[2])):
            self.stages.append([_Block(sd, "%s.%d" % (prefix, i), dev, split) for i in range(blocks)])
        self.top = [_Block(sd, "RCNN_top.0.%d" % i, dev, split, f16=self.f16) for i in range(layers[3])]

        def lin(name):
            return Pair.from_float(f32(sd[name + ".weight"]), split), f32(sd[name + ".bias"])
        self.rpn_q_w, _ = lin("rpn_adapt_q_layer")      # biases cancel under the mean-centering (dana.py:125,141)
        self.rpn_k_w, _ = lin("rpn_adapt_k_layer")
        self.rcnn_q_w, _ = lin("rcnn_adapt_q_layer")
        self.rcnn_k_w, _ = lin("rcnn_adapt_k_layer")
        self.rpn_un_w, self.rpn_un_b = f32(sd["rpn_unary_layer.weight"]).view(-1), f32(sd["rpn_unary_layer.bias"])
        self.rcnn_un_w, self.rcnn_un_b = f32(sd["rcnn_unary_layer.weight"]).view(-1), f32(sd["rcnn_unary_layer.bias"])
        if self.semantic_enhance:
            self.ba_w, self.ba_b = f32(sd["rpn_channel_k_layer.weight"]).view(-1), f32(sd["rpn_channel_k_layer.bias"])
        else:
            self.ba_w = self.ba_b = None
        self.rpn_conv = _Conv(sd["RCNN_rpn.RPN_Conv.weight"], None, dev, split, sd["RCNN_rpn.RPN_Conv.bias"], f16=self.f16)
        w_cb = torch.cat([sd["RCNN_rpn.RPN_cls_score.weight"], sd["RCNN_rpn.RPN_bbox_pred.weight"]], 0)
        b_cb = torch.cat([sd["RCNN_rpn.RPN_cls_score.bias"], sd["RCNN_rpn.RPN_bbox_pred.bias"]], 0)
        assert w_cb.shape[0] == 6 * self.num_a, "RPN head does not match the anchor configuration"
        self.rpn_out = _Conv(w_cb, None, dev, split, b_cb, f16=self.f16)
        wt = f32(sd["rcnn_transform_layer.weight"])                       # [64, 2048] on cat[query, dense]
        self.tr_wq = Pair.from_float(wt[:, :1024].contiguous(), split)
        self.tr_wd = Pair.from_float(wt[:, 1024:].contiguous(), split)
        self.tr_b = f32(sd["rcnn_transform_layer.bias"])
        # positional encoding of the head's query side (dana.py:259) folded through the two projections that read it:
        # (x + PE) W^T = x W^T + PE W^T, a [49, n_out] per-bin bias (built once in fp64), so the RoIAlign output
        # is written once instead of once plain and once with PE added
        pe49 = positional_encoding(49).double()
        self.pe_q = (pe49 @ sd["rcnn_adapt_q_layer.weight"].double().cpu().t()).float().to(dev).contiguous()
        self.pe_t = (pe49 @ wt[:, :1024].double().cpu().t() + self.tr_b.double().cpu()).float().to(dev).contiguous()
        # both projections read the same [R*49, 1024] pooled rows: ONE GEMM with the 256 + 64 output columns side by side
        self.head_qt_w = Pair.from_float(torch.cat([f32(sd["rcnn_adapt_q_layer.weight"]), wt[:, :1024]], 0).contiguous(), split)
        self.pe_qt = torch.cat([self.pe_q, self.pe_t], 1).contiguous()
        self.ffn1_w, self.ffn1_b = lin("output_score_layer.linear1")
        self.ffn2_w, self.ffn2_b = lin("output_score_layer.linear2")
        self.bbox_w, self.bbox_b = lin("RCNN_bbox_pred")

    # stage boundaries of forward(), in order; _mark(name) closes stage `name`
    STAGES = ("trunk", "cisa_rpn", "rpn_proposals", "roi_align", "layer4_bbox", "head_cisa")

    def _mark(self, name):
        if self.stage_events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.stage_events.append((name, ev))
            order = ("start",) + self.STAGES
            i = order.index(name)
            ops.STAGE = order[i + 1] if i + 1 < len(order) else ""

    def pe(self, n):
        if n not in self._pe:
            self._pe[n] = positional_encoding(n).to(self.device).contiguous()
        return self._pe[n]

    # ------------------------------------------------------------------ trunk
    def _conv(self, x, c: _Conv, stride=1, relu=True, res=None, out=None):
        return ops.conv_nhwc(x, c.w, c.n_out, ksize=c.ksize, stride=stride, scale=c.scale, bias=c.bias, res=res,
                             relu=relu, out=out, split=self.split, out_f16=x.is_f16)

    def _bottleneck(self, x, blk: _Block, stride, out=None):
        y = self._conv(x, blk.c1, stride=stride)
        y = self._conv(y, blk.c2)
        res = self._conv(x, blk.down, stride=stride, relu=False) if blk.down is not None else x
        return self._conv(y, blk.c3, res=res, out=out)

    def trunk(self, im_nchw, out=None):
        """RCNN_base (dana.py:344-345).  NCHW fp32 image batch -> NHWC pair, stride 16, 1024 channels.
        `out` optionally receives the last block's output (e.g. a channel slice of the RPN input).
        DANA_TRUNK_CHUNK=n (experiment): depth-first over chunks of n images, so that a chunk's layer1/2
        activations (38 MB per 600x1000 image at 256 channels) stay L2-resident between consecutive layers."""
        n = im_nchw.shape[0]
        nstreams = int(os.environ.get("DANA_TRUNK_STREAMS", "1"))
        if nstreams > 1 and n >= nstreams and self.stage_events is None and ops.GEMM_TRACE is None:
            # experiment: the batch in `nstreams` chunks on as many streams (tail waves of one chunk's layer overlap the
            # other chunk's launches)
            if out is None:
                qh, qw = self._trunk_hw(im_nchw.shape[2], im_nchw.shape[3])
                out = Pair.empty((n, qh, qw, 1024), self.device, self.split)
            cur = torch.cuda.current_stream()
            if not hasattr(self, "_tstreams"):
                self._tstreams = {}
            key = cur.cuda_stream
            if key not in self._tstreams:
                self._tstreams[key] = [torch.cuda.Stream(device=self.device) for _ in range(nstreams - 1)]
            per = (n + nstreams - 1) // nstreams
            for j in range(nstreams):
                i0, i1 = j * per, min(n, (j + 1) * per)
                if i0 >= i1:
                    break
                if j == 0:
                    self._trunk(im_nchw[i0:i1], out=out[i0:i1])
                else:
                    st = self._tstreams[key][j - 1]
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        self._trunk(im_nchw[i0:i1], out=out[i0:i1])
            for st in self._tstreams[key]:
                cur.wait_stream(st)
            return out
        chunk = int(os.environ.get("DANA_TRUNK_CHUNK", "0"))
        if chunk > 0 and n > chunk:
            if out is None:
                qh, qw = self._trunk_hw(im_nchw.shape[2], im_nchw.shape[3])
                out = Pair.empty((n, qh, qw, 1024), self.device, self.split)
            for i0 in range(0, n, chunk):
                self._trunk(im_nchw[i0:i0 + chunk], out=out[i0:i0 + chunk])
            return out
        return self._trunk(im_nchw, out=out)

    def _trunk(self, im_nchw, out=None):
        x = ops.stem(im_nchw, self.stem_w, self.stem_scale, self.stem_bias, split=self.split)
        for si, blocks in enumerate(self.stages):
            for bi, blk in enumerate(blocks):
                last = (si == len(self.stages) - 1) and (bi == len(blocks) - 1)
                x = self._bottleneck(x, blk, 2 if (bi == 0 and si > 0) else 1, out=out if last else None)
        return x

    def layer4(self, pooled: Pair):
        """RCNN_top (dana.py:346): [R,7,7,1024] -> [R,4,4,2048]."""
        x = pooled
        for bi, blk in enumerate(self.top):
            x = self._bottleneck(x, blk, 2 if bi == 0 else 1)
        return x

    # ------------------------------------------------------------------ attention
    FUSED_SOFTMAX_MAX_NS = 256 if os.environ.get("DANA_WIDE_SOFTMAX") == "0" else 512

    @classmethod
    def seg_pitch(cls, ns):
        """Column pitch of one shot's key segment in P / V^T: 8-aligned where the fused softmax epilogue is used
        (up to 256 keys: one 256-column tile; 257..512 keys: the wide single-accumulator tile)."""
        return (ns + 7) // 8 * 8 if ns <= cls.FUSED_SOFTMAX_MAX_NS else ns

    def _attention(self, qc: Pair, kc_all: Pair, vt_all: Pair, rbar_all, set_index, sets, batch, ns, out: Pair,
                   res_f32=None):
        """CISA contractions for one support set (dana.py:142-150 / :273-281).
        qc [batch*rows, 256] centred queries; kc_all [batch*sets*K*ns, 256]; vt_all [batch*sets, C, pitch];
        rbar_all [batch*sets, C].  Writes the attended feature into `out` [batch*rows, C] (any row pitch).
        ns <= 512: the logits GEMM carries the softmax in its epilogue (no fp32 logits in HBM);
        larger segments: logits GEMM -> attn_softmax kernel."""
        k = self.n_shot
        d = qc.hi.shape[1]
        c = vt_all.hi.shape[1]
        pitch = vt_all.hi.shape[2]
        sp = self.seg_pitch(ns)
        kn = k * ns
        rows_total = qc.hi.shape[0]
        kc = kc_all[set_index * kn:]
        if ns <= self.FUSED_SOFTMAX_MAX_NS:
            p = Pair.empty((rows_total, pitch), self.device, self.split)
            if pitch > k * sp:                       # row-pitch padding beyond the last segment must be finite
                p.hi[:, k * sp:].zero_()
                if p.lo is not None:
                    p.lo[:, k * sp:].zero_()
            ops.linear(qc, kc, kn, alpha=1.0 / math.sqrt(d), out=p, batch=batch, b_batch_stride=sets * kn * d,
                       softmax_ns=ns, softmax_pitch=sp)
        else:
            logits = torch.empty((rows_total, pitch), dtype=torch.float32, device=self.device)
            ops.linear(qc, kc, kn, alpha=1.0 / math.sqrt(d), out_f32=logits, batch=batch, b_batch_stride=sets * kn * d)
            p = ops.attn_softmax(logits, k, ns, split=self.split)
        kk = k * sp
        p_view = Pair(p.hi[:, :kk], None if p.lo is None else p.lo[:, :kk])
        vt = vt_all[set_index:]
        vt2 = Pair(vt.hi.view(-1, pitch), None if vt.lo is None else vt.lo.view(-1, pitch))
        ops.linear(p_view, vt2, c, alpha=1.0 / k, bias=rbar_all[set_index:], bias_sn=sets * c, out=out, batch=batch,
                   b_batch_stride=sets * c * pitch, res_f32=res_f32)
        return out

    def rpn_attention(self, base2d: Pair, dense_out: Pair, b, nq, sup: Pair, sets=1):
        """RPN-level BA + CISA (dana.py:117-151).  base2d [B*nq, 1024] (any row pitch): the query feature;
        dense_out [B*nq, 1024] (any row pitch; bf16 pair or one fp16 plane) receives the attended support feature --
        when both are channel halves of one [B,h,w,2048] buffer the `cat` of :154 comes for free.
        sup [B*sets*K, hs, ws, 1024]: support maps, image-major; set 0 of every image drives the block."""
        maps, sh, sw, c = sup.hi.shape
        ns = sh * sw
        # the same kernels sequenced from Python: debugging aid, and the path bench.py's per-launch GEMM trace uses
        # (the events of the trace sit around the Python-level launches)
        if os.environ.get("DANA_CISA_PY") == "1" or ops.GEMM_TRACE is not None:
            return self._rpn_attention_py(base2d, dense_out, b, nq, sup, sets)
        ops.cisa_fwd(base2d, sup.view(maps, ns, c), self.pe(ns), self.n_shot, sets, b, wq=self.rpn_q_w, wk=self.rpn_k_w,
                     un_w=self.rpn_un_w, un_b=self.rpn_un_b, unary_gamma=self.unary_gamma, ba_w=self.ba_w, ba_b=self.ba_b,
                     gamma=self.channel_gamma, out=dense_out)

    def _rpn_attention_py(self, base2d: Pair, dense_out: Pair, b, nq, sup: Pair, sets=1):
        split, k, dev = self.split, self.n_shot, self.device
        maps, sh, sw, c = sup.hi.shape
        ns = sh * sw
        # support side, all sets*K maps at once (:126-147)
        pitch = (k * self.seg_pitch(ns) + 7) // 8 * 8
        vc, vt, rbar = ops.support_prepare(sup.view(maps, ns, c), self.pe(ns), k, ba_w=self.ba_w, ba_b=self.ba_b,
                                           gamma=self.channel_gamma, un_w=self.rpn_un_w, un_b=self.rpn_un_b,
                                           unary_gamma=self.unary_gamma, vt_pitch=pitch, seg_pitch=self.seg_pitch(ns),
                                           split=split)
        kc = ops.linear(vc, self.rpn_k_w, 256, split=split)
        # query side (:118,124-125)
        q = torch.empty((b * nq, 256), dtype=torch.float32, device=dev)
        ops.linear(base2d, self.rpn_q_w, 256, out_f32=q)
        qc = ops.center_rows(q, b, nq, split=split)
        self._attention(qc, kc, vt, rbar, 0, sets, b, ns, dense_out)

    @torch.no_grad()
    def ba_cisa_block(self, base_feat_nchw, support_feat_nchw):
        """The lifted BA+CISA block (BASELINE.json configs 1 and 5): base_feat [B,1024,h,w] fp32,
        support_feat [B,K,1024,hs,ws] fp32 (positive set) -> dense support feature [B,1024,h,w] fp32."""
        b, c, h, w = base_feat_nchw.shape
        k = support_feat_nchw.shape[1]
        assert k == self.n_shot and c == 1024
        base = ops.split_f32(base_feat_nchw.permute(0, 2, 3, 1).contiguous(), self.split)
        dense = Pair.empty((b * h * w, 1024), self.device, self.split)
        hs, ws = support_feat_nchw.shape[3], support_feat_nchw.shape[4]
        sup = ops.split_f32(support_feat_nchw.reshape(b * k, c, hs, ws).permute(0, 2, 3, 1).contiguous(), self.split)
        self.rpn_attention(base.view(b * h * w, 1024), dense, b, h * w, sup, 1)
        return ops.merge_pair(dense).view(b, h, w, 1024).permute(0, 3, 1, 2)

    def _head_support_side(self, sup: Pair, b, sets, pooling_size, want, extra):
        """Support side of the per-RoI CISA head (dana.py:114,255-276,288): depends on the support maps only, so
        forward() runs it on a side stream while the RPN / proposal stages occupy the main one.
        Returns (kc_h centred k-projections, zt key-major transformed values, c64 row-constant term)."""
        split, k, dev = self.split, self.n_shot, self.device
        maps, sh, sw, c = sup.hi.shape
        bins = pooling_size * pooling_size
        sp_k = sh - pooling_size + 1
        s_pooled = ops.avgpool(sup, sp_k)                                 # dana.py:114  [maps,7,7,C] fp32
        if "support_pooled" in want:
            extra["support_pooled"] = s_pooled.permute(0, 3, 1, 2)
        pitch_h = (k * self.seg_pitch(bins) + 7) // 8 * 8
        sp_h = self.seg_pitch(bins)
        vc_h, _vt_h, rbar_h, cbar = ops.support_prepare(s_pooled.view(maps, bins, c), self.pe(bins), k,
                                                        un_w=self.rcnn_un_w, un_b=self.rcnn_un_b,
                                                        unary_gamma=self.unary_gamma, vt_pitch=pitch_h, seg_pitch=sp_h,
                                                        split=split, want_cbar=True)
        kc_h = ops.linear(vc_h, self.rcnn_k_w, 256, split=split)
        # (P V) W^T = P (V W^T): the dense half of the 2048 -> 64 transform (:288) is applied to the support values
        # BEFORE the attention-weighted sum (:281), so the [R*49, 1024] attended feature (241 MB per support set) is
        # never materialised.  V = Vc + 1 m^T (m = column mean) and the rows of P sum to one, hence
        #   dense W_d^T = (1/K) sum_k P_k (Vc_k W_d^T) + (rbar + mean_k m_k) W_d^T      (cbar = rbar + mean_k m_k)
        z = torch.empty((maps * bins, 64), dtype=torch.float32, device=dev)
        ops.linear(vc_h, self.tr_wd, 64, out_f32=z)
        zt = ops.transpose_segments(z.view(maps, bins, 64), k, sp_h, pitch_h, split=split)   # [B*sets, 64, pitch]
        c64 = torch.empty((b * sets, 64), dtype=torch.float32, device=dev)
        ops.linear(cbar, self.tr_wd, 64, out_f32=c64)
        return kc_h, zt, c64

    # ------------------------------------------------------------------ sibling model FSOD
    @torch.no_grad()
    def fsod_attention_feature(self, im_data, support_ims):
        """The attention-RPN input of the sibling model FSOD (lib/model/framework/fsod.py:90-112) on this engine's
        trunk: shot-mean of the positive support maps, AvgPool2d(14) -> a 7x7 kernel per channel, depth-wise
        cross-correlation with the query feature.  im_data [B,3,H,W], support_ims [B,K,3,Hs,Ws] -> NCHW fp32
        [B,1024,h-6,w-6] (what fsod.py feeds to RCNN_rpn).  Only this block of FSOD is built."""
        k = self.n_shot
        base = self.trunk(im_data.float().contiguous())
        sup = self.encode_supports(support_ims[:, :k])
        maps, sh, sw, c = sup.hi.shape
        pooled = ops.avgpool(sup, sh - 6)                              # [B*K,7,7,C]  (linear: commutes with the shot mean)
        kern = ops.group_mean(pooled, k)                               # [B,7,7,C]
        heat, _ = ops.depthwise_xcorr(base, kern, want_f32=True)
        return heat.permute(0, 3, 1, 2)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def encode_supports(self, support_ims):
        """Support-feature cache: at test time the support crops of a class are fixed
        (inference_loader.py:61-71), so their trunk features can be computed once and passed to forward() as
        `support_feats` instead of re-running RCNN_base on them for every query (38 GF per 3-shot query).
        support_ims [B, sets*K, 3, Hs, Ws] -> NHWC pair [B*sets*K, hs, ws, 1024]."""
        return self.trunk(support_ims.reshape(-1, *support_ims.shape[2:]).float().contiguous())

    @torch.no_grad()
    def forward(self, im_data, im_info, support_ims, pre_nms_top_n=6000, post_nms_top_n=300, nms_thresh=0.7,
                pooling_size=7, want=None, teacher=None, support_feats=None, rois_hook=None):
        """Eval forward.  im_data [B,3,H,W] fp32, im_info [B,3], support_ims [B, sets*K, 3, Hs, Ws].
        Returns (rois [B,post,5], cls_prob [sets*B*post, 2], bbox_pred [B*post, 4]); with `want` (a set of
        stage names) also a dict of intermediates exported in the reference's layouts.
        rois_hook (training branch): called as rois_hook(rois [B,post,5]) after the proposal layer, returns the rois
        the head runs on (the proposal-target layer's [B,R,5] sample, dana.py:166-170)."""
        dev, split, k = self.device, self.split, self.n_shot
        b = im_data.shape[0]
        n_sup = support_ims.shape[1] if support_feats is None else support_feats.hi.shape[0] // b
        assert n_sup % k == 0, "support_ims must hold sets*n_shot crops per image"
        sets = n_sup // k
        want = want or ()
        extra = {}

        # ---- trunk over the query and the support crops (dana.py:98,111)
        self._mark("start")
        qh, qw = self._trunk_hw(im_data.shape[2], im_data.shape[3])
        nq = qh * qw
        # the support trunk on the side stream next to the query trunk (DANA_TRUNK_SIDE=0: off), so that one
        # layer's tail wave overlaps the other trunk's next launch: 739 -> 787 images/s (mixed), 620 -> 649 (bf16x3)
        sup_early = None
        trunk_side = (os.environ.get("DANA_TRUNK_SIDE", "1") != "0" and support_feats is None and self.stage_events is None
                      and ops.GEMM_TRACE is None)
        if trunk_side:
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                sup_early = self.encode_supports(support_ims)
        if self.f16:
            # the RPN input [base | dense] is one fp16 plane; the trunk keeps its own bf16-pair output (the q-projection
            # and RoIAlign read it at full precision)
            corr16 = torch.empty((b, qh, qw, 2048), dtype=torch.float16, device=dev)
            corr = None
            base = self.trunk(im_data.float().contiguous())
            dense_out = Pair(corr16.view(b * nq, 2048)[:, 1024:])
        else:
            corr = Pair.empty((b, qh, qw, 2048), dev, split)         # [base | dense]: removes the cat (:154)
            base = self.trunk(im_data.float().contiguous(), out=corr[..., :1024])
            dense_out = Pair(corr.hi.view(b * nq, 2048)[:, 1024:],
                             None if corr.lo is None else corr.lo.view(b * nq, 2048)[:, 1024:])
        if sup_early is not None:
            torch.cuda.current_stream().wait_stream(self._side)
            sup_early.hi.record_stream(torch.cuda.current_stream())
            if sup_early.lo is not None:
                sup_early.lo.record_stream(torch.cuda.current_stream())
        sup = support_feats if support_feats is not None else (sup_early if sup_early is not None else
                                                               self.encode_supports(support_ims))
        maps, sh, sw, c = sup.hi.shape
        if sh != sw:
            raise ValueError("support feature maps must be square (got %dx%d): AvgPool2d(%d) of dana.py:42 assumes "
                             "20x20 maps" % (sh, sw, sh - pooling_size + 1))
        ns = sh * sw
        if "base_feat" in want:
            extra["base_feat"] = ops.merge_pair(base).permute(0, 3, 1, 2)
        if "support_feat" in want:
            extra["support_feat"] = ops.merge_pair(sup).permute(0, 3, 1, 2)
        if teacher and "base_feat" in teacher:                       # teacher forcing for per-stage parity
            tb = ops.split_f32(teacher["base_feat"].permute(0, 2, 3, 1).contiguous(), split)
            base.hi.copy_(tb.hi)
            if split:
                base.lo.copy_(tb.lo)
        if teacher and "support_feat" in teacher:
            sup = ops.split_f32(teacher["support_feat"].reshape(maps, c, sh, sw).permute(0, 2, 3, 1).contiguous(), split)

        self._mark("trunk")
        # The head's support side needs only the support maps.  Its dozen small launches (low occupancy: 24 maps) run on
        # a side stream UNDER the proposal layer's sort + NMS (one CTA per image, 0.27 ms with 144 SMs idle) -- the fork
        # is placed right before that stage (captured as parallel branches of the CUDA graph).  Single stream while
        # bench.py's instrumented step records per-launch / per-stage events, or with DANA_SIDE_STREAM=0.
        head_sup, side = None, None
        use_side = self.stage_events is None and ops.GEMM_TRACE is None and os.environ.get("DANA_SIDE_STREAM", "1") != "0"
        if self.f16:
            base2d = base.view(b * nq, 1024)
        else:
            base2d = Pair(corr.hi.view(b * nq, 2048)[:, :1024], None if corr.lo is None else corr.lo.view(b * nq, 2048)[:, :1024])
        self.rpn_attention(base2d, dense_out, b, nq, sup, sets)
        if "dense" in want:
            if self.f16:        # export at full precision: the same contractions once more into a bf16 pair (tests only)
                dpair = Pair.empty((b * nq, 1024), dev, split)
                self.rpn_attention(base2d, dpair, b, nq, sup, sets)
                extra["dense"] = ops.merge_pair(dpair).view(b, qh, qw, 1024).permute(0, 3, 1, 2)
            else:
                extra["dense"] = ops.merge_pair(corr).permute(0, 3, 1, 2)[:, 1024:]
        if teacher and "dense" in teacher:
            if self.f16:
                corr16[..., 1024:].copy_(teacher["dense"].permute(0, 2, 3, 1))
            else:
                td = ops.split_f32(teacher["dense"].permute(0, 2, 3, 1).contiguous(), split)
                corr.hi[..., 1024:].copy_(td.hi)
                if split:
                    corr.lo[..., 1024:].copy_(td.lo)

        self._mark("cisa_rpn")
        # ---- RoIAlign input (fp32 NHWC) -- and, in the mixed mode, the query half of the fp16 RPN input
        if self.f16:
            base_f32 = ops.merge_pair(base, out16=corr16[..., :1024])
            rpn_in = Pair(corr16)
        else:
            base_f32 = ops.merge_pair(Pair(corr.hi[..., :1024], None if corr.lo is None else corr.lo[..., :1024]))
            rpn_in = corr

        # ---- RPN head + proposal layer (rpn.py:58-78)
        r1 = self._conv(rpn_in, self.rpn_conv, relu=True)
        rpn_raw = torch.empty((b, qh, qw, 6 * self.num_a), dtype=torch.float32, device=dev)
        ops.conv_nhwc(r1, self.rpn_out.w, 6 * self.num_a, ksize=1, bias=self.rpn_out.bias, out_f32=rpn_raw)
        fg, deltas = ops.rpn_fg_prob(rpn_raw, self.num_a)
        if use_side:
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            side = self._side
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                head_sup = self._head_support_side(sup, b, sets, pooling_size, want, extra)
        if "rpn_fg" in want:
            extra["rpn_fg"], extra["rpn_deltas"] = fg, deltas
        if teacher and "rpn_fg" in teacher:
            fg, deltas = teacher["rpn_fg"].contiguous(), teacher["rpn_deltas"].contiguous()
        hwa = nq * self.num_a
        if self._prop_ws is None or self._prop_ws.key != (b, hwa, pre_nms_top_n):
            self._prop_ws = ops.ProposalWorkspace(b, hwa, pre_nms_top_n, dev)
        rois = ops.proposals(fg, deltas, self.base_anchors, im_info.to(dev).float(), qh, qw, self.feat_stride,
                             pre_nms_top_n, post_nms_top_n, nms_thresh, workspace=self._prop_ws)
        if teacher and "rois" in teacher:
            rois = teacher["rois"].to(dev).float().contiguous()
        if "rpn_raw" in want:
            extra["rpn_raw"] = rpn_raw
        if rois_hook is not None:
            rois = rois_hook(rois).to(dev).float().contiguous()

        self._mark("rpn_proposals")
        # ---- RoIAlign on the query feature (dana.py:183)
        r = rois.shape[0] * rois.shape[1]
        bins = pooling_size * pooling_size
        need_f32 = "pooled" in want
        pooled16 = None
        if pooling_size == 7:   # one pass: pooled bf16 pair (head CISA) [+ fp16 plane (layer4 input, mixed mode)]
            res4 = ops.roi_align_head(base_f32, rois.view(-1, 5), 1.0 / 16.0, 0, want_f32=need_f32, want_pair=True,
                                      split=split, want_f16=self.f16)
            pooled_f32, pooled = res4[0], res4[1]
            if self.f16:
                pooled16 = res4[3]
        else:
            pooled_f32, pooled = ops.roi_align_nhwc(base_f32, rois.view(-1, 5), 1.0 / 16.0, pooling_size, 0, split=split)
        if need_f32:
            extra["pooled"] = pooled_f32.permute(0, 3, 1, 2)
        if teacher and "pooled" in teacher:
            pooled_f32 = teacher["pooled"].permute(0, 2, 3, 1).contiguous()
            pooled = ops.split_f32(pooled_f32, split)
            pooled16 = None
        if self.f16 and pooled16 is None:
            pooled16 = Pair.from_float_f16(pooled_f32 if pooled_f32 is not None else ops.merge_pair(pooled))

        self._mark("roi_align")
        # ---- head: box regression (dana.py:246,387-389)
        top = self.layer4(pooled16 if self.f16 else pooled)
        fc7_f32, fc7 = ops.spatial_mean(top.view(r, top.hi.shape[1] * top.hi.shape[2], top.hi.shape[3]), split=split)
        bbox_pred = torch.empty((r, 4), dtype=torch.float32, device=dev)
        ops.linear(fc7, self.bbox_w, 4, bias=self.bbox_b, out_f32=bbox_pred)
        if "fc7" in want:
            extra["fc7"] = fc7_f32

        self._mark("layer4_bbox")
        # ---- head: per-RoI CISA (dana.py:247-290); support projections hoisted out of the RoI loop
        if head_sup is None:
            head_sup = self._head_support_side(sup, b, sets, pooling_size, want, extra)
        elif side is not None:
            torch.cuda.current_stream().wait_stream(side)                # join the side branch
        kc_h, zt, c64 = head_sup
        # query side: the positional encoding of :259 enters as the per-bin bias PE W^T of the two projections
        pooled2d = pooled.view(r * bins, c)
        if bins == 49:
            qt = torch.empty((r * bins, 320), dtype=torch.float32, device=dev)
            ops.linear(pooled2d, self.head_qt_w, 320, out_f32=qt, row_bias=self.pe_qt)        # :259,266 and query half of :288
            q_h, t_q = qt[:, :256], qt[:, 256:]
        else:
            q_h = torch.empty((r * bins, 256), dtype=torch.float32, device=dev)
            t_q = torch.empty((r * bins, 64), dtype=torch.float32, device=dev)
            qpe = Pair.empty((r * bins, c), dev, split)
            ops.add_pe_split(pooled_f32 if pooled_f32 is not None else ops.merge_pair(pooled), self.pe(bins), bins, qpe, c)
            ops.linear(qpe, self.rcnn_q_w, 256, out_f32=q_h)
            ops.linear(qpe, self.tr_wq, 64, bias=self.tr_b, out_f32=t_q)
        qc_h = ops.center_rows(q_h, r, bins, split=split)                 # :267
        cls_scores = torch.empty((sets * r, 2), dtype=torch.float32, device=dev)
        for s in range(sets):
            t = Pair.empty((r * bins, 64), dev, split)
            self._attention(qc_h, kc_h, zt, c64, s, sets, b, bins, t, res_f32=t_q)   # :273-288, dense half folded in
            hid = ops.linear(t.view(r, bins * 64), self.ffn1_w, 1024, bias=self.ffn1_b, relu=True, split=split)
            ops.linear(hid, self.ffn2_w, 2, bias=self.ffn2_b, out_f32=cls_scores[s * r:(s + 1) * r])
        cls_prob = ops.softmax2(cls_scores)                               # :290
        self._mark("head_cisa")
        if "cls_score" in want:
            extra["cls_score"] = cls_scores
        if want:
            return rois, cls_prob, bbox_pred, extra
        return rois, cls_prob, bbox_pred

    @staticmethod
    def _trunk_hw(h, w):
        def down(v):
            v = (v - 1) // 2 + 1          # conv1 7x7/2 pad 3
            v = (v - 2) // 2 + 1          # maxpool 3x3/2 ceil
            v = (v - 1) // 2 + 1          # layer2
            return (v - 1) // 2 + 1       # layer3
        return down(h), down(w)
