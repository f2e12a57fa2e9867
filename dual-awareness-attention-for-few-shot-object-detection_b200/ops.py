"""torch-tensor front end of the C ABI: every function here only validates shapes, allocates
outputs / workspaces with torch (device memory + current stream are torch's job) and hands raw
pointers to libdana_b200.so.  No math happens in Python."""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ConvBwdArgs, ConvGemmArgs, check


# Kernel-launch accounting (bench.py's gpu_launches) and optional per-GEMM CUDA-event trace
# (bench.py's roofline: list of (start_event, end_event, algorithmic_flops) on the launching stream).
LAUNCHES = 0
GEMM_TRACE = None
STAGE = ""          # set by DanaEngine._mark while a trace is recorded: the forward stage the next launches belong to


def _count(n):
    global LAUNCHES
    LAUNCHES += n


# stream-K scratch of the GEMM kernel: one zero-initialised buffer per (device, stream) -- launches on different
# streams may overlap, and the partial tiles / ready flags of two concurrent launches must not share storage --
# and a launch counter that serves as the epoch of the ready flags
_SK_WS = {}
_SK_EPOCH = 0


def _sk_workspace(device):
    global _SK_EPOCH
    index = device.index if device.index is not None else torch._C._cuda_getDevice()
    key = (device.type, index, torch._C._cuda_getCurrentRawStream(index))
    ws = _SK_WS.get(key)
    if ws is None:
        nbytes = _lib.load().dana_conv_gemm_workspace_bytes()
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        ws[:4096].zero_()      # only the ready flags need a defined start; the kernel hands them back as zero
        _SK_WS[key] = ws
    _SK_EPOCH = _SK_EPOCH % 0x7FFFFFF0 + 1
    return ws, _SK_EPOCH


def _raw_stream():
    # torch.cuda.current_stream() costs ~15 us of python per call (device-index resolution, Stream object); the raw
    # handle is what the C ABI takes, and at ~1500 launches per training step the difference is a third of the step
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _stream():
    return ctypes.c_void_p(_raw_stream())


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.DanaError("dana_b200 ops need CUDA tensors (there is no CPU fallback)")


class Pair:
    """bf16 (hi, lo) planes of an activation / weight: x ~= hi + lo.  lo is None in plain-bf16 mode.
    A single IEEE fp16 plane (hi.dtype == float16, lo None) is the operand format of the layers that run one
    MMA per product in the mixed-precision mode."""
    __slots__ = ("hi", "lo")

    def __init__(self, hi, lo=None):
        self.hi, self.lo = hi, lo

    @staticmethod
    def empty_f16(shape, device):
        return Pair(torch.empty(shape, dtype=torch.float16, device=device))

    @staticmethod
    def from_float_f16(x):
        """Host-side (torch) fp16 rounding, saturating: weights at load time and tests."""
        return Pair(x.clamp(-65504.0, 65504.0).to(torch.float16).contiguous())

    @property
    def is_f16(self):
        return self.hi.dtype == torch.float16

    @staticmethod
    def empty(shape, device, split=True):
        if not split:
            return Pair(torch.empty(shape, dtype=torch.bfloat16, device=device))
        # both planes from one allocation (the second one 16-byte aligned): half the allocator calls of a step
        n = 1
        for d in shape:
            n *= int(d)
        pad = (n + 7) // 8 * 8
        buf = torch.empty((2 * pad,), dtype=torch.bfloat16, device=device)
        return Pair(buf[:n].view(shape), buf[pad:pad + n].view(shape))

    @staticmethod
    def zeros(shape, device, split=True):
        hi = torch.zeros(shape, dtype=torch.bfloat16, device=device)
        lo = torch.zeros(shape, dtype=torch.bfloat16, device=device) if split else None
        return Pair(hi, lo)

    @staticmethod
    def from_float(x, split=True):
        """Host-side (torch) split, used for weights at load time and in tests."""
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16) if split else None
        return Pair(hi.contiguous(), None if lo is None else lo.contiguous())

    def float(self):
        return self.hi.float() if self.lo is None else self.hi.float() + self.lo.float()

    @property
    def shape(self):
        return self.hi.shape

    def view(self, *shape):
        return Pair(self.hi.view(*shape), None if self.lo is None else self.lo.view(*shape))

    def __getitem__(self, idx):
        return Pair(self.hi[idx], None if self.lo is None else self.lo[idx])


# --------------------------------------------------------------------------- tensor-core GEMM / conv
def conv_gemm(a: Pair, a_dims, a_strides, w: Pair, n_out: int, out_dims, o_strides, *, out: Optional[Pair] = None,
              out_f32=None, taps=(1, 1, 0, 0), scale=None, bias=None, bias_sn=0, res: Optional[Pair] = None,
              res_f32=None, r_strides=(0, 0, 0), relu=False, alpha=1.0, b_pitch=None, b_batch_stride=0,
              tile=(0, 0, 0), softmax_ns=0, softmax_pitch=0):
    """Raw call of dana_conv_gemm; see include/dana_b200.h for the argument meaning."""
    _need_cuda(a.hi, w.hi)
    args = ConvGemmArgs()
    args.a_hi, args.a_lo = _p(a.hi), _p(a.lo)
    args.a_c, args.a_w, args.a_h, args.a_n = [int(v) for v in a_dims]
    args.a_sx, args.a_sy, args.a_sn = [int(v) for v in a_strides]
    args.taps_r, args.taps_s, args.pad_y, args.pad_x = taps
    args.b_hi, args.b_lo = _p(w.hi), _p(w.lo)
    args.b_pitch = int(b_pitch if b_pitch is not None else w.hi.stride(-2))
    args.b_batch_stride = int(b_batch_stride)
    args.n_out = int(n_out)
    args.tile_w, args.tile_h, args.tile_n = tile
    args.out_w, args.out_h, args.out_n = [int(v) for v in out_dims]
    args.o_sx, args.o_sy, args.o_sn = [int(v) for v in o_strides]
    args.out_hi = _p(out.hi) if out is not None else None
    args.out_lo = _p(out.lo) if out is not None else None
    args.out_f32 = _p(out_f32)
    args.scale, args.bias, args.bias_sn = _p(scale), _p(bias), int(bias_sn)
    args.res_hi = _p(res.hi) if res is not None else None
    args.res_lo = _p(res.lo) if res is not None else None
    args.res_f32 = _p(res_f32)
    args.r_sx, args.r_sy, args.r_sn = [int(v) for v in r_strides]
    args.alpha = float(alpha)
    args.relu = 1 if relu else 0
    if a.is_f16 != w.is_f16:
        raise _lib.DanaError("dana_conv_gemm: A and B must both be fp16 planes or both bf16 planes")
    args.ab_f16 = 1 if a.is_f16 else 0
    args.io_f16 = 1 if (out is not None and out.is_f16) else 0
    if res is not None and res.is_f16 != bool(args.io_f16):
        raise _lib.DanaError("dana_conv_gemm: the residual must have the output's plane format")
    ws, epoch = _sk_workspace(a.hi.device)
    args.workspace, args.workspace_bytes, args.sk_epoch = _p(ws), ws.numel(), epoch
    args.softmax_ns, args.softmax_pitch = int(softmax_ns), int(softmax_pitch)
    _count(1)
    if GEMM_TRACE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.load().dana_conv_gemm(ctypes.byref(args), _stream()), "dana_conv_gemm")
        e1.record()
        m = args.out_w * args.out_h * args.out_n
        kk = args.taps_r * args.taps_s * args.a_c
        planes = 2 if a.lo is not None else 1
        # algorithmic bytes: every operand element once (activation, weights, residual) + the output once
        nb = args.out_n if b_batch_stride else 1
        out_b = (4 if out_f32 is not None else 0) + ((2 * (2 if out.lo is not None else 1)) if out is not None else 0)
        res_b = (4 if res_f32 is not None else 0) + ((2 * (2 if res.lo is not None else 1)) if res is not None else 0)
        a_elems = args.a_w * args.a_h * args.a_n * args.a_c          # the (strided) input view, each pixel once
        byts = 2.0 * planes * (a_elems + nb * args.n_out * kk) + m * args.n_out * (out_b + res_b)
        GEMM_TRACE.append((e0, e1, 2.0 * m * args.n_out * kk, (m, args.n_out, kk, args.taps_r * args.taps_s),
                           {"stage": STAGE, "mma_per_product": 3 if planes == 2 else 1, "bytes": byts}))
        return
    check(_lib.load().dana_conv_gemm(ctypes.byref(args), _stream()), "dana_conv_gemm")


def conv_nhwc(x: Pair, w: Pair, n_out: int, *, ksize=1, stride=1, scale=None, bias=None, res: Optional[Pair] = None,
              relu=False, out: Optional[Pair] = None, out_f32=None, out_channel_offset=0, split=True, out_f16=False):
    """Conv2d (1x1 any stride, or 3x3 stride 1 pad 1) + per-channel scale/bias (+residual) (+ReLU) on an NHWC
    activation pair [N,H,W,C] (resnet.py:71-76 Bottleneck convs + frozen BN).  Returns the output pair."""
    n, h, wd, c = x.hi.shape
    sn, sy, sx, sc = x.hi.stride()
    assert sc == 1
    if ksize == 1:
        oh, ow = (h - 1) // stride + 1, (wd - 1) // stride + 1
        a_dims = (c, ow, oh, n)
        a_strides = (sx * stride, sy * stride, sn)
        taps = (1, 1, 0, 0)
    else:
        assert ksize == 3 and stride == 1
        oh, ow = h, wd
        a_dims = (c, wd, h, n)
        a_strides = (sx, sy, sn)
        taps = (3, 3, 1, 1)
    if out is None and out_f32 is None:
        out = Pair.empty_f16((n, oh, ow, n_out), x.hi.device) if out_f16 else \
            Pair.empty((n, oh, ow, n_out), x.hi.device, split=split)
    ref = out.hi if out is not None else out_f32
    osn, osy, osx, osc = ref.stride()
    assert osc == 1
    o_strides = (osx, osy, osn)
    r_strides = (0, 0, 0)
    if res is not None:
        rsn, rsy, rsx, _ = res.hi.stride()
        r_strides = (rsx, rsy, rsn)
    if out_channel_offset:
        out_v = Pair(out.hi[..., out_channel_offset:], None if out.lo is None else out.lo[..., out_channel_offset:])
    else:
        out_v = out
    conv_gemm(x, a_dims, a_strides, w, n_out, (ow, oh, n), o_strides, out=out_v, out_f32=out_f32, taps=taps,
              scale=scale, bias=bias, res=res, r_strides=r_strides, relu=relu)
    return out


def linear(x: Pair, w: Pair, n_out: int, *, bias=None, scale=None, relu=False, alpha=1.0, out: Optional[Pair] = None,
           out_f32=None, split=True, batch=1, b_batch_stride=0, bias_sn=0, res_f32=None, softmax_ns=0,
           softmax_pitch=0, row_bias=None):
    """y = alpha * x @ w.T (* scale) + bias on a row-major pair x [batch*rows, K] (nn.Linear / torch.bmm,
    dana.py:124,140,142,147).  With batch > 1 the rows are split evenly and w / bias may differ per batch.
    row_bias: fp32 [period, n_out] added to row i as row_bias[i % period] (the positional-encoding term of a
    projection, (x + PE) W^T = x W^T + PE W^T): the rows are presented to the kernel as a (period x rows/period)
    grid and the table as a residual with a zero stride over the second index."""
    rows_total, k = x.hi.shape
    pitch = x.hi.stride(0)
    assert rows_total % batch == 0
    rows = rows_total // batch
    if out is None and out_f32 is None:
        out = Pair.empty((rows_total, n_out), x.hi.device, split=split)
    ref = out.hi if out is not None else out_f32
    opitch = ref.stride(0)
    if row_bias is not None:
        period = row_bias.shape[0]
        assert batch == 1 and res_f32 is None and rows % period == 0 and row_bias.shape[1] == n_out
        groups = rows // period
        conv_gemm(x, (k, period, groups, 1), (pitch, pitch * period, pitch * rows), w, n_out, (period, groups, 1),
                  (opitch, opitch * period, opitch * rows), out=out, out_f32=out_f32, scale=scale, bias=bias, relu=relu,
                  alpha=alpha, res_f32=row_bias, r_strides=(row_bias.stride(0), 0, 0))
        return out if out is not None else out_f32
    r_strides = (0, 0, 0)
    if res_f32 is not None:
        rp = res_f32.stride(0)
        r_strides = (rp, rp * rows, rp * rows)
    conv_gemm(x, (k, rows, 1, batch), (pitch, pitch * rows, pitch * rows), w, n_out, (rows, 1, batch),
              (opitch, opitch * rows, opitch * rows), out=out, out_f32=out_f32, scale=scale, bias=bias,
              bias_sn=bias_sn, relu=relu, alpha=alpha, b_batch_stride=b_batch_stride, res_f32=res_f32,
              r_strides=r_strides, softmax_ns=softmax_ns, softmax_pitch=softmax_pitch)
    return out if out is not None else out_f32


# --------------------------------------------------------------------------- NMS / proposals / RoIAlign
def nms(boxes, scores, thresh: float):
    """Greedy NMS, kept input indices ascending (int64).  Mirrors model._C.nms (csrc/nms.h:10-28)."""
    _need_cuda(boxes, scores)
    n = boxes.shape[0]
    dev = boxes.device
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device="cpu")  # reference: empty CPU long (nms.h:17-18)
    boxes = boxes.contiguous().float()
    scores = scores.contiguous().float()
    lib = _lib.load()
    wsb = lib.dana_nms_workspace_bytes(n)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    keep = torch.empty((n,), dtype=torch.int64, device=dev)
    count = torch.zeros((1,), dtype=torch.int32, device=dev)
    _count(9)
    check(lib.dana_nms(_p(boxes), _p(scores), n, float(thresh), _p(keep), _p(count), _p(ws), wsb, _stream()),
          "dana_nms")
    return keep[: int(count.item())]


class ProposalWorkspace:
    """Reusable device scratch of the proposal layer for one (batch, H*W*A, pre_nms_top_n)."""

    def __init__(self, batch, hwa, pre_nms_top_n, device):
        self.key = (batch, hwa, pre_nms_top_n)
        self.bytes = _lib.load().dana_proposals_workspace_bytes(batch, hwa, pre_nms_top_n)
        self.buf = torch.empty((self.bytes,), dtype=torch.uint8, device=device)


def proposals(fg_scores, deltas, base_anchors, im_info, feat_h, feat_w, feat_stride, pre_nms_top_n, post_nms_top_n,
              nms_thresh, workspace: Optional[ProposalWorkspace] = None, want_scores=False):
    """_ProposalLayer.forward (proposal_layer.py:49-190) -> rois [B, post, 5] (zero padded)."""
    _need_cuda(fg_scores, deltas, base_anchors, im_info)
    b = fg_scores.shape[0]
    num_a = base_anchors.shape[0]
    hwa = feat_h * feat_w * num_a
    assert fg_scores.shape == (b, hwa) and deltas.shape == (b, hwa, 4)
    dev = fg_scores.device
    if workspace is None or workspace.key != (b, hwa, pre_nms_top_n):
        workspace = ProposalWorkspace(b, hwa, pre_nms_top_n, dev)
    rois = torch.empty((b, post_nms_top_n, 5), dtype=torch.float32, device=dev)
    sc = torch.empty((b, post_nms_top_n), dtype=torch.float32, device=dev) if want_scores else None
    cnt = torch.empty((b,), dtype=torch.int32, device=dev)
    _count(10)
    check(_lib.load().dana_proposals(_p(fg_scores.contiguous()), _p(deltas.contiguous()),
                                     _p(base_anchors.contiguous().float()), _p(im_info.contiguous().float()), b,
                                     feat_h, feat_w, num_a, feat_stride, pre_nms_top_n, post_nms_top_n,
                                     float(nms_thresh), _p(rois), _p(sc), _p(cnt), _p(workspace.buf), workspace.bytes,
                                     _stream()), "dana_proposals")
    if want_scores:
        return rois, sc, cnt
    return rois


def roi_align_forward(inp, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio):
    """model._C.roi_align_forward (csrc/ROIAlign.h:11-27): NCHW fp32 in -> [R,C,ph,pw] fp32."""
    _need_cuda(inp, rois)
    inp = inp.contiguous().float()
    rois = rois.contiguous().float()
    b, c, h, w = inp.shape
    r = rois.shape[0]
    out = torch.empty((r, c, pooled_h, pooled_w), dtype=torch.float32, device=inp.device)
    if r == 0:
        return out
    lib = _lib.load()
    wsb = lib.dana_roi_align_workspace_bytes(b, c, h, w, 0)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=inp.device)
    _count(2)                                                                  # transpose + gather
    check(lib.dana_roi_align_forward(_p(inp), _p(rois), r, b, c, h, w, pooled_h, pooled_w, float(spatial_scale),
                                     int(sampling_ratio), 0, _p(out), None, None, _p(ws), wsb, _stream()),
          "dana_roi_align_forward")
    return out


def roi_align_nhwc(feat_nhwc, rois, spatial_scale, pooled, sampling_ratio, *, want_f32=True, want_pair=True,
                   split=True):
    """Pipeline variant: NHWC fp32 map [B,H,W,C] -> [R,ph,pw,C] as fp32 and/or bf16 pair."""
    _need_cuda(feat_nhwc, rois)
    b, h, w, c = feat_nhwc.shape
    r = rois.shape[0]
    dev = feat_nhwc.device
    out = torch.empty((r, pooled, pooled, c), dtype=torch.float32, device=dev) if want_f32 else None
    pair = Pair.empty((r, pooled, pooled, c), dev, split=split) if want_pair else None
    _count(1)
    check(_lib.load().dana_roi_align_forward(_p(feat_nhwc), _p(rois), r, b, c, h, w, pooled, pooled,
                                             float(spatial_scale), int(sampling_ratio), 1, _p(out),
                                             _p(pair.hi) if pair else None,
                                             _p(pair.lo) if (pair and pair.lo is not None) else None, None, 0,
                                             _stream()), "dana_roi_align_forward")
    return out, pair


def roi_align_head(feat_nhwc, rois, spatial_scale, sampling_ratio, *, pe=None, want_f32=False, want_pair=True,
                   want_qpe=False, split=True, want_f16=False):
    """Head RoIAlign (7x7): NHWC fp32 map [B,H,W,C] -> [R,7,7,C] as fp32 / bf16 pair / pair of value + pe[bin]
    (dana.py:183 and the positional-encoded query of :259 in one pass) / one fp16 plane.
    Returns (f32, pair, qpe_pair) -- and the fp16 plane as a fourth item when want_f16."""
    _need_cuda(feat_nhwc, rois)
    b, h, w, c = feat_nhwc.shape
    r = rois.shape[0]
    dev = feat_nhwc.device
    out = torch.empty((r, 7, 7, c), dtype=torch.float32, device=dev) if want_f32 else None
    pair = Pair.empty((r, 7, 7, c), dev, split=split) if want_pair else None
    qpe = Pair.empty((r, 7, 7, c), dev, split=split) if want_qpe else None
    h16 = Pair.empty_f16((r, 7, 7, c), dev) if want_f16 else None
    _count(1)
    check(_lib.load().dana_roi_align_head(_p(feat_nhwc), _p(rois), r, b, c, h, w, float(spatial_scale),
                                          int(sampling_ratio), _p(out), _p(pair.hi) if pair else None,
                                          _p(pair.lo) if (pair and pair.lo is not None) else None, _p(pe),
                                          _p(qpe.hi) if qpe else None,
                                          _p(qpe.lo) if (qpe and qpe.lo is not None) else None,
                                          _p(h16.hi) if h16 else None, _stream()),
          "dana_roi_align_head")
    if want_f16:
        return out, pair, qpe, h16
    return out, pair, qpe


def roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, batch, channels, height, width,
                       sampling_ratio):
    """model._C.roi_align_backward (csrc/ROIAlign.h:29-45)."""
    _need_cuda(grad, rois)
    grad = grad.contiguous().float()
    rois = rois.contiguous().float()
    gin = torch.empty((batch, channels, height, width), dtype=torch.float32, device=grad.device)
    lib = _lib.load()
    wsb = lib.dana_roi_align_backward_workspace_bytes(rois.shape[0], batch, channels, height, width, pooled_h, pooled_w, 0)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=grad.device)
    _count(4)                                                   # transposes in / out, memset, scatter
    check(lib.dana_roi_align_backward(_p(grad), _p(rois), rois.shape[0], batch, channels, height, width,
                                      pooled_h, pooled_w, float(spatial_scale), int(sampling_ratio), 0, _p(gin), _p(ws),
                                      wsb, _stream()), "dana_roi_align_backward")
    return gin


def roi_align_backward_nhwc(grad_r49c, rois, spatial_scale, batch, height, width, sampling_ratio):
    """Pipeline-layout backward of the 7x7 RoIAlign: grad [R,49,C] fp32 -> grad of the NHWC map [B,H,W,C] fp32."""
    _need_cuda(grad_r49c, rois)
    r, bins, c = grad_r49c.shape
    assert bins == 49
    gin = torch.empty((batch, height, width, c), dtype=torch.float32, device=grad_r49c.device)
    _count(2)
    check(_lib.load().dana_roi_align_backward(_p(grad_r49c.contiguous().float()), _p(rois.contiguous().float()), r, batch,
                                              c, height, width, 7, 7, float(spatial_scale), int(sampling_ratio), 1,
                                              _p(gin), None, 0, _stream()), "dana_roi_align_backward")
    return gin


# --------------------------------------------------------------------------- CUDA-core stages
def pack_stem_weight(w, split=True):
    """conv1 weight [64,3,7,7] -> B operand [64, 4*64] of the space-to-depth form (load time, host side):
    K index = d_y*64 + d_x*16 + (sy*2+sx)*3 + c  holds  w[co, c, 2*d_y+sy-1, 2*d_x+sx-1]  (0 when out of range)."""
    co = w.shape[0]
    packed = torch.zeros((co, 4, 4, 16), dtype=torch.float32, device=w.device)
    for dy in range(4):
        for dx in range(4):
            for sy in range(2):
                for sx in range(2):
                    ky, kx = 2 * dy + sy - 1, 2 * dx + sx - 1
                    if 0 <= ky < 7 and 0 <= kx < 7:
                        q = (sy * 2 + sx) * 3
                        packed[:, dy, dx, q:q + 3] = w[:, :, ky, kx]
    return Pair.from_float(packed.reshape(co, 256).contiguous(), split)


def stem(im_nchw, w_packed: Pair, scale, bias, split=True):
    """conv1 + frozen BN + ReLU + ceil-mode max-pool (resnet.py:109-113): space-to-depth repack, 4-tap
    tensor-core GEMM with the BN/ReLU epilogue, NHWC max-pool.  im [B,3,H,W] fp32 -> pair [B,Hp,Wp,64]."""
    b, c, h, w = im_nchw.shape
    assert c == 3
    dev = im_nchw.device
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    wp = w2 + 4
    s2d = Pair.empty((b, h2, wp, 16), dev, split=split)
    lib = _lib.load()
    _count(1)
    check(lib.dana_stem_s2d(_p(im_nchw.contiguous()), b, h, w, _p(s2d.hi), _p(s2d.lo), _stream()), "dana_stem_s2d")
    conv = Pair.empty((b, h2, w2, 64), dev, split=split)
    conv_gemm(s2d, (64, w2, h2, b), (16, wp * 16, h2 * wp * 16), w_packed, 64, (w2, h2, b),
              (64, w2 * 64, h2 * w2 * 64), out=conv, taps=(4, 1, 2, 0), scale=scale, bias=bias, relu=True)
    ph, pw = (h2 - 2) // 2 + 1, (w2 - 2) // 2 + 1
    out = Pair.empty((b, ph, pw, 64), dev, split=split)
    _count(1)
    check(lib.dana_maxpool3x3s2(_p(conv.hi), _p(conv.lo), b, h2, w2, 64, _p(out.hi), _p(out.lo), _stream()),
          "dana_maxpool3x3s2")
    return out


def avgpool(x: Pair, k: int):
    n, h, w, c = x.hi.shape
    out = torch.empty((n, h - k + 1, w - k + 1, c), dtype=torch.float32, device=x.hi.device)
    _count(1)
    check(_lib.load().dana_avgpool(_p(x.hi), _p(x.lo), n, h, w, c, k, _p(out), _stream()), "dana_avgpool")
    return out


def support_prepare(x, pe, shots, *, ba_w=None, ba_b=None, gamma=0.1, un_w, un_b, unary_gamma=0.1, vt_pitch,
                    seg_pitch=None, split=True, want_cbar=False):
    """Support side of BA+CISA.  x: Pair or fp32 tensor [maps, ns, c].  Returns (vc pair [maps*ns, c],
    vt pair [sets, c, vt_pitch], rbar fp32 [sets, c]) and, with want_cbar, the pair [sets, c] of
    rbar + mean over shots of the column means."""
    if isinstance(x, Pair):
        maps, ns, c = x.hi.shape
        dev = x.hi.device
        in_hi, in_lo, in_f32 = _p(x.hi), _p(x.lo), None
    else:
        maps, ns, c = x.shape
        dev = x.device
        in_hi, in_lo, in_f32 = None, None, _p(x)
    sets = maps // shots
    f = dict(dtype=torch.float32, device=dev)
    v = torch.empty((maps, ns, c), **f)
    logit = torch.empty((maps, ns), **f)
    g = torch.empty((maps, c), **f)
    r = torch.empty((maps, c), **f)
    colmean = torch.empty((maps, c), **f)
    vc = Pair.empty((maps * ns, c), dev, split=split)
    vt = Pair.empty((sets, c, vt_pitch), dev, split=split)       # pad columns are zeroed by the C entry
    rbar = torch.empty((sets, c), **f)
    cbar = Pair.empty((sets, c), dev, split=split) if want_cbar else None
    _count(4)                                  # logits, weighted sums, finalize, rbar
    check(_lib.load().dana_support_prepare(in_hi, in_lo, in_f32, _p(pe), maps, shots, ns, c, _p(ba_w), _p(ba_b),
                                           float(gamma), _p(un_w), _p(un_b), float(unary_gamma), _p(v), _p(logit),
                                           _p(g), _p(r), _p(colmean), _p(vc.hi), _p(vc.lo), _p(vt.hi), _p(vt.lo),
                                           int(vt_pitch), int(seg_pitch if seg_pitch is not None else ns), _p(rbar),
                                           _p(cbar.hi) if cbar else None,
                                           _p(cbar.lo) if (cbar and cbar.lo is not None) else None,
                                           _stream()), "dana_support_prepare")
    if want_cbar:
        return vc, vt, rbar, cbar
    return vc, vt, rbar


# scratch of dana_cisa_fwd, one per (device, stream, size)
_CISA_WS = {}


def cisa_fwd(q: Pair, sup: Pair, pe, shots, sets, batch, *, wq: Pair, wk: Pair, un_w, un_b, unary_gamma=0.1, ba_w=None,
             ba_b=None, gamma=0.1, out: Pair):
    """RPN-level BA + CISA block in ONE C call (dana_cisa_fwd; dana.py:117-151).
    q [batch*nq, c] query feature (any row pitch); sup [batch*sets*shots, ns, c] support maps (set 0 of every image
    drives the block); out [batch*nq, c] (any row pitch): bf16 pair, or one fp16 plane."""
    _need_cuda(q.hi, sup.hi, out.hi)
    rows, c = q.hi.shape
    maps, ns, c2 = sup.hi.shape
    d = wq.hi.shape[0]
    assert c2 == c and rows % batch == 0 and maps == batch * sets * shots
    assert sup.hi.is_contiguous() and (sup.lo is None or sup.lo.is_contiguous())
    nq = rows // batch
    lib = _lib.load()
    dev = q.hi.device
    nbytes = lib.dana_cisa_workspace_bytes(batch, nq, sets, shots, ns, c, d)
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    key = (index, torch.cuda.current_stream(index).cuda_stream, nbytes)
    ws = _CISA_WS.get(key)
    if ws is None:
        ws = _CISA_WS[key] = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    a = _lib.CisaArgs()
    a.q_hi, a.q_lo, a.q_pitch = _p(q.hi), _p(q.lo), q.hi.stride(0)
    a.s_hi, a.s_lo = _p(sup.hi), _p(sup.lo)
    a.batch, a.nq, a.sets, a.shots, a.ns, a.c, a.d = batch, nq, sets, shots, ns, c, d
    a.pe = _p(pe)
    a.wq_hi, a.wq_lo, a.wk_hi, a.wk_lo = _p(wq.hi), _p(wq.lo), _p(wk.hi), _p(wk.lo)
    a.un_w, a.un_b, a.unary_gamma = _p(un_w), _p(un_b), float(unary_gamma)
    a.ba_w, a.ba_b, a.gamma = _p(ba_w), _p(ba_b), float(gamma)
    a.out_hi, a.out_lo, a.out_pitch = _p(out.hi), _p(out.lo), out.hi.stride(0)
    a.out_f16 = 1 if out.is_f16 else 0
    a.workspace, a.workspace_bytes = _p(ws), nbytes
    _count(14)     # 4 support kernels, 4 GEMMs (softmax fused), 3 centring kernels, 3 memsets
    check(lib.dana_cisa_fwd(ctypes.byref(a), _stream()), "dana_cisa_fwd")
    return out


def transpose_segments(x_f32, shots, seg_pitch, vt_pitch, split=True):
    """x [maps, ns, c] fp32 -> pair [maps/shots, c, vt_pitch] with element (set, ch, slot*seg_pitch + n), pads zero."""
    _need_cuda(x_f32)
    maps, ns, c = x_f32.shape
    out = Pair.empty((maps // shots, c, vt_pitch), x_f32.device, split=split)
    _count(1)
    check(_lib.load().dana_transpose_segments(_p(x_f32.contiguous()), maps, shots, ns, c, int(seg_pitch), int(vt_pitch),
                                              _p(out.hi), _p(out.lo), _stream()), "dana_transpose_segments")
    return out


def center_rows(x_f32, groups, group_rows, split=True):
    """x - mean over the rows of each group (dana.py:125,141,267,272) -> pair.  x may be a column slice of a wider
    fp32 matrix (row pitch > columns) when the groups are small (the head: 49 rows)."""
    c = x_f32.shape[-1]
    dev = x_f32.device
    out = Pair.empty((groups * group_rows, c), dev, split=split)
    if x_f32.dim() == 2 and x_f32.stride(0) != c:
        _count(1)
        check(_lib.load().dana_center_rows_pitched(_p(x_f32), x_f32.stride(0), groups, group_rows, c, _p(out.hi),
                                                   _p(out.lo), _stream()), "dana_center_rows_pitched")
        return out
    # scratch of the large-group path: final column sums + per-block partial sums (fixed-order reduction, no atomics)
    sums = torch.empty((groups * (1 + (group_rows + 63) // 64), c), dtype=torch.float32, device=dev) if group_rows > 256 else None
    _count(3 if group_rows > 256 else 1)
    check(_lib.load().dana_center_rows(_p(x_f32), groups, group_rows, c, _p(out.hi), _p(out.lo), _p(sums), _stream()),
          "dana_center_rows")
    return out


def attn_softmax(logits_f32, segs, ns, split=True):
    rows, pitch = logits_f32.shape
    out = Pair.empty((rows, pitch), logits_f32.device, split=split)
    _count(1)
    check(_lib.load().dana_attn_softmax(_p(logits_f32), rows, segs, ns, pitch, _p(out.hi), _p(out.lo), _stream()),
          "dana_attn_softmax")
    return out


def rpn_fg_prob(rpn_out_f32, num_a):
    """rpn_out_f32 [B,H,W,6A] fp32 -> fg [B, H*W*A], deltas [B, H*W*A, 4]."""
    b, h, w, pitch = rpn_out_f32.shape
    dev = rpn_out_f32.device
    fg = torch.empty((b, h * w * num_a), dtype=torch.float32, device=dev)
    deltas = torch.empty((b, h * w * num_a, 4), dtype=torch.float32, device=dev)
    _count(1)
    check(_lib.load().dana_rpn_fg_prob(_p(rpn_out_f32), b * h * w, num_a, pitch, _p(fg), _p(deltas), _stream()),
          "dana_rpn_fg_prob")
    return fg, deltas


def add_pe_split(x_f32, pe, period, out: Pair, out_pitch):
    rows = x_f32.numel() // x_f32.shape[-1]
    c = x_f32.shape[-1]
    _count(1)
    check(_lib.load().dana_add_pe_split(_p(x_f32), _p(pe), rows, c, period, out_pitch, _p(out.hi), _p(out.lo),
                                        _stream()), "dana_add_pe_split")
    return out


def split_f32(x_f32, split=True):
    out = Pair.empty(tuple(x_f32.shape), x_f32.device, split=split)
    _count(1)
    check(_lib.load().dana_split_f32(_p(x_f32.contiguous()), x_f32.numel(), _p(out.hi), _p(out.lo), _stream()),
          "dana_split_f32")
    return out


def merge_pair(x: Pair, out16=None, want_f32=True):
    """pair [..., c] (last dim contiguous, uniform row pitch) -> contiguous fp32 of the same shape; `out16`
    (an fp16 tensor view [..., c] with a uniform row pitch) also receives the values as one fp16 plane."""
    shape = tuple(x.hi.shape)
    c = shape[-1]
    rows = x.hi.numel() // c
    pitch = x.hi.stride(-2) if x.hi.dim() > 1 else c
    out = torch.empty(shape, dtype=torch.float32, device=x.hi.device) if want_f32 else None
    p16 = 0
    if out16 is not None:
        assert out16.dtype == torch.float16 and out16.shape == x.hi.shape and out16.stride(-1) == 1
        p16 = out16.stride(-2) if out16.dim() > 1 else c
    _count(1)
    check(_lib.load().dana_merge_pair(_p(x.hi), _p(x.lo), rows, c, pitch, _p(out), _p(out16), p16, _stream()),
          "dana_merge_pair")
    return out


def spatial_mean(x: Pair, split=True):
    """[items, sp, c] (bf16 pair or one fp16 plane) -> (fp32 [items, c], bf16 pair [items, c])."""
    items, sp, c = x.hi.shape
    dev = x.hi.device
    out = torch.empty((items, c), dtype=torch.float32, device=dev)
    pair = Pair.empty((items, c), dev, split=split)
    _count(1)
    check(_lib.load().dana_spatial_mean(_p(x.hi), _p(x.lo), 1 if x.is_f16 else 0, items, sp, c, _p(out), _p(pair.hi),
                                        _p(pair.lo), _stream()), "dana_spatial_mean")
    return out, pair


def softmax2(x_f32):
    out = torch.empty_like(x_f32)
    _count(1)
    check(_lib.load().dana_softmax2(_p(x_f32.contiguous()), x_f32.shape[0], _p(out), _stream()), "dana_softmax2")
    return out


def nhwc_pair_to_nchw(x: Pair):
    n, h, w, c = x.hi.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.hi.device)
    _count(1)
    check(_lib.load().dana_nhwc_pair_to_nchw(_p(x.hi), _p(x.lo), n, c, h * w, _p(out), _stream()),
          "dana_nhwc_pair_to_nchw")
    return out


# --------------------------------------------------------------------------- training branch: losses
def rpn_losses(rpn_out_f32, labels_i8, bbox_targets, inside_w, outside_w, num_a):
    """rpn.py:96-116.  rpn_out_f32 [B,H,W,6A] fp32 (the RPN head's raw output), targets in (y, x, a) order
    (dana_b200.targets.anchor_targets) -> fp32 [2] = (rpn_loss_cls, rpn_loss_box)."""
    _need_cuda(rpn_out_f32, labels_i8, bbox_targets, inside_w, outside_w)
    b, h, w, pitch = rpn_out_f32.shape
    assert labels_i8.dtype == torch.int8 and labels_i8.numel() == b * h * w * num_a
    lib = _lib.load()
    dev = rpn_out_f32.device
    wsb = lib.dana_rpn_losses_workspace_bytes()
    ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
    out = torch.empty((2,), dtype=torch.float32, device=dev)
    _count(2)
    check(lib.dana_rpn_losses(_p(rpn_out_f32), b, b * h * w, num_a, pitch, _p(labels_i8.contiguous()),
                              _p(bbox_targets.contiguous()), _p(inside_w.contiguous()), _p(outside_w.contiguous()),
                              _p(out), _p(ws), wsb, _stream()), "dana_rpn_losses")
    return out


def rcnn_losses(cls_scores, labels, bbox_pred, bbox_targets, inside_w, outside_w):
    """dana.py:199-215.  cls_scores [2R,2] (positive-support rows then negative-support rows), labels fp32 [R],
    bbox tensors [R,4] -> fp32 [2] = (RCNN_loss_cls, RCNN_loss_bbox)."""
    _need_cuda(cls_scores, labels, bbox_pred, bbox_targets, inside_w, outside_w)
    r = labels.numel()
    assert cls_scores.shape == (2 * r, 2) and bbox_pred.shape == (r, 4)
    out = torch.empty((2,), dtype=torch.float32, device=cls_scores.device)
    _count(1)
    check(_lib.load().dana_rcnn_losses(_p(cls_scores.contiguous()), _p(labels.contiguous().float()), r,
                                       _p(bbox_pred.contiguous()), _p(bbox_targets.contiguous().float()),
                                       _p(inside_w.contiguous().float()), _p(outside_w.contiguous().float()), _p(out),
                                       _stream()), "dana_rcnn_losses")
    return out


# --------------------------------------------------------------------------- sibling model FSOD: attention-RPN feature
def group_mean(x_f32, k):
    """[groups*k, ...] fp32 -> [groups, ...]: mean over consecutive groups of k (fsod.py:96,103)."""
    _need_cuda(x_f32)
    groups = x_f32.shape[0] // k
    n = x_f32[0].numel()
    out = torch.empty((groups,) + tuple(x_f32.shape[1:]), dtype=torch.float32, device=x_f32.device)
    _count(1)
    check(_lib.load().dana_group_mean(_p(x_f32.contiguous()), groups, k, n, _p(out), _stream()), "dana_group_mean")
    return out


def depthwise_xcorr(x: Pair, kernel_f32, *, want_f32=True, want_pair=False, split=True):
    """x [B,h,w,C] pair, kernel [B,kh,kw,C] fp32 -> [B,h-kh+1,w-kw+1,C] (fsod.py:106-112).  Returns (f32, pair)."""
    _need_cuda(x.hi, kernel_f32)
    b, h, w, c = x.hi.shape
    kb, kh, kw, kc = kernel_f32.shape
    assert kb == b and kc == c and x.hi.is_contiguous()
    dev = x.hi.device
    out = torch.empty((b, h - kh + 1, w - kw + 1, c), dtype=torch.float32, device=dev) if want_f32 else None
    pair = Pair.empty((b, h - kh + 1, w - kw + 1, c), dev, split=split) if want_pair else None
    _count(1)
    check(_lib.load().dana_depthwise_xcorr(_p(x.hi), _p(x.lo), b, h, w, c, _p(kernel_f32.contiguous()), kh, kw, _p(out),
                                           _p(pair.hi) if pair else None,
                                           _p(pair.lo) if (pair and pair.lo is not None) else None, _stream()),
          "dana_depthwise_xcorr")
    return out, pair


# --------------------------------------------------------------------------- backward-pass layout kernels
def grad_prepare(grad_f32, relu_out=None, *, want_f32=False, want_pair=True, want_t=True):
    """g' = grad * (relu_out > 0) on an fp32 [..., C] gradient (C % 4 == 0).  Returns (g' fp32 | None, NHWC pair | None,
    channel-major pair [C, pitch] viewed [C, pixels] | None): the operands of the data- and weight-gradient GEMMs."""
    _need_cuda(grad_f32)
    g = grad_f32.contiguous()
    c = g.shape[-1]
    pixels = g.numel() // c
    dev = g.device
    out_f32 = torch.empty_like(g) if want_f32 else None
    pair = Pair.empty(g.shape, dev) if want_pair else None
    pitch = (pixels + 7) // 8 * 8
    tp = Pair.empty((c, pitch), dev) if want_t else None
    y = None if relu_out is None else relu_out.contiguous()
    _count(1)
    check(_lib.load().dana_grad_prepare(_p(g), _p(y), pixels, c, _p(out_f32), _p(pair.hi) if pair else None,
                                        _p(pair.lo) if pair else None, _p(tp.hi) if tp else None,
                                        _p(tp.lo) if tp else None, pitch, _stream()), "dana_grad_prepare")
    return out_f32, pair, (tp[:, :pixels] if tp else None)


def im2col_t(x: Pair, ksize=1, stride=1):
    """Channel-major im2col of an NHWC pair [N,H,W,C] -> pair [taps*C, pixels] (a view of a [taps*C, pitch] buffer)."""
    _need_cuda(x.hi)
    n, h, w, c = x.hi.shape
    sn, sy, sx, sc = x.hi.stride()
    assert sc == 1 and (x.lo is None or x.lo.stride() == x.hi.stride())
    oh, ow = (h, w) if ksize == 3 else ((h - 1) // stride + 1, (w - 1) // stride + 1)
    pixels = n * oh * ow
    pitch = (pixels + 7) // 8 * 8
    out = Pair.empty((ksize * ksize * c, pitch), x.hi.device, split=x.lo is not None)
    _count(1)
    check(_lib.load().dana_im2col_t(_p(x.hi), _p(x.lo), n, h, w, c, sn, sy, sx, ksize, stride, _p(out.hi), _p(out.lo),
                                    pitch, _stream()), "dana_im2col_t")
    return out[:, :pixels]


def pack_conv_weight(w, scale=None, want_dgrad=True):
    """fp32 [Cout,Cin,kh,kw] (x scale[Cout]) -> (forward pair [Cout, kh*kw*Cin], data-gradient pair [Cin, kh*kw*Cout])."""
    _need_cuda(w)
    w = w.detach().contiguous()
    co, ci, kh, kw = w.shape
    fwd = Pair.empty((co, kh * kw * ci), w.device)
    dg = Pair.empty((ci, kh * kw * co), w.device) if want_dgrad else None
    _count(1)
    check(_lib.load().dana_pack_conv_weight(_p(w), _p(scale), co, ci, kh, kw, _p(fwd.hi), _p(fwd.lo),
                                            _p(dg.hi) if dg else None, _p(dg.lo) if dg else None, _stream()),
          "dana_pack_conv_weight")
    return fwd, dg


def unpack_conv_wgrad(wgrad, scale, co, ci, kh, kw, accumulate_into=None):
    """Weight-gradient GEMM output [Cout, kh*kw*Cin] -> contiguous fp32 [Cout,Cin,kh,kw] (x scale[Cout])."""
    out = accumulate_into if accumulate_into is not None else \
        torch.empty((co, ci, kh, kw), dtype=torch.float32, device=wgrad.device)
    _count(1)
    check(_lib.load().dana_unpack_conv_wgrad(_p(wgrad), _p(scale), co, ci, kh * kw, _p(out),
                                             1 if accumulate_into is not None else 0, _stream()),
          "dana_unpack_conv_wgrad")
    return out


_BWD_WS = {}


def conv_backward(grad_out, relu_out, x: Pair, w_dgrad: Optional[Pair], scale, ksize, stride, *, need_dx, need_dw,
                  want_dres, dw_into=None):
    """Backward of y = relu?(conv(x, W*s) + t (+ res)) in ONE C call (dana_conv_backward): -> (dx | None, dw | None,
    dres | None).  grad_out fp32 [N,OH,OW,Cout]; x the saved NHWC input pair; dw_into: an fp32 [Cout,Cin,k,k] tensor
    that receives dw += (gradient arenas; a parameter used by several calls) instead of a fresh tensor."""
    _need_cuda(grad_out, x.hi)
    g = grad_out.contiguous()
    n, h, w, ci = x.hi.shape
    co = g.shape[-1]
    dev = g.device
    lib = _lib.load()
    sn, sy, sx, sc = x.hi.stride()
    assert sc == 1 and x.lo.stride() == x.hi.stride()
    a = ConvBwdArgs()
    a.batch, a.height, a.width, a.in_channels, a.out_channels, a.ksize, a.stride = n, h, w, ci, co, ksize, stride
    a.grad_out = g.data_ptr()
    a.relu_out = None if relu_out is None else relu_out.data_ptr()
    a.x_hi, a.x_lo = x.hi.data_ptr(), x.lo.data_ptr()
    a.x_sn, a.x_sy, a.x_sx = sn, sy, sx
    dx = dw = dres = None
    if need_dx:
        dx = torch.empty((n, h, w, ci), dtype=torch.float32, device=dev)
        a.dx = dx.data_ptr()
        a.wd_hi, a.wd_lo = w_dgrad.hi.data_ptr(), w_dgrad.lo.data_ptr()
    if need_dw:
        dw = dw_into if dw_into is not None else torch.empty((co, ci, ksize, ksize), dtype=torch.float32, device=dev)
        a.dw = dw.data_ptr()
        a.dw_accumulate = 1 if dw_into is not None else 0
    if want_dres:
        dres = torch.empty_like(g)
        a.dres = dres.data_ptr()
    a.scale = None if scale is None else scale.data_ptr()
    need = lib.dana_conv_backward_workspace_bytes(n, h, w, ci, co, ksize, stride)
    key = (dev.index, _raw_stream())
    ws = _BWD_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _BWD_WS[key] = torch.empty((int(need * 1.25),), dtype=torch.uint8, device=dev)   # grows, then stays
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    global _SK_EPOCH
    gws, epoch = _sk_workspace(dev)
    _SK_EPOCH = _SK_EPOCH % 0x7FFFFFF0 + 1            # the call uses two consecutive epochs
    a.gemm_workspace, a.gemm_workspace_bytes, a.sk_epoch = gws.data_ptr(), gws.numel(), epoch
    _count(1 + (1 if need_dx else 0) + (3 if need_dw else 0))
    check(lib.dana_conv_backward(ctypes.byref(a), _stream()), "dana_conv_backward")
    return dx, (dw if dw_into is None else None), dres


def sgd_momentum(param_flat, grad_flat, mom_flat, lr, momentum, weight_decay, grad_scale=1.0):
    """In-place torch.optim.SGD step (train.py:89,139) on flat fp32 buffers."""
    _need_cuda(param_flat, grad_flat, mom_flat)
    _count(1)
    check(_lib.load().dana_sgd_momentum(_p(param_flat), _p(grad_flat), _p(mom_flat), param_flat.numel(), float(lr),
                                        float(momentum), float(weight_decay), float(grad_scale), _stream()),
          "dana_sgd_momentum")


def device_error():
    return _lib.load().dana_device_error()
