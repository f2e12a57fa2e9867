"""torch.autograd bindings of the tensor-core kernels for the TRAINING step (BASELINE
configs[3]): every convolution / linear / batched product of the training branch runs forward AND backward on
dana_conv_gemm (tcgen05), the way train.py:138 `loss.backward()` runs them on cuDNN / cuBLAS in the reference.

  conv (+ frozen-BN fold) (+ residual) (+ ReLU)   resnet.py:66-102 (Bottleneck), rpn.py:58-72, dana.py:124,140,288
      forward      y = relu(conv(x, W * s) + t + res)                one implicit GEMM, fp32 + bf16-pair outputs
      data grad    dx = conv^T(g', W * s),  g' = g * (y > 0)         the same kernel on transposed / rotated weights;
                                                                     a strided 1x1 writes every s-th pixel of a zeroed dx
      weight grad  dW[co][tap][ci] = sum_p g'[p][co] x[p+tap][ci]    one K-major GEMM over the pixel index on the
                                                                     channel-major operands written by dana_grad_prepare
                                                                     and dana_im2col_t (stream-K spreads the long K)
  bmm_nt   c[b] = a[b] @ b[b]^T                   dana.py:142,147,273,278 (torch.bmm) -- three batched GEMMs fwd + bwd

Activations cross the autograd graph as fp32 NHWC tensors; the bf16 (hi, lo) operand planes of a tensor produced by
one of these functions ride along as `tensor._dana_pair` so that the consumer does not split again.  Operands are
split-bf16 everywhere (three MMAs per product, fp32-equivalent), gradients included."""
import torch

from . import ops
from .ops import Pair

_W_CACHE = {}
_LAST_PAIR = None


def begin_step():
    """Drop the packed-weight cache (weights change with every optimiser step)."""
    _W_CACHE.clear()


def _packed(w, scale, need_dgrad=True):
    """Packed operand planes of a conv / linear weight [Cout, Cin, k, k] with the frozen-BN scale folded in: the
    forward / weight-gradient layout and the rotated data-gradient layout, one kernel (dana_pack_conv_weight)."""
    key = (w.data_ptr(), w._version, 0 if scale is None else scale.data_ptr())
    ent = _W_CACHE.get(key)
    if ent is None:
        fwd, dg = ops.pack_conv_weight(w.detach().float(), scale, want_dgrad=need_dgrad)
        ent = {"fwd": fwd, "dgrad": dg, "w": w}
        _W_CACHE[key] = ent
    return ent


def _dgrad_weight(ent, scale):
    if ent["dgrad"] is None:
        ent["dgrad"] = ops.pack_conv_weight(ent["w"].detach().float(), scale, want_dgrad=True)[1]
    return ent["dgrad"]


def pair_of(x):
    """Operand planes of an fp32 NHWC activation (cached on the tensor when a dana op produced it)."""
    p = getattr(x, "_dana_pair", None)
    if p is None:
        p = ops.split_f32(x.detach().contiguous())
    return p


# Direct gradient sink (train_step.SGDTrainer): when set, the weight gradient of a LEAF parameter is accumulated by the
# unpack kernel straight into `sink(param)` (its slice of the gradient arena) instead of travelling through
# AccumulateGrad as a fresh tensor (one allocation + one add launch per use); `done(param)` is called once the last of
# the parameter's uses of this step has written (what a post-accumulate hook would signal).
_SINK = None
_DONE = None
_USES = {}


_FORK = False
_AUX = {}
_FORKED = {}
_KEEP = []


def set_direct_grads(sink=None, done=None, fork_wgrad=False):
    """fork_wgrad: run the weight-gradient half of every conv backward on an auxiliary stream (only with a sink and
    without a `done` callback: the gradients are complete after join_forks(), not when the last use returns)."""
    global _SINK, _DONE, _FORK
    _SINK, _DONE = sink, done
    _FORK = bool(fork_wgrad) and sink is not None and done is None
    _USES.clear()


def _aux_stream(main):
    key = main.cuda_stream
    st = _AUX.get(key)
    if st is None:
        st = _AUX[key] = torch.cuda.Stream(device=main.device)
    return st


def join_forks():
    """Make the current stream wait for every auxiliary weight-gradient stream, then release the kept inputs."""
    cur = torch.cuda.current_stream()
    for st in _FORKED.values():          # only the streams forked since the last join (a capture must not wait on others)
        cur.wait_stream(st)
    _FORKED.clear()
    _KEEP.clear()


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, res, scale, relu, ksize, stride, xp, rp, leaf):
        global _LAST_PAIR
        ent = _packed(w, scale, need_dgrad=ctx.needs_input_grad[0])
        n, h, wd, ci = x.shape
        co = w.shape[0]
        oh, ow = (h, wd) if ksize == 3 else ((h - 1) // stride + 1, (wd - 1) // stride + 1)
        y = torch.empty((n, oh, ow, co), dtype=torch.float32, device=x.device)
        yp = Pair.empty((n, oh, ow, co), x.device)
        if xp is None:
            xp = ops.split_f32(x.detach().contiguous())
        if res is not None and rp is None:
            rp = ops.split_f32(res.detach().contiguous())
        ops.conv_nhwc(xp, ent["fwd"], co, ksize=ksize, stride=stride, bias=None if bias is None else bias.detach(),
                      res=rp, relu=relu, out=yp, out_f32=y)
        ctx.ent, ctx.xp, ctx.relu, ctx.ksize, ctx.stride, ctx.scale = ent, xp, relu, ksize, stride, scale
        ctx.leaf = None
        if leaf is not None and _SINK is not None and ctx.needs_input_grad[1] and _SINK(leaf) is not None:
            ctx.leaf = leaf
            _USES[id(leaf)] = _USES.get(id(leaf), 0) + 1
        ctx.save_for_backward(y if relu else None)
        _LAST_PAIR = yp
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        need_x, need_w, need_b, need_r = ctx.needs_input_grad[:4]
        want_dres = (need_r and ctx.relu) or need_b
        leaf = ctx.leaf
        into = _SINK(leaf).view(gy.shape[-1], -1, ctx.ksize, ctx.ksize) if (leaf is not None and need_w) else None
        if _FORK and into is not None and need_x:
            # weight gradient on an auxiliary stream: it is off the critical path (nothing in backward consumes it), the
            # data gradient is the chain every earlier layer waits for -- and a single-episode launch fills only part of
            # the GPU.  The two halves share nothing but their inputs (each runs its own grad_prepare into its own
            # per-stream workspace); the inputs are kept alive until join_forks().
            main = torch.cuda.current_stream()
            aux = _aux_stream(main)
            gy = gy.contiguous()
            aux.wait_stream(main)
            _FORKED[aux.cuda_stream] = aux
            with torch.cuda.stream(aux):
                ops.conv_backward(gy, y, ctx.xp, None, ctx.scale, ctx.ksize, ctx.stride, need_dx=False, need_dw=True,
                                  want_dres=False, dw_into=into)
            _KEEP.append((gy, y, ctx.xp))
            dx, dw, gf = ops.conv_backward(gy, y, ctx.xp, _dgrad_weight(ctx.ent, ctx.scale), ctx.scale, ctx.ksize,
                                           ctx.stride, need_dx=True, need_dw=False, want_dres=want_dres)
        else:
            dx, dw, gf = ops.conv_backward(gy, y, ctx.xp, _dgrad_weight(ctx.ent, ctx.scale) if need_x else None, ctx.scale,
                                           ctx.ksize, ctx.stride, need_dx=need_x, need_dw=need_w, want_dres=want_dres,
                                           dw_into=into)
        if leaf is not None and need_w:
            _USES[id(leaf)] -= 1
            if _USES[id(leaf)] == 0 and _DONE is not None:
                _DONE(leaf)
        dr = (gf if ctx.relu else gy) if need_r else None
        db = gf.sum(dim=(0, 1, 2)) if need_b else None
        return dx, dw, db, dr, None, None, None, None, None, None, None


def conv(x, w, bias=None, res=None, scale=None, relu=False, ksize=1, stride=1, leaf=None):
    """y = relu?(conv(x, w * scale) + bias + res) on fp32 NHWC tensors; w [Cout, Cin, k, k].  leaf: the Parameter behind
    `w` when w is a reshaped view of it (nn.Linear weights), for the direct gradient sink."""
    xp = getattr(x, "_dana_pair", None)
    rp = None if res is None else getattr(res, "_dana_pair", None)
    if leaf is None and w.is_leaf:
        leaf = w
    y = _ConvFn.apply(x, w, bias, res, scale, relu, ksize, stride, xp, rp, leaf)
    y._dana_pair = _LAST_PAIR
    return y


def linear(x, w, bias=None, relu=False):
    """nn.Linear on [..., K] (K % 8 == 0): the 1x1 convolution of a one-row image."""
    lead = x.shape[:-1]
    co = w.shape[0]
    pad = (-co) % 8
    if pad:
        # the C -> 1 layers (unary / channel attention, dana.py:133,144): zero rows up to 8 outputs, so that the
        # gradient GEMMs see a 16-byte row pitch; the extra columns are sliced off (and receive zero gradient)
        w = torch.cat([w, w.new_zeros(pad, w.shape[1])], 0)
        if bias is not None:
            bias = torch.cat([bias, bias.new_zeros(pad)], 0)
    x4 = x.reshape(1, 1, -1, x.shape[-1])
    if hasattr(x, "_dana_pair") and x.is_contiguous():
        x4._dana_pair = x._dana_pair.view(*x4.shape)
    y = conv(x4, w.view(w.shape[0], w.shape[1], 1, 1), bias=bias, relu=relu, leaf=w if (w.is_leaf and not pad) else None)
    out = y.view(*lead, w.shape[0])
    if pad:
        return out[..., :co]
    out._dana_pair = y._dana_pair.view(*out.shape)
    return out


def _operand(t):
    """Split a [batch, rows, K] fp32 tensor into GEMM operand planes [batch*rows, K]; K is padded to a multiple of 8
    in memory (16-byte row pitch for the tensor map) while the logical extent stays K (out-of-bounds reads are zero)."""
    b, r, k = t.shape
    t = t.detach()
    if k % 8 == 0:
        return ops.split_f32(t.reshape(b * r, k).contiguous())
    kp = (k + 7) // 8 * 8
    buf = torch.zeros((b * r, kp), dtype=torch.float32, device=t.device)
    buf[:, :k] = t.reshape(b * r, k)
    return ops.split_f32(buf)[:, :k]


class _BmmNTFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _bmm_nt_run(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = db = None
        if ctx.needs_input_grad[0]:
            da = _bmm_nt_run(g, b.transpose(1, 2))           # [B,M,N] x [B,K,N]^T
        if ctx.needs_input_grad[1]:
            db = _bmm_nt_run(g.transpose(1, 2), a.transpose(1, 2))   # [B,N,M] x [B,K,M]^T
        return da, db


def _bmm_nt_run(a, b):
    bsz, m, k = a.shape
    n = b.shape[1]
    ap, bp = _operand(a), _operand(b)
    out = torch.empty((bsz, m, n), dtype=torch.float32, device=a.device)
    ops.linear(ap, bp, n, out_f32=out.view(bsz * m, n), batch=bsz, b_batch_stride=n * bp.hi.stride(0))
    return out


def bmm_nt(a, b):
    """torch.bmm(a, b.transpose(1, 2)) on the tensor-core GEMM, forward and backward."""
    return _BmmNTFn.apply(a, b)
