"""One training step of train.py:120-139 with one process per GPU: forward of the training branch with its autograd
graph (train_model.py), `loss.backward()` on the tcgen05 gradient GEMMs (autograd_ops.py), the gradient all-reduce that
replaces nn.DataParallel's replica sum (train.py:104-105,131-138) and torch.optim.SGD's update (train.py:89,139).

Memory layout: the trainable parameters, their gradients and their momentum live in three flat fp32 ARENAS per
parameter group (every tensor a 16-byte aligned slice, in reverse registration order = roughly the order backward
produces the gradients).  That makes
  * zeroing the gradients one memset per arena,
  * the all-reduce a sequence of NCCL calls directly on contiguous arena ranges ("buckets", ~25 MB), each launched from a
    post-accumulate hook the moment its last gradient is in -- no pack / unpack copies, the collective of the head's
    gradients overlaps the backward of the trunk,
  * the SGD update ONE fused kernel per arena (dana_sgd_momentum), with the 1/world of the gradient average folded in.

Parameter groups as train.py:78-89: biases take lr * (DOUBLE_BIAS + 1) and weight decay only if BIAS_DECAY, everything
else lr and WEIGHT_DECAY; momentum cfg.TRAIN.MOMENTUM."""
import torch
import torch.distributed as dist

from . import ops
from .config import cfg


class ParamArena:
    """Flat storage of a parameter group: params / grads / momentum as slices of three fp32 buffers."""

    def __init__(self, named, lr, weight_decay, bucket_bytes):
        self.lr, self.weight_decay = lr, weight_decay
        self.named = list(named)
        dev = self.named[0][1].device
        offs, o = [], 0
        for _, p in self.named:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4                       # 16-byte aligned slices
        self.total = o
        self.param = torch.zeros(o, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(o, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(o, dtype=torch.float32, device=dev)
        self.offsets = offs
        with torch.no_grad():
            for (_, p), off in zip(self.named, offs):
                n = p.numel()
                self.param[off:off + n].copy_(p.detach().reshape(-1).float())
                p.data = self.param[off:off + n].view(p.shape)
                p.grad = self.grad[off:off + n].view(p.shape)
        # buckets: contiguous ranges of whole tensors, closed when they reach bucket_bytes
        self.buckets, start, count = [], 0, 0                  # (first index, last index + 1, begin, end)
        first = 0
        for i, ((_, p), off) in enumerate(zip(self.named, offs)):
            count += p.numel() * 4
            end = off + (p.numel() + 3) // 4 * 4
            if count >= bucket_bytes or i == len(self.named) - 1:
                self.buckets.append((first, i + 1, start, end))
                first, start, count = i + 1, end, 0

    def rebind_grads(self):
        """autograd may replace .grad (e.g. after p.grad = None by the caller): point it back into the arena."""
        for (_, p), off in zip(self.named, self.offsets):
            g = p.grad
            if g is None or g.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + p.numel()].view(p.shape)


class ArenaGradAllReduce:
    """Bucketed, overlapped gradient all-reduce over the arenas (sum; the 1/world goes into the SGD kernel)."""

    def __init__(self, arenas, process_group=None):
        self.arenas = arenas
        self.group = process_group
        self.order = []                                         # (arena index, bucket index) in launch order
        self._bucket_of = {}
        for ai, a in enumerate(arenas):
            for bi, (i0, i1, _, _) in enumerate(a.buckets):
                self.order.append((ai, bi))
                for (_, p) in a.named[i0:i1]:
                    self._bucket_of[id(p)] = len(self.order) - 1
        self._hooks = []
        self._pending, self._work, self._next = [], [], 0

    def arm(self):
        self.disarm()
        self._pending = [self.arenas[ai].buckets[bi][1] - self.arenas[ai].buckets[bi][0] for ai, bi in self.order]
        self._work = [None] * len(self.order)
        self._next = 0
        for a in self.arenas:
            for _, p in a.named:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def disarm(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def _on_grad(self, p):
        k = self._bucket_of[id(p)]
        self._pending[k] -= 1
        if self._pending[k] < 0:
            raise RuntimeError("ArenaGradAllReduce: a gradient arrived twice after one arm()")
        # collectives must be issued in the same order on every rank: strictly in bucket order (as DDP does)
        while self._next < len(self.order) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def _launch(self, k):
        ai, bi = self.order[k]
        a = self.arenas[ai]
        _, _, begin, end = a.buckets[bi]
        self._work[k] = dist.all_reduce(a.grad[begin:end], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        for k in range(len(self.order)):                        # buckets whose hooks never completed, still in order
            if self._work[k] is None:
                self._launch(k)
        self._next = len(self.order)
        for w in self._work:
            w.wait()
        self.disarm()


class SGDTrainer:
    def __init__(self, net, lr=None, momentum=None, weight_decay=None, bucket_bytes=25 << 20):
        self.net = net
        self.lr = cfg.TRAIN.LEARNING_RATE if lr is None else lr
        self.momentum = cfg.TRAIN.MOMENTUM if momentum is None else momentum
        wd = cfg.TRAIN.WEIGHT_DECAY if weight_decay is None else weight_decay
        named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        named.reverse()                                         # backward reaches the last layers first
        weights = [(n, p) for n, p in named if "bias" not in n]
        biases = [(n, p) for n, p in named if "bias" in n]
        self.arenas = [ParamArena(weights, self.lr, wd, bucket_bytes)]
        if biases:
            self.arenas.append(ParamArena(biases, self.lr * (float(cfg.TRAIN.DOUBLE_BIAS) + 1.0),
                                          wd if cfg.TRAIN.BIAS_DECAY else 0.0, bucket_bytes))
        self.groups = [(n, p) for a in self.arenas for n, p in a.named]
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.sync = ArenaGradAllReduce(self.arenas) if self.world > 1 else None
        self.events = None               # optional: list receiving (name, CUDA event) marks of one step

    def _mark(self, name):
        if self.events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.events.append((name, e))

    def step(self, im_data, im_info, gt_boxes, num_boxes, support_ims):
        """-> (loss, (rpn_loss_cls, rpn_loss_box, RCNN_loss_cls, RCNN_loss_bbox)) as detached 0-dim CUDA tensors."""
        for a in self.arenas:
            a.rebind_grads()
            a.grad.zero_()
        self._mark("start")
        out = self.net(im_data, im_info, gt_boxes, num_boxes, support_ims)
        losses = out[3:7]
        loss = losses[0].mean() + losses[1].mean() + losses[2].mean() + losses[3].mean()      # train.py:136-137
        self._mark("forward")
        if self.sync is not None:
            self.sync.arm()
        loss.backward()
        self._mark("backward")
        if self.sync is not None:
            self.sync.finish()
        self._mark("allreduce")
        for a in self.arenas:
            ops.sgd_momentum(a.param, a.grad, a.mom, a.lr, self.momentum, a.weight_decay, grad_scale=1.0 / self.world)
            # the update happened behind autograd's back: bump the version counters so that every cache keyed on them
            # (packed weights, the eval engine of DAnARCNN) sees new weights
            torch.autograd.graph.increment_version([p for _, p in a.named])
        self._mark("sgd")
        return loss.detach(), tuple(l.detach() for l in losses)
