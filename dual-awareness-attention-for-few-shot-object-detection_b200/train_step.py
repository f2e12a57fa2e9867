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
import os

import torch
import torch.distributed as dist

from . import autograd_ops, ops, train_model
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
        self._seen = set()
        for a in self.arenas:
            for _, p in a.named:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def disarm(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def _on_grad(self, p):
        # A parameter whose gradient is written by the direct sink (autograd_ops) is signalled by its last use AND,
        # depending on the torch version, by the post-accumulate hook of its (gradient-less) AccumulateGrad node: the
        # first signal counts
        if id(p) in self._seen:
            return
        self._seen.add(id(p))
        k = self._bucket_of[id(p)]
        self._pending[k] -= 1
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


class _CapturedStep:
    """The device side of one training step as two CUDA graphs around the host's target layers:
         graph 1   gradient arenas zeroed, weights packed, frozen stem, layer2-3 on query and supports, RPN-level
                   attention, RPN head, proposal layer                                   (TrainGraph.part1)
         host      anchor targets (drawn while graph 1 runs), proposals D2H, proposal targets, targets H2D
         graph 2   RPN losses, RoIAlign, layer4, both head passes, R-CNN losses, and the WHOLE backward (through the
                   autograd graph that the capture of graph 1 recorded), gradients accumulated into the arenas
       ~3800 launches issued from Python become two graph launches: the eager step is host-bound (47 ms of device work
       in 84 ms), the captured one is not.  The all-reduce and the two SGD launches follow graph 2."""

    def __init__(self, trainer, im_data, im_info, gt_boxes, support_ims):
        # parallel graph branches (the two trunks, the three head branches) on side streams; eager steps stay
        # single-stream: their hook-launched collectives are ordered after the current stream only, and they are launch-bound
        train_model.SIDE_STREAM = True
        try:
            self._build(trainer, im_data, im_info, gt_boxes, support_ims)
        finally:
            train_model.SIDE_STREAM = False
            autograd_ops.set_direct_grads(None, None)

    def _build(self, trainer, im_data, im_info, gt_boxes, support_ims):
        import numpy as np

        from .anchors import generate_anchors
        from .train_model import FrozenStem, TrainGraph
        net = trainer.net
        dev = im_data.device
        named = dict(net.named_parameters())
        named.update(dict(net.named_buffers()))
        self.tg = TrainGraph(named, num_layers=net.num_layers, n_shot=net.n_shot, semantic_enhance=net.semantic_enhance,
                             channel_gamma=net.channel_gamma, unary_gamma=net.unary_gamma)
        self.stem = FrozenStem(named, dev)
        self.anchors = torch.from_numpy(generate_anchors(ratios=tuple(cfg.ANCHOR_RATIOS),
                                                         scales=tuple(cfg.ANCHOR_SCALES))).float().to(dev)
        self.anchors_host = self.anchors.cpu().numpy()
        self.stride = cfg.FEAT_STRIDE[0]
        self.s_im, self.s_info, self.s_sup = im_data.detach().clone(), im_info.detach().float().clone(), support_ims.detach().clone()
        sink = lambda p: trainer._grad_of.get(id(p))  # noqa: E731
        fork = os.environ.get("DANA_TRAIN_FORK_WGRAD", "0") == "1"   # measured: no gain (33.2 vs 32.8 ms), opt-in
        rng_state = np.random.get_state()          # capture draws targets too: leave the caller's RNG stream untouched

        def host_targets(st, gt_host, info_host):
            anchor_t = self.tg.anchor_targets_host(st["qh"], st["qw"], gt_host, info_host, self.anchors_host, self.stride)
            sample = self.tg.proposal_targets_host(st["rois_all"].cpu().numpy(), gt_host)
            return ([torch.from_numpy(np.ascontiguousarray(t)) for t in anchor_t],
                    [torch.from_numpy(np.ascontiguousarray(t)).float() for t in sample])
        gt_host = gt_boxes.detach().float().cpu().numpy()
        info_host = im_info.detach().float().cpu().numpy()
        # AccumulateGrad nodes of the bias parameters are created by the first warm-up pass on the side stream and
        # reused on the capture stream: intended, both are ordered by the waits below
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):              # warm-up outside capture: lazy workspaces, kernel attributes, PE tables
            for _ in range(2):
                autograd_ops.set_direct_grads(sink, None, fork_wgrad=fork)
                st = self.tg.part1(self.stem, self.s_im, self.s_info, self.s_sup, self.anchors, self.stride)
                a_t, smp = host_targets(st, gt_host, info_host)
                out = self.tg.part2(st, [t.to(dev) for t in a_t], [t.to(dev) for t in smp])
                (out[3] + out[4] + out[5] + out[6]).backward()
                autograd_ops.join_forks()
                autograd_ops.set_direct_grads(None, None)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.s_anchor = [t.to(dev) for t in a_t]
        self.s_sample = [t.to(dev) for t in smp]
        pool = torch.cuda.graph_pool_handle()
        self.g1, self.g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        autograd_ops.set_direct_grads(sink, None, fork_wgrad=fork)
        launches0 = ops.LAUNCHES
        try:
            with torch.cuda.graph(self.g1, pool=pool):
                for a in trainer.arenas:
                    a.grad.zero_()
                self.st = self.tg.part1(self.stem, self.s_im, self.s_info, self.s_sup, self.anchors, self.stride)
            with torch.cuda.graph(self.g2, pool=pool):
                out = self.tg.part2(self.st, self.s_anchor, self.s_sample)
                self.losses = out[3:7]
                self.loss = out[3] + out[4] + out[5] + out[6]
                self.loss.backward()
                autograd_ops.join_forks()
        finally:
            autograd_ops.set_direct_grads(None, None)
        self.launches = ops.LAUNCHES - launches0       # own kernels recorded in the two graphs (replayed every step)
        self.out = out
        np.random.set_state(rng_state)
        for a in trainer.arenas:                   # the warm-up / capture passes left gradients behind
            a.grad.zero_()

    def run(self, im_data, im_info, gt_boxes, support_ims):
        import numpy as np
        for dst, src in ((self.s_im, im_data), (self.s_info, im_info), (self.s_sup, support_ims)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        gt_host = gt_boxes.detach().float().cpu().numpy()        # before the replay: a D2H copy waits for the stream
        info_host = im_info.detach().float().cpu().numpy()
        self.g1.replay()
        # anchor targets first (they do not depend on the network: drawn while graph 1 runs), then proposal targets:
        # the order of the reference's numpy RNG draws
        anchor_t = self.tg.anchor_targets_host(self.st["qh"], self.st["qw"], gt_host, info_host, self.anchors_host, self.stride)
        sample = self.tg.proposal_targets_host(self.st["rois_all"].cpu().numpy(), gt_host)
        for dst, src in zip(self.s_anchor, anchor_t):
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src)), non_blocking=True)
        for dst, src in zip(self.s_sample, sample):
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).float(), non_blocking=True)
        self.g2.replay()
        ops.LAUNCHES += self.launches
        return self.out


class SGDTrainer:
    def __init__(self, net, lr=None, momentum=None, weight_decay=None, bucket_bytes=25 << 20, cuda_graph=False):
        self.net = net
        self.cuda_graph = cuda_graph
        self._captured = {}
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
        self._grad_of = {}               # id(param) -> its slice of the gradient arena (autograd_ops' direct sink)
        for a in self.arenas:
            for (_, p), off in zip(a.named, a.offsets):
                self._grad_of[id(p)] = a.grad[off:off + p.numel()].view(p.shape)

    def _mark(self, name):
        if self.events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.events.append((name, e))

    def _finish(self):
        """All-reduce of the arenas after backward (the captured step: not overlapped, ~1 ms on NVSwitch) + SGD."""
        if self.world > 1:
            work = [dist.all_reduce(a.grad[begin:end], op=dist.ReduceOp.SUM, async_op=True)
                    for a in self.arenas for (_, _, begin, end) in a.buckets]
            for w in work:
                w.wait()
        for a in self.arenas:
            ops.sgd_momentum(a.param, a.grad, a.mom, a.lr, self.momentum, a.weight_decay, grad_scale=1.0 / self.world)
            torch.autograd.graph.increment_version([p for _, p in a.named])

    def step_captured(self, im_data, im_info, gt_boxes, num_boxes, support_ims):
        """step() with the device work replayed from two CUDA graphs (_CapturedStep); one capture per input shape."""
        if self.net.n_way != 2 or support_ims.shape[1] != 2 * self.net.n_shot:
            raise ValueError("train mode expects n_way = 2 and 2 * n_shot support crops per image")
        key = (tuple(im_data.shape), tuple(support_ims.shape), tuple(gt_boxes.shape))
        cap = self._captured.get(key)
        if cap is None:
            for a in self.arenas:
                a.rebind_grads()
            try:
                cap = _CapturedStep(self, im_data, im_info, gt_boxes, support_ims)
            except Exception as e:  # noqa: BLE001  -- capture is an optimisation, never a requirement
                import warnings
                warnings.warn("dana_b200: CUDA graph capture of the training step failed (%r); running eagerly" % (e,))
                cap = False
            self._captured[key] = cap
        if cap is False:
            return self.step(im_data, im_info, gt_boxes, num_boxes, support_ims, _eager=True)
        self._mark("start")
        out = cap.run(im_data, im_info, gt_boxes, support_ims)
        self._mark("graphs")
        self._finish()
        self._mark("allreduce+sgd")
        return cap.loss.detach().clone(), tuple(l.detach().clone() for l in cap.losses)

    def step(self, im_data, im_info, gt_boxes, num_boxes, support_ims, _eager=False):
        """-> (loss, (rpn_loss_cls, rpn_loss_box, RCNN_loss_cls, RCNN_loss_bbox)) as detached 0-dim CUDA tensors."""
        if self.cuda_graph and not _eager:
            return self.step_captured(im_data, im_info, gt_boxes, num_boxes, support_ims)
        for a in self.arenas:
            a.rebind_grads()
            a.grad.zero_()
        self._mark("start")
        # weight gradients of the conv / linear functions are accumulated by the unpack kernel straight into the arena;
        # completion is signalled to the all-reduce like a post-accumulate hook would
        autograd_ops.set_direct_grads(lambda p: self._grad_of.get(id(p)),
                                      self.sync._on_grad if self.sync is not None else None)
        out = self.net(im_data, im_info, gt_boxes, num_boxes, support_ims)
        losses = out[3:7]
        loss = losses[0].mean() + losses[1].mean() + losses[2].mean() + losses[3].mean()      # train.py:136-137
        self._mark("forward")
        if self.sync is not None:
            self.sync.arm()
        loss.backward()
        autograd_ops.set_direct_grads(None, None)
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
