"""Gradient synchronisation of the training step (groundwork for row a15 -- the training kernels
themselves are not built yet).

The reference trains with `nn.DataParallel` (train.py:104-105): one process, replicas scatter the batch, backward sums
the replica gradients on GPU 0, and the four losses are `.mean()`-ed over replicas (train.py:131-138) -- i.e. every
parameter receives the AVERAGE of the per-replica gradients.  With one process per GPU the same result is one
all-reduce (sum) per step divided by the world size, after which every rank applies the identical SGD step
(train.py:89).

`BucketedGradAllReduce` packs the gradients into flat fp32 buckets in REVERSE parameter order (roughly the order
backward produces them), launches the asynchronous all-reduce of a bucket the moment its last gradient has been
accumulated (post-accumulate-grad hooks), and unpacks on `finish()`: the collective of the head's gradients overlaps
the backward of the trunk.  37.1 M fp32 = 148 MB for res50 -> six 25 MB buckets.  NCCL over NVLink on the
GPU box, gloo in the CPU tests (tests/test_host_logic.py)."""
import torch
import torch.distributed as dist


class BucketedGradAllReduce:
    def __init__(self, params, bucket_bytes=25 << 20, process_group=None, average=True):
        self.group = process_group
        self.average = average
        self.params = [p for p in params if p.requires_grad]
        # buckets in reverse registration order, closed when they reach bucket_bytes
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self._flat = [None] * len(self.buckets)
        self._pending = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._hooks = []
        self._armed = False
        self._next = 0

    # ---- overlap mode: call arm() before backward, finish() after it
    def arm(self):
        """Register per-parameter hooks for ONE backward pass; a bucket is reduced as soon as all of its parameters
        have their gradient."""
        self.disarm()
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)
        self._next = 0                      # buckets [0, _next) have been launched
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._armed = True

    def disarm(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        self._armed = False

    def _on_grad(self, p):
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] < 0:
            raise RuntimeError("BucketedGradAllReduce: a gradient arrived twice after one arm() -- call arm() before "
                               "every backward pass")
        # Collectives must be issued in the same order on every rank, and hook completion order is not that: a
        # parameter unused on one rank never fires its hook there.  Buckets are therefore launched strictly in index
        # order (as DDP does): bucket i goes out once it is complete AND buckets 0..i-1 have gone out.
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def _launch(self, bi):
        bucket = self.buckets[bi]
        dev = bucket[0].device
        n = sum(p.numel() for p in bucket)
        flat = self._flat[bi]
        if flat is None or flat.device != dev or flat.numel() != n:
            flat = self._flat[bi] = torch.empty(n, dtype=torch.float32, device=dev)
        o = 0
        for p in bucket:
            g = p.grad if p.grad is not None else torch.zeros_like(p)     # unused parameters contribute zeros
            flat[o:o + p.numel()].copy_(g.reshape(-1))
            o += p.numel()
        self._work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket (launching those whose hooks never fired, e.g. parameters without a gradient this
        step) and write the averaged gradients back into `.grad`."""
        world = dist.get_world_size(self.group)
        for bi, bucket in enumerate(self.buckets):      # the rest, still in index order
            if self._work[bi] is None:
                self._launch(bi)
        self._next = len(self.buckets)
        for bi, bucket in enumerate(self.buckets):
            self._work[bi].wait()
            flat = self._flat[bi]
            if self.average:
                flat.div_(world)
            o = 0
            for p in bucket:
                g = flat[o:o + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                o += p.numel()
        self.disarm()

    # ---- simple mode: everything after backward
    def reduce_now(self):
        self.disarm()
        self._work = [None] * len(self.buckets)
        self.finish()
