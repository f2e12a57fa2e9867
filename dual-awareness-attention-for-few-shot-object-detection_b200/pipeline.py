"""Host-side input pipeline helper: overlaps the host->device copy of the next episode batch with the
forward of the current one (the reference copies into persistent holders and then runs, inference.py:91-103,
which serialises a ~58 MB PCIe copy in front of every step).  Two sets of device holders, one copy stream."""
import torch


class EpisodePrefetcher:
    """for tensors in prefetcher.run(batches): model(*tensors) ...

    `batches` yields tuples of pinned host tensors with identical shapes.  Each step receives device
    tensors whose H2D copy was issued one step earlier on a side stream."""

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = depth
        self.stream = torch.cuda.Stream(device=self.device)
        self.holders = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]

    def _issue(self, slot, host_tensors):
        if self.holders[slot] is None:
            self.holders[slot] = [torch.empty_like(t, device=self.device) for t in host_tensors]
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.free[slot])          # the step that last used this slot is done
            for dst, src in zip(self.holders[slot], host_tensors):
                dst.copy_(src, non_blocking=True)
            self.ready[slot].record(self.stream)

    def run(self, batches):
        it = iter(batches)
        pending = []
        main = torch.cuda.current_stream(self.device)
        for slot in range(self.depth):
            self.free[slot].record(main)
        nxt = 0
        for _ in range(self.depth - 1):
            b = next(it, None)
            if b is None:
                break
            self._issue(nxt % self.depth, b)
            pending.append(nxt % self.depth)
            nxt += 1
        while pending:
            b = next(it, None)
            if b is not None:
                self._issue(nxt % self.depth, b)
                pending.append(nxt % self.depth)
                nxt += 1
            slot = pending.pop(0)
            main.wait_event(self.ready[slot])
            yield self.holders[slot]
            self.free[slot].record(main)                     # recorded after the consumer enqueued its work


class ResultFetcher:
    """Device -> host read of a step's results without stalling the GPU: the copies go into pinned host buffers
    (`depth` rotating sets) right behind the step on the main stream, and the host waits for them only after the
    NEXT step has been enqueued.  The reference reads every result with blocking `.cpu()` calls (inference.py:104-107),
    which leaves the GPU idle while python prepares the following step.

        fetch = ResultFetcher(device)
        pending = None
        for tensors in prefetcher.run(batches):
            out = model(*tensors)[:3]
            ticket = fetch.start(out)            # async D2H of (rois, cls_prob, bbox_pred)
            if pending is not None:
                host = fetch.wait(pending)       # results of the previous step, now on the host
            pending = ticket
        host = fetch.wait(pending)
    """

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.depth = depth
        self.buffers = [None] * depth
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.count = 0

    def start(self, tensors):
        slot = self.count % self.depth
        self.count += 1
        if self.buffers[slot] is None or any(b.shape != t.shape or b.dtype != t.dtype
                                             for b, t in zip(self.buffers[slot], tensors)):
            self.buffers[slot] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in tensors]
        for dst, src in zip(self.buffers[slot], tensors):
            dst.copy_(src, non_blocking=True)
        self.events[slot].record(torch.cuda.current_stream(self.device))
        return slot

    def wait(self, slot):
        """Host tensors of the step that `start` returned `slot` for (valid until the slot comes round again)."""
        self.events[slot].synchronize()
        return self.buffers[slot]
