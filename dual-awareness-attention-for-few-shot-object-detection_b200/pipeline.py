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
