"""Host -> device feed for the train / eval loops.

The reference moves every batch with a blocking `.cuda()` on the compute stream after its CPU
pre-processing (SU/ModelTraining.py:579-587, ED/Main_MMR_SegModel.py:690-694), so the copy and the step
never overlap.  `DevicePrefetcher` wraps any iterable of (pinned) host tensors and keeps ONE batch in
flight on a copy stream: batch i+1 crosses PCIe / NVLink-C2C while the kernels of batch i run.  The two
batches live in two fixed sets of device buffers (no allocator traffic on the hot path); a set is
overwritten only after the stream that consumed it has passed the point where the consumer asked for
the next batch.  With uint8 HWC frames (`model.set_input_normalization`) the copy is 4x smaller as well.
"""
import torch


class DevicePrefetcher:
    def __init__(self, loader, device=None):
        self.loader = loader
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs = [None, None]
        self._free = [None, None]

    def _stage(self, batch, slot):
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])
            old = self._bufs[slot]
            dst = []
            for i, t in enumerate(batch):
                if not torch.is_tensor(t):
                    dst.append(t)
                    continue
                d = old[i] if old is not None and i < len(old) and torch.is_tensor(old[i]) else None
                if d is None or d.shape != t.shape or d.dtype != t.dtype:
                    d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                d.copy_(t, non_blocking=True)
                dst.append(d)
            self._bufs[slot] = dst
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return tuple(dst), ev, slot

    def __iter__(self):
        it = iter(self.loader)
        try:
            nxt = self._stage(next(it), 0)
        except StopIteration:
            return
        while nxt is not None:
            (batch, ev, slot), nxt = nxt, None
            try:
                nxt = self._stage(next(it), slot ^ 1)      # the next batch starts moving before this one is used
            except StopIteration:
                pass
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            yield batch
            done = torch.cuda.Event()                       # everything the consumer launched on this batch
            done.record(torch.cuda.current_stream(self.device))
            self._free[slot] = done

    def __len__(self):
        return len(self.loader)


class ResultReader:
    """Per-step device -> host read of a small result (the loss, a confusion matrix) that does not drain the
    GPU queue.  `loss.item()` after every step (SU/ModelTraining.py:619-622) stalls the host until the step has
    finished and the GPU then idles while the host enqueues the next step (~0.4 ms of an 11 ms step).  `push(t)`
    enqueues an asynchronous copy of `t` into pinned memory behind the step's kernels and returns the values of
    the steps that have completed `lag` pushes ago (waiting on their copy events only); `flush()` returns the
    rest.  Every step's value still reaches the host, in order, one step late."""

    def __init__(self, lag=1):
        self.lag = lag
        self._slots = []      # (pinned tensor, event), oldest first
        self._pool = []

    def push(self, t):
        t = t.detach()
        host = None
        for i, h in enumerate(self._pool):
            if h.shape == t.shape and h.dtype == t.dtype:
                host = self._pool.pop(i)
                break
        if host is None:
            host = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        host.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(t.device))
        self._slots.append((host, ev))
        out = []
        while len(self._slots) > self.lag:
            out.append(self._take())
        return out

    def _take(self):
        host, ev = self._slots.pop(0)
        ev.synchronize()
        val = host.clone()
        self._pool.append(host)
        return val

    def flush(self):
        out = []
        while self._slots:
            out.append(self._take())
        return out
