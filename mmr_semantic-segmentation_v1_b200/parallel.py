"""Batch-sharded data-parallel training: one process per GPU, gradients all-reduced with NCCL
over NVLink in buckets that start while the backward plan is still running.

The reference is single-process; semantics follow torch DDP defaults: per-replica
BatchNorm statistics, loss mean per replica, gradients averaged over ranks, parameters
broadcast from rank 0 at construction (or, when the module is moved to its device after wrapping, before
the first forward).
"""
import os

import torch
import torch.distributed as dist


class DistributedDataParallel(torch.nn.Module):
    def __init__(self, module, bucket_mb=None, process_group=None, tail_mb=None):
        super().__init__()
        self.module = module
        self.pg = process_group
        # MMR_DDP_BUCKET_MB / MMR_DDP_TAIL_MB: measurement knobs (defaults: 24 MB buckets, 2 MB tail bucket: measured
        # on 2 and 8 GPUs, 8 -> 24 MB buckets: 11.34 -> 11.20 ms and 11.59 -> 11.48 ms per step; fewer cuts of the backward graph)
        bucket_mb = float(os.environ.get("MMR_DDP_BUCKET_MB", "24")) if bucket_mb is None else bucket_mb
        tail_mb = float(os.environ.get("MMR_DDP_TAIL_MB", "2")) if tail_mb is None else tail_mb
        self.bucket_elems = int(bucket_mb * 1024 * 1024 / 4)
        self.tail_elems = int(tail_mb * 1024 * 1024 / 4)
        self.world = dist.get_world_size(self.pg) if dist.is_initialized() else 1
        self._buckets = None
        self._side = None
        self._cuts = {}
        module._on_grads_ready = self._ready
        module._after_backward = self._finish
        module._grad_cuts = self.cuts
        # torch DDP broadcasts rank 0's parameters and buffers in its constructor: do the same as soon as the
        # module sits on its device (otherwise step 1's forward would run on per-rank weights and the
        # broadcast would land in the middle of the first backward)
        first = next(iter(module.parameters()), None)
        if first is not None and (first.is_cuda or not torch.cuda.is_available()):
            module._ensure_flat(first.device)
            self._setup()

    def forward(self, *a, **k):
        if self._buckets is None:       # the module was moved to its device after wrapping
            first = next(iter(self.module.parameters()))
            self.module._ensure_flat(first.device)
            self._setup()
        elif self.module._gflat is not self._gflat_seen:   # re-flattened (moved again): rebuild the buckets
            self._cuts = {}
            self._setup()
        return self.module(*a, **k)

    # ---- bucket plan: contiguous slices of the flat gradient buffer, closed in backward order
    def plan_buckets(self, names, offsets, sizes, total):
        """names in flat order; returns [(lo, hi, set(names))] covering [0,total).  Buckets are closed in
        backward order (from the end of the flat buffer); the one that completes LAST -- the front of the
        buffer, the first layers of the encoder -- is cut down to `tail_elems`, because its all-reduce is the
        only one nothing is left to overlap with."""
        buckets, hi, cur = [], total, set()
        for name in reversed(names):
            cur.add(name)
            lo = offsets[name]
            if hi - lo >= self.bucket_elems:
                buckets.append((lo, hi, cur))
                hi, cur = lo, set()
        if cur:
            buckets.append((0, hi, cur))
        lo, hi, cur = buckets[-1]
        if hi - lo > 2 * self.tail_elems:
            front = [n for n in names if n in cur and offsets[n] < lo + self.tail_elems]
            rest = cur - set(front)
            if front and rest:
                cut = min(offsets[n] for n in rest)
                buckets[-1:] = [(cut, hi, rest), (lo, cut, set(front))]
        return buckets

    def _setup(self):
        m = self.module
        names = m._flat_names
        sizes = {n: p.numel() for n, p in m.named_parameters()}
        self._buckets = self.plan_buckets(names, m._flat_offsets, sizes, m._gflat.numel())
        self._pending = None
        self._gflat_seen = m._gflat
        # high priority: a bucket's all-reduce gets its few CTAs at the next kernel boundary of the backward
        # instead of waiting for idle SMs behind the persistent conv kernels
        self._side = torch.cuda.Stream(device=m._gflat.device, priority=-1) if m._gflat.is_cuda else None
        self._works = []
        # broadcast_parameters: rank 0's weights and BN buffers
        if self.world > 1:
            dist.broadcast(m._flat, 0, group=self.pg)
            for b in m.buffers():
                dist.broadcast(b, 0, group=self.pg)

    def cuts(self, hooks):
        """hooks: the engine's [(position in the backward launch list, [parameter names])].  Returns the
        positions at which some bucket has all its gradients: the backward plan is replayed as one CUDA
        graph per stretch between two of them, and the bucket's all-reduce starts at the cut."""
        if self._buckets is None:
            self._setup()
        key = id(hooks)
        if self._cuts.get(key) is None:
            ready_at = {}
            for pos, names in hooks:
                for n in names:
                    ready_at.setdefault(n, pos)
            last = max(pos for pos, _ in hooks)
            self._cuts[key] = {max(ready_at.get(n, last) for n in names) for _, _, names in self._buckets}
        return self._cuts[key]

    def sync_parameters(self):
        """Call once after the model is on its device and flattened."""
        if self._buckets is None:
            self._setup()

    def _ready(self, names):
        if self._buckets is None:
            self._setup()
        if self._pending is None:
            self._pending = [set(b[2]) for b in self._buckets]
        for i, pend in enumerate(self._pending):
            if not pend:
                continue
            pend.difference_update(names)
            if not pend:
                self._launch(i)

    def _launch(self, i):
        lo, hi, _ = self._buckets[i]
        g = self.module._gflat[lo:hi]
        if self.world == 1 or os.environ.get("MMR_DDP_NO_COMM"):   # NO_COMM: attribution runs (N ranks, no exchange)
            return
        if self._side is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._side.wait_event(ev)
            with torch.cuda.stream(self._side):
                self._works.append(dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
        else:  # gloo (CPU tests): no AVG
            w = dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._works.append((w, g))

    def _finish(self):
        if self._pending is not None:
            for i, pend in enumerate(self._pending):
                if pend:  # parameters that received no gradient callback: reduce anyway
                    self._launch(i)
        for w in self._works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world)
            else:
                w.wait()
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        self._works, self._pending = [], None
