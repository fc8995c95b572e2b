"""N>1 host logic on CPU: world_size-2 gloo processes run the bucketed gradient all-reduce of
parallel.DistributedDataParallel on a stand-in model (flat gradient buffer + readiness callbacks
in backward order) and must end with the rank-mean in every bucket."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakePlanModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Parameter(torch.zeros(1000))
        self.b = torch.nn.Parameter(torch.zeros(3000))
        self.c = torch.nn.Parameter(torch.zeros(10))
        self._on_grads_ready = None
        self._after_backward = None
        self._grad_cuts = None
        names = ["a", "b", "c"]
        offs, tot = {}, 0
        for n in names:
            offs[n] = tot
            tot += (getattr(self, n).numel() + 3) // 4 * 4
        self._flat = torch.zeros(tot)
        self._gflat = torch.zeros(tot)
        self._flat_names, self._flat_offsets = names, offs

    def _ensure_flat(self, device):
        pass


def _worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmrseg_b200.parallel import DistributedDataParallel
    m = _FakePlanModel()
    m._flat.fill_(float(rank + 1))
    ddp = DistributedDataParallel(m, bucket_mb=0.008)      # ~2k floats per bucket -> several buckets
    assert torch.all(m._flat == 1.0)                        # rank 0's parameters everywhere, from the constructor on
    ddp.sync_parameters()                                   # (idempotent)
    assert torch.all(m._flat == 1.0)
    assert len(ddp._buckets) >= 2
    covered = sorted((lo, hi) for lo, hi, _ in ddp._buckets)
    assert covered[0][0] == 0 and covered[-1][1] == m._gflat.numel()
    assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    # the bucket that completes last (front of the buffer) is cut down to the tail size
    tail = DistributedDataParallel(_FakePlanModel(), bucket_mb=0.016, tail_mb=0.002)
    tb = tail.plan_buckets(m._flat_names, m._flat_offsets, None, m._gflat.numel())
    cov = sorted((lo, hi) for lo, hi, _ in tb)
    assert cov[0][0] == 0 and cov[-1][1] == m._gflat.numel()
    assert all(cov[i][1] == cov[i + 1][0] for i in range(len(cov) - 1))
    assert tb[-1][0] == 0 and tb[-1][2] == {"a"} and tb[-1][1] == 1000      # only the first tensor is left in it
    for step in range(2):
        m._gflat.copy_(torch.arange(m._gflat.numel(), dtype=torch.float32) * (rank + 1))
        for names in (["c"], ["b"], ["a"]):                 # backward order: last parameter first
            m._on_grads_ready(names)
        m._after_backward()
        want = torch.arange(m._gflat.numel(), dtype=torch.float32) * (sum(range(1, world + 1)) / world)
        assert torch.allclose(m._gflat, want), (rank, step)
    # cut positions of the engine's backward launch list: every bucket closes at the hook of its last
    # parameter; coalescing the hooks between two cuts (engine.Engine.backward) still reduces everything
    hooks = [(3, ["c"]), (7, ["b"]), (9, ["a"])]
    cuts = m._grad_cuts(hooks)
    assert cuts and cuts <= {3, 7, 9} and 9 in cuts
    m._gflat.copy_(torch.arange(m._gflat.numel(), dtype=torch.float32) * (rank + 1))
    names = []
    for pos, ns in hooks:
        names += ns
        if pos in cuts:
            m._on_grads_ready(names)
            names = []
    assert not names
    m._after_backward()
    assert torch.allclose(m._gflat, want), rank
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
