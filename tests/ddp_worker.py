"""torchrun worker of tests/test_ddp_gpu.py: N ranks, one GPU each, NCCL.

Checks, on every rank, that the gradients `mmrseg_b200.parallel.DistributedDataParallel` leaves after
backward equal the rank-mean of the per-shard gradients computed WITHOUT any collective (every rank recomputes
all N shards locally on an unwrapped copy of the model: per-replica BatchNorm statistics, loss mean per
replica, torch DDP's semantics), that the parameters were broadcast from rank 0 in the constructor, and that
two DDP training steps keep the replicas bit-identical."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    from mmrseg_b200.models import UnetPlusPlus
    from mmrseg_b200.optim import FusedAdam
    from mmrseg_b200.parallel import DistributedDataParallel
    from tests.helpers import rel, synthetic_batch
    ds = os.environ.get("MMR_DDP_DS", "0") == "1"
    per_rank, hw, classes = 2, 128, 2
    crit = DiceCrossEntropyLoss(0.5)

    def loss_of(out, y):
        outs = out if isinstance(out, list) else [out]
        return sum(crit(o, y) for o in outs) / len(outs)

    # different initial weights per rank: the constructor must broadcast rank 0's
    torch.manual_seed(100 + rank)
    model = UnetPlusPlus("resnet18", classes=classes, deep_supervision=ds).to(dev).train()
    ddp = DistributedDataParallel(model, bucket_mb=4.0)
    torch.manual_seed(100)
    ref0 = UnetPlusPlus("resnet18", classes=classes, deep_supervision=ds).to(dev).train()
    for (k, a), (_, b) in zip(model.state_dict().items(), ref0.state_dict().items()):
        assert torch.equal(a, b), "rank %d: %s was not broadcast from rank 0" % (rank, k)

    x, y = synthetic_batch(per_rank * world, classes, hw, hw)       # the global batch, identical on every rank
    x, y = x.to(dev), y.to(dev)
    shard = slice(rank * per_rank, (rank + 1) * per_rank)
    # (1) DDP backward on this rank's shard
    loss_of(ddp(x[shard]), y[shard]).backward()
    got = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    # (2) no collective: every shard on the unwrapped reference copy, gradients averaged in rank order
    want = None
    for r in range(world):
        for p in ref0.parameters():
            p.grad = None
        sl = slice(r * per_rank, (r + 1) * per_rank)
        loss_of(ref0(x[sl]), y[sl]).backward()
        g = {k: p.grad.detach().clone() for k, p in ref0.named_parameters()}
        want = g if want is None else {k: want[k] + g[k] for k in g}
    worst = 0.0
    for k in want:
        w = want[k] / world
        worst = max(worst, rel(got[k], w))
        assert torch.allclose(got[k], w, rtol=1e-5, atol=1e-8), (rank, k, rel(got[k], w))
    # (3) two optimiser steps under DDP: replicas stay identical (the all-reduced gradients are)
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    for _ in range(2):
        for p in model.parameters():
            p.grad = None
        loss_of(ddp(x[shard]), y[shard]).backward()
        opt.step()
    flat = model.flat_parameters()[0]
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for r in range(1, world):
        assert torch.equal(gathered[0], gathered[r]), "replica %d diverged from replica 0" % r
    dist.barrier()
    if rank == 0:
        print("ddp ok: world %d, deep_supervision %s, worst grad rel err %.2e" % (world, ds, worst), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
