"""Multi-GPU: problem instances (batch columns) are independent through all K layers, so GPU g of G owns
columns [g*B/G, (g+1)*B/G) of X/Z/E/L/T; A and the weights are replicated.

Inference needs no communication.  Training needs one sum-allreduce of the parameter gradients per step
(NCCL over NVLink on GPUs; gloo in the CPU tests).  The reference has no distributed code at all (only a
commented-out nn.DataParallel, main_syn_l1l1_scalar.py:238)."""
import torch
import torch.distributed as dist


def column_shard(B, rank, world):
    """Half-open column range of `rank`; ranges tile [0, B) and differ in size by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(B, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_gradients(parameters, group=None, average=False, bucket_bytes=64 << 20):
    """Sum (or average) .grad of `parameters` across ranks, coalesced into flat fp32 buckets so the
    per-layer (1,1) scalars do not each pay a collective launch.  Returns the number of collectives."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    grads = [p.grad for p in parameters if p.grad is not None]
    calls = 0
    i = 0
    while i < len(grads):
        bucket, size = [], 0
        while i < len(grads) and (not bucket or size + grads[i].numel() * 4 <= bucket_bytes):
            bucket.append(grads[i])
            size += grads[i].numel() * 4
            i += 1
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        calls += 1
        if average:
            flat.div_(world)
        views, off = [], 0
        for g in bucket:
            n = g.numel()
            views.append(flat[off:off + n].view_as(g))
            off += n
        torch._foreach_copy_(bucket, views)           # one multi-tensor launch instead of one copy per parameter
    return calls


class ShardedTrainer(object):
    """Data-parallel training step over column shards: each rank runs forward+backward on its own columns
    with a loss normalised by the GLOBAL batch, gradients are sum-allreduced, every rank takes the same
    optimizer step (weights stay replicated without a broadcast)."""

    def __init__(self, model, optimizer, group=None):
        self.model, self.optimizer, self.group = model, optimizer, group

    def step(self, x_local, loss_fn, global_batch):
        """loss_fn(outputs, x_local) must return the SUM over local columns; it is divided by
        `global_batch` here so the allreduced gradient equals the single-process gradient of the mean."""
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(x_local)
        loss = loss_fn(out, x_local) / float(global_batch)
        loss.backward()
        allreduce_gradients(list(self.model.parameters()), self.group)
        self.optimizer.step()
        total = loss.detach().clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(total, group=self.group)
        return total
