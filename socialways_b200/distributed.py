"""Multi-GPU plumbing of the training step: one process per GPU, scenes of every
mini-batch sharded over ranks, one flat-buffer all-reduce (NCCL over NVLink; gloo in the CPU tests)
per optimiser step.  The reference has no distributed code; the parity rule is that the summed
per-rank gradients equal the single-process gradients of nn.MSELoss over the GLOBAL mini-batch
(train.py:484-488,514-516): every rank divides its sum of squares by the global element count.

Backend-agnostic on purpose (pure host logic + torch.distributed), so it is unit-tested with
world_size-2 gloo on CPU (tests/test_distributed_cpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_scenes(sub_batches, world_size, rank):
    """Contiguous block of scenes for `rank`, balanced by agent count (ties to the pairwise cost for
    equal-sized scenes).  Returns (agent_lo, agent_hi, local sub_batches rebased to agent_lo)."""
    sb = np.asarray(sub_batches, dtype=np.int64).reshape(-1, 2)
    if world_size == 1:
        return int(sb[0, 0]), int(sb[-1, 1]), sb - sb[0, 0]
    sizes = sb[:, 1] - sb[:, 0]
    total = int(sizes.sum())
    ends = np.cumsum(sizes)
    # scene s goes to the rank whose [r, r+1) * total / world interval contains its mid-point
    mids = ends - sizes / 2.0
    owner = np.minimum((mids * world_size / total).astype(np.int64), world_size - 1)
    mine = np.nonzero(owner == rank)[0]
    if len(mine) == 0:
        return 0, 0, np.zeros((0, 2), dtype=np.int64)
    lo, hi = int(sb[mine[0], 0]), int(sb[mine[-1], 1])
    return lo, hi, sb[mine] - lo


def global_mse(a, b, global_numel):
    """This rank's share of nn.MSELoss over the global batch: local sum of squares / GLOBAL element count."""
    return ((a - b) ** 2).sum() / float(global_numel)


def allreduce_grads(params, world_size):
    """Sum the gradients of `params` over ranks through ONE flat fp32 buffer (G: 86 122 floats,
    D: 27 939 floats -- latency-bound; params without a gradient contribute zeros)."""
    if world_size == 1:
        return
    params = [p for p in params]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def allreduce_scalars(values, device, world_size):
    if world_size == 1:
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
