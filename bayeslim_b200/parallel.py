"""
Multi-GPU execution of the RIME path: one process per GPU (torch.distributed, NCCL over
NVLink/NVSwitch), work sharded by time and/or baseline group, and ONE all-reduce of the
parameter gradients per backward.

This replaces the reference's single-process data parallelism
(``optim.DistributedLogProb``, bayeslim/optim.py:1391-1628), which broadcasts parameters
with ``tensor.to(device)`` (:1517-1523) and sums gradients on one root device
(:1558-1564).  The forward needs no communication: the (time, baseline-group) units of
``RIME``'s own minibatch grid (rime_model.py:261-274) produce disjoint visibilities, and each
rank scores its own shard of the data.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_units(n_units, rank, world_size, weights=None):
    """Contiguous split of `n_units` work units over ranks, balanced by `weights`
    (e.g. sources above the horizon per time).  Returns the index list of this rank."""
    if weights is None:
        weights = np.ones(n_units)
    w = np.asarray(weights, dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(w)])
    total = csum[-1]
    bounds = [int(np.searchsorted(csum, total * r / world_size, side='left'))
              for r in range(world_size + 1)]
    bounds[0], bounds[-1] = 0, n_units
    for r in range(1, world_size + 1):
        bounds[r] = max(bounds[r], bounds[r - 1])
    return list(range(bounds[rank], bounds[rank + 1]))


def shard_rime_batches(rime, rank, world_size):
    """Batch indices (rime.batch_idx values) owned by `rank`: the reference's minibatch grid
    [(time group, baseline group)] split contiguously."""
    return shard_units(rime.Nbatch, rank, world_size)


def allreduce_gradients(params, group=None):
    """Sum the .grad of every parameter over ranks with a single flat all-reduce.

    params: iterable of tensors (leaf parameters).  Missing grads count as zero.  The bucket is
    float64 if any gradient is, else float32; NCCL's ring/tree order is fixed for a given
    world size, so the result is reproducible run to run."""
    params = [p for p in params if p is not None]
    if not params:
        return 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    dev = params[0].device
    dtype = torch.float64 if any(p.dtype == torch.float64 for p in params) else torch.float32
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes), dtype=dtype, device=dev)
    off = 0
    for p, n in zip(params, sizes):
        if p.grad is not None:
            flat[off:off + n] = p.grad.reshape(-1).to(dtype)
        off += n
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p, n in zip(params, sizes):
        g = flat[off:off + n].reshape(p.shape).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return flat.numel() * flat.element_size()


def broadcast_parameters(params, src=0, group=None):
    """Make every rank start from rank `src`'s parameter values (one flat broadcast)."""
    params = [p for p in params if p is not None]
    if not params or not (dist.is_available() and dist.is_initialized()):
        return
    flat = torch.cat([p.detach().reshape(-1).to(torch.float64) for p in params])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].reshape(p.shape).to(p.dtype))
            off += n
