"""
Multi-GPU execution of the RIME path: one process per GPU (torch.distributed, NCCL over
NVLink/NVSwitch), work sharded by time and/or baseline group, and one all-reduce of the
parameter gradients per backward (in place on the large cotangents, no staging copies).

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


SMALL_GRAD_BYTES = 1 << 20      # gradients below this size share one bucket


def allreduce_gradients(params, group=None):
    """Sum the .grad of every parameter over ranks.  Large gradients (the sky and beam-map
    cotangents, 0.1 - 3 GB) are reduced IN PLACE, one all-reduce each, with no staging copy; the
    small ones (antenna positions, Airy diameters) share one flat bucket.  All reductions are
    launched asynchronously on NCCL's stream and waited for together.  Missing grads count as
    zero.  NCCL's ring / tree order is fixed for a given world size, so the result is
    reproducible run to run.  Returns the number of bytes reduced."""
    params = [p for p in params if p is not None]
    if not params:
        return 0
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    big = [p for p in params if p.grad.numel() * p.grad.element_size() >= SMALL_GRAD_BYTES
           and p.grad.is_contiguous()]
    small = [p for p in params if not any(p is q for q in big)]
    works, nbytes = [], 0
    for p in big:
        works.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group, async_op=True))
        nbytes += p.grad.numel() * p.grad.element_size()
    flat = None
    if small:
        dtype = torch.float64 if any(p.dtype == torch.float64 for p in small) else torch.float32
        flat = torch.cat([p.grad.reshape(-1).to(dtype) for p in small])
        works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True))
        nbytes += flat.numel() * flat.element_size()
    for w in works:
        w.wait()
    if flat is not None:
        off = 0
        for p in small:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].reshape(p.shape).to(p.dtype))
            off += n
    return nbytes


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
