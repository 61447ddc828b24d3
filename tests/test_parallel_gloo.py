"""
N > 1 path on CPU: world_size-2 ``gloo`` process group.  Each rank simulates its shard of the
time axis (kernels replaced by the test double, tests/cpu_double.py), the gradients are summed
with ``parallel.allreduce_gradients`` (one flat all-reduce), and every rank must end up with
the single-process gradient -- the invariant shown by the reference's
minibatching_and_distributed_training notebook (full batch == distributed).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from bayeslim_b200 import parallel
    from tests import model_cases as mc
    from tests.cpu_double import emulated_kernels
    from tests.oracle_cases import load
    g = load("rime_pixel_interp")
    with emulated_kernels():
        rime, leaves = mc.build_pixel_interp(g, 'cpu')
        # shard the reference's own minibatch grid: one time per batch
        rime.setup_sim_times([np.asarray([t]) for t in g["times"]])
        mine = parallel.shard_rime_batches(rime, rank, world)
        G = torch.as_tensor(g["G"])
        params = [leaves["sky"], leaves["beam"], leaves["antvecs"]]
        for b in mine:
            rime.batch_idx = b
            V = rime().data
            Gb = G[:, :, :, rime.time_group_id:rime.time_group_id + 1]
            torch.sum(Gb.real * V.real + Gb.imag * V.imag).backward()
        nbytes = parallel.allreduce_gradients(params)
    assert nbytes == sum(p.numel() for p in params) * 8
    torch.save([p.grad.clone() for p in params], os.path.join(out_dir, "grads_%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process(tmp_path):
    from tests.oracle_cases import load
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = load("rime_pixel_interp")
    g0 = torch.load(os.path.join(tmp_path, "grads_0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "grads_1.pt"))
    for a, b, key in zip(g0, g1, ("grad_sky", "grad_beam", "grad_antvecs")):
        assert torch.equal(a, b)                      # every rank holds the reduced gradient
        ref = torch.as_tensor(g[key])
        assert float((a - ref).abs().max() / ref.abs().max()) < 1e-9, key


def _worker_bl_groups(rank, world, port, out_dir):
    """Rank r simulates baseline group r of the reference's minibatch grid for all times."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from bayeslim_b200 import parallel
    from tests import model_cases as mc
    from tests.cpu_double import emulated_kernels
    from tests.oracle_cases import load
    g = load("rime_pixel_interp")
    with emulated_kernels():
        rime, leaves = mc.build_pixel_interp(g, 'cpu')
        bls = list(rime.sim_bls)
        half = len(bls) // 2
        rime.setup_sim_bls([bls[:half], bls[half:]])
        assert rime.Nbatch == 2
        G = torch.as_tensor(g["G"])
        params = [leaves["sky"], leaves["beam"], leaves["antvecs"]]
        rime.batch_idx = parallel.shard_rime_batches(rime, rank, world)[0]
        V = rime().data
        sl = slice(0, half) if rank == 0 else slice(half, len(bls))
        Gb = G[:, :, sl]
        torch.sum(Gb.real * V.real + Gb.imag * V.imag).backward()
        parallel.allreduce_gradients(params)
    torch.save([p.grad.clone() for p in params], os.path.join(out_dir, "bgrads_%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_baseline_group_sharding(tmp_path):
    """Baseline-group sharding (BASELINE config 5): the summed gradients of two ranks that each
    own one baseline group equal the single-process golden gradients."""
    from tests.oracle_cases import load
    port = _free_port()
    mp.spawn(_worker_bl_groups, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = load("rime_pixel_interp")
    g0 = torch.load(os.path.join(tmp_path, "bgrads_0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "bgrads_1.pt"))
    for a, b, key in zip(g0, g1, ("grad_sky", "grad_beam", "grad_antvecs")):
        assert torch.equal(a, b)
        ref = torch.as_tensor(g[key])
        assert float((a - ref).abs().max() / ref.abs().max()) < 1e-9, key


def test_allreduce_is_noop_without_process_group():
    sys.path.insert(0, ROOT)
    from bayeslim_b200 import parallel
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert parallel.allreduce_gradients([p]) == 0
    assert (p.grad == 2).all()
