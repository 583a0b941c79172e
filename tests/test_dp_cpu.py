"""CPU (gloo, world_size 2): host-side logic of the data-parallel path -- episode sharding,
the mean all-reduce wrapper and the GradAverager call pattern of one update."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcrl_b200.parallel import GradAverager, allreduce_mean, local_batch, shard_of_episode


def test_shards_are_balanced_and_disjoint():
    for world in (1, 2, 4, 8):
        owners = [shard_of_episode(e, world) for e in range(1000)]
        counts = np.bincount(owners, minlength=world)
        assert counts.max() - counts.min() <= 1 and set(owners) == set(range(world))
    assert local_batch(65536, 8) == 8192
    with pytest.raises(ValueError):
        local_batch(100, 8)


class _FakeAgent:
    """Stands in for the CUDA agent: flat 'gradient buffers' that live on the CPU."""

    def __init__(self, rank):
        self.grads = {0: torch.full((1000,), float(rank + 1)), 1: torch.arange(10.0) * (rank + 1)}
        self.metrics = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0]) * (rank + 1)

    def grad_tensor(self, net):
        return self.grads[net]

    def metrics_tensor(self):
        return self.metrics


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = torch.tensor([float(rank), 10.0 * (rank + 1)])
        allreduce_mean(t)
        ag = _FakeAgent(rank)
        avg = GradAverager(ag)
        avg.average((1,))          # critic phase
        avg.average((0,))          # actor phase
        avg.average_metrics()
        # the normaliser's moments all-gather: rank order, every rank sees every rank's moments
        from gcrl_b200.normalizer import RunningNormalizer
        nz = RunningNormalizer.__new__(RunningNormalizer)      # host plumbing only: no CUDA object behind it
        nz.device_index, nz.size = 0, 3
        nz.enable_data_parallel()
        mine = np.full((3, 3), float(rank + 1)) * np.array([1.0, 10.0, 100.0])
        gathered = nz._gather(mine)
        out.put((rank, t.tolist(), ag.grads[0][:3].tolist(), ag.grads[1][:3].tolist(), ag.metrics[:2].tolist(),
                 avg.calls, gathered.tolist()))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_mean_allreduce_and_averager():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, t, g0, g1, m, calls, gathered in res:
        want = [(np.full((3, 3), float(r + 1)) * np.array([1.0, 10.0, 100.0])).tolist() for r in range(2)]
        assert gathered == want                      # [world, dim, 3], rank order, identical on every rank
        assert t == [0.5, 15.0]                      # mean over ranks
        assert g0 == [1.5, 1.5, 1.5]                 # (1 + 2) / 2
        assert g1 == [0.0, 1.5, 3.0]
        assert m == [1.5, 3.0]
        assert calls == 2
