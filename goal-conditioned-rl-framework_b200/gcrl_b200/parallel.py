"""Data-parallel plumbing: one process per GPU, episode-sharded buffer, averaged gradients.

The reference has no distributed code; this is the north-star's multi-GPU
layout: every rank owns the episodes ``episode_id % world == rank`` (all future-goal lookups stay
GPU-local), samples ``B`` positions of its own shard, and the flat gradient buffers of the C
library are averaged with NCCL between the backward and optimiser phases
(``gcrl_agent_update_phase``, include/gcrl_b200.h).
"""
from __future__ import annotations


class _CudaArray:
    """Minimal ``__cuda_array_interface__`` carrier so torch can alias library-owned memory."""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, n, device_index):
    import torch
    return torch.as_tensor(_CudaArray(ptr, n), device=torch.device("cuda", device_index))


def shard_of_episode(episode_id: int, world_size: int) -> int:
    """Owner rank of an episode: round-robin keeps the shards equal-sized (+-1 episode), so a
    uniform local draw on every rank is a uniform draw over the global buffer."""
    return int(episode_id) % int(world_size)


def local_batch(global_batch: int, world_size: int) -> int:
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    return global_batch // world_size


def allreduce_mean(tensor, group=None):
    """In-place mean over the ranks of ``group``.  NCCL averages natively; other backends
    (gloo in the CPU tests) sum, then divide."""
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(tensor, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
        tensor.div_(dist.get_world_size(group))
    return tensor


class GradAverager:
    def __init__(self, agent, group=None, fn=None):
        self.agent, self.group = agent, group
        self.fn = fn or (lambda t: allreduce_mean(t, group))
        self._grads = {}
        self._metrics = None
        self.calls = 0

    def average(self, nets):
        for net in nets:
            t = self._grads.get(net)
            if t is None:
                t = self._grads[net] = self.agent.grad_tensor(net)
            self.fn(t)
            self.calls += 1

    def average_metrics(self):
        if self._metrics is None:
            self._metrics = self.agent.metrics_tensor()
        self.fn(self._metrics)
