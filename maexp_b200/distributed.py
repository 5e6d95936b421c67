"""Data-parallel plumbing (host side): env sharding and the learner's gradient all-reduce.

Environments are independent units (no cross-env state in default_environment.py), so rank r
owns the contiguous global env ids [r*E, (r+1)*E) and step / observation / replay need no
collective.  The only exchange is one all-reduce over the learner's flat gradient buffer.
"""
import os


def rank_info():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def first_global_env_id(rank, envs_per_rank):
    return rank * envs_per_rank


def make_grad_all_reduce(dist, group=None):
    """Returns callable(grads) -> world_size that sums `grads` in place over the process group.
    LearnerGroup.learn divides by the returned world size inside the clip+RMSprop kernel."""
    world = dist.get_world_size(group)

    def all_reduce(grads):
        dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=group)
        return world
    return all_reduce


def broadcast_tensors(dist, tensors, group=None, src=0):
    """Rank `src` sends every tensor of the list (parameters, target parameters, optimizer state) to the other ranks, in
    place: data-parallel replicas must start from the same state."""
    for t in tensors:
        dist.broadcast(t, src=src, group=group)


def replica_spread(dist, params, group=None):
    """max |params_rank - params_rank0| over all elements and ranks: 0.0 when the replicas are in sync."""
    ref = params.clone()
    dist.broadcast(ref, src=0, group=group)
    spread = (params - ref).abs().max().reshape(1)
    dist.all_reduce(spread, op=dist.ReduceOp.MAX, group=group)
    return float(spread)


def max_over_ranks(dist, values, device):
    """Element-wise maximum of a list of floats over all ranks (device-side timings)."""
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(dist, values, device):
    import torch
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]
