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
