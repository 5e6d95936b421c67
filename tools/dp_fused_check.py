"""Multi-GPU check + timing of the fused data-parallel apply (maexp_learner_apply_data_parallel) against the NCCL path:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_fused_check.py
(1) K updates of twin learners (same weights, same per-rank batches) through both paths: parameters agree to fp32 rounding
    and every rank holds identical parameters; (2) microseconds per update of both paths at B = 256 x 8 agents."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from maexp_b200 import _lib
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import GraphedUpdate, LearnerGroup

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev)
dist.barrier(device_ids=[local]); torch.cuda.synchronize()
sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
cfg = baseline_config(5)
G, B = 8, int(os.environ.get("B", "256"))
cfg = cfg.merged_with(batch_size=B)


def make(fused):
    torch.manual_seed(7)
    g = LearnerGroup(cfg, [3, 7, 7], 4, G, B, device=dev)
    g.broadcast(dist)
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    batch = {"obs": (torch.rand(G, B, g.obs_stride, generator=gen) < 0.2).float().to(dev),
             "next_obs": (torch.rand(G, B, g.obs_stride, generator=gen) < 0.2).float().to(dev),
             "actions": torch.randint(0, 4, (G, B), generator=gen).to(torch.int32).to(dev),
             "rewards": torch.randint(-1, 2, (G, B), generator=gen).float().to(dev),
             "dones": torch.zeros(G, B, dtype=torch.uint8, device=dev)}
    return g, GraphedUpdate(g, batch, dist=dist, fused_collective=fused)


ga, ua = make(True)
gb, ub = make(False)
assert ua.peers is not None, "fused path unavailable"
for k in range(6):
    ua.step()
    if os.environ.get("DEBUG"): torch.cuda.synchronize(); print(rank, "fused step", k, flush=True)
    ub.step()
    if os.environ.get("DEBUG"): torch.cuda.synchronize(); print(rank, "nccl step", k, flush=True)
torch.cuda.synchronize()
pa, pb = ga.params(), gb.params()
diff = float((pa - pb).abs().max())
fresh = make(False)[0]  # keep the handle alive: params() is a zero-copy view of library memory
moved = float((pa - fresh.params()).abs().max())
spread_a, spread_b = ga.replica_spread(dist), gb.replica_spread(dist)
norm_diff = float((ga.grad_norm - gb.grad_norm).abs().max() / gb.grad_norm.abs().max())


def timed(u, n=200):
    for _ in range(10): u.step()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): u.step()
    b.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


t_fused, t_nccl = timed(ua), timed(ub)


class Part:
    def __init__(self, fn): self.step = fn


t_grads_symm, t_grads_own = timed(Part(ua.g_grads.replay)), timed(Part(ub.g_grads.replay))
t_apply_fused = timed(Part(ua._apply))
t_allreduce = timed(Part(ub._reduce))
t_apply_own = timed(Part(ub.g_apply.replay))
if rank == 0:
    print(json.dumps({"world": world, "batch_per_gpu": B, "fused_vs_nccl_max_abs_param_diff": diff, "params_moved_by": moved,
                      "grad_norm_rel_diff": norm_diff, "replica_spread_fused": spread_a, "replica_spread_nccl": spread_b,
                      "multimem": bool(ua.peers.multicast_grads), "us_per_update_fused": t_fused, "us_per_update_nccl": t_nccl,
                      "us_grads_graph_symmetric_buffer": t_grads_symm, "us_grads_graph_own_buffer": t_grads_own,
                      "us_fused_exchange_apply_kernel": t_apply_fused, "us_nccl_all_reduce": t_allreduce,
                      "us_clip_rmsprop_graph": t_apply_own}))
    # the two paths add the ranks' gradients in different orders; RMSprop's first steps amplify last-bit differences of
    # near-zero gradients (see tests/test_learner_gpu.py), so the parameters are compared loosely and the norms tightly
    assert spread_a == 0.0 and norm_diff <= 1e-4 and diff <= 5e-3 and moved > 1e-3, "fused data-parallel apply disagrees with the NCCL path"
dist.barrier()
dist.destroy_process_group()
