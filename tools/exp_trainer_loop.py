"""GPU experiment: where does a trainer loop step go at full batch size? (not part of the product path)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maexp_b200.config import baseline_config
from maexp_b200.core.handlers.trainers import generate_trainer
from tests.cuda_backend import make_env

E = int(os.environ.get("E", "65536"))
cfg = baseline_config(int(os.environ.get("CFG", "5"))).merged_with(populate_steps=4 * E, capacity=8 * E, max_episode_length=200)
t = generate_trainer(cfg, make_env(cfg, E, auto_reset=True))
if os.environ.get("CHUNK"):
    from maexp_b200.core.agents import BatchedActor
    t.actor = BatchedActor(cfg, t.env, t.learner, chunk=int(os.environ["CHUNK"]))
t.populate(4)
torch.cuda.synchronize()


def timed(fn, n=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


print("loop_step ms      ", timed(t.loop_step))
print("  sample+learn ms ", timed(lambda: (t.buffer.sample(t.batch, seed=1, tick=t.tick), t.learner.learn(t.batch, active=t.active))))
print("  refresh ms      ", timed(t.actor.refresh))
h = t.buffer.head
print("  act ms          ", timed(lambda: t.actor.act(t.buffer.obs[h], 0.1, 5, positions=None if t.buffer.positions is None else t.buffer.positions[h])))
print("  env.step ms     ", timed(lambda: t.env.step(t.buffer.next_outputs(), tick=t._tick())))
print("agent-steps/s in the loop:", E * t.env.N / (timed(t.loop_step) * 1e-3))

import collections
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t.actor.act(t.buffer.obs[h], 0.1, 5, positions=None if t.buffer.positions is None else t.buffer.positions[h])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:70]][0] += 1
        agg[e.name[:70]][1] += e.time_range.end - e.time_range.start
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{k:72s} n {n:4d}  total {us / 1e3:8.3f} ms  avg {us / n:8.1f} us")

print("---- kernels of one loop_step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    t.loop_step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:70]][0] += 1
        agg[e.name[:70]][1] += e.time_range.end - e.time_range.start
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:72s} n {n:4d}  total {us / 1e3:8.3f} ms  avg {us / n:8.1f} us")

print("---- host profile of 20 loop steps")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    t.loop_step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
