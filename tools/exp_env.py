"""GPU experiment: where does the env step time go? (not part of the product path)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maexp_b200.config import baseline_config
from tests.cuda_backend import make_env

def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3  # us

E = 65536
env = make_env(baseline_config(5), E, auto_reset=True)
ring = torch.empty((8, E, 8, env.slot_floats), dtype=torch.float32, device="cuda")
out = env.new_outputs()
env.reset(out, tick=0)
state = {"t": 1}
def step(obs):
    o = dict(out); o["obs"] = obs
    env.step(o, tick=state["t"]); state["t"] += 1
print("slot MB", ring[0].numel() * 4 / 1e6)
i = [0]
def full():
    i[0] += 1; step(ring[i[0] % 8])
print("step full           us", timeit(full))
print("step no-obs         us", timeit(lambda: step(None)))
def fill():
    i[0] += 1; ring[i[0] % 8].fill_(1.0)
t = timeit(fill); print("torch fill_ slot    us", t, "GB/s", ring[0].numel() * 4 / t / 1e3)
def cpy():
    i[0] += 1; ring[i[0] % 8].copy_(ring[(i[0] + 4) % 8])
t = timeit(cpy); print("torch copy slot     us", t, "GB/s (r+w)", 2 * ring[0].numel() * 4 / t / 1e3)
