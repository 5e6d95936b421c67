"""GPU experiment: time one C51 team update (not part of the product path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import LearnerGroup

cfg = baseline_config(int(os.environ.get("CFG", "5")))
G, B = 8, int(cfg.batch_size)
g = LearnerGroup(cfg, [3, 7, 7], 4, G, B)
batch = {"obs": torch.rand(G, B, g.obs_stride, device="cuda").round(), "next_obs": torch.rand(G, B, g.obs_stride, device="cuda").round(),
         "actions": torch.randint(0, 4, (G, B), dtype=torch.int32, device="cuda"), "rewards": torch.zeros(G, B, device="cuda"),
         "dones": torch.zeros(G, B, dtype=torch.uint8, device="cuda")}
for _ in range(3):
    g.learn(batch)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    g.learn(batch)
b.record(); torch.cuda.synchronize()
print("eager us/update", a.elapsed_time(b) / 20 * 1e3)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    g.learn(batch)
a.record()
for _ in range(50):
    gr.replay()
b.record(); torch.cuda.synchronize()
print("graph us/update", a.elapsed_time(b) / 50 * 1e3)
