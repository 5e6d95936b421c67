import os, sys, torch
sys.path.insert(0, "/root/repo")
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import LearnerGroup
cfg = baseline_config(5)
G, B = 8, 256
g = LearnerGroup(cfg, [3, 7, 7], 4, G, B)
obs = torch.rand(G, B, g.obs_stride, device="cuda").round()
for _ in range(3): g.q_values(obs)
torch.cuda.synchronize()
