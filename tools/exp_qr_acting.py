import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import LearnerGroup
B, G = 65536, 8
cfg = baseline_config(3).merged_with(batch_size=B)
g = LearnerGroup(cfg, [3, 7, 7], 4, G, B, obs_stride=152)
obs = torch.zeros((G, B, 152)); obs[..., :147] = (torch.rand((G, B, 147)) < 0.25).float(); obs = obs.cuda()
for _ in range(3): q = g.q_values(obs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): q = g.q_values(obs)
b.record(); torch.cuda.synchronize()
print("QR-DQN q_values: %.1f us per call" % (a.elapsed_time(b) / 10 * 1e3))
