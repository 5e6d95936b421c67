"""GPU experiment: the acting-batch Q forward of BASELINE configs[4] (C51) -- fused trunk + tc_heads_kernel.

    python tools/exp_tc_heads.py            # times q_values on 8 agents x 8192 observations
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import LearnerGroup

B, G = int(os.environ.get("EXP_B", "8192")), 8
cfg = baseline_config(5).merged_with(batch_size=B)
g = LearnerGroup(cfg, [3, 7, 7], 4, G, B, obs_stride=152)
gen = torch.Generator().manual_seed(1)
obs = torch.zeros((G, B, 152))
obs[..., :147] = (torch.rand((G, B, 147), generator=gen) < 0.25).float()
obs = obs.cuda()
for _ in range(3):
    q = g.q_values(obs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get("EXP_STEPS", "20"))
a.record()
for _ in range(n):
    q = g.q_values(obs)
b.record()
torch.cuda.synchronize()
print("q_values: %.1f us per call (%d x %d rows), checksum %.6f" % (a.elapsed_time(b) / n * 1e3, G, B, float(q.double().sum())))
