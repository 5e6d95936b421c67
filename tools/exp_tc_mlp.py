"""GPU experiment: time of the C51 acting forward at B = 256 x 8 agents (fused trunk + tcgen05 MLP heads + expectation) under
the tc_mlp kernel's debug switches; the MLP kernel is the only part that changes.   python tools/exp_tc_mlp.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import LearnerGroup
cfg = baseline_config(5)
G, B = 8, 256
g = LearnerGroup(cfg, [3, 7, 7], 4, G, B)
obs = torch.rand(G, B, g.obs_stride, device="cuda").round()
for _ in range(5): g.q_values(obs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100): g.q_values(obs)
b.record(); torch.cuda.synchronize()
print("%%.2f" %% (a.elapsed_time(b) / 100 * 1e3))
''' % ROOT
for env in ({"MAEXP_TC_MLP": "0"}, {}, {"MAEXP_TC_MLP_DBG": "1"}, {"MAEXP_TC_MLP_DBG": "2"}, {"MAEXP_TC_MLP_DBG": "4"},
            {"MAEXP_TC_MLP_DBG": "3"}, {"MAEXP_TC_MLP_DBG": "7"}):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **env), capture_output=True, text=True)
    print(env, "us per q_values call:", r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
