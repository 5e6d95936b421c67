"""Times maexp_heatmap_observation_add on the BASELINE configs[4] slot (65536 envs x 8 agents, 310 MB):
   python tools/exp_heatmap_obs.py [envs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maexp_b200.config import baseline_config  # noqa: E402
from tests.cuda_backend import make_env  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = make_env(baseline_config(5), E, auto_reset=True)
outs = [env.new_outputs() for _ in range(2)]  # two slots: 620 MB > L2
for o in outs:
    env.reset(o)
    env.step(o)
heat = env.new_heatmap_observation()
for i in range(5):
    env.accumulate_heatmap_observation(outs[i & 1], heat)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
K = 50
for i in range(K):
    env.accumulate_heatmap_observation(outs[i & 1], heat)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) * 1e3 / K
read = E * env.N * env.K * 4  # algorithmic bytes: every observation element once
print({"envs": E, "us_per_call": round(us, 2), "algorithmic_GBps": round(read / us / 1e3, 1),
       "slot_GBps": round(E * env.N * env.slot_floats * 4 / us / 1e3, 1)})
