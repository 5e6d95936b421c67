"""GPU helper: N timed env steps of one BASELINE config (for ncu captures and knob experiments).

    EXP_CFG=5 EXP_ENVS=65536 EXP_STEPS=200 python tools/exp_env_one.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maexp_b200.config import baseline_config
from tests.cuda_backend import make_env
cfg_id, E, steps = int(os.environ.get("EXP_CFG", "5")), int(os.environ.get("EXP_ENVS", "65536")), int(os.environ.get("EXP_STEPS", "200"))
env = make_env(baseline_config(cfg_id), E, auto_reset=True)
ring = [env.new_slot() for _ in range(4)]
out = env.new_outputs()
env.reset(out, tick=0)
t = [1]
def step():
    o = dict(out); o["obs"] = ring[t[0] % 4]
    env.step(o, tick=t[0]); t[0] += 1
for _ in range(10): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps): step()
b.record(); torch.cuda.synchronize()
print("lanes" if env.uses_lanes_kernel() else "lane-per-env", "%.2f us/step" % (a.elapsed_time(b) / steps * 1e3))
