"""GPU experiment: env step time at the benchmark shape under the kernel's environment-variable knobs.

    python tools/exp_env_knobs.py MAEXP_STAGGER_NS=0,2000,4000 MAEXP_EPW=16,32

Every combination runs in a fresh subprocess (the knobs are read at maexp_env_create)."""
import itertools, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from maexp_b200.config import baseline_config
from tests.cuda_backend import make_env
cfg_id, E = int(os.environ.get("EXP_CFG", "5")), int(os.environ.get("EXP_ENVS", "65536"))
env = make_env(baseline_config(cfg_id), E, auto_reset=True)
ring = [env.new_slot() for _ in range(4)]
out = env.new_outputs()
env.reset(out, tick=0)
t = [1]
def step():
    o = dict(out); o["obs"] = ring[t[0] %% 4]
    env.step(o, tick=t[0]); t[0] += 1
for _ in range(10): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200): step()
b.record(); torch.cuda.synchronize()
print("%%.2f" %% (a.elapsed_time(b) / 200 * 1e3))
''' % ROOT
knobs = [a.split("=") for a in sys.argv[1:]]
names = [k for k, _ in knobs]
for combo in itertools.product(*[v.split(",") for _, v in knobs]):
    env = dict(os.environ, **dict(zip(names, combo)))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(dict(zip(names, combo)), "us/step:", r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
