"""GPU measurement of the fused replay sampler (maexp_replay_sample): time and achieved HBM GB/s at the learner's batch
(B = 256, latency-bound) and at large batches where the gather is HBM-bound.   python tools/exp_replay.py
bytes per sampled transition and agent = 2 * slot_floats * 4 read + the same written (+ ~6 B meta)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maexp_b200.config import baseline_config
from maexp_b200.core.utils.buffer import DeviceReplayRing
from tests.cuda_backend import make_env

E = int(os.environ.get("EXP_ENVS", "65536"))
env = make_env(baseline_config(5), E, auto_reset=True)
ring = DeviceReplayRing(env, 8)
env.reset(ring.next_outputs(), tick=0); ring.commit()
for t in range(1, 8):
    env.step(ring.next_outputs(), tick=t); ring.commit()
peak = 6551.7
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for B in (256, 4096, 65536, 262144):
    batch = ring.new_batch(B)
    for i in range(5): ring.sample(batch, seed=1, tick=i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    a.record()
    for i in range(n): ring.sample(batch, seed=1, tick=10 + i)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    moved = B * env.N * (4 * env.slot_floats * 4 + 6)
    print(json.dumps({"kernel": "replay_sample_kernel", "batch": B, "agents": env.N, "us": us, "bytes": moved,
                      "GB/s": moved / us / 1e3, "frac_of_measured_hbm_peak": moved / us / 1e3 / peak}), flush=True)
    ring.check(batch)
