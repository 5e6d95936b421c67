"""GPU experiment: cost of the optional per-step outputs and of the Python call path (not part of the product path)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maexp_b200 import _lib
from maexp_b200.config import baseline_config
from maexp_b200.core.utils.buffer import DeviceReplayRing
from tests.cuda_backend import make_env

E = 65536
env = make_env(baseline_config(5), E, auto_reset=True)
ring = DeviceReplayRing(env, 8, keep_positions=True)
env.reset(ring.next_outputs(), tick=0)
ring.commit()
tick = [0]


def timed(fn, n=300, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, (t1 - t0) / n * 1e6


def full():
    tick[0] += 1
    env.step(ring.next_outputs(), tick=tick[0])
    ring.commit()


def drop(*names):
    def fn():
        tick[0] += 1
        out = ring.next_outputs()
        for n in names:
            out.pop(n)
        env.step(out, tick=tick[0])
        ring.commit()
    return fn


structs = []
for s in range(ring.S):
    ring.head = s - 1
    structs.append(env._out_struct(ring.next_outputs()))
stream = env._stream()


def prepared():
    tick[0] += 1
    s = structs[tick[0] % ring.S]
    env.lib.maexp_env_step(env.h, tick[0], None, None, None, C.byref(s), stream)


for name, fn in (("ring outputs (bench loop)", full), ("without positions", drop("positions")),
                 ("without positions, actions_taken", drop("positions", "actions_taken")),
                 ("obs only", drop("positions", "actions_taken", "rewards", "dones", "was_reset")),
                 ("prepared structs, same outputs as bench", prepared)):
    gpu, host = timed(fn)
    print(f"{name:45s} gpu {gpu:7.2f} us/step   host issue {host:7.2f} us/step", flush=True)
