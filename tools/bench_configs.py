"""GPU measurement of every BASELINE config: env step + observation throughput against
the HBM roofline, and the learner update of the config's brain.  Prints one JSON line per config.

    python tools/bench_configs.py [--envs 65536] [--steps 200]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import LearnerGroup
from maexp_b200.core.utils.buffer import DeviceReplayRing
from tests.cuda_backend import make_env


def peak_gbs():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"])
    except Exception:
        return 6551.7


def b_alg(cfg, env):
    """obs write (x2 for merged) + objects grid r/w per env + 12 B scalars + past positions."""
    C, V, N = env.C, env.V, env.N
    obs = 4 * C * V * V * (2 if cfg.observation_area_mask == "merged" else 1)
    past = 4 * int(cfg.past_step) if "transition" in cfg.agent_view_method else 0
    return obs + 2 * env.T * env.X * env.Y / N + 12 + past


def timed(fn, steps, warm=10):
    for i in range(warm):
        fn(i)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(steps):
        fn(warm + i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--configs", default="1,2,3,4,5")
    args = ap.parse_args()
    peak = peak_gbs()
    for idx in [int(v) for v in args.configs.split(",")]:
        cfg = baseline_config(idx)
        E = args.envs
        env = make_env(cfg, E, auto_reset=True)
        slots = 4
        ring = DeviceReplayRing(env, slots)
        rel = [env.new_slot() for _ in range(2)] if env.merged else None
        env.reset(ring.next_outputs(), tick=0)
        ring.commit()
        tick = [0]

        def step(i):
            tick[0] += 1
            out = ring.next_outputs()
            if rel is not None:
                out["obs_rel"] = rel[i % 2]
            env.step(out, tick=tick[0])
            ring.commit()
        us = timed(step, args.steps)
        N = env.N
        rate = E * N / (us * 1e-6)
        balg = b_alg(cfg, env)
        line = {"config": idx, "map": cfg.map.name, "envs": E, "agents": N, "channels": env.C, "visible_range": env.V,
                "noise": cfg.observation_noise, "area_mask": cfg.observation_area_mask, "us_per_step": us,
                "agent_steps_per_sec": rate, "b_alg": balg, "achieved_gbs": balg * rate / 1e9, "peak_gbs": peak,
                "frac": balg * rate / 1e9 / peak, "slot_mb": env.slot_floats * 4 * N * E / 1e6}
        # learner update of the config's own brain (cfg3: QR-DQN; IQN is measured by the caller via --configs)
        if not (cfg.observation_area_mask == "merged" and cfg.observation_noise):
            B = int(cfg.batch_size)
            group = LearnerGroup(cfg, env.model_obs_shape, 4, N, B, obs_stride=env.slot_floats, device=env.device)
            batch = ring.new_batch(B)
            ring.sample(batch, seed=921, tick=0)
            group.learn(batch)
            torch.cuda.synchronize()
            g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                group.grads(batch)
            with torch.cuda.graph(g2):
                group.apply(1.0)

            def update(i):
                ring.sample(batch, seed=921, tick=i)
                g1.replay()
                g2.replay()
            uus = timed(update, min(args.steps, 100), warm=5)
            line["learner"] = {"brain": cfg.brain, "model": cfg.model.name, "batch": B, "us_per_team_update": uus,
                               "agent_network_updates_per_sec": N / (uus * 1e-6), "params_per_agent": group.P,
                               "loss": float(group.loss.mean())}
            del group, g1, g2
        print(json.dumps(line), flush=True)
        del env, ring, rel
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
