"""GPU demo: does the batched trainer learn?  Greedy evaluation reward before and after a short training run.

    python tools/train_demo.py [--config 5] [--envs 1024] [--epochs 30]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maexp_b200.config import baseline_config
from maexp_b200.core.handlers.evaluators import generate_evaluator
from maexp_b200.core.handlers.trainers import generate_trainer
from tests.cuda_backend import make_env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--episode", type=int, default=100)
    ap.add_argument("--map", default=None, help="e.g. dynamic_simple (20 x 20, moving obstacles) instead of the config's map")
    ap.add_argument("--attention", default=None, help="hard: HardShrinkBlock with --af-lambd (attention models)")
    ap.add_argument("--af-lambd", type=float, default=0.02)
    args = ap.parse_args()
    base = baseline_config(args.config)
    if args.map:
        base = base.merged_with(map=dict(name=args.map, SIZE_X=20, SIZE_Y=20))
    if args.attention:
        base["model"] = dict(base.model, attention=args.attention, af_lambd=args.af_lambd)
    cfg = base.merged_with(max_episode_length=args.episode, populate_steps=4 * args.envs,
                                                    capacity=64 * args.envs, synchronize_frequency=1, epsilon_decay=0.9,
                                                    validate_epochs=2, observation_noise=False)
    env = make_env(cfg, args.envs, auto_reset=True)
    trainer = generate_trainer(cfg, env)

    def evaluate():
        ev = generate_evaluator(cfg, make_env(cfg, args.envs, auto_reset=False))
        ev.learner.params().copy_(trainer.learner.params())
        h = ev.run()
        return sum(r["episode/total_reward"] for r in h) / len(h), sum(r["episode/objects_completed"] for r in h) / len(h)

    before = evaluate()
    t0 = time.perf_counter()
    hist = trainer.run(max_epochs=args.epochs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    after = evaluate()
    steps = args.epochs * args.episode
    print(json.dumps({"config": args.config, "map": cfg.map["name"], "attention": cfg.model.get("attention"), "envs": args.envs, "epochs": args.epochs, "updates": steps,
                      "train_seconds": dt, "agent_steps_per_sec_in_training": steps * args.envs * env.N / dt,
                      "greedy_reward_before": before[0], "greedy_reward_after": after[0],
                      "objects_completed_before": before[1], "objects_completed_after": after[1],
                      "train_reward_first_epoch": hist[0]["reward_per_env"], "train_reward_last_epoch": hist[-1]["reward_per_env"],
                      "loss_first": hist[0]["loss"], "loss_last": hist[-1]["loss"]}))


if __name__ == "__main__":
    main()
