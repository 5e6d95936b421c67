"""Greedy evaluation loop (core/handlers/evaluators/__init__.py:23-52, abstract_evaluator.py:37-66,
default_evaluator.py:11-55) on the batched CUDA environment.

The reference evaluates one environment for `validate_epochs` episodes with epsilon = 0 and logs the
step / episode rewards and the environment counters to wandb.  Here every episode is run on all E
environments of a `BatchedEnvironment` at once; the scalars the reference logs per episode
(abstract_loop_handler.py:407-428) are returned as a list of dicts (per-environment means).
"""
import os

import torch

from ...agents import BatchedActor
from ...agents.brains import LearnerGroup
from ...environments.batched import BatchedEnvironment
from ...utils.buffer import generate_buffer
from ...utils.logging import initialize_logging

logger = initialize_logging(__name__)


class DefaultEvaluator:
    def __init__(self, config, environment, num_envs=None, first_global_env_id=0):
        self.config = config
        world = environment.world
        self.env = environment if isinstance(environment, BatchedEnvironment) else BatchedEnvironment(
            config, world, int(num_envs or config.get("num_envs") or 1024), auto_reset=False,
            first_global_env_id=first_global_env_id)
        env = self.env
        self.max_epochs = int(config.validate_epochs)
        self.ring = generate_buffer(config.merged_with(capacity=2 * env.E), env)  # two slots: current and next state
        # weights live in a LearnerGroup sized like the actor's chunks (no optimizer step is ever taken)
        self.learner = LearnerGroup(config, env.model_obs_shape, 4, env.N, 1, obs_stride=env.slot_floats, device=env.device)
        self.actor = BatchedActor(config, env, self.learner)
        if env.merged and config.observation_noise:  # see DefaultTrainer
            self.learner.set_los(env.los_table)
            self.actor.group.set_los(env.los_table)
        self.tick = 0
        self.episode_count = 0
        self.history = []

    def _tick(self):
        self.tick += 1
        return self.tick

    # ---- abstract_loop_handler.py:168-180
    def load_state_dict(self, directory=None):
        directory = directory or self.config.get("pretrained_weight_path")
        if not directory:
            logger.warning("pretrained_weight_path is not set: evaluating freshly initialised networks")
            return
        for agent_id in range(self.env.N):
            self.learner.load_state_dict(agent_id, torch.load(os.path.join(directory, f"agent{agent_id}.pth")))
        self.learner.synchronize()

    @torch.no_grad()
    def run(self, max_epochs=None):
        env, ring = self.env, self.ring
        self.actor.refresh()
        for _ in range(int(max_epochs if max_epochs is not None else self.max_epochs)):
            env.reset(ring.next_outputs(), tick=self._tick())
            ring.commit()
            reward_sum = torch.zeros((env.N,), device=env.device)
            for _ in range(int(self.config.max_episode_length)):
                head = ring.head
                positions = None if ring.positions is None else ring.positions[head]
                actions = self.actor.act(ring.obs[head], 0.0, self.tick, positions=positions, ring=ring)  # greedy
                out = ring.next_outputs()
                env.step(out, actions=actions, tick=self._tick())
                ring.commit()
                reward_sum += out["rewards"].sum(dim=0)
            cnt = env.state()["counters"].float().mean(dim=0).tolist()  # reset() zeroes them (default_environment.py:65-69)
            self.history.append({"episode": self.episode_count,
                                 "episode/total_reward": float(reward_sum.sum()) / env.E,
                                 "agent_reward": (reward_sum / env.E).tolist(),
                                 "episode/objects_generated": cnt[0], "episode/objects_completed": cnt[1],
                                 "episode/agents_collided": cnt[2], "episode/walls_collided": cnt[3]})
            self.episode_count += 1
        return self.history


def generate_evaluator(config, environment, **kw):
    if config.evaluator in ("default", "attention"):  # the attention evaluator adds attention-map images only
        return DefaultEvaluator(config=config, environment=environment, **kw)
    logger.warning(f"Unexpected evaluator is given. config.evaluator: {config.evaluator}")
    raise ValueError()
