"""`generate_trainer` (core/handlers/trainers/__init__.py:20-40): the reference's hot loop
(abstract_loop_handler.py:79-102, default_trainer.py:18-99, abstract_trainer.py:45-104) for E batched
environments per GPU.

Per step, as in the reference: one learner update of every agent's network on a sampled batch, then one
epsilon-greedy environment step appended to the replay ring.  Per episode (max_episode_length steps):
epsilon decay and, every `synchronize_frequency` episodes, a hard target sync.
"""
import os

import torch

from ...agents import BatchedActor
from ...agents.brains import GraphedUpdate, LearnerGroup
from ...environments.batched import BatchedEnvironment
from ...utils.buffer import generate_buffer
from ...utils.logging import initialize_logging

logger = initialize_logging(__name__)


class DefaultTrainer:
    def __init__(self, config, environment, num_envs=None, all_reduce=None, first_global_env_id=0):
        self.config = config
        world = environment.world
        self.env = environment if isinstance(environment, BatchedEnvironment) else BatchedEnvironment(
            config, world, int(num_envs or config.get("num_envs") or 1024), auto_reset=True,
            first_global_env_id=first_global_env_id)
        env = self.env
        self.buffer = generate_buffer(config, env)
        self.learner = LearnerGroup(config, env.model_obs_shape, 4, env.N, int(config.batch_size),
                                    obs_stride=env.slot_floats, device=env.device)
        self.actor = BatchedActor(config, env, self.learner)
        self.batch = self.buffer.new_batch(int(config.batch_size))
        self.active = 1 - self.actor.no_task  # agent_tasks "-1" agents are never trained (abstract_trainer.py:49-51)
        self.epsilon = float(config.epsilon_initial)
        self.all_reduce = all_reduce
        # single-GPU updates replay from CUDA graphs (about a hundred launches per update otherwise); the IQN learners draw
        # their quantile fractions with torch.rand per update and the data-parallel path takes the caller's all-reduce: eager
        self.graphed = None
        if all_reduce is None and self.learner.kind != "iqn" and config.get("cuda_graph_update", True):
            self.graphed = GraphedUpdate(self.learner, self.batch, active=self.active)
        self.tick = 0
        self.episode_count = 0
        self.global_step = 0
        self.history = []

    # ---- abstract_loop_handler.py:51-57
    def populate(self, steps):
        self.env.reset(self.buffer.next_outputs(), tick=self._tick())
        self.buffer.commit()
        for _ in range(steps):
            self.env.step(self.buffer.next_outputs(), tick=self._tick())  # random policy (epsilon = 1)
            self.buffer.commit()

    def _tick(self):
        self.tick += 1
        return self.tick

    # ---- default_trainer.py:18-85
    def loop_step(self):
        self.buffer.sample(self.batch, seed=int(self.config.seed), tick=self.tick)
        if self.graphed is not None:
            loss = self.graphed.step()
        else:
            loss = self.learner.learn(self.batch, active=self.active, all_reduce=self.all_reduce)
        self.actor.refresh()
        head = self.buffer.head
        positions = None if self.buffer.positions is None else self.buffer.positions[head]
        actions = self.actor.act(self.buffer.obs[head], self.epsilon, self.tick, positions=positions)
        out = self.buffer.next_outputs()
        self.env.step(out, actions=actions, tick=self._tick())
        self.buffer.commit()
        self.global_step += 1
        return loss, out["rewards"]

    # ---- default_trainer.py:87-99
    def loop_epoch_end(self):
        self.epsilon = max(float(self.config.epsilon_end), self.epsilon * float(self.config.epsilon_decay))
        if self.episode_count % int(self.config.synchronize_frequency) == 0:
            self.learner.synchronize()
        self.episode_count += 1

    def run(self, max_epochs=None):
        cfg = self.config
        self.populate(max(2, int(cfg.populate_steps) // max(1, self.env.E)))
        for _ in range(int(max_epochs if max_epochs is not None else cfg.max_epochs)):
            reward_sum = torch.zeros((), device=self.env.device)
            loss_sum = torch.zeros((), device=self.env.device)
            for _ in range(int(cfg.max_episode_length)):
                loss, rewards = self.loop_step()
                reward_sum += rewards.sum()
                loss_sum += loss.sum()
            self.loop_epoch_end()
            self.history.append({"episode": self.episode_count, "epsilon": self.epsilon,
                                 "reward_per_env": float(reward_sum) / self.env.E, "loss": float(loss_sum)})
        return self.history

    # ---- abstract_loop_handler.py:149-166 (local files instead of wandb artifacts)
    def save_state_dict(self, directory="."):
        os.makedirs(directory, exist_ok=True)
        for agent_id in range(self.env.N):
            torch.save(self.learner.state_dict(agent_id), os.path.join(directory, f"agent{agent_id}.pth"))

    def load_state_dict(self, directory="."):
        for agent_id in range(self.env.N):
            self.learner.load_state_dict(agent_id, torch.load(os.path.join(directory, f"agent{agent_id}.pth")))
        self.learner.synchronize()


def generate_trainer(config, environment, **kw):
    if config.trainer in ("default", "attention"):  # attention_trainer.py only adds attention-map logging
        return DefaultTrainer(config=config, environment=environment, **kw)
    logger.warning(f"Unexpected trainer is given. config.trainer: {config.trainer}")
    raise ValueError()
