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
    """dist: the torch.distributed module of an initialised process group (NCCL on the GPUs, one process per GPU) for the
    data-parallel learner — every rank steps its own environments (pass first_global_env_id = rank * num_envs) and the
    per-agent networks are replicas: rank 0's initial weights are broadcast, each update all-reduces the flat gradient
    buffer between the two CUDA graphs of GraphedUpdate.  all_reduce: legacy callable(grads) -> world size (eager path)."""

    def __init__(self, config, environment, num_envs=None, all_reduce=None, first_global_env_id=0, dist=None,
                 process_group=None, log_path=None):
        self.config = config
        world = environment.world
        self.env = environment if isinstance(environment, BatchedEnvironment) else BatchedEnvironment(
            config, world, int(num_envs or config.get("num_envs") or 1024), auto_reset=True,
            first_global_env_id=first_global_env_id)
        env = self.env
        self.buffer = generate_buffer(config, env)
        self.learner = LearnerGroup(config, env.model_obs_shape, 4, env.N, int(config.batch_size),
                                    obs_stride=env.slot_floats, device=env.device)
        self.dist, self.process_group = dist, process_group
        if dist is not None:
            self.learner.broadcast(dist, process_group)  # replicas start from rank 0's weights
        self.actor = BatchedActor(config, env, self.learner)
        if env.merged and config.observation_noise:
            # the merged models read the area channel of the noise-free "relative" key (merged_observtion_handler.py:27,41-47):
            # rebuilt from the environment's line-of-sight table, the ring only keeps the noisy crop
            self.learner.set_los(env.los_table)
            self.actor.group.set_los(env.los_table)
        self.batch = self.buffer.new_batch(int(config.batch_size))
        self.active = 1 - self.actor.no_task  # agent_tasks "-1" agents are never trained (abstract_trainer.py:49-51)
        self.epsilon = float(config.epsilon_initial)
        self.all_reduce = all_reduce
        # updates replay from CUDA graphs (about a hundred launches per update otherwise) — single GPU, data parallel
        # (the NCCL all-reduce runs between the gradient graph and the optimizer graph) and IQN (fractions in fixed
        # buffers refilled by one Philox launch); only a caller-supplied all_reduce callable takes the eager path
        self.graphed = None
        if all_reduce is None and config.get("cuda_graph_update", True):
            self.graphed = GraphedUpdate(self.learner, self.batch, active=self.active, dist=dist, process_group=process_group)
        elif all_reduce is None and dist is not None:
            from .... import distributed as mdist
            self.all_reduce = mdist.make_grad_all_reduce(dist, process_group)
        self.tick = 0
        self.episode_count = 0
        self.global_step = 0
        self.history = []
        self.log_path = log_path
        self.episode_return = torch.zeros((env.E,), dtype=torch.float32, device=env.device)

    # ---- abstract_loop_handler.py:45-57
    def reset(self):
        """AbstractLoopHandler.reset: a fresh episode in every environment (one ring slot, marked was_reset)."""
        self.env.reset(self.buffer.next_outputs(), tick=self._tick())
        self.buffer.commit()
        self.episode_return.zero_()

    def populate(self, steps):
        self.reset()
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
        actions = self.actor.act(self.buffer.obs[head], self.epsilon, self.tick, positions=positions, ring=self.buffer)
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
        self.buffer.check(self.batch)  # the replay sampler's error flag (one host sync per episode)
        self.reset()  # default_trainer.py:99: every epoch is one episode
        self.episode_count += 1

    # ---- abstract_loop_handler.py:407-428 (wandb.log of the episode scalars) as one JSON line per episode
    def log_scalar(self):
        if not self.history:
            return
        row = self.history[-1]
        if self.log_path and (self.dist is None or self.dist.get_rank(self.process_group) == 0):
            import json
            with open(self.log_path, "a") as f:
                f.write(json.dumps({"step": self.global_step, **{f"episode/{k}": v for k, v in row.items()}}) + "\n")

    def run(self, max_epochs=None):
        """AbstractLoopHandler.run (abstract_loop_handler.py:79-102).  An epoch is one episode of max_episode_length steps in
        every environment, closed by reset() as in the reference; an environment whose objects run out before that
        (keep_objects_num off) restarts on its own inside the epoch, where the reference would end the epoch early."""
        cfg = self.config
        self.populate(max(2, int(cfg.populate_steps) // max(1, self.env.E)))
        self.reset()
        for _ in range(int(max_epochs if max_epochs is not None else cfg.max_epochs)):
            loss_sum = torch.zeros((), device=self.env.device)
            for _ in range(int(cfg.max_episode_length)):
                loss, rewards = self.loop_step()
                self.episode_return += rewards.sum(dim=1)
                loss_sum += loss.sum()
            stats = self.env.stats()
            if stats["error_flag"]:
                raise RuntimeError("the respawn sampler hit its attempt cap (objects_area too small, SURVEY.md §E-8)")
            reward = float(self.episode_return.mean())  # the episode's return, summed over the team, mean over environments
            self.loop_epoch_end()
            self.history.append({"episode": self.episode_count, "epsilon": self.epsilon, "reward_per_env": reward,
                                 "loss": float(loss_sum)})
            self.log_scalar()
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
        if self.dist is not None:
            self.learner.broadcast(self.dist, self.process_group)


def generate_trainer(config, environment, **kw):
    if config.trainer in ("default", "attention"):  # attention_trainer.py only adds attention-map logging
        return DefaultTrainer(config=config, environment=environment, **kw)
    logger.warning(f"Unexpected trainer is given. config.trainer: {config.trainer}")
    raise ValueError()
