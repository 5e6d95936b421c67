"""`generate_environment` (core/environments/__init__.py:37-71) on top of the CUDA step kernel.

`DefaultEnvironment` keeps the reference's single-environment interface — `reset() -> obs_n`,
`step(action_n, order) -> (reward_n, done_n, obs_n)` with per-agent observation dicts
(default_environment.py:64-175, observation_handler.py:95-105) — by driving a one-env
`BatchedEnvironment`.  Training at scale uses `BatchedEnvironment` directly (see trainers).
"""
import numpy as np
import torch

from ..handlers.observations import generate_observation_handler
from ..handlers.observations.masks import COORDINATE_KEYS
from ..utils.logging import initialize_logging
from .batched import BatchedEnvironment

logger = initialize_logging(__name__)

__all__ = ["BatchedEnvironment", "DefaultEnvironment", "SharedMeanRewardEnvironment", "SharedMaxRewardEnvironment",
           "SequentialEnvironment", "JunctionalEnvironment", "ObservationStatsEnvironment", "generate_environment"]


class DefaultEnvironment:
    reward_environment = "default"

    def __init__(self, config, world, reference_compat=False):
        self.config, self.world = config, world
        self.num_agents = int(config.num_agents)
        self.agents = list(range(self.num_agents))
        self.observation_handler = generate_observation_handler(config=config, world=world)
        self.action_space = [4 for _ in self.agents]  # default_environment.py:27
        self.observation_space = [self.observation_handler.observation_space for _ in self.agents]
        self.view_method = config.observation_area_mask
        cfg = config if config.environment == self.reward_environment else config.merged_with(
            environment=self.reward_environment)
        self.batched = BatchedEnvironment(cfg, world, 1, auto_reset=False, shared_max_is_mean=reference_compat)
        self.out = self.batched.new_outputs(extras=True)
        self.current_step = 0
        self.objects_generated = self.objects_completed = self.agents_collided = self.walls_collided = 0
        self._heat = self.batched.new_heatmap_observation()  # default_environment.py:93-95
        # the six integer visit statistics of the running episode (default_environment.py:70-92), written by the step
        # kernel through maexp_obs_out.heatmaps, and their sums over episodes (:36-61, accumulate_heatmap :249-257)
        self._hm_buf, self._hm_split = self.batched.new_heatmaps()
        self.out["heatmaps"] = self._hm_buf
        self._hm_acc = torch.zeros_like(self._hm_buf)
        self.heatmap_accumulated_observation = np.zeros(tuple(self._heat.shape), np.float32)

    @property
    def heatmap_observation(self):
        """[N, *observation_space] sums of this episode's step observations (default_environment.py:93,238-247)."""
        return self._heat.cpu().numpy()

    def __getattr__(self, name):
        # heatmap_agents, heatmap_complete, heatmap_objects, heatmap_objects_left, heatmap_wall_collision,
        # heatmap_agents_collision and their heatmap_accumulated_* sums, as numpy int32 arrays
        if name.startswith("heatmap_") and "_hm_split" in self.__dict__:
            acc = name.startswith("heatmap_accumulated_")
            key = "heatmap_" + name[len("heatmap_accumulated_"):] if acc else name
            views = self._hm_split(self._hm_acc) if acc else self._hm_split()
            if key in views:
                return views[key].cpu().numpy()
        raise AttributeError(name)

    def accumulate_heatmap(self):
        """default_environment.py:249-257."""
        self._hm_acc += self._hm_buf
        self.heatmap_accumulated_observation += self.heatmap_observation

    # ---- reference structures
    def _obs_n(self):
        b, o = self.batched, self.out
        torch.cuda.synchronize(b.device)
        obs = b.view(o["obs"])[0].cpu()
        rel = b.view(o["obs_rel"])[0].cpu() if b.merged else None
        pos = o["positions"][0].cpu()
        coords = o["coordinates"][0].cpu().numpy()
        objects = o["objects"][0].cpu()
        dest = torch.from_numpy(b.destination_masks[o["destination"][0].cpu().numpy().astype(np.int64)])
        view = "local" if self.view_method in ("local", "merged") else "relative"
        obs_n = []
        for i in range(self.num_agents):
            d = {view: obs[i], "coordinates": {k: int(v) for k, v in zip(COORDINATE_KEYS, coords[i])},
                 "agents": pos.clone(), "objects": objects[i], "destination": dest[i]}
            if rel is not None:
                d["relative"] = rel[i]
            obs_n.append(d)
        st = b.state()
        cnt = st["counters"][0].cpu().tolist()
        self.objects_generated, self.objects_completed, self.agents_collided, self.walls_collided = cnt
        self.current_step = int(st["steps"][0])
        return obs_n

    def reset(self):
        self._hm_buf.zero_()  # default_environment.py:70-96 re-creates the maps; reset() itself counts the new objects
        self.batched.reset(self.out)
        self._heat.zero_()
        return self._obs_n()

    def step(self, action_n, order):
        dev = self.batched.device
        actions = torch.as_tensor(np.asarray(action_n, np.int8).reshape(1, -1)).to(dev)
        order = torch.as_tensor(np.asarray(order, np.uint8).reshape(1, -1)).to(dev)
        self.batched.step(self.out, actions=actions, order=order)
        self.batched.accumulate_heatmap_observation(self.out, self._heat)
        obs_n = self._obs_n()
        rewards = self.out["rewards"][0].cpu().tolist()
        dones = [bool(d) for d in self.out["dones"][0].cpu().tolist()]
        return rewards, dones, obs_n


class SharedMeanRewardEnvironment(DefaultEnvironment):
    """shared_reward_environment.py:15-20 (with the `order` argument the shipped class lacks)."""
    reward_environment = "shared_mean_reward"


class SharedMaxRewardEnvironment(DefaultEnvironment):
    """shared_reward_environment.py:23-28.  The reference computes the MEAN here; pass reference_compat=True to
    reproduce that, the default is the true maximum."""
    reward_environment = "shared_max_reward"


class SequentialEnvironment(DefaultEnvironment):
    """sequential_environment.py:20-64: object types are picked up as a chain t -> t+1 on the same cell."""
    reward_environment = "sequential"


class JunctionalEnvironment(DefaultEnvironment):
    """junctional_environment.py:20-71: the chain forks after type 1 depending on the agent."""
    reward_environment = "junctional"


class ObservationStatsEnvironment(DefaultEnvironment):
    """observation_stats_environment.py:16-61: the default step plus `observation_stats`, per-agent sums of every step
    observation that no reset zeroes (local view: the [C,V,V] crop; relative view: the decoded [C,X,Y] canvas; the
    reference leaves the merged view out, :46-52)."""
    reward_environment = "default"

    def __init__(self, config, world, reference_compat=False):
        super().__init__(config, world, reference_compat)
        self._stats = self.batched.new_heatmap_observation()

    @property
    def observation_stats(self):
        return list(self._stats.cpu().numpy())

    def step(self, action_n, order):
        result = super().step(action_n, order)
        if self.view_method in ("relative", "local"):
            self.batched.accumulate_heatmap_observation(self.out, self._stats)
        return result


def generate_environment(config, world):
    if config.environment == "default":
        return DefaultEnvironment(config=config, world=world)
    if config.environment == "shared_mean_reward":
        return SharedMeanRewardEnvironment(config=config, world=world)
    if config.environment == "shared_max_reward":
        return SharedMaxRewardEnvironment(config=config, world=world)
    if config.environment == "sequential":
        return SequentialEnvironment(config=config, world=world)
    if config.environment == "junctional":
        return JunctionalEnvironment(config=config, world=world)
    if config.environment == "observation_stats":  # core/environments/__init__.py:52-53
        return ObservationStatsEnvironment(config=config, world=world)
    logger.warning(f"Unexpected environment is given. config.environment: {config.environment}")
    raise ValueError()
