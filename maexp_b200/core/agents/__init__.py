"""`generate_agents` (core/agents/__init__.py:22-54) and the batched epsilon-greedy actor.

`DefaultAgent` keeps the reference's per-agent interface; `BatchedActor` selects the actions of
every (env, agent) pair on the GPU in chunks (Q forward + maexp_select_actions).
"""
import ctypes as C
import random

import torch

from ... import _lib
from ..utils.logging import initialize_logging
from .brains import LearnerGroup, generate_brain

logger = initialize_logging(__name__)


class DefaultAgent:
    """abstract_agent.py:15-46 + default_agent.py:12-19."""

    def __init__(self, config, obs_shape, act_size):
        self.config, self.obs_shape, self.act_size = config, obs_shape, act_size
        self.brain = generate_brain(config=config, obs_shape=obs_shape, act_size=act_size)

    def get_random_action(self):
        return int(random.random() * self.act_size)

    def get_action(self, state, epsilon):
        if random.random() < epsilon:
            return self.get_random_action()
        return self.brain.get_action(state)

    def learn(self, state, action, reward, done, next_state):
        return self.brain.learn(state, action, reward, done, next_state)

    def synchronize_brain(self):
        self.brain.synchronize_network()


class AttentionAgent(DefaultAgent):
    """attention_agent.py:13-27: the brain returns (action, attns); the maps of each group are stacked over the blocks."""

    def get_action_attns(self, state, epsilon):
        action, attns = self.brain.get_action(state)
        attns = [torch.stack(attn) for attn in attns]
        if random.random() < epsilon:
            action = self.get_random_action()
        return action, attns

    def get_action(self, state, epsilon):
        return self.get_action_attns(state, epsilon)[0]


def generate_agents(config, observation_space, action_space):
    if config.agent_type == "default":
        return [DefaultAgent(config, obs, act) for obs, act in zip(observation_space, action_space)]
    if config.agent_type == "attention":
        return [AttentionAgent(config, obs, act) for obs, act in zip(observation_space, action_space)]
    logger.warning(f"Unexpected agent type is given. config.agent_type: {config.agent_type}")
    raise ValueError()


class BatchedActor:
    """Epsilon-greedy actions for all environments of a BatchedEnvironment from a LearnerGroup's online nets."""

    def __init__(self, config, env, learner, chunk=None):
        self.lib = _lib.load()
        self.env, self.learner = env, learner
        if chunk is None:
            # observations per agent and forward call.  The conv / MLP families keep ~3 KB of activations per observation, so
            # one call takes the whole slot (measured at 65 536 envs x 8 agents, C51: 3.69 / 3.28 / 3.12 / 2.82 ms per act() at
            # chunks of 8 192 / 16 384 / 32 768 / 65 536 -- fewer launches and no partial last wave of 128-row tiles); the
            # transformer families keep ~100 KB per observation (tokens x embedding per block) and stay at 8 192.
            # Relative / merged views decode a [C, X, Y] canvas per observation: also 8 192.
            attention = str(config.model.name) in ("da3", "da6", "da3_iqn", "da6_iqn")
            chunk = 8192 if attention or config.observation_area_mask != "local" else 65536
        self.chunk = min(int(chunk), env.E)
        self.group = LearnerGroup(config, env.model_obs_shape, 4, env.N, self.chunk, obs_stride=env.slot_floats,
                                  device=env.device)
        self.actions = torch.zeros((env.E, env.N), dtype=torch.int8, device=env.device)
        self.no_task = torch.tensor([1 if str(t) == "-1" else 0 for t in list(config.agent_tasks)[:env.N]],
                                    dtype=torch.uint8, device=env.device)
        self.seed = int(config.seed) & 0xFFFFFFFF

    def refresh(self):
        """Copy the learner's online parameters (the actor is a separate handle sized for acting batches)."""
        self.group.params().copy_(self.learner.params())

    def act(self, obs_slot, epsilon, tick, positions=None, ring=None):
        """obs_slot fp32 [N, E, slot] (an agent-major ring slot) -> int8 actions [E, N] on the device.
        positions int8 [E, N, 2]: the step's agent cells, needed by the relative / merged models; ring: the
        DeviceReplayRing whose newest slot is obs_slot, for the models that read the extra canvas channels."""
        env, stream = self.env, C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)
        for e0 in range(0, env.E, self.chunk):
            b = min(self.chunk, env.E - e0)
            obs = obs_slot[:, e0:e0 + b].contiguous()  # [N, b, slot]
            pos = None
            if self.group.area_mask != "local":
                pos = positions[e0:e0 + b].transpose(0, 1).contiguous()  # [N, b, 2]
            canvas = ring.canvas_inputs(e0, b) if ring is not None and any(self.group.canvas_reads) else None
            q = self.group.q_values(obs, pos=pos, canvas=canvas)
            _lib.check(self.lib.maexp_select_actions(C.c_void_p(q.data_ptr()), C.c_void_p(self.actions.data_ptr()),
                                                     env.N, b, 4, e0, float(epsilon), self.seed,
                                                     env.first_global_env_id, tick & 0xFFFFFFFF, C.c_void_p(self.no_task.data_ptr()), stream))
        return self.actions
