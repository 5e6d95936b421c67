"""`generate_observation_handler` (core/handlers/observations/__init__.py:21-28).

The observation *builder* is the CUDA step kernel; these classes carry what the rest of the
reference reads from a handler — channel count, observation_space, the visibility table — and
the static decoders, which run on the GPU (maexp_decode_relative).
"""
import ctypes as C

import torch

from .... import _lib
from ...utils.logging import initialize_logging
from .masks import COORDINATE_KEYS, generate_observation_area_mask, window_bounds

logger = initialize_logging(__name__)

__all__ = ["ObservationHandler", "MergedObservationHandler", "generate_observation_handler"]


def agent_channels(config):
    """get_channel of core/handlers/observations/agents/*.py"""
    n, p = int(config.num_agents), int(config.get("past_step") or 0)
    table = {"default": 1, "simple": 1, "individual": n, "transition": 1 + p, "individual_transition": n + 1 + p}
    if config.agent_view_method not in table:
        logger.warning(f"Unexpected agent_view_method is given. config.agent_view_method: {config.agent_view_method}")
        raise ValueError()
    return table[config.agent_view_method]


def object_channels(config):
    """get_channel of core/handlers/observations/objects/*.py"""
    table = {"default": 1, "simple": 1, "object_types": int(config.type_objects)}
    if config.object_view_method not in table:
        logger.warning(f"Unexpected object_view_method is given. config.object_view_method: {config.object_view_method}")
        raise ValueError()
    return table[config.object_view_method]


class ObservationHandler:
    """observation_handler.py:25-66."""

    def __init__(self, config, world):
        self.config, self.world = config, world
        self.view_method = config.observation_area_mask
        if self.view_method == "local":
            self.observation_size = [config.visible_range, config.visible_range]
        elif self.view_method == "relative":
            self.observation_size = [world.map.SIZE_X, world.map.SIZE_Y]
        else:
            logger.warning(f"Unexpected view_method is given. self.view_method: {self.view_method}")
            raise ValueError()
        self.observation_area_mask = generate_observation_area_mask(config=config, world=world)

    @property
    def get_channel(self):
        return agent_channels(self.config) + object_channels(self.config) + 1

    @property
    def observation_space(self):
        return [self.get_channel, *self.observation_size]

    def get_mask_coordinates(self, pos_x, pos_y):
        return window_bounds(int(pos_x), int(pos_y), self.world.map.SIZE_X, self.world.map.SIZE_Y,
                             int(self.config.visible_range))

    @staticmethod
    def decode_relative_state(state, observation_size, only_area_plus_one=False):
        """observation_handler.py:150-216 on the GPU.  state["relative"]: fp32 [C,V,V] or [B,C,V,V] (CUDA),
        state["agents"] or state["positions"]: the observing agents' index-space positions ([2] / [B,2])."""
        assert "relative" in state.keys()
        x = state["relative"]
        single = x.dim() == 3
        if x.dim() not in (3, 4):
            logger.warning(f"Unexpected state length is given. len(state['relative'].shape): {x.dim()}")
            raise ValueError()
        if not x.is_cuda:
            raise _lib.MaexpError("decode_relative_state needs CUDA tensors (there is no CPU fallback)")
        x = x.unsqueeze(0) if single else x
        B, Cn, V, _ = x.shape
        pos = state["positions"].reshape(B, 2).to(torch.int8).contiguous()
        rows = x.reshape(B, -1).contiguous().float()
        X, Y = int(observation_size[0]), int(observation_size[1])
        out = torch.empty((B, 1 if only_area_plus_one else Cn, X, Y), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        _lib.check(lib.maexp_decode_relative(C.c_void_p(rows.data_ptr()), C.c_void_p(pos.data_ptr()),
                                             C.c_void_p(out.data_ptr()), B, Cn, V, X, Y, rows.shape[1],
                                             int(only_area_plus_one),
                                             C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return out[0] if single else out


class MergedObservationHandler:
    """merged_observtion_handler.py:19-40: a local handler (with noise) plus a relative one (without)."""

    def __init__(self, config, world):
        self.local_view_observation_handler = ObservationHandler(config.merged_with(observation_area_mask="local"), world)
        self.relative_view_observation_handler = ObservationHandler(
            config.merged_with(observation_area_mask="relative", observation_noise=False), world)

    @property
    def get_channel(self):
        return self.local_view_observation_handler.get_channel

    @property
    def observation_space(self):
        return [self.local_view_observation_handler.observation_space[0],
                *self.relative_view_observation_handler.observation_space[1:]]


def generate_observation_handler(config, world):
    if config.observation_area_mask == "merged":
        return MergedObservationHandler(config=config, world=world)
    return ObservationHandler(config=config, world=world)
