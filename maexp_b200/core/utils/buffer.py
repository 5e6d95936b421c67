"""Device-resident replay ring (replaces core/utils/buffer.py:36-63 + core/utils/dataset.py:24-29).

The reference keeps a deque of python Experience tuples and collates a batch per step; here
every env step writes its observations, rewards, dones and actions straight into slot `head` of
ring tensors in HBM, and `sample()` is one fused pick+gather launch (maexp_replay_sample).
"""
import ctypes as C

import torch

from ... import _lib
from .logging import initialize_logging

logger = initialize_logging(__name__)


class DeviceReplayRing:
    def __init__(self, env, num_slots, keep_positions=False):
        self.lib = _lib.load()
        self.env, self.S = env, int(num_slots)
        E, N, d = env.E, env.N, env.device
        self.obs = torch.zeros((self.S, N, E, env.slot_floats), dtype=torch.float32, device=d)  # agent-major slots
        self.actions = torch.zeros((self.S, E, N), dtype=torch.int8, device=d)
        self.rewards = torch.zeros((self.S, E, N), dtype=torch.float32, device=d)
        self.dones = torch.zeros((self.S, E, N), dtype=torch.uint8, device=d)
        self.was_reset = torch.ones((self.S, E), dtype=torch.uint8, device=d)
        # the agents' cells stand in for the `coordinates` dict, which only the relative / merged models read
        # (observation_handler.py:150-216); local-view rings do not spend a store per agent on them
        self.positions = None if env.config.observation_area_mask == "local" and not keep_positions else \
            torch.zeros((self.S, E, N, 2), dtype=torch.int8, device=d)
        self.head, self.count = -1, 0

    def __len__(self):
        """Number of stored whole-team transitions (valid or not), as ReplayBuffer.__len__."""
        return max(0, self.count - 1) * self.env.E

    def next_outputs(self):
        """Output dict for env.step / env.reset that targets the next ring slot; call commit() after."""
        s = (self.head + 1) % self.S
        out = {"obs": self.obs[s], "rewards": self.rewards[s], "dones": self.dones[s],
               "was_reset": self.was_reset[s], "actions_taken": self.actions[s]}
        if self.positions is not None:
            out["positions"] = self.positions[s]
        return out

    def commit(self):
        self.head = (self.head + 1) % self.S
        self.count = min(self.count + 1, self.S)

    def new_batch(self, batch_size):
        N, d, sf = self.env.N, self.env.device, self.env.slot_floats
        return {"obs": torch.empty((N, batch_size, sf), dtype=torch.float32, device=d),
                "next_obs": torch.empty((N, batch_size, sf), dtype=torch.float32, device=d),
                "actions": torch.empty((N, batch_size), dtype=torch.int32, device=d),
                "rewards": torch.empty((N, batch_size), dtype=torch.float32, device=d),
                "dones": torch.empty((N, batch_size), dtype=torch.uint8, device=d),
                "picked": torch.empty((batch_size, 2), dtype=torch.int32, device=d),
                "pos": torch.empty((N, batch_size, 2), dtype=torch.int8, device=d),
                "next_pos": torch.empty((N, batch_size, 2), dtype=torch.int8, device=d),
                "error_flag": torch.zeros((1,), dtype=torch.int32, device=d)}

    def sample(self, batch, seed=0, tick=0, indices=None):
        """Fills `batch` (from new_batch) with uniformly sampled transitions; indices: int32 [B,2]."""
        B = batch["actions"].shape[1]
        view = _lib.ReplayView(self.obs.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(),
                               self.dones.data_ptr(), self.was_reset.data_ptr(),
                               None if self.positions is None else self.positions.data_ptr(), self.S, self.env.E, self.env.N,
                               self.env.slot_floats, self.head, self.count)
        out = _lib.Batch(*[batch[k].data_ptr() if k in batch else None
                           for k in ("obs", "next_obs", "actions", "rewards", "dones", "picked", "pos", "next_pos",
                                     "error_flag")])
        idx = None if indices is None else C.c_void_p(indices.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)
        _lib.check(self.lib.maexp_replay_sample(C.byref(view), B, seed & 0xFFFFFFFF, tick & 0xFFFFFFFF, idx,
                                                C.byref(out), stream))
        return batch

    @staticmethod
    def check(batch):
        """Raises when a sample() into `batch` could not find a valid transition or was given bad picks (synchronises;
        the trainers call it once per epoch).  The reference cannot fail this way: its deque only holds transitions."""
        flag = int(batch["error_flag"].item()) if "error_flag" in batch else 0
        if flag:
            batch["error_flag"].zero_()
            why = [text for bit, text in ((_lib.REPLAY_ERR_NO_VALID, "no valid transition in the ring (every stored step is "
                                           "followed by a reset: the ring is too short for the episode length)"),
                                          (_lib.REPLAY_ERR_INDEX_RANGE, "an injected (slot, env) pick lies outside the ring"),
                                          (_lib.REPLAY_ERR_INDEX_INVALID, "an injected pick names a transition into a reset slot"))
                   if flag & bit]
            raise _lib.MaexpError("maexp_replay_sample: " + "; ".join(why))


def generate_buffer(config, env=None):
    """core/utils/buffer.py:21-33.  `capacity` counts whole-team transitions, as in the reference."""
    if config.buffer == "default":
        slots = max(2, min(int(config.capacity) // max(1, env.E) + 1, 1 << 16))
        return DeviceReplayRing(env, slots)
    logger.warning(f"Unexpected buffer is given. config.buffer: {config.buffer}")
    raise ValueError()
