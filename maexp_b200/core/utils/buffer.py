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
        # inputs of the extra canvas channels the relative / DA6 models concatenate (state["objects"], state["agents"],
        # state["destination"], observation_handler.py:95-105): the per-agent objects snapshot as bit rows (env-minor, as the
        # step kernel stores them), every agent's cell (= positions) and the index of the agent's destination mask
        cfg = env.config
        canvas = cfg.observation_area_mask != "local"
        self.objects_bits = self.destination = self.dest_table = None
        if canvas and cfg.get("objects_channel"):
            self.objects_bits = torch.zeros((self.S, N, env.T, env.X, E), dtype=torch.int32, device=d)
        if canvas and cfg.get("destination_channel"):
            self.destination = torch.zeros((self.S, E, N), dtype=torch.uint8, device=d)
            masks = torch.from_numpy(env.destination_masks != 0).to(torch.int64)  # [M, X, Y]
            rows = (masks << torch.arange(env.Y, dtype=torch.int64)).sum(dim=-1)
            self.dest_table = rows.to(torch.int32).contiguous().to(d)
        self.keep_agents = bool(canvas and cfg.get("agents_channel"))
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
        if self.objects_bits is not None:
            out["objects_bits"] = self.objects_bits[s]
        if self.destination is not None:
            out["destination"] = self.destination[s]
        return out

    def commit(self):
        self.head = (self.head + 1) % self.S
        self.count = min(self.count + 1, self.S)

    def new_batch(self, batch_size):
        N, d, sf = self.env.N, self.env.device, self.env.slot_floats
        b = {"obs": torch.empty((N, batch_size, sf), dtype=torch.float32, device=d),
                "next_obs": torch.empty((N, batch_size, sf), dtype=torch.float32, device=d),
                "actions": torch.empty((N, batch_size), dtype=torch.int32, device=d),
                "rewards": torch.empty((N, batch_size), dtype=torch.float32, device=d),
                "dones": torch.empty((N, batch_size), dtype=torch.uint8, device=d),
                "picked": torch.empty((batch_size, 2), dtype=torch.int32, device=d),
                "pos": torch.empty((N, batch_size, 2), dtype=torch.int8, device=d),
                "next_pos": torch.empty((N, batch_size, 2), dtype=torch.int8, device=d),
                "error_flag": torch.zeros((1,), dtype=torch.int32, device=d)}
        e = self.env
        if self.objects_bits is not None:
            b["obj_rows"] = torch.zeros((N, batch_size, e.T, e.X), dtype=torch.int32, device=d)
            b["next_obj_rows"] = torch.zeros((N, batch_size, e.T, e.X), dtype=torch.int32, device=d)
        if self.keep_agents:
            b["agents"] = torch.zeros((N, batch_size, N, 2), dtype=torch.int8, device=d)
            b["next_agents"] = torch.zeros((N, batch_size, N, 2), dtype=torch.int8, device=d)
        if self.destination is not None:
            b["dest_rows"] = torch.zeros((N, batch_size, e.X), dtype=torch.int32, device=d)
            b["next_dest_rows"] = torch.zeros((N, batch_size, e.X), dtype=torch.int32, device=d)
        return b

    def sample(self, batch, seed=0, tick=0, indices=None):
        """Fills `batch` (from new_batch) with uniformly sampled DISTINCT transitions (buffer.py:61, replace=False);
        indices: int32 [B,2] injected (slot, env) picks."""
        B = batch["actions"].shape[1]
        view = _lib.ReplayView(self.obs.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(),
                               self.dones.data_ptr(), self.was_reset.data_ptr(),
                               None if self.positions is None else self.positions.data_ptr(), self.S, self.env.E, self.env.N,
                               self.env.slot_floats, self.head, self.count,
                               None if self.objects_bits is None else self.objects_bits.data_ptr(),
                               None if self.destination is None else self.destination.data_ptr(),
                               None if self.dest_table is None else self.dest_table.data_ptr(), self.env.T, self.env.X)
        out = _lib.Batch(*[batch[k].data_ptr() if k in batch else None
                           for k in ("obs", "next_obs", "actions", "rewards", "dones", "picked", "pos", "next_pos",
                                     "error_flag", "obj_rows", "next_obj_rows", "agents", "next_agents", "dest_rows",
                                     "next_dest_rows")])
        idx = None if indices is None else C.c_void_p(indices.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)
        _lib.check(self.lib.maexp_replay_sample(C.byref(view), B, seed & 0xFFFFFFFF, tick & 0xFFFFFFFF, idx,
                                                C.byref(out), stream))
        return batch

    def canvas_inputs(self, e0, b):
        """The extra canvas-channel inputs of the newest slot's states for envs [e0, e0 + b), in the learner's layout
        (what BatchedActor hands to LearnerGroup.q_values): obj_rows [N,b,T,X], agents [N,b,N,2], dest_rows [N,b,X]."""
        out, h, N = {}, self.head, self.env.N
        if self.objects_bits is not None:
            out["obj_rows"] = self.objects_bits[h, :, :, :, e0:e0 + b].permute(0, 3, 1, 2).contiguous()
        if self.keep_agents:
            out["agents"] = self.positions[h, e0:e0 + b].unsqueeze(0).expand(N, b, N, 2).contiguous()
        if self.destination is not None:
            idx = self.destination[h, e0:e0 + b].transpose(0, 1).long()  # [N, b]
            out["dest_rows"] = self.dest_table[idx].contiguous()
        return out

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
                                          (_lib.REPLAY_ERR_INDEX_INVALID, "an injected pick names a transition into a reset slot"),
                                          (_lib.REPLAY_ERR_NOT_UNIQUE, "the ring holds fewer distinct transitions than the batch "
                                           "(np.random.choice(replace=False) raises ValueError here)"))
                   if flag & bit]
            raise _lib.MaexpError("maexp_replay_sample: " + "; ".join(why))


def generate_buffer(config, env=None):
    """core/utils/buffer.py:21-33.  `capacity` counts whole-team transitions, as in the reference."""
    if config.buffer == "default":
        slots = max(2, min(int(config.capacity) // max(1, env.E) + 1, 1 << 16))
        return DeviceReplayRing(env, slots)
    logger.warning(f"Unexpected buffer is given. config.buffer: {config.buffer}")
    raise ValueError()
