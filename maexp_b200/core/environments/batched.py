"""E batched gridworld environments on one GPU, driven through the C-ABI.

This is the object behind the reference-style `generate_environment` (see __init__.py): it
owns the opaque `maexp_env` handle and the torch tensors the kernels write into.
PyTorch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from ..handlers.observations.masks import generate_observation_area_mask

AGENT_VIEWS = {"default": 0, "simple": 1, "individual": 2, "transition": 3, "individual_transition": 4}
OBJECT_VIEWS = {"default": 0, "simple": 1, "object_types": 2}
AREA_MASKS = {"local": 0, "relative": 1, "merged": 2}
NOISES = {False: 0, None: 0, "flip": 1, "sensing_dist": 2, "threshold_sensing_dist": 3, "flat": 4}
REWARD_MODES = {"default": 0, "shared_mean_reward": 1, "shared_max_reward": 2, "sequential": 0, "junctional": 0,
                "observation_stats": 0}  # observation_stats_environment.py:16-61 steps exactly like the default one
ENV_KINDS = {"sequential": 1, "junctional": 2}  # sequential_environment.py:20, junctional_environment.py:20


def noise_constants(config):
    """Per-ring fp32 constants exactly as the reference derives them
    (flip_observation_noise.py:28; dist_observation_noise.py:60,66)."""
    rng = np.zeros(_lib.MAX_RINGS, np.float32)
    std = np.zeros(_lib.MAX_RINGS, np.float32)
    thr = np.zeros(_lib.MAX_RINGS, np.float32)
    for i, p in enumerate(list(config.get("flip_noise_probabilities") or [])[:_lib.MAX_RINGS]):
        rng[i] = np.float32(0.5) / (np.float32(p) - np.float32(1.0))
    ratio = float(config.get("noise_threshold_ratio") or 0.0)
    for i, s in enumerate(list(config.get("std_distribution") or [])[:_lib.MAX_RINGS]):
        std[i] = np.float32(s)
        thr[i] = np.float32(float(s) * ratio)
    return rng, std, thr


def validate(config, world_map):
    """Up-front checks for the reference's silent failure modes."""
    V, N, T = int(config.visible_range), int(config.num_agents), int(config.type_objects)
    if config.agent_view_method == "self":
        raise ValueError("agent_view_method=self is not supported (IndexError in the reference, SURVEY.md §E-7)")
    noise = config.observation_noise
    if noise == "flip" and len(config.flip_noise_probabilities) < V // 2:
        raise ValueError(f"flip noise needs {V // 2} flip_noise_probabilities for visible_range={V} (§E-6)")
    if noise in ("sensing_dist", "threshold_sensing_dist") and len(config.std_distribution) < V // 2:
        raise ValueError(f"sensing noise needs {V // 2} std_distribution entries for visible_range={V}")
    if noise and config.observation_area_mask == "relative":
        raise ValueError("observation noise with observation_area_mask=relative fails in the reference (appendix C)")
    if len(config.init_xys) < N or len(config.agent_tasks) < N:
        raise ValueError("init_xys / agent_tasks shorter than num_agents")
    if config.shuffle_init_xys and len(config.init_xys) > _lib.MAX_AGENTS:
        # default_environment.py:33,97-98 shuffles ALL len(init_xys) rows and uses the first num_agents of the order
        raise ValueError(f"shuffle_init_xys supports at most {_lib.MAX_AGENTS} init_xys rows")
    env_name = config.environment
    if env_name == "junctional" and T < 5:
        raise ValueError("environment=junctional places object types 2, 3 and 4 (junctional_environment.py:44-51): type_objects >= 5")
    cells = set()
    rows = list(config.init_xys) if config.shuffle_init_xys else list(config.init_xys)[:N]
    for xy in rows:
        ix, iy = world_map.coord2ind(xy)
        if world_map.wall_matrix[ix, iy]:
            raise ValueError(f"init_xys {list(xy)} is on a wall")
        cells.add((int(ix), int(iy)))
    if len(cells) < len(rows):
        raise ValueError("init_xys must be distinct")
    free = (world_map.wall_matrix == 0)
    for t in range(T if env_name not in ("sequential", "junctional") else 1):  # those two only ever sample type 0
        if (world_map.objects_area_matrix[t] * free).sum() < int(config.num_objects) + N:
            # the reference's rejection sampler would never terminate (§E-8)
            raise ValueError(f"objects_area of type {t} cannot hold num_objects={config.num_objects}")
    for task in list(config.agent_tasks)[:N]:
        if str(task) != "-1" and any(int(ch) >= T for ch in str(task)):
            raise ValueError(f"agent task {task!r} names an object type >= type_objects={T}")


class BatchedEnvironment:
    def __init__(self, config, world, num_envs, device="cuda:0", auto_reset=True, first_global_env_id=0,
                 shared_max_is_mean=False):
        if not torch.cuda.is_available():
            raise _lib.MaexpError("BatchedEnvironment needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.config, self.world, self.map = config, world, world.map
        validate(config, self.map)
        self.device = torch.device(device)
        self.E, self.N, self.T = int(num_envs), int(config.num_agents), int(config.type_objects)
        self.X, self.Y, self.V = self.map.SIZE_X, self.map.SIZE_Y, int(config.visible_range)
        self.merged = config.observation_area_mask == "merged"
        self.tick = 0
        self.first_global_env_id = int(first_global_env_id)  # Philox keys are global: shards reproduce a single-GPU run

        cfg = _lib.EnvCfg()
        cfg.num_envs, cfg.num_agents, cfg.type_objects = self.E, self.N, self.T
        cfg.size_x, cfg.size_y, cfg.visible_range = self.X, self.Y, self.V
        cfg.past_step = int(config.get("past_step") or 0)
        cfg.agent_view = AGENT_VIEWS[config.agent_view_method]
        cfg.object_view = OBJECT_VIEWS[config.object_view_method]
        cfg.area_mask = AREA_MASKS[config.observation_area_mask]
        cfg.noise = NOISES[config.observation_noise]
        cfg.reward_mode = REWARD_MODES[config.environment]
        cfg.env_kind = ENV_KINDS.get(config.environment, 0)
        cfg.dynamic_obstacles = int(config.map.name == "dynamic_simple")  # dynamic_simple_map.py:29-63
        cfg.max_episode_length = int(config.max_episode_length)
        cfg.num_objects = int(config.num_objects)
        cfg.keep_objects_num = int(bool(config.keep_objects_num))
        cfg.shuffle_init_xys = int(bool(config.shuffle_init_xys))
        cfg.shared_max_is_mean = int(bool(shared_max_is_mean))
        cfg.auto_reset = int(bool(auto_reset))
        dest = np.ascontiguousarray(self.map.destination_masks(), np.int8)
        cfg.num_destinations = dest.shape[0]
        cfg.random_destinations = int(bool(self.map.RANDOM_DESTINATIONS))
        cfg.seed = int(config.seed) & 0xFFFFFFFF
        cfg.first_global_env_id = int(first_global_env_id)
        rows = list(config.init_xys) if config.shuffle_init_xys else list(config.init_xys)[:self.N]
        cfg.num_init_xys = len(rows)
        for i, xy in enumerate(rows):
            ix, iy = self.map.coord2ind(xy)
            cfg.init_xys[i][0], cfg.init_xys[i][1] = int(ix), int(iy)
        for i in range(self.N):
            task = str(config.agent_tasks[i])
            for q in range(_lib.MAX_TYPES):
                cfg.task_order[i][q] = -1
            if task == "-1":
                cfg.task_mask[i] = -1
                continue
            cfg.task_mask[i] = 0
            for q, ch in enumerate(task[:_lib.MAX_TYPES]):
                cfg.task_order[i][q] = int(ch)
                cfg.task_mask[i] |= 1 << int(ch)
        rng, std, thr = noise_constants(config)
        for i in range(_lib.MAX_RINGS):
            cfg.flip_noise_range[i], cfg.std_distribution[i], cfg.noise_threshold[i] = rng[i], std[i], thr[i]

        mask_cfg = config if not self.merged else config.merged_with(observation_area_mask="local")
        self.los_table = np.ascontiguousarray(generate_observation_area_mask(mask_cfg, world), np.int8)
        self._tables_keep = [np.ascontiguousarray(self.map.wall_matrix, np.int8),
                             np.ascontiguousarray(self.map.objects_area_matrix, np.int8),
                             np.ascontiguousarray(self.map.noise_area_matrix, np.int8), dest, self.los_table]
        tb = _lib.Tables(*[a.ctypes.data for a in self._tables_keep])
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.maexp_env_create(C.byref(cfg), C.byref(tb), C.byref(handle)))
        self.h = handle
        self.C = self.lib.maexp_env_channels(self.h)
        self.slot_floats = self.lib.maexp_env_slot_floats(self.h)
        self.K = self.C * self.V * self.V
        self.destination_masks = dest

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            self.lib.maexp_env_destroy(h)

    # ------------------------------------------------------------------ buffers
    @property
    def model_obs_shape(self):
        """obs_shape the reference hands to generate_brain: [C,V,V], or [C,X,Y] for the relative / merged views
        (observation_handler.py:60-70, merged_observtion_handler.py:37-39)."""
        if self.config.observation_area_mask == "local":
            return [self.C, self.V, self.V]
        return [self.C, int(self.map.SIZE_X), int(self.map.SIZE_Y)]

    def new_heatmaps(self):
        """Zeroed int32 buffer for out["heatmaps"] and a function splitting it into the reference's named arrays
        (default_environment.py:70-96)."""
        n = int(self.lib.maexp_env_heatmap_ints(self.h))
        buf = torch.zeros((n,), dtype=torch.int32, device=self.device)
        N, T, X, Y = self.N, self.T, self.X, self.Y

        def split(b=buf):
            sizes = [("heatmap_agents", (N, X, Y)), ("heatmap_complete", (N, X, Y)), ("heatmap_objects", (T, X, Y)),
                     ("heatmap_objects_left", (T, X, Y)), ("heatmap_wall_collision", (X, Y)),
                     ("heatmap_agents_collision", (X, Y))]
            out, off = {}, 0
            for name, shape in sizes:
                cnt = int(np.prod(shape))
                out[name] = b[off:off + cnt].view(*shape)
                off += cnt
            return out
        return buf, split

    def new_heatmap_observation(self):
        """Zeroed fp32 [N, *observation_space] for `accumulate_heatmap_observation` (default_environment.py:93-95)."""
        return torch.zeros((self.N, *self.model_obs_shape), dtype=torch.float32, device=self.device)

    def accumulate_heatmap_observation(self, out, heat, chunk_envs=8192):
        """heatmap_observation += the observations `step(out)` just wrote, summed over the environments that were
        not reset by that call (default_environment.py:238-247: the local observation, or the decoded relative
        canvas on the relative / merged views — the merged view needs out["positions"])."""
        assert heat.is_cuda and heat.is_contiguous() and heat.dtype == torch.float32
        E, N, S = self.E, self.N, self.slot_floats
        skip = self._p(out["was_reset"])
        if self.config.observation_area_mask == "local":
            _lib.check(self.lib.maexp_heatmap_observation_add(self._p(out["obs"]), E, N, self.K, S, E * S, skip,
                                                              self._p(heat), self._stream()))
            return heat
        if out.get("positions") is None:
            raise ValueError("heatmap_observation on the relative / merged views needs out['positions'] (new_outputs(extras=True))")
        slot = out["obs_rel"] if self.merged else out["obs"]
        pos = out["positions"].transpose(0, 1).contiguous()  # [N, E, 2], the slot's row order
        X, Y = self.X, self.Y
        canvas = torch.empty((min(E, chunk_envs), self.C, X, Y), dtype=torch.float32, device=self.device)
        for n in range(N):
            for e0 in range(0, E, chunk_envs):
                rows = min(chunk_envs, E - e0)
                _lib.check(self.lib.maexp_decode_relative(self._p(slot[n, e0:]), self._p(pos[n, e0:]), self._p(canvas), rows,
                                                          self.C, self.V, X, Y, S, 0, self._stream()))
                _lib.check(self.lib.maexp_heatmap_observation_add(
                    self._p(canvas), rows, 1, self.C * X * Y, self.C * X * Y, 0, self._p(out["was_reset"][e0:]),
                    self._p(heat[n]), self._stream()))
        return heat

    def new_slot(self):
        """One replay-ring slot, agent-major: fp32 [N, E, slot_floats] (slot_floats = C*V*V padded to 32 bytes)."""
        return torch.empty((self.N, self.E, self.slot_floats), dtype=torch.float32, device=self.device)

    def new_outputs(self, extras=False):
        d = self.device
        o = {
            "obs": self.new_slot(),
            "obs_rel": self.new_slot() if self.merged else None,
            "rewards": torch.zeros((self.E, self.N), dtype=torch.float32, device=d),
            "dones": torch.zeros((self.E, self.N), dtype=torch.uint8, device=d),
            "was_reset": torch.zeros((self.E,), dtype=torch.uint8, device=d),
        }
        if extras:
            o["coordinates"] = torch.zeros((self.E, self.N, 12), dtype=torch.int32, device=d)
            o["positions"] = torch.zeros((self.E, self.N, 2), dtype=torch.int8, device=d)
            o["objects"] = torch.zeros((self.E, self.N, self.T, self.X, self.Y), dtype=torch.int8, device=d)
            o["destination"] = torch.zeros((self.E, self.N), dtype=torch.uint8, device=d)
        return o

    def view(self, slot):
        """Dense env-major [E, N, C, V, V] copy of an (agent-major) ring slot, without the alignment padding."""
        return slot[..., :self.K].reshape(self.N, self.E, self.C, self.V, self.V).transpose(0, 1).contiguous()

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _out_struct(self, out):
        s = _lib.ObsOut()
        for name in ("obs", "obs_rel", "rewards", "dones", "was_reset", "coordinates", "positions", "objects",
                     "destination", "actions_taken", "rewards_i8", "done_env", "heatmaps", "objects_bits"):
            t = out.get(name)
            if t is not None:
                assert t.is_cuda and t.is_contiguous()
                setattr(s, name, t.data_ptr())
        return s

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _next_tick(self, tick):
        if tick is None:
            tick = self.tick
        self.tick = tick + 1
        return tick & 0xFFFFFFFF

    # ------------------------------------------------------------------ calls
    def reset(self, out, tick=None, env_mask=None, noise_tape=None):
        """DefaultEnvironment.reset for the masked envs (default_environment.py:64-116)."""
        s = self._out_struct(out)
        _lib.check(self.lib.maexp_env_reset(self.h, self._next_tick(tick), self._p(env_mask), self._p(noise_tape),
                                            C.byref(s), self._stream()))
        return out

    def step(self, out, actions=None, order=None, tick=None, noise_tape=None):
        """DefaultEnvironment.step(action_n, order) for all envs (default_environment.py:153-175).
        actions int8 [E,N] / order uint8 [E,N] device tensors; None = on-chip Philox draws."""
        s = self._out_struct(out)
        _lib.check(self.lib.maexp_env_step(self.h, self._next_tick(tick), self._p(actions), self._p(order),
                                           self._p(noise_tape), C.byref(s), self._stream()))
        return out

    def step_host(self, actions_host, obs_slot, rewards_host, dones_host, was_reset_host, tick=None):
        """Host-buffer entry point (pinned torch CPU tensors): H2D actions, step, D2H rewards/dones."""
        def hp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        _lib.check(self.lib.maexp_env_step_host(self.h, self._next_tick(tick), hp(actions_host),
                                                C.c_void_p(obs_slot.data_ptr()), hp(rewards_host), hp(dones_host),
                                                hp(was_reset_host), self._stream()))

    def step_host_async(self, actions_host, obs_slot, rewards_host, dones_host, was_reset_host, tick=None):
        """Pipelined step_host: returns a ticket at once; the step's host buffers are valid after host_wait(ticket).
        Use pinned tensors and a distinct set per in-flight step (3 deep)."""
        def hp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        ticket = C.c_uint64()
        _lib.check(self.lib.maexp_env_step_host_async(self.h, self._next_tick(tick), hp(actions_host),
                                                      C.c_void_p(obs_slot.data_ptr()), hp(rewards_host), hp(dones_host),
                                                      hp(was_reset_host), self._stream(), C.byref(ticket)))
        return ticket.value

    def step_host_compact_async(self, actions_host, obs_slot, rewards_i8_host, done_env_host, was_reset_host, tick=None):
        """step_host_async with int8 rewards [E,N] and one done flag per env (default reward mode only)."""
        def hp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        ticket = C.c_uint64()
        _lib.check(self.lib.maexp_env_step_host_compact_async(self.h, self._next_tick(tick), hp(actions_host),
                                                              C.c_void_p(obs_slot.data_ptr()), hp(rewards_i8_host),
                                                              hp(done_env_host), hp(was_reset_host), self._stream(),
                                                              C.byref(ticket)))
        return ticket.value

    def host_wait(self, ticket=0):
        _lib.check(self.lib.maexp_env_host_wait(self.h, int(ticket)))

    def state(self):
        d = self.device
        pos = torch.zeros((self.E, self.N, 2), dtype=torch.int8, device=d)
        obj = torch.zeros((self.E, self.T, self.X, self.Y), dtype=torch.int8, device=d)
        cnt = torch.zeros((self.E, 4), dtype=torch.int32, device=d)
        steps = torch.zeros((self.E,), dtype=torch.int32, device=d)
        dest = torch.zeros((self.E, self.N, self.X, self.Y), dtype=torch.int8, device=d)
        _lib.check(self.lib.maexp_env_get_state(self.h, self._p(pos), self._p(obj), self._p(cnt), self._p(steps),
                                                self._p(dest), self._stream()))
        return {"pos": pos, "objects": obj, "counters": cnt, "steps": steps, "destination": dest}

    def set_state(self, pos=None, objects=None):
        if pos is not None:
            pos = torch.as_tensor(np.asarray(pos, np.int8)).reshape(self.E, self.N, 2).contiguous().to(self.device)
        if objects is not None:
            objects = torch.as_tensor(np.asarray(objects, np.int8)).reshape(self.E, self.T, self.X, self.Y)
            objects = objects.contiguous().to(self.device)
        _lib.check(self.lib.maexp_env_set_state(self.h, self._p(pos), self._p(objects), self._stream()))
        torch.cuda.synchronize(self.device)

    def uses_lanes_kernel(self):
        """True when plain step / reset calls run the agent-parallel kernel (csrc/env_lanes.cu)."""
        return bool(self.lib.maexp_env_agent_parallel(self.h))

    def stats(self):
        steps, resets, err = C.c_uint64(), C.c_uint64(), C.c_int32()
        _lib.check(self.lib.maexp_env_stats(self.h, C.byref(steps), C.byref(resets), C.byref(err), self._stream()))
        return {"env_steps": steps.value, "resets": resets.value, "error_flag": err.value}
