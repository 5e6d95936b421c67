"""ctypes front-end of the CPU oracle (oracle/maexp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py.  Never imported by maexp_b200/.

The config argument is a plain mapping with the reference's own keys
(configs/experiments/*.yaml), e.g. ``num_agents``,
``visible_range``, ``agent_view_method`` ...
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_AGENTS, MAX_TYPES, MAX_RINGS = 16, 8, 4
MAPS = {"simple": 0, "four_rooms": 1, "three_rooms": 2, "four_rectangle": 3, "central_room": 4,
        "central_room_destination": 5, "central_room_large": 6, "central_room_large_destination": 7, "two_central_rooms": 8,
        "dynamic_simple": 9}
AGENT_VIEWS = {"default": 0, "simple": 1, "individual": 2, "transition": 3, "individual_transition": 4}
OBJECT_VIEWS = {"default": 0, "simple": 1, "object_types": 2}
MASKS = {"local": 0, "relative": 1, "merged": 2}
NOISES = {False: 0, None: 0, "flip": 1, "sensing_dist": 2, "threshold_sensing_dist": 3, "flat": 4}
REWARDS = {"default": 0, "shared_mean_reward": 1, "shared_max_reward": 2, "sequential": 0, "junctional": 0,
           "observation_stats": 0}  # observation_stats_environment.py: DefaultEnvironment.step plus a never-zeroed observation sum
ENV_KINDS = {"sequential": 1, "junctional": 2}
COORD_KEYS = ("obs_x_min", "obs_x_max", "obs_y_min", "obs_y_max", "area_x_min", "area_x_max", "area_y_min",
              "area_y_max", "map_x_min", "map_x_max", "map_y_min", "map_y_max")


class OrcCfg(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int32), ("num_agents", C.c_int32), ("type_objects", C.c_int32), ("map_id", C.c_int32),
        ("size_x", C.c_int32), ("size_y", C.c_int32), ("visible_range", C.c_int32), ("past_step", C.c_int32),
        ("agent_view", C.c_int32), ("object_view", C.c_int32), ("area_mask", C.c_int32), ("noise", C.c_int32),
        ("reward_mode", C.c_int32), ("transparent_observation", C.c_int32), ("max_episode_length", C.c_int32),
        ("num_objects", C.c_int32), ("keep_objects_num", C.c_int32), ("shuffle_init_xys", C.c_int32),
        ("shared_max_is_mean", C.c_int32), ("auto_reset", C.c_int32), ("seed", C.c_uint32),
        ("first_global_env_id", C.c_uint32), ("init_xys", (C.c_int32 * 2) * MAX_AGENTS),
        ("task_mask", C.c_int32 * MAX_AGENTS), ("task_order", (C.c_int32 * MAX_TYPES) * MAX_AGENTS),
        ("flip_noise_range", C.c_float * MAX_RINGS), ("std_distribution", C.c_float * MAX_RINGS),
        ("noise_threshold", C.c_float * MAX_RINGS), ("env_kind", C.c_int32),
        ("num_init_xys", C.c_int32),
    ]


class OrcOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs", "obs_rel", "rewards", "dones", "was_reset", "coordinates",
                                           "agents", "objects", "destination", "noise_tape", "heatmaps")]


def build(force=False):
    """Compile oracle/libmaexp_oracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libmaexp_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("maexp_oracle.c", "maexp_oracle.h", "philox.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcCfg)]
        L.orc_destroy.argtypes = [C.c_void_p]
        for f in ("orc_channels", "orc_obs_elems", "orc_size_x", "orc_size_y"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.orc_get_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.orc_reset.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(OrcOut)]
        L.orc_step.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(OrcOut), C.c_int]
        L.orc_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_philox.argtypes = [C.c_uint32] * 6 + [C.c_void_p]
        _LIB = L
    return _LIB


def philox(k0, k1, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(k0 & 0xFFFFFFFF, k1 & 0xFFFFFFFF, c0 & 0xFFFFFFFF, c1 & 0xFFFFFFFF, c2 & 0xFFFFFFFF,
                     c3 & 0xFFFFFFFF, out)
    return [int(v) for v in out]


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def noise_tables(cfg):
    """fp32 per-ring constants exactly as the reference derives them.

    flip_observation_noise.py:28: 0.5 / (torch.tensor(p) - 1) in fp32.
    dist_observation_noise.py:60,66: std*ratio is a python double that torch compares
    against fp32 noise in fp32 (wrapped-number promotion)."""
    rng = np.zeros(MAX_RINGS, np.float32)
    std = np.zeros(MAX_RINGS, np.float32)
    thr = np.zeros(MAX_RINGS, np.float32)
    probs = _get(cfg, "flip_noise_probabilities") or []
    for i, p in enumerate(list(probs)[:MAX_RINGS]):
        rng[i] = np.float32(0.5) / (np.float32(p) - np.float32(1.0))
    stds = _get(cfg, "std_distribution") or []
    ratio = _get(cfg, "noise_threshold_ratio") or 0.0
    for i, s in enumerate(list(stds)[:MAX_RINGS]):
        std[i] = np.float32(s)
        thr[i] = np.float32(float(s) * float(ratio))
    return rng, std, thr


def make_cfg(cfg, num_envs=1, auto_reset=False, first_global_env_id=0, shared_max_is_mean=False):
    c = OrcCfg()
    m = _get(cfg, "map")
    c.num_envs = num_envs
    c.num_agents = N = int(_get(cfg, "num_agents"))
    c.type_objects = int(_get(cfg, "type_objects"))
    c.map_id = MAPS[_get(m, "name")]
    c.size_x, c.size_y = int(_get(m, "SIZE_X", 0) or 0), int(_get(m, "SIZE_Y", 0) or 0)
    c.visible_range = int(_get(cfg, "visible_range"))
    c.past_step = int(_get(cfg, "past_step", 0) or 0)
    c.agent_view = AGENT_VIEWS[_get(cfg, "agent_view_method")]
    c.object_view = OBJECT_VIEWS[_get(cfg, "object_view_method")]
    c.area_mask = MASKS[_get(cfg, "observation_area_mask")]
    c.noise = NOISES[_get(cfg, "observation_noise")]
    c.reward_mode = REWARDS[_get(cfg, "environment")]
    c.env_kind = ENV_KINDS.get(_get(cfg, "environment"), 0)
    c.transparent_observation = int(bool(_get(cfg, "transparent_observation")))
    c.max_episode_length = int(_get(cfg, "max_episode_length"))
    c.num_objects = int(_get(cfg, "num_objects"))
    c.keep_objects_num = int(bool(_get(cfg, "keep_objects_num")))
    c.shuffle_init_xys = int(bool(_get(cfg, "shuffle_init_xys")))
    c.shared_max_is_mean = int(bool(shared_max_is_mean))
    c.auto_reset = int(bool(auto_reset))
    c.seed = int(_get(cfg, "seed")) & 0xFFFFFFFF
    c.first_global_env_id = first_global_env_id
    c.num_init_xys = len(_get(cfg, "init_xys"))
    for i, xy in enumerate(_get(cfg, "init_xys")):
        c.init_xys[i][0], c.init_xys[i][1] = int(xy[0]), int(xy[1])
    tasks = _get(cfg, "agent_tasks")
    for i in range(N):
        task = str(tasks[i])
        for q in range(MAX_TYPES):
            c.task_order[i][q] = -1
        if task == "-1":
            c.task_mask[i] = -1
            continue
        mask = 0
        for q, ch in enumerate(task):
            c.task_order[i][q] = int(ch)
            mask |= 1 << int(ch)
        c.task_mask[i] = mask
    rng, std, thr = noise_tables(cfg)
    for i in range(MAX_RINGS):
        c.flip_noise_range[i], c.std_distribution[i], c.noise_threshold[i] = rng[i], std[i], thr[i]
    return c


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """E independent environments stepped by the C restatement."""

    def __init__(self, cfg, num_envs=1, auto_reset=False, first_global_env_id=0, extras=False,
                 shared_max_is_mean=False):
        self.L = lib()
        self.c = make_cfg(cfg, num_envs, auto_reset, first_global_env_id, shared_max_is_mean)
        self.h = self.L.orc_create(C.byref(self.c))
        if not self.h:
            raise ValueError("oracle: unsupported configuration")
        self.E, self.N, self.T, self.V = num_envs, self.c.num_agents, self.c.type_objects, self.c.visible_range
        self.Cn = self.L.orc_channels(self.h)
        self.X, self.Y = self.L.orc_size_x(self.h), self.L.orc_size_y(self.h)
        self.merged = self.c.area_mask == MASKS["merged"]
        self.extras = extras
        self.heatmaps = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def tables(self):
        wall = np.zeros((self.X, self.Y), np.int8)
        area = np.zeros((self.T, self.X, self.Y), np.int8)
        los = np.zeros((self.X, self.Y, self.V, self.V), np.int8)
        self.L.orc_get_tables(self.h, _ptr(wall), _ptr(area), _ptr(los))
        return wall, area, los

    def _out(self, noise_tape):
        E, N, T, V, X, Y, Cn = self.E, self.N, self.T, self.V, self.X, self.Y, self.Cn
        o = {
            "obs": np.zeros((E, N, Cn, V, V), np.float32),
            "obs_rel": np.zeros((E, N, Cn, V, V), np.float32) if self.merged else None,
            "rewards": np.zeros((E, N), np.float32),
            "dones": np.zeros((E, N), np.uint8),
            "was_reset": np.zeros((E,), np.uint8),
        }
        if self.extras:
            o["coordinates"] = np.zeros((E, N, 12), np.int32)
            o["agents"] = np.zeros((E, N, N, 2), np.int8)
            o["objects"] = np.zeros((E, N, T, X, Y), np.int8)
            o["destination"] = np.zeros((E, N, X, Y), np.int8)
        s = OrcOut()
        for k, v in o.items():
            setattr(s, k, None if v is None else v.ctypes.data)
        if noise_tape is not None:
            noise_tape = np.ascontiguousarray(noise_tape, np.float32)
            assert noise_tape.shape == (E, N, Cn, V, V)
            s.noise_tape = noise_tape.ctypes.data
        o["_keep"] = noise_tape
        if self.heatmaps is not None:
            s.heatmaps = self.heatmaps.ctypes.data
        return o, s

    def enable_heatmaps(self):
        """Allocates the visit statistics (summed over envs) and returns the named views of default_environment.py:70-96."""
        N, T, X, Y = self.N, self.T, self.X, self.Y
        self.heatmaps = np.zeros(((2 * N + 2 * T + 2) * X * Y,), np.int32)
        out, off = {}, 0
        for name, shape in (("heatmap_agents", (N, X, Y)), ("heatmap_complete", (N, X, Y)), ("heatmap_objects", (T, X, Y)),
                            ("heatmap_objects_left", (T, X, Y)), ("heatmap_wall_collision", (X, Y)),
                            ("heatmap_agents_collision", (X, Y))):
            cnt = int(np.prod(shape))
            out[name] = self.heatmaps[off:off + cnt].reshape(shape)
            off += cnt
        return out

    def heatmap_observation_add(self, heat, out):
        """default_environment.py:238-247 for the outputs of one step (extras=True): heat[n] += every stepped
        environment's observation of agent n — the local [C,V,V] crop, or decode_relative_state
        (observation_handler.py:150-177) of the relative view on the relative / merged masks.  Environments reset by
        the call are skipped (their observation comes from observation_handler.reset, :114).  Plain numpy loops."""
        local = self.c.area_mask == MASKS["local"]
        src = out["obs_rel"] if self.merged else out["obs"]
        for e in range(self.E):
            if out["was_reset"][e]:
                continue
            for n in range(self.N):
                if local:
                    heat[n] += src[e, n]
                    continue
                c = dict(zip(COORD_KEYS, (int(v) for v in out["coordinates"][e, n])))
                canvas = np.zeros((self.Cn, self.X, self.Y), np.float32)
                canvas[-1] -= 1
                canvas[:, c["map_x_min"]:c["map_x_max"], c["map_y_min"]:c["map_y_max"]] = \
                    src[e, n][:, c["obs_x_min"]:c["obs_x_max"], c["obs_y_min"]:c["obs_y_max"]]
                heat[n] += canvas
        return heat

    def reset(self, tick, env_mask=None, noise_tape=None):
        o, s = self._out(noise_tape)
        if env_mask is not None:
            env_mask = np.ascontiguousarray(env_mask, np.uint8)
        self.L.orc_reset(self.h, tick, _ptr(env_mask), C.byref(s))
        return o

    def step(self, tick, actions=None, order=None, noise_tape=None, threads=1):
        o, s = self._out(noise_tape)
        if actions is not None:
            actions = np.ascontiguousarray(actions, np.int8).reshape(self.E, self.N)
        if order is not None:
            order = np.ascontiguousarray(order, np.uint8).reshape(self.E, self.N)
        self.L.orc_step(self.h, tick, _ptr(actions), _ptr(order), C.byref(s), threads)
        return o

    def step_nocopy(self, tick, threads=1):
        """Random-policy step with no outputs allocated per call except the observation
        buffer given at construction time (cpu_baseline timing)."""
        if not hasattr(self, "_bench_out"):
            self._bench_out = self._out(None)
        o, s = self._bench_out
        self.L.orc_step(self.h, tick, None, None, C.byref(s), threads)
        return o

    def state(self):
        E, N, T, X, Y = self.E, self.N, self.T, self.X, self.Y
        pos = np.zeros((E, N, 2), np.int8)
        obj = np.zeros((E, T, X, Y), np.int8)
        cnt = np.zeros((E, 4), np.int32)
        steps = np.zeros((E,), np.int32)
        dest = np.zeros((E, N, X, Y), np.int8)
        self.L.orc_get_state(self.h, _ptr(pos), _ptr(obj), _ptr(cnt), _ptr(steps), _ptr(dest))
        return {"pos": pos, "objects": obj, "counters": cnt, "steps": steps, "destination": dest}

    def set_state(self, pos=None, objects=None):
        if pos is not None:
            pos = np.ascontiguousarray(pos, np.int8).reshape(self.E, self.N, 2)
        if objects is not None:
            objects = np.ascontiguousarray(objects, np.int8).reshape(self.E, self.T, self.X, self.Y)
        self.L.orc_set_state(self.h, _ptr(pos), _ptr(objects))
