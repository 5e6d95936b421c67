"""Drives the LIVE reference (read-only under /root/reference) with injected draws.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: the GPU box has no
/root/reference.  It is used (a) by oracle/make_golden.py to write tests/golden/*.npz and
(b) by tests/test_oracle_vs_reference.py (skipped when the reference is absent) to pin the
C oracle against the reference itself.

How the reference is made replayable: the module-level
``random`` of core/environments/default_environment.py and core/maps/three_rooms_map.py is
replaced by a tape object whose numbers come from the build's Philox spec
(oracle/philox.h); the ``torch`` name inside the noise modules is replaced by a proxy whose
``rand`` / ``empty().normal_`` return slices of a per-agent draw field.  Actions and the
move order are passed to ``env.step`` explicitly.  The reference's own code is executed
unmodified.
"""
import contextlib
import io
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("MAEXP_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))

STREAM_ORDER, STREAM_ACTION, STREAM_RESPAWN, STREAM_INIT_PERM, STREAM_DEST, STREAM_OBJECTS, STREAM_NOISE = range(7)
CONSTRUCT_TICK = 0xFFFFFFFF
COORD_KEYS = ("obs_x_min", "obs_x_max", "obs_y_min", "obs_y_max", "area_x_min", "area_x_max", "area_y_min",
              "area_y_max", "map_x_min", "map_x_max", "map_y_min", "map_y_max")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


def _import_reference():
    shim = os.path.join(_HERE, "_shims")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)


# ---------------------------------------------------------------- Philox (pure python)

def philox4x32_10(k0, k1, c0, c1, c2, c3):
    """Independent pure-python statement of the same published algorithm as oracle/philox.h."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def draw(seed, env, tick, stream, d, sub=0):
    return philox4x32_10(seed, env, tick, stream, d >> 2, sub)[d & 3]


def fisher_yates(seq, seed, env, tick, stream, d0=0):
    """In-place shuffle used by every tape; returns the next draw index."""
    d = d0
    for k in range(len(seq) - 1, 0, -1):
        j = (draw(seed, env, tick, stream, d) * (k + 1)) >> 32
        d += 1
        seq[k], seq[j] = seq[j], seq[k]
    return d


def philox_order(n, seed, env, tick):
    order = list(range(n))
    fisher_yates(order, seed, env, tick, STREAM_ORDER)
    return order


def philox_actions(n, seed, env, tick):
    return [draw(seed, env, tick, STREAM_ACTION, i) >> 30 for i in range(n)]


def uniform_field(n_elems, seed, env, tick, agent):
    """Uniform draws of the flip noise: 16 bits each, eight per Philox call -- element k: call k >> 3, word (k & 7) >> 1, the low
    half for even k; u = h / 65536 (DESIGN.md RNG specification; oracle/maexp_oracle.c noise_draw)."""
    out = np.empty(n_elems, np.float32)
    for c in range((n_elems + 7) // 8):
        w = philox4x32_10(seed, env, tick, STREAM_NOISE, c, agent)
        for j in range(8):
            if 8 * c + j < n_elems:
                h = (w[j >> 1] >> 16) if (j & 1) else (w[j >> 1] & 0xFFFF)
                out[8 * c + j] = np.float32(h) * np.float32(1.0 / 65536.0)
    return out


class EnvTape:
    """Replaces ``random`` in core/environments/default_environment.py."""

    def __init__(self, seed, env):
        self.seed, self.env = seed, env
        self.begin(CONSTRUCT_TICK, STREAM_OBJECTS)

    def begin(self, tick, stream):
        self.tick, self.stream, self.calls = tick, stream, 0

    def random(self):  # default_environment.py:137-138 — two calls per attempt
        attempt, second = self.calls >> 1, self.calls & 1
        self.calls += 1
        w = philox4x32_10(self.seed, self.env, self.tick, self.stream, attempt >> 1, 0)
        return w[(attempt & 1) * 2 + second] / 4294967296.0

    def shuffle(self, x):  # default_environment.py:98
        fisher_yates(x, self.seed, self.env, self.tick, STREAM_INIT_PERM)


class MapTape:
    """Replaces ``random`` in core/maps/three_rooms_map.py."""

    def __init__(self, seed, env):
        self.seed, self.env, self.tick, self.d = seed, env, CONSTRUCT_TICK, 0

    def begin(self, tick):
        self.tick, self.d = tick, 0

    def shuffle(self, x):  # three_rooms_map.py:52
        self.d = fisher_yates(x, self.seed, self.env, self.tick, STREAM_DEST, self.d)

    def choice(self, x):  # three_rooms_map.py:55
        i = (draw(self.seed, self.env, self.tick, STREAM_DEST, self.d) * len(x)) >> 32
        self.d += 1
        return x[i]


class _Pending:
    def __init__(self, proxy, shape):
        self.proxy, self.shape = proxy, shape

    def normal_(self, mean=0, std=1):
        import torch
        z = self.proxy.field_slice(self.shape)
        return torch.from_numpy(z * np.float32(std) + np.float32(mean)) if mean else torch.from_numpy(z * np.float32(std))


class TorchProxy:
    """Stands in for the name ``torch`` inside one reference noise module."""

    def __init__(self, runner):
        import torch
        self._torch, self._runner = torch, runner

    def __getattr__(self, name):
        return getattr(self._torch, name)

    def field_slice(self, shape):
        field = self._runner.current_field()  # [C,V,V]
        c, n, _ = shape
        i = (field.shape[1] - n) // 2
        return np.ascontiguousarray(field[:c, i:i + n, i:i + n])

    def rand(self, shape):
        return self._torch.from_numpy(self.field_slice(tuple(shape)))

    def empty(self, shape):
        return _Pending(self, tuple(shape))


def to_dictconfig(cfg):
    _import_reference()
    from omegaconf import DictConfig
    return DictConfig({k: (dict(v) if isinstance(v, dict) else v) for k, v in dict(cfg).items()})


class RefEnv:
    """One live reference environment with global env id ``env_id``."""

    def __init__(self, cfg, env_id=0):
        _import_reference()
        with contextlib.redirect_stderr(io.StringIO()):  # tqdm bar of the mask generator
            import core.environments.default_environment as de
            import core.maps.central_room_destination_map as crd
            import core.maps.central_room_large_destination_map as crld
            import core.maps.three_rooms_map as trm
            from core.environments import generate_environment
            from core.environments.default_environment import DefaultEnvironment
            from core.handlers.observations.noises import (dist_observation_noise, flat_observation_noise,
                                                           flip_observation_noise)
            from core.maps import generate_map
            from core.worlds import generate_world

            self.cfg = to_dictconfig(cfg)
            self.seed = int(self.cfg.seed) & 0xFFFFFFFF
            self.env_id = env_id
            self.env_tape, self.map_tape = EnvTape(self.seed, env_id), MapTape(self.seed, env_id)
            self._mods = (de, trm, crd, crld)
            self._noise_mods = (flip_observation_noise, dist_observation_noise, flat_observation_noise)
            self._proxy = TorchProxy(self)
            self._field = None
            self._noise_tape = None
            self._tick = CONSTRUCT_TICK
            with self._patched():
                world_map = generate_map(self.cfg)
                world = generate_world(self.cfg, world_map)
                name = self.cfg.environment
                if name in ("shared_mean_reward", "shared_max_reward"):
                    # the shipped classes lack `order` in step(); same arithmetic, fixed signature
                    class _Shared(DefaultEnvironment):
                        def step(self, action_n, order):
                            reward_n, done_n, obs_n = super().step(action_n, order)
                            shared = [np.asarray(reward_n).mean() for _ in range(self.num_agents)]
                            return shared, done_n, obs_n
                    self.env = _Shared(config=self.cfg, world=world)
                else:
                    self.env = generate_environment(self.cfg, world)
        self.N = int(self.cfg.num_agents)
        handler = self.env.observation_handler
        self.merged = self.cfg.observation_area_mask == "merged"
        noisy = handler.local_view_observation_handler if self.merged else handler
        self._wrap_noise(noisy.observation_noise)
        self.C, self.V = handler.get_channel, int(self.cfg.visible_range)

    @contextlib.contextmanager
    def _patched(self):
        de, map_mods = self._mods[0], self._mods[1:]
        saved = (de.random, [m.random for m in map_mods], [m.torch for m in self._noise_mods])
        de.random = self.env_tape
        for m in map_mods:  # three_rooms_map.py:52-55 and the central_room*_destination maps draw from `random`
            m.random = self.map_tape
        for m in self._noise_mods:
            m.torch = self._proxy
        try:
            yield
        finally:
            de.random = saved[0]
            for m, r in zip(map_mods, saved[1]):
                m.random = r
            for m, t in zip(self._noise_mods, saved[2]):
                m.torch = t

    def _wrap_noise(self, noise):
        orig = noise.add_noise

        def add_noise(obs, agent, agent_id):
            self._agent = agent_id
            return orig(obs, agent, agent_id)
        noise.add_noise = add_noise

    def current_field(self):
        K = self.C * self.V * self.V
        if self._noise_tape is not None:
            return np.asarray(self._noise_tape[self._agent], np.float32).reshape(self.C, self.V, self.V)
        return uniform_field(K, self.seed, self.env_id, self._tick, self._agent).reshape(self.C, self.V, self.V)

    def tables(self):
        m = self.env.world.map
        h = self.env.observation_handler
        h = h.local_view_observation_handler if self.merged else h
        return (np.array(m.wall_matrix, np.int8), np.array(m.objects_area_matrix, np.int8),
                np.array(h.observation_area_mask, np.int8))

    def _pack(self, rewards, dones, obs_n, was_reset):
        N = self.N
        view = "local" if self.cfg.observation_area_mask in ("local", "merged") else "relative"
        out = {
            "obs": np.stack([o[view].detach().numpy().astype(np.float32) for o in obs_n]),
            "obs_rel": np.stack([o["relative"].detach().numpy().astype(np.float32) for o in obs_n]) if self.merged else None,
            "rewards": np.asarray(rewards, np.float64),
            "dones": np.asarray(dones, np.uint8),
            "was_reset": np.uint8(was_reset),
            "coordinates": np.array([[int(o["coordinates"][k]) for k in COORD_KEYS] for o in obs_n], np.int32),
            "agents": np.stack([o["agents"].numpy().astype(np.int8) for o in obs_n]),
            "objects": np.stack([o["objects"].numpy().astype(np.int8) for o in obs_n]),
            "destination": np.stack([o["destination"].numpy().astype(np.int8) for o in obs_n]),
        }
        assert out["agents"].shape == (N, N, 2)
        return out

    def reset(self, tick, noise_tape=None):
        self._tick, self._noise_tape = tick, noise_tape
        self.env_tape.begin(tick, STREAM_OBJECTS)
        self.map_tape.begin(tick)
        with self._patched():
            obs_n = self.env.reset()
        return self._pack([0.0] * self.N, [0] * self.N, obs_n, 1)

    def step(self, tick, actions=None, order=None, noise_tape=None):
        self._tick, self._noise_tape = tick, noise_tape
        if actions is None:
            actions = philox_actions(self.N, self.seed, self.env_id, tick)
        if order is None:
            order = philox_order(self.N, self.seed, self.env_id, tick)
        self.env_tape.begin(tick, STREAM_RESPAWN)
        with self._patched():
            rewards, dones, obs_n = self.env.step([int(a) for a in actions], [int(o) for o in order])
        return self._pack(rewards, dones, obs_n, 0)

    def state(self):
        env = self.env
        m = env.world.map
        pos = np.array([m.coord2ind(a.xy) for a in env.agents], np.int8)
        return {"pos": pos, "objects": np.array(m.objects_matrix, np.int8),
                "counters": np.array([env.objects_generated, env.objects_completed, env.agents_collided,
                                      env.walls_collided], np.int32),
                "steps": np.int32(env.current_step), "destination": np.array(m.destination_area_matrix, np.int8)}

    def set_state(self, pos=None, objects=None):
        """Hand-built scenario (index-space positions), mirrors Oracle.set_state."""
        env = self.env
        m = env.world.map
        if pos is not None:
            m.reset_agents()
            for i, (ix, iy) in enumerate(np.asarray(pos).reshape(self.N, 2)):
                env.agents[i].move(m.ind2coord((int(ix), int(iy))))
                m.agents_matrix[i, int(ix), int(iy)] = 1
        if objects is not None:
            m.objects_matrix[...] = np.asarray(objects, np.int8).reshape(m.objects_matrix.shape)
