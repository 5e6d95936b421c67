"""Line-of-sight visibility table and window/map coordinate helper (host side, init time).

Provides what `generate_observation_area_mask` / `generate_observation_mask_coordinate`
return in the reference (core/handlers/observations/masks/__init__.py:26-69,
masks/local_area_mask.py:18-97,100-332): an int8 table ``mask[X, Y, V, V]`` with 1 = visible,
0 = hidden for an agent standing on map cell (x, y), used for both `local` and `relative`.

The reference fills each window with four order-dependent sweeps whose x- and y-variants are
mirror images of each other.  Here a sweep is written once against a `_Probe` that can look at
the map transposed, and run twice (plain, then transposed) — same visiting order, same
results (checked bit-for-bit against the reference's tables for every map and V in
tests/test_host_tables.py).
"""
import numpy as np

from ...utils.logging import initialize_logging

logger = initialize_logging(__name__)

HIDDEN, VISIBLE = -1, 0
SIGNS = (-1, 1)


class _Probe:
    """Window/map access in centred offsets (u along the sweep axis, v across it)."""

    def __init__(self, wall, x, y, visible_range, transposed):
        self.wall, self.x, self.y = wall, x, y
        self.r = visible_range // 2
        self.t = transposed
        self.window = None

    def _xy(self, u, v):
        cx, cy = (v, u) if self.t else (u, v)
        return self.x + cx, self.y - cy, self.r + cx, self.r - cy  # map x, map y, window a, window b

    def outside(self, u, v):
        mx, my, _, _ = self._xy(u, v)
        return mx < 0 or mx >= self.wall.shape[0] or my < 0 or my >= self.wall.shape[1]

    def is_wall(self, u, v):
        mx, my, _, _ = self._xy(u, v)
        return self.wall[mx, my] == 1  # negative indices wrap exactly like the reference's numpy lookups

    def mark(self, u, v, value):
        _, _, a, b = self._xy(u, v)
        self.window[a, b] = value


def _fan_sweep(p):
    """Diagonal fans on both sides of the sweep axis (local_area_mask.py:107-146 / :150-191)."""
    for du in (-1, 0, 1):
        for dv in SIGNS:
            for drift in SIGNS:
                for j in range(p.r):
                    u, v = du + j * drift, dv * (1 + j)
                    if p.outside(u, v):
                        p.mark(u, v, HIDDEN)
                        break
                    if j == 0:
                        if du == 0 and not p.is_wall(u + drift, v):
                            p.mark(u, v, VISIBLE)  # straight cell next to an open corner
                            continue
                        if du != 0 and drift != du:
                            if p.is_wall(u - du, v):
                                break  # corner cell shadowed by its inner neighbour
                            continue
                    if p.is_wall(u, v):
                        p.mark(u, v, HIDDEN)
                        break
                    p.mark(u, v, VISIBLE)


def _ray(p, cells, stop_outside):
    for u, v in cells:
        if p.outside(u, v):
            p.mark(u, v, HIDDEN)
            if stop_outside:
                break
            continue
        if p.is_wall(u, v):
            p.mark(u, v, HIDDEN)
            break
        p.mark(u, v, VISIBLE)


def _corridor_sweep(p):
    """Walk along the agent's own line and the two adjacent ones (local_area_mask.py:193-261 / :263-330)."""
    for su in SIGNS:
        for d in range(p.r + 1):
            if p.is_wall(d * su, 0):
                p.mark(d * su, 0, HIDDEN)
                break
            for sv in SIGNS:
                _ray(p, [(d * su, dv * sv) for dv in range(2)], stop_outside=False)
        for sv in SIGNS:
            _ray(p, [(d * su, sv) for d in range(p.r + 1)], stop_outside=False)


def generate_each_mask(x, y, wall, visible_range):
    """One V x V window for an agent on map cell (x, y): 0 visible, -1 hidden."""
    r = visible_range // 2
    window = np.full((visible_range, visible_range), HIDDEN, np.int8)
    window[r, r] = VISIBLE
    for sweep in (_fan_sweep, _corridor_sweep):
        for transposed in (False, True):
            probe = _Probe(wall, x, y, visible_range, transposed)
            probe.window = window
            sweep(probe)
    return window


def window_bounds(px, py, size_x, size_y, visible_range):
    """local_area_mask.py:18-47 for an index-space position: the 12 `coordinates` values."""
    r = visible_range // 2
    obs_x_min, obs_x_max = abs(min(0, px - r)), visible_range - max(0, px + r - (size_x - 1))
    obs_y_min, obs_y_max = abs(min(0, py - r)), visible_range - max(0, py + r - (size_y - 1))
    return {
        "obs_x_min": obs_x_min, "obs_x_max": obs_x_max, "obs_y_min": obs_y_min, "obs_y_max": obs_y_max,
        "area_x_min": obs_x_min, "area_x_max": obs_x_max, "area_y_min": obs_y_min, "area_y_max": obs_y_max,
        "map_x_min": max(0, px - r), "map_x_max": min(size_x - 1, px + r) + 1,
        "map_y_min": max(0, py - r), "map_y_max": min(size_y - 1, py + r) + 1,
    }


COORDINATE_KEYS = ("obs_x_min", "obs_x_max", "obs_y_min", "obs_y_max", "area_x_min", "area_x_max",
                   "area_y_min", "area_y_max", "map_x_min", "map_x_max", "map_y_min", "map_y_max")

_CACHE = {}


def generate_local_area_mask(config, world):
    """int8 [X, Y, V, V], 1 visible / 0 hidden (local_area_mask.py:50-97).  Border cells stay 0."""
    wall = world.map.wall_matrix
    X, Y = wall.shape
    V = int(config.visible_range)
    key = (wall.tobytes(), X, Y, V, bool(config.transparent_observation))
    if key in _CACHE:
        return _CACHE[key]
    table = np.full((X, Y, V, V), HIDDEN, np.int8)
    for x in range(1, X - 1):
        for y in range(1, Y - 1):
            if config.transparent_observation:
                b = window_bounds(x, y, X, Y, V)
                table[x, y, b["obs_x_min"]:b["obs_x_max"], b["obs_y_min"]:b["obs_y_max"]] = \
                    -wall[b["map_x_min"]:b["map_x_max"], b["map_y_min"]:b["map_y_max"]]
            else:
                table[x, y] = generate_each_mask(x, y, wall, V)
    table = table + 1
    _CACHE[key] = table
    return table


def generate_observation_area_mask(config, world):
    """masks/__init__.py:26-43: `local` and `relative` share the local table."""
    if config.observation_area_mask in ("local", "relative"):
        return generate_local_area_mask(config=config, world=world)
    logger.warning(f"Unexpected area mask generator is given. config.observation_area_mask: "
                   f"{config.observation_area_mask}")
    raise ValueError()
