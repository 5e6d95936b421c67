"""Static gridworld maps (host side): walls, object areas, destination areas.

Mirrors `generate_map` / `AbstractMap` of the reference (core/maps/__init__.py:25-61,
core/maps/abstract_map.py:12-81).  The maps are pure data here: every map is a small
declarative description (wall runs with gaps, area boxes) rasterised into the reference's
int8 matrices, which the CUDA library then packs into bit rows (maexp_env_create).
"""
import numpy as np

from ..utils.logging import initialize_logging

logger = initialize_logging(__name__)


class AbstractMap:
    """Attribute-compatible with core/maps/abstract_map.py:12-81."""

    #: (axis, line, cells) — wall cells on row/column `line`; axis "x": wall[cells, line], "y": wall[line, cells]
    WALL_RUNS = ()
    #: {type_objects: [(type, x0, x1, y0, y1), ...]} half-open boxes, negative = relative to SIZE
    AREAS = {}
    #: True: destination areas are redrawn at every reset from `destination_masks()`
    RANDOM_DESTINATIONS = False

    def __init__(self, config, size_x, size_y):
        self.config = config
        self.num_agents = config.num_agents
        self.type_objects = config.type_objects
        self.SIZE_X, self.SIZE_Y = int(size_x), int(size_y)
        X, Y = self.SIZE_X, self.SIZE_Y
        self.wall_matrix = np.zeros((X, Y), np.int8)
        self.agents_matrix = np.zeros((self.num_agents, X, Y), np.int8)
        self.objects_matrix = np.zeros((self.type_objects, X, Y), np.int8)
        self.objects_area_matrix = np.zeros((self.type_objects, X, Y), np.int8)
        self.destination_area_matrix = np.ones((self.num_agents, X, Y), np.int8)
        self.locate_walls()
        self.set_objects_area()
        self.set_noise_area()

    # ---- coordinate transforms, abstract_map.py:44-64
    def coord2ind(self, position, size_x=None, size_y=None):
        sx, sy = size_x or self.SIZE_X, size_y or self.SIZE_Y
        return np.array([sx // 2 + position[0], sy // 2 - position[1]], dtype=np.int8)

    def ind2coord(self, position, size_x=None, size_y=None):
        sx, sy = size_x or self.SIZE_X, size_y or self.SIZE_Y
        return np.array([position[0] - sx // 2, sy // 2 - position[1]], dtype=np.int8)

    # ---- rasterisation
    def _border(self):
        self.wall_matrix[[0, self.SIZE_X - 1], :] = 1
        self.wall_matrix[:, [0, self.SIZE_Y - 1]] = 1

    def locate_walls(self):
        self._border()
        for axis, line, cells in self.WALL_RUNS:
            cells = np.asarray(list(cells))
            if axis == "x":
                self.wall_matrix[cells, line] = 1
            else:
                self.wall_matrix[line, cells] = 1

    def set_objects_area(self):
        X, Y = self.SIZE_X, self.SIZE_Y
        for t, x0, x1, y0, y1 in self._areas().get(self.type_objects, ()):
            x1 = x1 if x1 > 0 else X + x1
            y1 = y1 if y1 > 0 else Y + y1
            self.objects_area_matrix[t, x0:x1, y0:y1] = 1

    def _areas(self):
        return self.AREAS

    def set_noise_area(self):
        self.noise_area_matrix = np.ones((self.SIZE_X, self.SIZE_Y), np.int8)

    def destination_masks(self):
        """Canonical destination masks [M, X, Y]; default: one all-ones mask (abstract_map.py:80-81)."""
        return np.ones((1, self.SIZE_X, self.SIZE_Y), np.int8)

    # ---- mutable-grid API kept for interface compatibility (state lives on the GPU)
    def reset(self):
        self.reset_agents()
        self.reset_objects()
        self.reset_destination_area()

    def reset_agents(self):
        self.agents_matrix = np.zeros_like(self.agents_matrix)

    def reset_objects(self):
        self.objects_matrix = np.zeros_like(self.objects_matrix)

    def reset_destination_area(self):
        self.destination_area_matrix = np.ones_like(self.destination_area_matrix)

    def step(self):
        pass


class SimpleMap(AbstractMap):
    """core/maps/simple_map.py:10-21"""

    def _areas(self):
        h = self.SIZE_X // 2
        return {1: [(0, 1, -1, 1, -1)], 2: [(0, 1, h, 1, -1), (1, h, -1, 1, -1)]}


class DynamicSimpleMap(SimpleMap):
    """core/maps/dynamic_simple_map.py:12-63.  The simple map with two one-cell obstacles that `world.step` moves
    one cell per step (abstract_world.py:40) and that `reset()` leaves alone.  On the device the two cells are a
    closed-form function of a per-environment world-step counter (`dynamic_obstacles` of maexp_env_cfg);
    `wall_matrix` here is the initial layout, from which the line-of-sight table is generated once, as in the
    reference.  `step()` restates the shuttle rule on the host matrices for interface compatibility."""

    def __init__(self, config, size_x, size_y):
        self.obstacles_pos = [[int(size_x) // 2, 1], [int(size_x) // 2, int(size_y) // 2]]
        self.obstacles_homing_flag = [False, False]
        super().__init__(config, size_x, size_y)

    def locate_walls(self):
        self._border()
        for x, y in self.obstacles_pos:
            self.wall_matrix[x, y] = 1

    def step(self):
        X, Y = self.SIZE_X, self.SIZE_Y
        far = ((X // 2, Y - 2), (X - 2, Y // 2))    # turning points: obstacle 0 shuttles along y, obstacle 1 along x
        home = ((X // 2, 1), (1, Y // 2))
        for x, y in self.obstacles_pos:
            self.wall_matrix[x, y] = 0
        for k, pos in enumerate(self.obstacles_pos):
            if tuple(pos) == far[k]:
                self.obstacles_homing_flag[k] = True
            elif tuple(pos) == home[k]:
                self.obstacles_homing_flag[k] = False
        for k, axis in enumerate((1, 0)):
            self.obstacles_pos[k][axis] += -1 if self.obstacles_homing_flag[k] else 1
        for x, y in self.obstacles_pos:
            self.wall_matrix[x, y] = 1


_FOUR_ROOMS_RUN = (1, 2, 3, 7, 8, 9, 10, 13, 14, 15, 16, 20, 21, 22)


class FourRoomsMap(AbstractMap):
    """core/maps/four_rooms_map.py:11-42"""
    WALL_RUNS = (("x", 10, _FOUR_ROOMS_RUN), ("x", 13, _FOUR_ROOMS_RUN),
                 ("y", 10, _FOUR_ROOMS_RUN), ("y", 13, _FOUR_ROOMS_RUN))
    AREAS = {1: [(0, 1, 10, 1, 10), (0, 1, 10, 14, -1), (0, 14, -1, 1, 10), (0, 14, -1, 14, -1)],
             2: [(0, 1, 10, 1, 10), (0, 1, 10, 14, -1), (1, 14, -1, 1, 10), (1, 14, -1, 14, -1)]}


class ThreeRoomsMap(AbstractMap):
    """core/maps/three_rooms_map.py:13-55"""
    WALL_RUNS = (("x", 10, (1, 2, 3, 7, 8, 9, 10)), ("x", 14, (1, 2, 3, 7, 8, 9, 10)),
                 ("y", 10, (1, 2, 3, 7, 8, 9, 10, 14, 15, 16, 17, 21, 22, 23)),
                 ("y", 14, (1, 2, 3, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 21, 22, 23)))
    AREAS = {1: [(0, 1, 10, 1, 10), (0, 1, 10, 15, 24), (0, 15, 24, 1, 24)],
             2: [(0, 1, 10, 1, 10), (0, 15, 24, 1, 24), (1, 1, 10, 15, 24), (1, 15, 24, 1, 24)]}
    RANDOM_DESTINATIONS = True

    def destination_masks(self):
        # three_rooms_map.py:43-51: two copies of each half-plane plus the full map, in this order
        X, Y, h = self.SIZE_X, self.SIZE_Y, self.SIZE_X // 2
        m = np.ones((9, X, Y), np.int8)
        m[0:2, :h, :] = 0
        m[2:4, h:, :] = 0
        m[4:6, :, :h] = 0
        m[6:8, :, h:] = 0
        return m


class FourRectangleMap(AbstractMap):
    """core/maps/four_rectangle.py:11-39"""
    AREAS = {1: [(0, 1, 19, 1, 10), (0, 1, 19, 14, -1), (0, 21, -1, 1, 10), (0, 21, -1, 14, -1)],
             2: [(0, 1, 19, 1, 10), (0, 1, 19, 14, -1), (1, 21, -1, 1, 10), (1, 21, -1, 14, -1)]}

    def locate_walls(self):
        self._border()
        w = self.wall_matrix
        w[[19, 20], :] = 1
        w[19:21, 11:13] = 0
        w[:, [10, 13]] = 1
        for lo, hi in ((8, 12), (28, 32)):
            w[lo:hi, 10] = 0
            w[lo:hi, 13] = 0


def _room(w, x0, x1, y0, y1):
    """Square room outline x0..x1 / y0..y1 (inclusive) with a 3-cell door in the middle of every side."""
    w[x0:x1 + 1, [y0, y1]] = 1
    w[[x0, x1], y0:y1 + 1] = 1
    mx, my = (x0 + x1) // 2, (y0 + y1) // 2
    w[mx - 1:mx + 2, [y0, y1]] = 0
    w[[x0, x1], my - 1:my + 2] = 0


def _half_planes(X, Y, inside):
    """[x < X//2, x >= X//2, y < Y//2, y >= Y//2] as masks holding `inside` there and 1 - inside elsewhere."""
    m = np.full((4, X, Y), 1 - inside, np.int8)
    m[0, :X // 2, :] = inside
    m[1, X // 2:, :] = inside
    m[2, :, :Y // 2] = inside
    m[3, :, Y // 2:] = inside
    return m


class CentralRoomMap(AbstractMap):
    """core/maps/central_room_map.py:10-52: one 9x9 room in the middle; only type 0 has an objects area."""

    def locate_walls(self):
        self._border()
        _room(self.wall_matrix, 8, 16, 8, 16)

    def set_objects_area(self):
        if 1 <= self.type_objects <= 5:
            self.objects_area_matrix[0, 1:24, 1:24] = 1
            self.objects_area_matrix[0, 8:17, 8:17] = 0


class CentralRoomDestinationMap(CentralRoomMap):
    """core/maps/central_room_destination_map.py:12-22: four half-planes blanked out plus the full map."""
    RANDOM_DESTINATIONS = True

    def destination_masks(self):
        return np.concatenate([_half_planes(self.SIZE_X, self.SIZE_Y, 0), np.ones((1, self.SIZE_X, self.SIZE_Y), np.int8)])


class CentralRoomLargeMap(AbstractMap):
    """core/maps/central_room_large_map.py:10-23"""

    def locate_walls(self):
        self._border()
        _room(self.wall_matrix, 20, 28, 8, 16)

    def set_objects_area(self):
        self.objects_area_matrix[0, 1:48, 1:24] = 1
        self.objects_area_matrix[0, 20:29, 8:17] = 0


class CentralRoomLargeDestinationMap(CentralRoomLargeMap):
    """core/maps/central_room_large_destination_map.py:12-22: the four half-planes themselves."""
    RANDOM_DESTINATIONS = True

    def destination_masks(self):
        return _half_planes(self.SIZE_X, self.SIZE_Y, 1)


class TwoCentralRoomsMap(AbstractMap):
    """core/maps/two_central_rooms_map.py:10-69: two rooms, fixed destination areas per agent pair."""

    def locate_walls(self):
        self._border()
        _room(self.wall_matrix, 8, 16, 8, 16)
        _room(self.wall_matrix, 33, 41, 8, 16)

    def set_objects_area(self):
        a = self.objects_area_matrix
        if self.type_objects == 1:
            a[0, 1:24, 1:24] = 1
            a[0, 8:17, 8:17] = 0
        elif self.type_objects == 2:
            a[0, 1:26, 1:13] = 1
            a[0, 8:17, 8:13] = 0
            a[0, 24:49, 12:24] = 1
            a[0, 33:42, 12:17] = 0
            a[1, 1:26, 12:24] = 1
            a[1, 8:17, 12:17] = 0
            a[1, 24:49, 1:13] = 1
            a[1, 33:42, 8:13] = 0

    def destination_masks(self):
        """One mask per agent (the step kernel maps agent k to mask min(k, M-1)); agents past the twelfth get zeros."""
        X, Y = self.SIZE_X, self.SIZE_Y
        pairs = np.zeros((6, X, Y), np.int8)
        pairs[0, 1:26, 1:13] = 1
        pairs[0, 24:49, 12:24] = 1
        pairs[1, 1:26, 12:24] = 1
        pairs[1, 24:49, 1:13] = 1
        pairs[2, 1:49, 1:13] = 1
        pairs[3, 1:49, 12:24] = 1
        pairs[4, 1:26, 1:24] = 1
        pairs[5, 24:49, 1:24] = 1
        n = min(self.num_agents, 12)
        masks = [pairs[k // 2] for k in range(n)]
        if self.num_agents > 12:
            masks.append(np.zeros((X, Y), np.int8))
        return np.stack(masks)

    def reset_destination_area(self):
        m = self.destination_masks()
        self.destination_area_matrix = np.stack([m[min(k, len(m) - 1)] for k in range(self.num_agents)])


_REGISTRY = {
    # name -> (class, fixed size or None for config.map.SIZE_X/SIZE_Y)  core/maps/__init__.py:26-39
    "simple": (SimpleMap, None),
    "dynamic_simple": (DynamicSimpleMap, None),
    "four_rooms": (FourRoomsMap, (24, 24)),
    "three_rooms": (ThreeRoomsMap, (25, 25)),
    "four_rectangle": (FourRectangleMap, (40, 24)),
    "central_room": (CentralRoomMap, (25, 25)),
    "two_central_rooms": (TwoCentralRoomsMap, (50, 25)),
    "central_room_destination": (CentralRoomDestinationMap, (25, 25)),
    "central_room_large": (CentralRoomLargeMap, (49, 25)),
    "central_room_large_destination": (CentralRoomLargeDestinationMap, (49, 25)),
}


def generate_map(config):
    name = config.map.name
    if name not in _REGISTRY:
        logger.warning(f"Unexpected map is given. config.map.name: {name}")
        raise ValueError()
    cls, size = _REGISTRY[name]
    size_x, size_y = size if size else (config.map.SIZE_X, config.map.SIZE_Y)
    return cls(config=config, size_x=size_x, size_y=size_y)
