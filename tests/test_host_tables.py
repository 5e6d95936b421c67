"""Host-side (product) maps and line-of-sight tables vs the oracle and the reference's hashes."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

from maexp_b200.config import make_config
from maexp_b200.core.handlers.observations.masks import generate_observation_area_mask, window_bounds
from maexp_b200.core.maps import generate_map
from maexp_b200.core.worlds import generate_world
from oracle.oracle import Oracle
from tests.test_oracle_golden import SURVEY_HASHES, TRANSPARENT, _sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("key", list(SURVEY_HASHES))
@pytest.mark.parametrize("V", [5, 7, 9])
@pytest.mark.parametrize("T", [1, 2])
def test_product_tables_match_oracle_and_reference_hashes(key, V, T):
    name, X, Y = key
    for transparent in (False, True):
        cfg = make_config(map=dict(name=name, SIZE_X=X, SIZE_Y=Y), visible_range=V, type_objects=T,
                          transparent_observation=transparent)
        world = generate_world(cfg, generate_map(cfg))
        table = generate_observation_area_mask(cfg, world)
        wall, area, los = Oracle(cfg, 1).tables()
        assert np.array_equal(world.map.wall_matrix, wall)
        assert np.array_equal(world.map.objects_area_matrix, area)
        assert np.array_equal(table, los)
        assert _sha(world.map.wall_matrix) == SURVEY_HASHES[key][0]
        if not transparent:
            assert (int(table.sum()), _sha(table)) == SURVEY_HASHES[key][1][V]
        elif name in TRANSPARENT:
            assert (int(table.sum()), _sha(table)) == TRANSPARENT[name][V]


def test_unknown_registry_keys_raise_value_error():
    with pytest.raises(ValueError):
        generate_map(make_config(map=dict(name="nope", SIZE_X=5, SIZE_Y=5)))
    cfg = make_config()
    with pytest.raises(ValueError):
        generate_world(cfg.merged_with(world="nope"), generate_map(cfg))


def test_window_bounds_edges():
    b = window_bounds(1, 22, 24, 24, 7)
    assert (b["obs_x_min"], b["obs_x_max"], b["map_x_min"], b["map_x_max"]) == (2, 7, 0, 5)
    assert (b["obs_y_min"], b["obs_y_max"], b["map_y_min"], b["map_y_max"]) == (0, 5, 19, 24)


def test_c_abi_library_exports_every_declared_symbol():
    """The .so must load without a GPU and export each function include/maexp_b200.h declares."""
    from maexp_b200 import _lib
    header = open(os.path.join(ROOT, "include", "maexp_b200.h")).read()
    declared = set(re.findall(r"\b(maexp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert _lib.load().maexp_version() >= 1


def test_product_path_has_no_oracle_import():
    """The oracle is test infrastructure: nothing under maexp_b200/ may import it."""
    for base, _, files in os.walk(os.path.join(ROOT, "maexp_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(base, f)


@pytest.mark.parametrize("name,X,Y", [("central_room", 25, 25), ("central_room_destination", 25, 25),
                                      ("central_room_large", 49, 25), ("central_room_large_destination", 49, 25),
                                      ("two_central_rooms", 50, 25)])
@pytest.mark.parametrize("V", [5, 7, 9])
def test_central_room_maps_match_oracle(name, X, Y, V):
    """SURVEY.md 8(f)-1 maps: walls, object areas and line-of-sight tables equal the C oracle's (which the
    fixtures pin against the live reference)."""
    for T in (1, 2, 5):
        cfg = make_config(map=dict(name=name, SIZE_X=X, SIZE_Y=Y), visible_range=V, type_objects=T,
                          flip_noise_probabilities=[0.2, 0.1, 0.05, 0.02])
        world = generate_world(cfg, generate_map(cfg))
        assert (world.map.SIZE_X, world.map.SIZE_Y) == (X, Y)
        wall, area, los = Oracle(cfg, 1).tables()
        assert np.array_equal(world.map.wall_matrix, wall)
        assert np.array_equal(world.map.objects_area_matrix, area)
        assert np.array_equal(generate_observation_area_mask(cfg, world), los)


def _triangle_cells(k, X, Y):
    """The step kernel's closed form of the two dynamic_simple obstacle cells after k world steps
    (env_kernel.cuh `place_obstacles`)."""
    L0, L1 = Y - 3, X - 3
    m0, m1 = k % (2 * L0), (k + X // 2 - 1) % (2 * L1)
    return [[X // 2, 1 + (m0 if m0 <= L0 else 2 * L0 - m0)], [1 + (m1 if m1 <= L1 else 2 * L1 - m1), Y // 2]]


@pytest.mark.parametrize("X,Y", [(20, 20), (10, 10), (9, 14), (5, 5)])
def test_dynamic_simple_obstacles_closed_form(X, Y):
    """dynamic_simple_map.py:29-63: the host restatement of the shuttle rule, the kernel's closed form and (in the
    build container) the reference's own map agree for several periods; the fixtures' wall table is the initial one."""
    cfg = make_config(map=dict(name="dynamic_simple", SIZE_X=X, SIZE_Y=Y), num_agents=1, init_xys=[[1, 1]])
    m = generate_map(cfg)
    ref = None
    from oracle import ref_runner as rr
    if rr.available():  # build container only: the reference's own class next to ours
        import importlib
        rr._import_reference()
        ref = importlib.import_module("core.maps.dynamic_simple_map").DynamicSimpleMap(cfg, X, Y)
    if ref is not None:
        assert np.array_equal(ref.wall_matrix, m.wall_matrix)
    for k in range(1, 5 * (X + Y)):
        m.step()
        assert m.obstacles_pos == _triangle_cells(k, X, Y), k
        want = np.zeros((X, Y), np.int8)
        want[[0, X - 1], :] = 1
        want[:, [0, Y - 1]] = 1
        for x, y in m.obstacles_pos:
            want[x, y] = 1
        assert np.array_equal(m.wall_matrix, want)
        if ref is not None:
            ref.step()
            assert np.array_equal(ref.wall_matrix, m.wall_matrix), k
