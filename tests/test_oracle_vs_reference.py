"""Live cross-check of the C oracle against the reference (build container only)."""
import numpy as np
import pytest

from oracle import ref_runner as rr
from oracle.cases import cases
from oracle.oracle import Oracle

pytestmark = pytest.mark.skipif(not rr.available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("name", ["cfg5_four_rooms_n8", "cfg2_four_rooms_transition_flip_sharedmean"])
def test_other_env_id_and_seed(name):
    """Fixtures use env id 3 / seed 921; the keying must hold for other ids and seeds."""
    cfg, _, _, okw = cases()[name]
    cfg = cfg.merged_with(seed=12345)
    ref = rr.RefEnv(cfg, env_id=70001)
    orc = Oracle(cfg, num_envs=1, first_global_env_id=70001, extras=True, **okw)
    a, b = ref.reset(7), orc.reset(7)
    assert np.array_equal(a["obs"], b["obs"][0])
    for tick in range(8, 40):
        a, b = ref.step(tick), orc.step(tick)
        for k in ("obs", "dones", "objects", "agents", "coordinates", "destination"):
            assert np.array_equal(a[k], b[k][0]), (tick, k)
        assert np.array_equal(a["rewards"].astype(np.float32), b["rewards"][0])


def test_explicit_actions_and_order_with_stay():
    """Injected actions (including STAY=4, never sampled by the random policy) and orders."""
    cfg = cases()["cfg5_four_rooms_n8"][0]
    ref = rr.RefEnv(cfg, env_id=1)
    orc = Oracle(cfg, num_envs=1, first_global_env_id=1, extras=True)
    rng = np.random.default_rng(0)
    ref.reset(0), orc.reset(0)
    for tick in range(1, 60):
        actions = rng.integers(0, 5, size=8)
        order = rng.permutation(8)
        a, b = ref.step(tick, actions=actions, order=order), orc.step(tick, actions=actions[None], order=order[None])
        for k in ("obs", "dones", "objects", "agents"):
            assert np.array_equal(a[k], b[k][0]), (tick, k)
        assert np.array_equal(a["rewards"].astype(np.float32), b["rewards"][0])


def test_fixed_scenario_pickup_and_collision():
    """Hand-checkable scenario in the spirit of core/environments/test_environment.py:63-72:
    agent 0 walks onto an object (reward 1, respawn), agent 1 walks into agent 0's old... wall."""
    cfg = cases()["cfg1_simple_mlp"][0].merged_with(shuffle_init_xys=False)
    ref = rr.RefEnv(cfg, env_id=0)
    orc = Oracle(cfg, num_envs=1, first_global_env_id=0, extras=True)
    ref.reset(0), orc.reset(0)
    pos = np.array([[5, 5], [1, 1]], np.int8)
    objects = np.zeros((1, 20, 20), np.int8)
    objects[0, 6, 5] = 1
    ref.set_state(pos, objects), orc.set_state(pos[None], objects[None])
    a = ref.step(1, actions=[0, 2], order=[0, 1])  # agent0 RIGHT onto the object, agent1 LEFT into the border wall
    b = orc.step(1, actions=np.array([[0, 2]]), order=np.array([[0, 1]]))
    assert list(a["rewards"]) == [1.0, -1.0]
    assert np.array_equal(a["rewards"].astype(np.float32), b["rewards"][0])
    assert np.array_equal(a["obs"], b["obs"][0]) and np.array_equal(a["objects"], b["objects"][0])
    assert a["objects"][0].sum() == 1 and a["objects"][0][0, 6, 5] == 0  # consumed, respawned elsewhere


@pytest.mark.parametrize("name", ["cfg5_four_rooms_n8", "sequential_central_room_t3", "junctional_chain_tasks"])
def test_heatmaps_match_reference(name):
    """Visit statistics of default_environment.py:70-96: the six integer heat-maps after 80 steps."""
    cfg = cases()[name][0]
    ref = rr.RefEnv(cfg, env_id=5)
    orc = Oracle(cfg, num_envs=1, first_global_env_id=5)
    hm = orc.enable_heatmaps()
    ref.reset(0), orc.reset(0)
    for tick in range(1, 81):
        a, b = ref.step(tick), orc.step(tick)
        assert np.array_equal(a["obs"], b["obs"][0])
    env = ref.env
    for key in ("heatmap_agents", "heatmap_complete", "heatmap_objects", "heatmap_objects_left", "heatmap_wall_collision",
                "heatmap_agents_collision"):
        assert np.array_equal(np.asarray(getattr(env, key)), hm[key]), key
    assert hm["heatmap_complete"].sum() > 0 and hm["heatmap_objects_left"].sum() > 0


@pytest.mark.parametrize("name", ["cfg5_four_rooms_n8", "relative_transparent", "cfg4_four_rectangle_merged_types",
                                  "cfg2_four_rooms_transition_flip_sharedmean"])
def test_heatmap_observation_matches_reference(name):
    """heatmap_observation (default_environment.py:93,238-247) on the local, relative and merged views: the sum of the
    step observations of an episode (decoded to the map canvas on the relative / merged views), zeroed by reset()."""
    cfg = cases()[name][0]
    ref = rr.RefEnv(cfg, env_id=5)
    orc = Oracle(cfg, num_envs=1, first_global_env_id=5, extras=True)
    ref.reset(0), orc.reset(0)
    heat = np.zeros(np.asarray(ref.env.heatmap_observation).shape, np.float32)
    for tick in range(1, 41):
        ref.step(tick)
        orc.heatmap_observation_add(heat, orc.step(tick))
    assert np.array_equal(np.asarray(ref.env.heatmap_observation, np.float32), heat)
    assert np.abs(heat).sum() > 0


@pytest.mark.parametrize("mask", ["local", "relative"])
def test_observation_stats_environment_matches_reference(mask):
    """environment=observation_stats (observation_stats_environment.py:16-61): the default step, heatmap_observation
    zeroed by every reset() and observation_stats carried across episodes."""
    cfg = cases()["relative_transparent"][0].merged_with(environment="observation_stats", observation_area_mask=mask)
    ref = rr.RefEnv(cfg, env_id=2)
    orc = Oracle(cfg, num_envs=1, first_global_env_id=2, extras=True)
    stats, tick = None, 0
    for episode in range(2):
        a, b = ref.reset(tick), orc.reset(tick)
        assert np.array_equal(a["obs"], b["obs"][0])
        shape = np.asarray(ref.env.heatmap_observation).shape  # the attribute is created by reset() (:93)
        stats = np.zeros(shape, np.float32) if stats is None else stats
        heat = np.zeros(shape, np.float32)
        for _ in range(15):
            tick += 1
            a, b = ref.step(tick), orc.step(tick)
            assert np.array_equal(a["obs"], b["obs"][0])
            orc.heatmap_observation_add(heat, b)
            orc.heatmap_observation_add(stats, b)
        tick += 1
        assert np.array_equal(np.asarray(ref.env.heatmap_observation, np.float32), heat)
    assert np.array_equal(np.stack(ref.env.observation_stats).astype(np.float32), stats)
