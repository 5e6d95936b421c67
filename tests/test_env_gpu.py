"""Parity of the CUDA step + observation path (through the C-ABI) — bit-exact.

1. against tests/golden/*.npz, i.e. the live reference's own outputs (E=1, injected draws);
2. against the C oracle on batches of environments with on-chip Philox draws and auto-reset;
3. at BASELINE.json's full size (65 536 envs x 8 agents) through size-independent invariants
   plus an oracle check of a strided sample of environments.
"""
import numpy as np
import pytest
import torch

from maexp_b200.config import Config, baseline_config, make_config
from oracle.cases import cases
from oracle.oracle import Oracle
from tests import replay as rp
from tests.cuda_backend import CudaBackend

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", rp.fixture_names())
def test_cuda_replays_reference_fixture(name):
    fix = rp.load(name)
    b = CudaBackend(fix["config"], 1, first_global_env_id=fix["env_id"], **fix["oracle_kwargs"])
    assert rp.replay(fix, b) == len(fix["ticks"])


def _compare(a, b, keys, what):
    for k in keys:
        if a.get(k) is None:
            continue
        if not np.array_equal(a[k], b[k]):
            idx = np.argwhere(a[k] != b[k])[0]
            raise AssertionError(f"{what}: {k} differs at {idx.tolist()}: cuda {a[k][tuple(idx)]} oracle {b[k][tuple(idx)]}")


BATCH_CASES = ["cfg1_simple_mlp", "cfg2_four_rooms_transition_flip_sharedmean", "cfg3_three_rooms",
               "cfg4_four_rectangle_merged_types", "cfg5_four_rooms_n8", "indtrans_t2_v9_tasks",
               "no_keep_small", "relative_transparent", "simple_views_v5", "sequential_central_room_t3",
               "junctional_central_room_t5_merged", "two_central_rooms_t2_n12", "central_room_large_destination_v9",
               "dynamic_simple_20", "dynamic_simple_small_t2", "shuffle_spare_init_xys"]


@pytest.mark.parametrize("name", BATCH_CASES)
@pytest.mark.parametrize("num_envs", [37, 300])
def test_cuda_batch_matches_oracle(name, num_envs):
    """Ragged batch sizes (not a multiple of the envs-per-warp / per-CTA tiling), Philox policy,
    auto-reset across an episode boundary."""
    cfg, _, _, okw = cases()[name]
    cfg = cfg.merged_with(max_episode_length=12)
    gid0 = 1000
    cuda = CudaBackend(cfg, num_envs, first_global_env_id=gid0, auto_reset=True, extras=True, **okw)
    orc = Oracle(cfg, num_envs, auto_reset=True, first_global_env_id=gid0, extras=True, **okw)
    keys = ("obs", "obs_rel", "rewards", "dones", "was_reset", "coordinates", "objects", "destination")
    _compare(cuda.reset(0), orc.reset(0), keys, f"{name} reset")
    for tick in range(1, 30):
        a, b = cuda.step(tick), orc.step(tick, threads=4)
        _compare(a, b, keys, f"{name} tick {tick}")
        sa, sb = cuda.state(), orc.state()
        _compare(sa, sb, ("pos", "objects", "counters", "steps", "destination"), f"{name} state tick {tick}")
    st = cuda.env.stats()
    assert st["error_flag"] == 0 and st["resets"] >= num_envs * 2  # initial reset + one auto-reset each


def test_dynamic_obstacles_out_of_lockstep():
    """dynamic_simple with episodes that end by exhaustion at different times: every environment keeps its own
    world-step count (a reset consumes a tick without moving the obstacles, dynamic_simple_map.py:29-63 /
    abstract_map.py:32-36), so the obstacle cells differ between environments of one launch."""
    from maexp_b200.config import make_config
    cfg = make_config(map=dict(name="dynamic_simple", SIZE_X=10, SIZE_Y=10), keep_objects_num=False, num_objects=2,
                      num_agents=6, max_episode_length=40,
                      init_xys=[[1, 1], [-1, 1], [1, -1], [-1, -1], [2, 2], [-2, -2]])
    E, gid0 = 150, 77
    cuda = CudaBackend(cfg, E, first_global_env_id=gid0, auto_reset=True, extras=True)
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=gid0, extras=True)
    keys = ("obs", "rewards", "dones", "was_reset", "coordinates", "objects")
    _compare(cuda.reset(0), orc.reset(0), keys, "reset")
    resets_per_tick = set()
    for tick in range(1, 160):
        a, b = cuda.step(tick), orc.step(tick, threads=4)
        _compare(a, b, keys, f"tick {tick}")
        resets_per_tick.add(int(a["was_reset"].sum()))
        if tick % 20 == 0:
            _compare(cuda.state(), orc.state(), ("pos", "objects", "counters", "steps"), f"state tick {tick}")
    assert len(resets_per_tick) > 3 and cuda.env.stats()["error_flag"] == 0  # the environments did drift apart


@pytest.mark.parametrize("epw", [8, 16, 32])
@pytest.mark.parametrize("name", ["cfg2_four_rooms_transition_flip_sharedmean", "cfg5_four_rooms_n8", "merged_flip_three_rooms"])
def test_every_envs_per_warp_instantiation(name, epw, monkeypatch):
    """The library picks 8 / 16 / 32 envs per warp from the batch size; force each instantiation on a ragged batch."""
    monkeypatch.setenv("MAEXP_EPW", str(epw))
    cfg, _, _, okw = cases()[name]
    cfg = cfg.merged_with(max_episode_length=9)
    cuda = CudaBackend(cfg, 133, first_global_env_id=77, auto_reset=True, extras=True, **okw)
    orc = Oracle(cfg, 133, auto_reset=True, first_global_env_id=77, extras=True, **okw)
    keys = ("obs", "obs_rel", "rewards", "dones", "was_reset", "coordinates", "objects", "destination")
    _compare(cuda.reset(0), orc.reset(0), keys, f"{name} reset")
    for tick in range(1, 22):
        _compare(cuda.step(tick), orc.step(tick, threads=4), keys, f"{name} epw {epw} tick {tick}")


@pytest.mark.parametrize("num_agents", [1, 3, 12, 16])
def test_agent_counts(num_agents):
    """1..16 agents (16 is the packing limit of the order / selection nibbles), including the eight-wide vector stores."""
    xy = [[4, 4], [-4, 4], [4, -4], [-4, -4], [6, 6], [-6, 6], [6, -6], [-6, -6], [4, 6], [-4, 6], [4, -6], [-4, -6],
          [6, 4], [-6, 4], [6, -4], [-6, -4]]
    cfg = make_config(num_agents=num_agents, init_xys=xy[:num_agents], agent_view_method="individual", max_episode_length=10)
    cuda = CudaBackend(cfg, 70, first_global_env_id=5, auto_reset=True, extras=True)
    orc = Oracle(cfg, 70, auto_reset=True, first_global_env_id=5, extras=True)
    keys = ("obs", "rewards", "dones", "was_reset", "coordinates", "objects", "destination")
    _compare(cuda.reset(0), orc.reset(0), keys, "reset")
    for tick in range(1, 25):
        _compare(cuda.step(tick), orc.step(tick, threads=4), keys, f"N={num_agents} tick {tick}")
        _compare(cuda.state(), orc.state(), ("pos", "objects", "counters", "steps"), f"N={num_agents} state {tick}")


def test_masked_reset_and_injected_actions():
    cfg = baseline_config(5)
    E = 64
    cuda = CudaBackend(cfg, E, first_global_env_id=5, extras=True)
    orc = Oracle(cfg, E, first_global_env_id=5, extras=True)
    rng = np.random.default_rng(1)
    keys = ("obs", "rewards", "dones", "objects")
    _compare(cuda.reset(0), orc.reset(0), keys, "reset")
    for tick in range(1, 20):
        actions = rng.integers(0, 5, size=(E, 8)).astype(np.int8)
        order = np.stack([rng.permutation(8) for _ in range(E)]).astype(np.uint8)
        _compare(cuda.step(tick, actions, order), orc.step(tick, actions, order), keys, f"tick {tick}")
        if tick % 7 == 0:
            mask = (rng.random(E) < 0.3).astype(np.uint8)
            a, b = cuda.reset(1000 + tick, env_mask=mask), orc.reset(1000 + tick, env_mask=mask)
            sel = mask.astype(bool)
            for k in keys:
                assert np.array_equal(a[k][sel], b[k][sel]), (tick, k)
            _compare(cuda.state(), orc.state(), ("pos", "objects", "counters", "steps"), f"masked reset {tick}")


def test_full_size_invariants_and_sampled_oracle():
    """BASELINE.json config 5 at full size: 65 536 envs x 8 agents on four_rooms."""
    cfg = baseline_config(5)
    E, N = 65536, 8
    cuda = CudaBackend(cfg, E, first_global_env_id=0, auto_reset=True, extras=False)
    env = cuda.env
    wall = torch.as_tensor(env.map.wall_matrix.astype(np.int64), device=env.device)
    out = cuda.out
    env.reset(out, tick=0)
    T = 12
    for tick in range(1, T + 1):
        env.step(out, tick=tick)
        st = env.state()
        pos = st["pos"].long()
        # agents never stand on walls and never overlap (abstract_world.py:56-70)
        assert int(wall[pos[..., 0], pos[..., 1]].sum()) == 0
        key = pos[..., 0] * 64 + pos[..., 1]
        assert int((key.sort(dim=1).values.diff(dim=1) == 0).sum()) == 0
        # keep_objects_num: object count conserved (default_environment.py:214-215)
        assert torch.equal(st["objects"].sum(dim=(1, 2, 3)), torch.full((E,), 25, device=env.device))
        # objects never sit on walls, agents or outside their area; generated = 25 + completed
        assert int((st["objects"][:, 0].long() * wall).sum()) == 0
        assert torch.equal(st["counters"][:, 0], 25 + st["counters"][:, 1])
        obs = env.view(out["obs"])
        # area channel in {-1, 0}; centre cell: visible and occupied by the agent itself
        assert bool(((obs[:, :, 2] == 0) | (obs[:, :, 2] == -1)).all())
        assert bool((obs[:, :, 0, 3, 3] == 1).all()) and bool((obs[:, :, 2, 3, 3] == 0).all())
        # hidden cells carry no agents / objects (default views multiply by the mask)
        hidden = obs[:, :, 2] == -1
        assert float(obs[:, :, 0][hidden].abs().sum()) == 0 and float(obs[:, :, 1][hidden].abs().sum()) == 0
        # rewards in {-1, 0, 1}
        r = out["rewards"]
        assert bool(((r == 0) | (r == 1) | (r == -1)).all())
    assert env.stats()["env_steps"] == E * T
    # strided sample of envs re-run on the oracle (Philox keys are per global env id)
    sample = np.arange(0, E, 4099)
    final_obs, final_objects = env.view(out["obs"]), env.state()["objects"]
    for gid in sample:
        orc = Oracle(cfg, 1, auto_reset=True, first_global_env_id=int(gid))
        orc.reset(0)
        for tick in range(1, T + 1):
            o = orc.step(tick)
        assert np.array_equal(o["obs"][0], final_obs[gid].cpu().numpy()), gid
        assert np.array_equal(o["rewards"][0], out["rewards"][gid].cpu().numpy())
        assert np.array_equal(orc.state()["objects"][0], final_objects[gid].cpu().numpy())


def test_sharded_envs_equal_single_gpu_run():
    """Two shards with first_global_env_id offsets reproduce one big batch."""
    cfg = baseline_config(5)
    whole = CudaBackend(cfg, 96, first_global_env_id=0, auto_reset=True, extras=False)
    lo = CudaBackend(cfg, 48, first_global_env_id=0, auto_reset=True, extras=False)
    hi = CudaBackend(cfg, 48, first_global_env_id=48, auto_reset=True, extras=False)
    for b in (whole, lo, hi):
        b.reset(0)
    for tick in range(1, 10):
        w, a, b = whole.step(tick), lo.step(tick), hi.step(tick)
    assert np.array_equal(w["obs"], np.concatenate([a["obs"], b["obs"]]))
    assert np.array_equal(w["rewards"], np.concatenate([a["rewards"], b["rewards"]]))


def test_host_buffer_entry_point():
    cfg = baseline_config(5)
    E, N = 256, 8
    b = CudaBackend(cfg, E, first_global_env_id=0, auto_reset=True, extras=False)
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=0)
    b.reset(0), orc.reset(0)
    rng = np.random.default_rng(3)
    actions = torch.empty((E, N), dtype=torch.int8).pin_memory()
    rewards = torch.empty((E, N), dtype=torch.float32).pin_memory()
    dones = torch.empty((E, N), dtype=torch.uint8).pin_memory()
    was_reset = torch.empty((E,), dtype=torch.uint8).pin_memory()
    slot = b.env.new_slot()
    for tick in range(1, 6):
        a = rng.integers(0, 4, size=(E, N)).astype(np.int8)
        actions.copy_(torch.from_numpy(a))
        b.env.step_host(actions, slot, rewards, dones, was_reset, tick=tick)
        # order is drawn on chip (Philox) in the host entry point
        o = orc.step(tick, actions=a)
        assert np.array_equal(rewards.numpy(), o["rewards"]) and np.array_equal(dones.numpy(), o["dones"])
        assert np.array_equal(b.env.view(slot).cpu().numpy(), o["obs"])


def test_pipelined_host_entry_point():
    """maexp_env_step_host_async: 9 steps issued back to back (3 in flight), every step's host results equal the oracle's."""
    cfg = baseline_config(5).merged_with(max_episode_length=5)
    E, N, K = 256, 8, 9
    b = CudaBackend(cfg, E, first_global_env_id=0, auto_reset=True, extras=False)
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=0)
    b.reset(0), orc.reset(0)
    rng = np.random.default_rng(5)
    acts = [torch.from_numpy(rng.integers(0, 4, size=(E, N)).astype(np.int8)).pin_memory() for _ in range(K)]
    rewards = [torch.empty((E, N), dtype=torch.float32).pin_memory() for _ in range(K)]
    dones = [torch.empty((E, N), dtype=torch.uint8).pin_memory() for _ in range(K)]
    was_reset = [torch.empty((E,), dtype=torch.uint8).pin_memory() for _ in range(K)]
    slots = [b.env.new_slot() for _ in range(K)]
    tickets = [b.env.step_host_async(acts[k], slots[k], rewards[k], dones[k], was_reset[k], tick=k + 1) for k in range(K)]
    assert tickets == list(range(1, K + 1))
    b.env.host_wait(tickets[2])
    b.env.host_wait()
    resets = 0
    for k in range(K):
        o = orc.step(k + 1, actions=acts[k].numpy())
        assert np.array_equal(rewards[k].numpy(), o["rewards"]) and np.array_equal(dones[k].numpy(), o["dones"]), k
        assert np.array_equal(was_reset[k].numpy(), o["was_reset"]), k
        assert np.array_equal(b.env.view(slots[k]).cpu().numpy(), o["obs"]), k
        resets += int(was_reset[k].sum())
    assert resets > 0


def test_compact_host_entry_point():
    """maexp_env_step_host_compact_async: int8 rewards and one done flag per env carry the same information."""
    cfg = baseline_config(5).merged_with(max_episode_length=5)
    E, N, K = 200, 8, 8
    b = CudaBackend(cfg, E, first_global_env_id=0, auto_reset=True, extras=False)
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=0)
    b.reset(0), orc.reset(0)
    rng = np.random.default_rng(6)
    acts = [torch.from_numpy(rng.integers(0, 4, size=(E, N)).astype(np.int8)).pin_memory() for _ in range(K)]
    rewards = [torch.empty((E, N), dtype=torch.int8).pin_memory() for _ in range(K)]
    done = [torch.empty((E,), dtype=torch.uint8).pin_memory() for _ in range(K)]
    was_reset = [torch.empty((E,), dtype=torch.uint8).pin_memory() for _ in range(K)]
    slot = b.env.new_slot()
    for k in range(K):
        b.env.step_host_compact_async(acts[k], slot, rewards[k], done[k], was_reset[k], tick=k + 1)
    b.env.host_wait()
    for k in range(K):
        o = orc.step(k + 1, actions=acts[k].numpy())
        assert np.array_equal(rewards[k].numpy().astype(np.float32), o["rewards"]), k
        assert np.array_equal(done[k].numpy(), o["dones"][:, 0]) and np.array_equal(was_reset[k].numpy(), o["was_reset"]), k
    shared = CudaBackend(baseline_config(2), 32, auto_reset=True, extras=False)
    with pytest.raises(Exception):
        shared.env.step_host_compact_async(None, shared.env.new_slot(), rewards[0], done[0], was_reset[0], tick=1)


def test_unsupported_configs_fail_loudly():
    from tests.cuda_backend import make_env
    with pytest.raises(ValueError):
        make_env(make_config(agent_view_method="self"), 1)
    with pytest.raises(ValueError):
        make_env(make_config(observation_noise="flip", visible_range=9), 1)  # needs 4 probabilities (§E-6)
    with pytest.raises(ValueError):
        make_env(make_config(map=dict(name="four_rooms", SIZE_X=24, SIZE_Y=24), type_objects=3,
                             agent_tasks=["0"] * 8), 1)  # empty objects_area: reference spins forever (§E-8)


@pytest.mark.parametrize("name", ["cfg5_four_rooms_n8", "junctional_chain_tasks", "cfg4_four_rectangle_merged_types"])
def test_heatmaps_match_oracle(name):
    """Optional visit statistics (maexp_obs_out.heatmaps) summed over 64 environments, with auto-resets."""
    from oracle.cases import cases
    cfg = cases()[name][0].merged_with(max_episode_length=12)
    E = 64
    b = CudaBackend(cfg, E, first_global_env_id=11, auto_reset=True, extras=False)
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=11)
    want = orc.enable_heatmaps()
    buf, split = b.env.new_heatmaps()
    b.out["heatmaps"] = buf
    b.reset(0), orc.reset(0)
    for tick in range(1, 40):
        got, ref = b.step(tick), orc.step(tick)
        assert np.array_equal(got["rewards"], ref["rewards"])
    for key, t in split().items():
        assert np.array_equal(t.cpu().numpy(), want[key]), key
    assert want["heatmap_agents"].sum() > 0 and want["heatmap_wall_collision"].sum() > 0


@pytest.mark.parametrize("name,E", [("cfg5_four_rooms_n8", 300), ("relative_transparent", 70),
                                    ("cfg4_four_rectangle_merged_types", 70),
                                    ("cfg2_four_rooms_transition_flip_sharedmean", 1000), ("sensing_dist", 64),
                                    ("merged_flip_three_rooms", 70)])  # 25 x 25 canvas: C*X*Y % 4 != 0, the scalar kernel
def test_heatmap_observation_matches_oracle(name, E):
    """heatmap_observation (default_environment.py:93,238-247) summed over a batch of environments with auto-resets:
    local, relative and merged views; bit-exact for the integer-valued observations (any summation order), 1e-5 with
    gaussian sensing noise (maexp_heatmap_observation_add + maexp_decode_relative)."""
    from oracle.cases import cases
    cfg = cases()[name][0].merged_with(max_episode_length=9)
    gaussian = cfg.observation_noise in ("sensing_dist", "threshold_sensing_dist")
    b = CudaBackend(cfg, E, first_global_env_id=11, auto_reset=True, extras=True)
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=11, extras=True)
    heat = b.env.new_heatmap_observation()
    want = np.zeros(tuple(heat.shape), np.float32)
    rng = np.random.default_rng(3)
    shape = (E, b.env.N, b.env.C, b.env.V, b.env.V)
    tape = rng.standard_normal(shape).astype(np.float32) if gaussian else None
    b.reset(0, noise_tape=tape), orc.reset(0, noise_tape=tape)
    resets = 0
    for tick in range(1, 25):
        tape = rng.standard_normal(shape).astype(np.float32) if gaussian else None
        got, ref = b.step(tick, noise_tape=tape), orc.step(tick, noise_tape=tape, threads=4)
        assert np.array_equal(got["obs"], ref["obs"])
        b.env.accumulate_heatmap_observation(b.out, heat, chunk_envs=48)
        orc.heatmap_observation_add(want, ref)
        resets += int(ref["was_reset"].sum())
    got = heat.cpu().numpy()
    if gaussian:
        assert np.allclose(got, want, rtol=1e-5, atol=1e-3)
    else:
        assert np.array_equal(got, want)
    assert resets >= E and np.abs(want).sum() > 0


def _random_config(rng):
    """A random valid configuration over the whole supported surface (maps, views, masks, noises, environments)."""
    maps = {
        "simple": (dict(name="simple", SIZE_X=int(rng.integers(12, 27)), SIZE_Y=int(rng.integers(12, 25))), (1, 2)),
        "four_rooms": (dict(name="four_rooms", SIZE_X=24, SIZE_Y=24), (1, 2)),
        "three_rooms": (dict(name="three_rooms", SIZE_X=25, SIZE_Y=25), (1, 2)),
        "four_rectangle": (dict(name="four_rectangle", SIZE_X=40, SIZE_Y=24), (1, 2)),
        "central_room": (dict(name="central_room", SIZE_X=25, SIZE_Y=25), (1,)),
        "central_room_destination": (dict(name="central_room_destination", SIZE_X=25, SIZE_Y=25), (1,)),
        "central_room_large": (dict(name="central_room_large", SIZE_X=49, SIZE_Y=25), (1,)),
        "central_room_large_destination": (dict(name="central_room_large_destination", SIZE_X=49, SIZE_Y=25), (1,)),
        "two_central_rooms": (dict(name="two_central_rooms", SIZE_X=50, SIZE_Y=25), (1, 2)),
    }
    name = str(rng.choice(list(maps)))
    m, types = maps[name]
    environment = str(rng.choice(["default", "default", "shared_mean_reward", "shared_max_reward", "sequential", "junctional"]))
    T = int(rng.choice(types))
    if environment == "sequential":
        T = int(rng.integers(1, 5)) if name.startswith("central_room") else T
    if environment == "junctional":
        if not name.startswith("central_room"):
            environment = "default"
        else:
            T = 5
    N = int(rng.integers(1, 11))
    xy = [[2, 2], [-2, 2], [2, -2], [-2, -2], [0, 2], [0, -2], [2, 0], [-2, 0], [3, 3], [-3, -3]]
    if name == "four_rooms" or name == "three_rooms" or name == "four_rectangle":
        xy = [[4, 4], [-4, 4], [4, -4], [-4, -4], [6, 6], [-6, 6], [6, -6], [-6, -6], [4, 6], [-4, 6]]
    tasks = []
    for _ in range(N):
        k = int(rng.integers(0, T + 1))
        tasks.append("-1" if rng.random() < 0.1 else "".join(str(d) for d in rng.permutation(T)[:max(1, k)]))
    mask = str(rng.choice(["local", "local", "relative", "merged"]))
    noise = rng.choice([False, False, "flip", "sensing_dist", "threshold_sensing_dist"])
    noise = False if (mask == "relative" or noise == "False") else (noise if noise is False else str(noise))
    V = int(rng.choice([5, 7, 9]))
    cfg = make_config(map=m, environment=environment, type_objects=T, num_agents=N, init_xys=xy[:N], agent_tasks=tasks,
                      agent_view_method=str(rng.choice(["default", "simple", "individual", "transition", "individual_transition"])),
                      object_view_method=str(rng.choice(["default", "simple", "object_types"])), past_step=int(rng.integers(1, 5)),
                      observation_area_mask=mask, observation_noise=noise, visible_range=V,
                      flip_noise_probabilities=[0.2, 0.1, 0.05, 0.02], std_distribution=[0.3, 0.2, 0.1, 0.05],
                      noise_threshold_ratio=0.7, transparent_observation=bool(rng.random() < 0.2),
                      keep_objects_num=bool(rng.random() < 0.7), num_objects=int(rng.integers(3, 40)),
                      shuffle_init_xys=bool(rng.random() < 0.7), max_episode_length=int(rng.integers(6, 30)))
    return cfg


@pytest.mark.parametrize("seed", range(8))
def test_random_dynamic_simple_and_spare_init_rows_match_oracle(seed):
    """The sweep above on the dynamic_simple map (moving obstacles, random sizes) with up to four spare init_xys rows
    under shuffle_init_xys: outputs and state bit-identical to the C oracle across resets."""
    rng = np.random.default_rng(5000 + seed)
    for _ in range(50):
        cfg = _random_config(rng)
        if cfg.environment in ("sequential", "junctional"):
            continue
        X, Y = int(rng.integers(9, 27)), int(rng.integers(9, 25))
        spare = [[1, 1], [-1, 1], [1, -1], [-1, -1]][:int(rng.integers(0, 5))]
        # the spare rows sit next to the centre; the drawn rows are the sweep's (+-2.. cells), all distinct
        rows = [xy for xy in list(cfg.init_xys) if max(abs(xy[0]), abs(xy[1])) < min(X, Y) // 2 - 1]
        if not rows:
            continue
        N = len(rows)
        cfg = cfg.merged_with(map=dict(name="dynamic_simple", SIZE_X=X, SIZE_Y=Y), num_agents=N, init_xys=rows + spare,
                              agent_tasks=list(cfg.agent_tasks)[:N], type_objects=cfg.type_objects)
        E, first = int(rng.choice([1, 17, 70])), int(rng.integers(0, 1000))
        try:
            cuda = CudaBackend(cfg, E, first_global_env_id=first, auto_reset=True, extras=True)
        except ValueError:
            continue
        break
    else:
        pytest.skip("no valid configuration drawn")
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=first, extras=True)
    gaussian = cfg.observation_noise in ("sensing_dist", "threshold_sensing_dist")
    shape = (E, cuda.env.N, cuda.env.C, cuda.env.V, cuda.env.V)
    keys = ("obs", "obs_rel", "rewards", "dones", "was_reset", "coordinates", "objects", "destination")
    tape = rng.standard_normal(shape).astype(np.float32) if gaussian else None
    _compare(cuda.reset(0, noise_tape=tape), orc.reset(0, noise_tape=tape), keys, f"seed {seed} reset {dict(cfg)}")
    for tick in range(1, 40):
        tape = rng.standard_normal(shape).astype(np.float32) if gaussian else None
        _compare(cuda.step(tick, noise_tape=tape), orc.step(tick, noise_tape=tape, threads=4), keys, f"seed {seed} tick {tick}")
        _compare(cuda.state(), orc.state(), ("pos", "objects", "counters", "steps", "destination"), f"seed {seed} state {tick}")
    assert cuda.env.stats()["error_flag"] == 0


@pytest.mark.parametrize("seed", range(24))
def test_random_configurations_match_oracle(seed):
    """Randomised sweep over maps x environments x views x masks x noises x sizes: every output and the state stay
    bit-identical to the C oracle across resets (gaussian noises are compared with taped draws)."""
    rng = np.random.default_rng(1000 + seed)
    for _ in range(50):
        cfg = _random_config(rng)
        E, first = int(rng.choice([1, 17, 70])), int(rng.integers(0, 1000))
        try:
            cuda = CudaBackend(cfg, E, first_global_env_id=first, auto_reset=True, extras=True)
        except ValueError:
            continue  # e.g. the object area cannot hold num_objects (the reference would spin forever)
        break
    else:
        pytest.skip("no valid configuration drawn")
    orc = Oracle(cfg, E, auto_reset=True, first_global_env_id=first, extras=True)
    gaussian = cfg.observation_noise in ("sensing_dist", "threshold_sensing_dist")
    shape = (E, cuda.env.N, cuda.env.C, cuda.env.V, cuda.env.V)
    keys = ("obs", "obs_rel", "rewards", "dones", "was_reset", "coordinates", "objects", "destination")
    tape = rng.standard_normal(shape).astype(np.float32) if gaussian else None
    _compare(cuda.reset(0, noise_tape=tape), orc.reset(0, noise_tape=tape), keys, f"seed {seed} reset {dict(cfg)}")
    for tick in range(1, 26):
        tape = rng.standard_normal(shape).astype(np.float32) if gaussian else None
        _compare(cuda.step(tick, noise_tape=tape), orc.step(tick, noise_tape=tape, threads=4), keys, f"seed {seed} tick {tick}")
        _compare(cuda.state(), orc.state(), ("pos", "objects", "counters", "steps", "destination"), f"seed {seed} state {tick}")
    assert cuda.env.stats()["error_flag"] == 0


@pytest.mark.parametrize("noise", ["sensing_dist", "threshold_sensing_dist", "flat", "flip"])
def test_on_chip_noise_stream_has_the_reference_distribution(noise):
    """The on-chip draws (Philox + Box-Muller with libm logf / cosf / sinf) are bit-exact with the reference only under
    injected tapes; without a tape the claim is distributional.  Merged view: `obs` carries the noise, `obs_rel` is the same
    observation without it (merged_observtion_handler.py:27,41-47), so their difference is the noise the kernel drew.
    sensing_dist: z = (obs - clean) / sigma_ring must be N(0,1) like normal_(0, std) in dist_observation_noise.py:33-35
    (moments, Kolmogorov-Smirnov distance, tail mass, no correlation between neighbours); threshold_sensing_dist keeps
    exactly the draws with |noise| > sigma * ratio (:53-75); flat is N(0, 0.03^2) on three channels
    (flat_observation_noise.py:28-34); flip: |rint(x + u * range)| of u ~ U[0,1) flips a cell with probability p_ring
    (flip_observation_noise.py:28-46) — checked through the flip rate per ring."""
    import math
    from tests.cuda_backend import make_env
    E = 4096
    ratio = 1.0
    cfg = baseline_config(5).merged_with(observation_area_mask="merged", observation_noise=noise, noise_threshold_ratio=ratio,
                                         std_distribution=[0.3, 0.2, 0.1], flip_noise_probabilities=[0.2, 0.1, 0.05])
    env = make_env(cfg, E, auto_reset=True)
    out = env.new_outputs()
    env.reset(out, tick=0)
    env.step(out, tick=1)
    noisy, clean = env.view(out["obs"]).double().cpu(), env.view(out["obs_rel"]).double().cpu()  # [E,N,C,V,V]
    V, C = env.V, env.C
    a = torch.arange(V)
    ring = torch.minimum(torch.minimum(a[:, None], a[None, :]), torch.minimum(V - 1 - a[:, None], V - 1 - a[None, :])).clamp(max=V // 2 - 1)
    diff = noisy - clean
    if noise == "flip":
        # channel 2 is shifted by +1 before and -1 after the rounding (flip_observation_noise.py:33,44): same flip rate
        flipped = (diff.abs() > 0.5).double()
        for r, p in enumerate([0.2, 0.1, 0.05]):
            rate = float(flipped[..., ring == r].mean())
            n = flipped[..., ring == r].numel()
            assert abs(rate - p) < 6 * math.sqrt(p * (1 - p) / n), (r, rate, p)
        return
    sigma = torch.tensor([0.3, 0.2, 0.1])[ring] if noise != "flat" else torch.full((V, V), 0.03)
    z = (diff / sigma).reshape(-1, C, V, V)
    if noise == "threshold_sensing_dist":
        kept = z != 0
        # P(|N(0,1)| > ratio) of the draws survive, and every survivor lies outside the threshold
        p = math.erfc(ratio / math.sqrt(2))
        assert abs(float(kept.double().mean()) - p) < 6 * math.sqrt(p * (1 - p) / kept.numel())
        assert float(z[kept].abs().min()) > ratio - 1e-6
        tail = z[kept]
        assert abs(float(tail.mean())) < 0.01 and abs(float((tail ** 2).mean()) - (1 + ratio * math.sqrt(2 / math.pi) * math.exp(-ratio ** 2 / 2) / p)) < 0.02
        return
    z = z.reshape(-1)
    n = z.numel()
    assert abs(float(z.mean())) < 5 / math.sqrt(n) + 2e-4
    assert abs(float(z.var()) - 1) < 5 * math.sqrt(2 / n) + 2e-3   # fp32 rounding of obs + noise widens the bound a little
    assert abs(float((z ** 3).mean())) < 0.02 and abs(float((z ** 4).mean()) - 3) < 0.05
    sample = z[torch.randperm(n, generator=torch.Generator().manual_seed(0))[:200_000]].sort()[0]
    cdf = 0.5 * (1 + torch.erf(sample / math.sqrt(2)))
    ks = float((cdf - torch.arange(1, sample.numel() + 1) / sample.numel()).abs().max())
    assert ks < 1.63 / math.sqrt(sample.numel()) + 1e-3, ks             # 1 % critical value of the KS statistic
    assert abs(float((z.abs() > 3).double().mean()) - 0.0026998) < 3e-4   # tail mass
    zz = z.reshape(-1, V * V)
    assert abs(float(torch.corrcoef(torch.stack([zz[:, 0], zz[:, 1]]))[0, 1])) < 0.01  # neighbouring elements are independent


# ---------------------------------------------------------------------------------------------------------------------
# env_step_lanes_kernel (csrc/env_lanes.cu): one lane per (environment, agent).  The library picks it per call when the
# shape allows it and no snapshot output is requested, so these tests ask for the plain outputs only.
def _lanes_outputs(env):
    d = env.device
    o = env.new_outputs(extras=False)
    o["positions"] = torch.zeros((env.E, env.N, 2), dtype=torch.int8, device=d)
    o["destination"] = torch.zeros((env.E, env.N), dtype=torch.uint8, device=d)
    o["actions_taken"] = torch.zeros((env.E, env.N), dtype=torch.int8, device=d)
    o["rewards_i8"] = torch.zeros((env.E, env.N), dtype=torch.int8, device=d)
    o["done_env"] = torch.zeros((env.E,), dtype=torch.uint8, device=d)
    return o


def _lanes_variants():
    c = cases()
    base = c["cfg5_four_rooms_n8"][0]
    out = {name: (c[name][0], c[name][3]) for name in
           ("cfg1_simple_mlp", "cfg3_three_rooms", "cfg5_four_rooms_n8", "simple_views_v5", "relative_transparent")}
    out["n8_no_keep_few_objects"] = (base.merged_with(keep_objects_num=False, num_objects=3), {})
    out["n8_shared_mean"] = (base.merged_with(environment="shared_mean_reward"), {})
    out["n8_shared_max"] = (base.merged_with(environment="shared_max_reward"), {"shared_max_is_mean": False})
    out["n8_tasks"] = (base.merged_with(agent_tasks=["0", "-1", "0", "0", "-1", "0", "0", "0"]), {})
    out["n8_dense_objects"] = (base.merged_with(num_objects=120), {})
    out["n4_four_rooms"] = (base.merged_with(num_agents=4, init_xys=[[4, 4], [-4, 4], [4, -4], [-4, -4]], num_objects=60), {})
    return out


@pytest.mark.parametrize("name", sorted(_lanes_variants()))
@pytest.mark.parametrize("num_envs", [37, 300])
def test_lanes_kernel_matches_oracle_and_lane_per_env_kernel(name, num_envs, monkeypatch):
    """Ragged batches through env_step_lanes_kernel against the oracle AND against env_step_kernel (MAEXP_LANES=0) run on
    the same keys: observations, rewards, dones, positions, actions, compact outputs and the whole state, every tick,
    across an auto-reset."""
    cfg, okw = _lanes_variants()[name]
    cfg = cfg.merged_with(max_episode_length=12)
    gid0 = 4242
    monkeypatch.setenv("MAEXP_LANES", "1")  # the default takes it for N = 8 only
    lanes = CudaBackend(cfg, num_envs, first_global_env_id=gid0, auto_reset=True, extras=False, **okw)
    assert lanes.env.uses_lanes_kernel(), "this shape should take the agent-parallel kernel"
    lanes.out = _lanes_outputs(lanes.env)
    monkeypatch.setenv("MAEXP_LANES", "0")
    plain = CudaBackend(cfg, num_envs, first_global_env_id=gid0, auto_reset=True, extras=False, **okw)
    assert not plain.env.uses_lanes_kernel()
    plain.out = _lanes_outputs(plain.env)
    orc = Oracle(cfg, num_envs, auto_reset=True, first_global_env_id=gid0, extras=True, **okw)
    keys = ("obs", "rewards", "dones", "was_reset")
    raw = ("obs", "rewards", "dones", "was_reset", "positions", "destination", "actions_taken", "rewards_i8", "done_env")

    def check(a, b, ref, what):
        _compare(a, ref, keys, f"{name} {what} (lanes vs oracle)")
        for k in raw:
            assert torch.equal(lanes.out[k], plain.out[k]), f"{name} {what}: {k} differs between the two kernels"
        _compare(lanes.state(), orc.state(), ("pos", "objects", "counters", "steps", "destination"), f"{name} state {what}")

    check(lanes.reset(0), plain.reset(0), orc.reset(0), "reset")
    for tick in range(1, 30):
        check(lanes.step(tick), plain.step(tick), orc.step(tick, threads=4), f"tick {tick}")
    st = lanes.env.stats()
    assert st["error_flag"] == 0 and st["resets"] >= num_envs * 2 and st == plain.env.stats()


def test_lanes_kernel_injected_actions_orders_and_masked_resets():
    cfg = baseline_config(5)
    E = 70
    cuda = CudaBackend(cfg, E, first_global_env_id=5, extras=False)
    assert cuda.env.uses_lanes_kernel()
    orc = Oracle(cfg, E, first_global_env_id=5, extras=True)
    rng = np.random.default_rng(3)
    keys = ("obs", "rewards", "dones")
    _compare(cuda.reset(0), orc.reset(0), keys, "reset")
    for tick in range(1, 24):
        actions = rng.integers(0, 5, size=(E, 8)).astype(np.int8)
        order = np.stack([rng.permutation(8) for _ in range(E)]).astype(np.uint8)
        _compare(cuda.step(tick, actions, order), orc.step(tick, actions, order), keys, f"tick {tick}")
        _compare(cuda.state(), orc.state(), ("pos", "objects", "counters", "steps"), f"state {tick}")
        if tick % 7 == 0:
            mask = (rng.random(E) < 0.3).astype(np.uint8)
            a, b = cuda.reset(1000 + tick, env_mask=mask), orc.reset(1000 + tick, env_mask=mask)
            sel = mask.astype(bool)
            for k in keys:
                assert np.array_equal(a[k][sel], b[k][sel]), (tick, k)
            _compare(cuda.state(), orc.state(), ("pos", "objects", "counters", "steps"), f"masked reset {tick}")


def test_lanes_kernel_full_size_equals_lane_per_env_kernel(monkeypatch):
    """BASELINE configs[4] at full size: both kernels from the same seed, 230 ticks (one auto-reset), every output equal."""
    cfg = baseline_config(5)
    E = 65536
    lanes = CudaBackend(cfg, E, first_global_env_id=0, auto_reset=True, extras=False)
    assert lanes.env.uses_lanes_kernel()
    monkeypatch.setenv("MAEXP_LANES", "0")
    plain = CudaBackend(cfg, E, first_global_env_id=0, auto_reset=True, extras=False)
    lanes.env.reset(lanes.out, tick=0)
    plain.env.reset(plain.out, tick=0)
    for tick in range(1, 231):
        lanes.env.step(lanes.out, tick=tick)
        plain.env.step(plain.out, tick=tick)
        if tick % 10 == 0 or tick > 195:
            for k in ("obs", "rewards", "dones", "was_reset"):
                assert torch.equal(lanes.out[k], plain.out[k]), (tick, k)
    sa, sb = lanes.env.state(), plain.env.state()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert lanes.env.stats() == plain.env.stats() and lanes.env.stats()["error_flag"] == 0
