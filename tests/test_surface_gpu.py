"""The reference-facing plugin surface on top of the kernels: single-env environment, decoders, actor, trainer."""
import numpy as np
import pytest
import torch

from maexp_b200.config import Config, baseline_config
from oracle.oracle import Oracle
from tests import replay as rp

pytestmark = pytest.mark.gpu


def _world(cfg):
    from maexp_b200.core.maps import generate_map
    from maexp_b200.core.worlds import generate_world
    return generate_world(cfg, generate_map(cfg))


def test_generate_environment_returns_reference_structures():
    """env.reset()/env.step(action_n, order) as the reference's trainers call them, against a fixture of the
    live reference (same actions and orders as the tape draws)."""
    from maexp_b200.core.environments import generate_environment
    from oracle import ref_runner as rr
    fix = rp.load("cfg5_four_rooms_n8")
    cfg = Config(fix["config"])
    env = generate_environment(cfg, _world(cfg))
    assert env.action_space == [4] * 8 and env.observation_space == [[3, 7, 7]] * 8
    # the single-env wrapper keys its draws with global env id 0; compare against the oracle with the same id
    orc = Oracle(cfg, 1, first_global_env_id=0, extras=True)
    obs_n = env.reset()
    want = orc.reset(0)
    assert len(obs_n) == 8 and set(obs_n[0]) == {"local", "coordinates", "agents", "objects", "destination"}
    for i in range(8):
        assert np.array_equal(obs_n[i]["local"].numpy(), want["obs"][0, i])
        assert np.array_equal(obs_n[i]["objects"].numpy(), want["objects"][0, i])
        assert obs_n[i]["coordinates"]["map_x_max"] == want["coordinates"][0, i, 9]
    rng = np.random.default_rng(0)
    for tick in range(1, 25):
        actions, order = rng.integers(0, 5, 8), rng.permutation(8)
        rewards, dones, obs_n = env.step(list(actions), list(order))
        want = orc.step(tick, actions=actions[None], order=order[None])
        assert rewards == want["rewards"][0].tolist() and dones == [bool(d) for d in want["dones"][0]]
        for i in range(8):
            assert np.array_equal(obs_n[i]["local"].numpy(), want["obs"][0, i])
            assert np.array_equal(obs_n[i]["agents"].numpy(), want["agents"][0, i])
    assert env.objects_completed == int(orc.state()["counters"][0, 1]) and env.current_step == 24
    with pytest.raises(ValueError):
        generate_environment(cfg.merged_with(environment="nope"), _world(cfg))


def test_single_environment_heatmap_attributes():
    """heatmap_* / heatmap_accumulated_* attributes and accumulate_heatmap() of DefaultEnvironment
    (default_environment.py:36-61,70-96,249-257) against the oracle's statistics (which no reset zeroes)."""
    from maexp_b200.config import make_config
    from maexp_b200.core.environments import generate_environment
    cfg = make_config(num_agents=4, type_objects=2, object_view_method="object_types")
    env = generate_environment(cfg, _world(cfg))
    keys = ("heatmap_agents", "heatmap_complete", "heatmap_objects", "heatmap_objects_left", "heatmap_wall_collision",
            "heatmap_agents_collision")
    orc = Oracle(cfg, 1, first_global_env_id=0)
    want = orc.enable_heatmaps()
    rng = np.random.default_rng(4)
    tick = 0
    for episode in range(2):
        env.reset(), orc.reset(tick)
        for _ in range(20):
            tick += 1
            actions, order = rng.integers(0, 5, 4), rng.permutation(4)
            env.step(list(actions), list(order))
            orc.step(tick, actions=actions[None], order=order[None])
        tick += 1
        for k in keys:  # this episode's maps + the accumulated earlier ones = the oracle's running totals
            acc = getattr(env, k.replace("heatmap_", "heatmap_accumulated_"))
            assert np.array_equal(acc + getattr(env, k), want[k]), (episode, k)
        first_obs = env.heatmap_observation.copy()
        env.accumulate_heatmap()
    assert want["heatmap_agents"].sum() > 0 and np.array_equal(env.heatmap_accumulated_agents, want["heatmap_agents"])
    assert np.abs(env.heatmap_accumulated_observation).sum() > np.abs(first_obs).sum() > 0
    env.reset()
    assert not env.heatmap_agents.any() and env.heatmap_objects.sum() == 2 * int(cfg.num_objects)
    with pytest.raises(AttributeError):
        env.heatmap_nope


@pytest.mark.parametrize("mask", ["local", "relative"])
def test_observation_stats_environment(mask):
    """generate_environment(environment=observation_stats) (observation_stats_environment.py:16-61): observation_stats
    accumulates across episodes, heatmap_observation (default_environment.py:93,238-247) restarts at reset()."""
    from maexp_b200.config import make_config
    from maexp_b200.core.environments import ObservationStatsEnvironment, generate_environment
    cfg = make_config(environment="observation_stats", observation_area_mask=mask, num_agents=4)
    env = generate_environment(cfg, _world(cfg))
    assert isinstance(env, ObservationStatsEnvironment)
    orc = Oracle(cfg, 1, first_global_env_id=0, extras=True)
    shape = (4, 3, 7, 7) if mask == "local" else (4, 3, 24, 24)
    stats = np.zeros(shape, np.float32)
    rng = np.random.default_rng(1)
    tick = 0
    for episode in range(2):
        env.reset(), orc.reset(tick)
        heat = np.zeros(shape, np.float32)
        assert not env.heatmap_observation.any()
        for _ in range(12):
            tick += 1
            actions, order = rng.integers(0, 5, 4), rng.permutation(4)
            env.step(list(actions), list(order))
            want = orc.step(tick, actions=actions[None], order=order[None])
            orc.heatmap_observation_add(heat, want)
            orc.heatmap_observation_add(stats, want)
        tick += 1
        assert np.array_equal(env.heatmap_observation, heat)
    assert np.array_equal(np.stack(env.observation_stats), stats) and np.abs(stats).sum() > 0


def test_decode_relative_state_matches_reference_semantics():
    """observation_handler.py:150-216 restated in numpy on the oracle's coordinates."""
    from maexp_b200.core.handlers.observations import ObservationHandler
    cfg = baseline_config(4)  # four_rectangle 40x24, merged, C=11
    orc = Oracle(cfg, 6, first_global_env_id=0, extras=True)
    orc.reset(0)
    out = orc.step(1)
    B = 6 * 8
    rel = torch.from_numpy(out["obs_rel"].reshape(B, 11, 7, 7)).cuda()
    pos = torch.from_numpy(np.stack([out["agents"][e, i, i] for e in range(6) for i in range(8)])).cuda()
    coords = out["coordinates"].reshape(B, 12)
    want = np.zeros((B, 11, 40, 24), np.float32)
    want[:, -1] -= 1
    for b in range(B):
        c = coords[b]
        want[b, :, c[8]:c[9], c[10]:c[11]] = out["obs_rel"].reshape(B, 11, 7, 7)[b, :, c[0]:c[1], c[2]:c[3]]
    got = ObservationHandler.decode_relative_state({"relative": rel, "positions": pos}, [40, 24])
    assert np.array_equal(got.cpu().numpy(), want)
    area = ObservationHandler.decode_relative_state({"relative": rel, "positions": pos}, [40, 24], only_area_plus_one=True)
    assert np.array_equal(area.cpu().numpy(), want[:, -1:] + 1)
    one = ObservationHandler.decode_relative_state({"relative": rel[3], "positions": pos[3]}, [40, 24])
    assert np.array_equal(one.cpu().numpy(), want[3])


def test_batched_actor_epsilon_greedy():
    from maexp_b200.core.agents import BatchedActor
    from maexp_b200.core.agents.brains import LearnerGroup
    from tests.cuda_backend import make_env
    cfg = baseline_config(5).merged_with(agent_tasks=["0"] * 7 + ["-1"])
    env = make_env(cfg, 300, auto_reset=True)
    out = env.new_outputs()
    env.reset(out, tick=0)
    learner = LearnerGroup(cfg, [3, 7, 7], 4, 8, 32, obs_stride=env.slot_floats)
    actor = BatchedActor(cfg, env, learner, chunk=128)  # 300 envs -> chunks of 128, 128, 44
    actor.refresh()
    greedy = actor.act(out["obs"], 0.0, tick=1).clone()
    q = torch.cat([learner_q for learner_q in [actor.group.q_values(out["obs"][:, e0:e0 + 100].contiguous())
                                               .clone() for e0 in (0, 100, 200)]], dim=1)  # [N, 300, A]
    want = q.argmax(dim=2).transpose(0, 1).to(torch.int8)
    assert torch.equal(greedy[:, :7], want[:, :7])  # epsilon = 0: argmax of the online net
    rnd = actor.act(out["obs"], 1.0, tick=2)
    assert int(rnd.min()) >= 0 and int(rnd.max()) <= 3
    counts = torch.bincount(rnd.flatten().long(), minlength=4).float()
    assert float(counts.min() / counts.sum()) > 0.2  # uniform over the four actions
    a3 = actor.act(out["obs"], 0.0, tick=3)[:, 7].clone()
    assert not torch.equal(a3, actor.act(out["obs"], 0.0, tick=4)[:, 7])  # the "-1" agent acts randomly


def test_trainer_runs_and_saves_reference_compatible_checkpoints(tmp_path):
    from maexp_b200.core.handlers.trainers import generate_trainer
    from tests.cuda_backend import make_env
    cfg = baseline_config(5).merged_with(max_episode_length=8, batch_size=32, populate_steps=4 * 64, capacity=64 * 16,
                                         synchronize_frequency=1, epsilon_decay=0.5)
    env = make_env(cfg, 64, auto_reset=True)
    trainer = generate_trainer(cfg, env)
    before = trainer.learner.params().clone()
    hist = trainer.run(max_epochs=2)
    assert len(hist) == 2 and hist[-1]["epsilon"] == 0.25 and np.isfinite(hist[-1]["loss"])
    assert not torch.equal(trainer.learner.params(), before)  # networks were updated
    assert torch.equal(trainer.learner.params(1), trainer.learner.params())  # target synchronised at episode end
    assert trainer.env.stats()["env_steps"] > 0
    trainer.save_state_dict(str(tmp_path))
    sd = torch.load(str(tmp_path / "agent3.pth"))
    assert set(sd) == {"conv.conv_list.0.weight", "conv.conv_list.0.bias", "conv.conv_post.weight", "conv.conv_post.bias",
                       "fc_V.fc_list.0.weight", "fc_V.fc_list.0.bias", "fc_V.fc_post.weight", "fc_V.fc_post.bias",
                       "fc_A.fc_list.0.weight", "fc_A.fc_list.0.bias", "fc_A.fc_post.weight", "fc_A.fc_post.bias"}
    assert tuple(sd["fc_A.fc_post.weight"].shape) == (204, 128) and tuple(sd["conv.conv_post.weight"].shape) == (64, 32, 2, 2)
    trainer.load_state_dict(str(tmp_path))
    with pytest.raises(ValueError):
        generate_trainer(cfg.merged_with(trainer="nope"), env)


def test_trainer_merged_view_da3_end_to_end():
    """BASELINE cfg4 (MergedDA3 on the merged view): the trainer feeds ring positions to the learner and the actor;
    the Q-values it acts on equal the torch oracle's on the same observations and positions."""
    from maexp_b200.core.handlers.trainers import generate_trainer
    from oracle.torch_learner import TorchLearner
    from tests.cuda_backend import make_env
    cfg = baseline_config(4).merged_with(max_episode_length=6, batch_size=16, populate_steps=4 * 32, capacity=32 * 16,
                                         synchronize_frequency=1)
    env = make_env(cfg, 32, auto_reset=True)
    trainer = generate_trainer(cfg, env)
    assert trainer.env.model_obs_shape == [trainer.env.C, int(cfg.map.SIZE_X), int(cfg.map.SIZE_Y)]
    before = trainer.learner.params().clone()
    hist = trainer.run(max_epochs=1)
    assert np.isfinite(hist[-1]["loss"]) and not torch.equal(trainer.learner.params(), before)
    ring, e, N = trainer.buffer, trainer.env, trainer.env.N
    head = ring.head
    obs_slot, positions = ring.obs[head], ring.positions[head]            # [N,E,slot], [E,N,2]
    q = trainer.learner.q_values(obs_slot[:, :16].contiguous(), pos=positions[:16].transpose(0, 1).contiguous()).cpu()
    dense = e.view(obs_slot).cpu()                                        # [E,N,C,V,V]
    for a in (0, N - 1):
        o = TorchLearner("da3", trainer.learner.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8, area_mask="merged",
                         size=[int(cfg.map.SIZE_X), int(cfg.map.SIZE_Y)], num_heads=int(cfg.model.num_heads))
        want = o.q_values((dense[:16, a], positions[:16, a].cpu().long()))
        assert float((q[a] - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-6
    sd = trainer.learner.state_dict(0)
    assert tuple(sd["drl_head.weight"].shape)[0] == 4 and "local_blocks.0.attn.qkv.weight" in sd and "relative_conv.conv_post.weight" in sd


def test_trainer_merged_view_under_observation_noise_reads_the_noise_free_area_channel():
    """BASELINE cfg4 with observation noise: the merged models take the area channel of the NOISE-FREE "relative" key
    (merged_observtion_handler.py:27,41-47; customs/da3.py:217-225).  The ring keeps only the noisy crop (env slot `obs`)
    and the trainer binds the environment's line-of-sight table; the Q-values the learner computes from the ring equal
    the torch oracle's on (noisy local crop, clean relative crop) — the clean crop is what the env writes to `obs_rel`."""
    from maexp_b200.core.handlers.trainers import generate_trainer
    from oracle.torch_learner import TorchLearner
    from tests.cuda_backend import make_env
    cfg = baseline_config(4).merged_with(max_episode_length=6, batch_size=16, populate_steps=4 * 32, capacity=32 * 16,
                                         synchronize_frequency=1, observation_noise="sensing_dist")
    env = make_env(cfg, 32, auto_reset=True)
    trainer = generate_trainer(cfg, env)
    before = trainer.learner.params().clone()
    hist = trainer.run(max_epochs=1)
    assert np.isfinite(hist[-1]["loss"]) and not torch.equal(trainer.learner.params(), before)
    e, N = trainer.env, trainer.env.N
    out = e.new_outputs(extras=True)   # one more step with both keys kept: noisy `obs`, clean `obs_rel`
    e.step(out, tick=10_000)
    noisy, clean, positions = e.view(out["obs"]).cpu(), e.view(out["obs_rel"]).cpu(), out["positions"]
    assert float((noisy - clean).abs().max()) > 1e-3                      # the noise is there ...
    assert torch.equal(clean[:, :, -1], clean[:, :, -1].round())         # ... and the relative key is free of it
    q = trainer.learner.q_values(out["obs"][:, :16].contiguous(), pos=positions[:16].transpose(0, 1).contiguous()).cpu()
    for a in (0, N - 1):
        o = TorchLearner("da3", trainer.learner.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8, area_mask="merged",
                         size=[int(cfg.map.SIZE_X), int(cfg.map.SIZE_Y)], num_heads=int(cfg.model.num_heads))
        pos = positions[:16, a].cpu().long()
        want = o.q_values((noisy[:16, a], pos, {"relative": clean[:16, a]}))
        assert float((q[a] - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-6
        wrong = o.q_values((noisy[:16, a], pos))                          # area channel taken from the noisy crop
        assert float((wrong - want).abs().max()) > 1e-4


def test_evaluator_greedy_rollouts_with_saved_weights(tmp_path):
    """generate_evaluator: epsilon = 0 episodes with weights written by the trainer; greedy actions of identical
    networks on identical envs are reproducible, and the counters the reference logs are returned."""
    from maexp_b200.core.handlers.evaluators import generate_evaluator
    from maexp_b200.core.handlers.trainers import generate_trainer
    from tests.cuda_backend import make_env
    cfg = baseline_config(2).merged_with(max_episode_length=12, batch_size=16, populate_steps=4 * 32, capacity=32 * 16,
                                         observation_noise=False, validate_epochs=2)
    trainer = generate_trainer(cfg, make_env(cfg, 32, auto_reset=True))
    trainer.run(max_epochs=1)
    trainer.save_state_dict(str(tmp_path))
    runs = []
    for _ in range(2):
        ev = generate_evaluator(cfg.merged_with(pretrained_weight_path=str(tmp_path)), make_env(cfg, 32, auto_reset=False))
        ev.load_state_dict()
        assert torch.equal(ev.learner.params(), trainer.learner.params())
        runs.append(ev.run())
    assert len(runs[0]) == 2 and runs[0] == runs[1]
    assert set(runs[0][0]) >= {"episode/total_reward", "episode/objects_completed", "episode/agents_collided"}
    assert runs[0][0]["episode/objects_generated"] >= cfg.num_objects
    with pytest.raises(ValueError):
        generate_evaluator(cfg.merged_with(evaluator="nope"), trainer.env)


@pytest.mark.parametrize("environment,brain,model,mask,types,tasks,flags", [
    ("sequential", "da3_iqn", "da3_iqn", "merged", 3, ["0", "0", "1", "1", "2", "2", "012", "01"], ""),
    ("junctional", "da3_iqn", "da6_iqn", "merged", 5, ["0", "0", "12", "12", "13", "13", "4", "4"], ""),
    ("sequential", "da6", "da6", "merged", 3, ["0", "0", "1", "1", "2", "2", "012", "01"], ""),
    ("default", "iqn", "iqn", "relative", 1, ["0"] * 8, ""),
    ("sequential", "da3", "da3", "relative", 2, ["0", "1"] * 4, ""),
    # objects_channel / agents_channel / destination_channel on, as configs/experiments/IJCNN2025/exp1/goapos/*.yaml
    # (Gpos+Opos+Apos: eda3_iqn, eda6_iqn, iqn, dqn, eda6_dqn, eda3_dqn) and KES2023/*.yaml (destination_channel)
    ("sequential", "da3_iqn", "da3_iqn", "relative", 3, ["0", "0", "1", "1", "2", "2", "012", "01"], "oa"),
    ("sequential", "da3_iqn", "da6_iqn", "merged", 3, ["0", "0", "1", "1", "2", "2", "012", "01"], "oa"),
    ("sequential", "iqn", "iqn", "relative", 3, ["0", "0", "1", "1", "2", "2", "012", "01"], "oa"),
    ("sequential", "da3_baseline", "da3_baseline", "relative", 3, ["0", "0", "1", "1", "2", "2", "012", "01"], "oa"),
    ("sequential", "da6", "da6", "merged", 2, ["0", "1"] * 4, "oa"),
    ("sequential", "da3", "da3", "relative", 3, ["0", "0", "1", "1", "2", "2", "012", "01"], "oa"),
    ("default", "da3_iqn", "da3_iqn", "relative", 1, ["0"] * 8, "d"),
    ("default", "da6", "da6", "merged", 1, ["0"] * 8, "d"),
])
def test_trainer_on_shipped_experiment_shapes(environment, brain, model, mask, types, tasks, flags):
    """The combinations of configs/experiments/IJCNN2025 (central_room, 8 agents, individual / object_types views):
    environment, brain, model and view wired through generate_trainer for a short run."""
    from maexp_b200.core.handlers.trainers import generate_trainer
    from tests.cuda_backend import make_env
    xy8 = [[-2, 2], [2, 2], [-2, -2], [2, -2], [0, 2], [0, -2], [2, 0], [-2, 0]]
    m = dict(name=model, embed_dim=32, patch_size=5, relative_patch_size=5, local_patch_size=1, block_loop=1, num_heads=2,
             mlp_ratio=2.0, attention="normal", af_lambd=0.5, hidden_channels=[8], output_channel=16, hidden_kernels=[2, 2],
             hidden_strides=[1, 1], conv_afs=["max_pool2d", "max_pool2d"], af_kernels=[2, 2], af_strides=[2, 2],
             hidden_layer_sizes=[32], fc_afs=["relu"], num_quantiles=8, num_quantiles_dash=8, num_cosines=8)
    from maexp_b200.config import make_config
    cfg = make_config(environment=environment, map=dict(name="central_room", SIZE_X=25, SIZE_Y=25), type_objects=types,
                      num_objects=40, agent_tasks=tasks, init_xys=xy8, agent_view_method="individual",
                      object_view_method="object_types", observation_area_mask=mask, agent_type="attention", trainer="attention",
                      brain=brain, model=m, max_episode_length=6, batch_size=8, populate_steps=4 * 16, capacity=16 * 16,
                      synchronize_frequency=1, objects_channel="o" in flags, agents_channel="a" in flags,
                      destination_channel="d" in flags)
    if "d" in flags:  # KES2023: one destination half per agent (central_room_destination_map.py)
        cfg = cfg.merged_with(map=dict(name="central_room_destination", SIZE_X=25, SIZE_Y=25))
    trainer = generate_trainer(cfg, make_env(cfg, 16, auto_reset=True))
    before = trainer.learner.params().clone()
    hist = trainer.run(max_epochs=1)
    assert np.isfinite(hist[-1]["loss"]) and not torch.equal(trainer.learner.params(), before)
    assert trainer.env.stats()["error_flag"] == 0
    if flags:
        _check_canvas_inputs_against_env_state(trainer, flags)


def _check_canvas_inputs_against_env_state(trainer, flags):
    """The ring's canvas inputs of the newest slot against the environment's state taps (reference dtypes): the last
    agent's objects snapshot is the objects_matrix after the step, every agent sees all agents' cells, and the destination
    rows are the agent's own destination_area_matrix.  Then a sampled batch against the ring, pick by pick."""
    from maexp_b200.core.agents.brains import Brain
    ring, env = trainer.buffer, trainer.env
    E, N, T, X, Y = env.E, env.N, env.T, env.X, env.Y
    st = env.state()
    inputs = ring.canvas_inputs(0, E)
    reads = trainer.learner.canvas_reads
    assert reads[0] == ("o" in flags) and reads[2] == ("d" in flags and trainer.config.model.name == "da3_iqn")
    if "o" in flags:
        rows = Brain._bit_rows(st["objects"].cpu())                       # [E, T, X]
        assert torch.equal(inputs["obj_rows"][N - 1].cpu(), rows)         # the last agent observes after every pickup
    if "a" in flags and reads[1]:
        assert torch.equal(inputs["agents"][0].cpu(), st["pos"].cpu()) and torch.equal(inputs["agents"][N - 1], inputs["agents"][0])
    if "d" in flags:
        assert ring.destination is not None
        if reads[2]:
            want = Brain._bit_rows(st["destination"].cpu()).transpose(0, 1)   # [N, E, X]
            assert torch.equal(inputs["dest_rows"].cpu(), want)
    batch = trainer.batch
    S = ring.S
    rng = np.random.default_rng(1)
    B = batch["actions"].shape[1]
    picks = np.stack([(ring.head - 1 - rng.integers(0, ring.count - 1, B)) % S, rng.integers(0, E, B)], 1).astype(np.int32)
    ring.sample(batch, indices=torch.from_numpy(picks).cuda())
    torch.cuda.synchronize()
    batch["error_flag"].zero_()  # injected picks may name transitions into reset slots
    for j, (s0, e) in enumerate(picks):
        s1 = (s0 + 1) % S
        if ring.objects_bits is not None:
            assert torch.equal(batch["obj_rows"][:, j], ring.objects_bits[s0, :, :, :, e])
            assert torch.equal(batch["next_obj_rows"][:, j], ring.objects_bits[s1, :, :, :, e])
        if ring.keep_agents:
            assert torch.equal(batch["agents"][:, j], ring.positions[s0, e].expand(N, N, 2))
            assert torch.equal(batch["next_agents"][:, j], ring.positions[s1, e].expand(N, N, 2))
        if ring.destination is not None:
            assert torch.equal(batch["dest_rows"][:, j], ring.dest_table[ring.destination[s0, e].long()])
            assert torch.equal(batch["next_dest_rows"][:, j], ring.dest_table[ring.destination[s1, e].long()])
