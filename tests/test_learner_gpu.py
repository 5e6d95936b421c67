"""Parity of the CUDA learner and replay sampler (through the C-ABI).

Tolerance (north_star): losses and Q-values within 1e-5 relative in fp32.  Updated weights are
checked with a looser bound because RMSprop's first steps normalise every gradient to about
+-lr/sqrt(1-alpha), which amplifies last-bit differences of near-zero gradients.
"""
import numpy as np
import pytest
import torch

from maexp_b200.config import Config, baseline_config
from oracle.make_golden_learner import configs as golden_configs
from oracle.torch_learner import TorchLearner
from tests.test_learner_golden import ATOMS, IQN_KINDS, KINDS, MASKED, base, load, make_oracle, oracle_inputs

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def make_group(kind, shape, G, B):
    from maexp_b200.core.agents.brains import LearnerGroup
    cfg, _ = golden_configs()[kind]
    return LearnerGroup(cfg.merged_with(batch_size=B), shape, 4, G, B)


def pack(group, x):
    """[G,B,C,V,V] dense -> [G,B,obs_stride] device rows (the replay sampler's layout)."""
    G, B = x.shape[:2]
    out = torch.zeros((G, B, group.obs_stride), dtype=torch.float32, device=group.device)
    out[:, :, : group.K] = x.reshape(G, B, -1).to(group.device)
    return out


def canvas_batch(extras, next_extras=None):
    """The reference's "objects" / "agents" / "destination" state keys ([G,B,...] stacks) in the library's layout."""
    from maexp_b200.core.agents.brains import Brain
    out = {}
    for pre, ex in (("", extras), ("next_", next_extras or {})):
        if "objects" in ex:
            out[pre + "obj_rows"] = Brain._bit_rows(ex["objects"]).contiguous().cuda()
        if "agents" in ex:
            out[pre + "agents"] = ex["agents"].to(torch.int8).contiguous().cuda()
        if "destination" in ex:
            out[pre + "dest_rows"] = Brain._bit_rows(ex["destination"]).contiguous().cuda()
    return out


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("kind", KINDS)
def test_cuda_learner_matches_reference_fixture(kind):
    fix = load(kind)
    b = fix["batch"]
    g = make_group(kind, fix["shape"], 1, b["obs"].shape[0])
    g.load_state_dict(0, fix["init"])
    g.load_state_dict(0, fix["target"], which=1)
    batch = {"obs": pack(g, b["obs"][None]), "next_obs": pack(g, b["next_obs"][None]),
             "actions": b["actions"].to(torch.int32)[None].contiguous().cuda(),
             "rewards": b["rewards"].float()[None].contiguous().cuda(),
             "dones": b["dones"].to(torch.uint8)[None].contiguous().cuda()}
    if base(kind) in MASKED:
        batch["pos"] = b["pos"].to(torch.int8)[None].contiguous().cuda()
        batch["next_pos"] = b["next_pos"].to(torch.int8)[None].contiguous().cuda()
        ex, next_ex = fix["extras"]
        batch.update(canvas_batch({k: v[None] for k, v in ex.items()}, {k: v[None] for k, v in next_ex.items()}))
    if fix["los"] is not None:  # *_noise: the stored crop is noisy, the area canvas comes from the line-of-sight table
        g.set_los(fix["los"])
    q = g.q_values(batch["obs"], tau_sel=fix["tau0"].cuda() if base(kind) in IQN_KINDS else None, pos=batch.get("pos"),
                   canvas=batch).cpu()[0]
    assert rel_err(q, fix["q0"]) < RTOL
    for u, want in enumerate(fix["losses"]):
        if base(kind) in IQN_KINDS:
            t, d, sel = fix["taus"][u]
            batch.update(taus=t[None].contiguous().cuda(), tau_dashes=d[None].contiguous().cuda(), tau_sel=sel.contiguous().cuda())
        loss = float(g.learn(batch).cpu()[0])
        assert abs(loss - want) <= RTOL * abs(want), (loss, want)
    got = g.state_dict(0)
    bad = 0
    total = 0
    for k, v in fix["final"].items():
        diff = (got[k] - v).abs()
        bad += int((diff > 1e-4 + 1e-4 * v.abs()).sum())
        total += v.numel()
    assert bad <= total * 1e-3, (bad, total)


@pytest.mark.parametrize("kind,B", [("dqn_conv_mlp", 32), ("c51", 256), ("qr", 256), ("iqn", 256), ("da3", 32),
                                    ("relative_conv_mlp", 32), ("merged_conv_mlp", 32), ("merged_da3", 32), ("da6", 32),
                                    ("da3_iqn", 32), ("merged_da3_iqn", 32), ("da6_iqn", 32), ("relative_da3", 32),
                                    ("relative_da3_iqn", 32), ("relative_iqn", 32), ("merged_iqn", 32),
                                    ("relative_conv_mlp_ch", 32), ("relative_da3_ch", 32), ("relative_iqn_ch", 32),
                                    ("relative_da3_iqn_ch", 32), ("da6_ch", 32), ("da6_iqn_ch", 32), ("merged_da3_iqn_ch", 32),
                                    ("merged_conv_mlp_noise", 32), ("merged_da3_noise", 32), ("da6_noise", 32)])
def test_cuda_learner_group_matches_torch_oracle(kind, B):
    """G independent networks with different weights and batches, BASELINE batch sizes."""
    G = 3
    fix = load(kind)
    C, V = fix["batch"]["obs"].shape[1:3]
    X, Y = fix["shape"][1:]
    g = make_group(kind, fix["shape"], G, B)
    gen = torch.Generator().manual_seed(11)
    oracles, batches = [], []
    for a in range(G):
        sd = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in fix["init"].items()}
        tg = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in fix["init"].items()}
        g.load_state_dict(a, sd)
        g.load_state_dict(a, tg, which=1)
        oracles.append(make_oracle(kind, fix, sd, tg))
        obs = (torch.rand((B, C, V, V), generator=gen) < 0.2).float()
        nxt = (torch.rand((B, C, V, V), generator=gen) < 0.2).float()
        obs[:, -1] -= 1
        nxt[:, -1] -= 1
        batches.append({"obs": obs, "next_obs": nxt, "actions": torch.randint(0, 4, (B,), generator=gen),
                        "rewards": torch.randint(-1, 2, (B,), generator=gen).double(),
                        "dones": torch.rand((B,), generator=gen) < 0.1})
        if base(kind) in MASKED:
            for key in ("pos", "next_pos"):
                batches[-1][key] = torch.stack([torch.randint(0, X, (B,), generator=gen), torch.randint(0, Y, (B,), generator=gen)], 1)
            ex = ({}, {})
            if kind.endswith("_ch"):  # random objects / agents / destination keys, shaped like the fixture's
                for d, fx in zip(ex, fix["extras"]):
                    T, Nteam = fx["objects"].shape[1], fx["agents"].shape[1]
                    d["objects"] = (torch.rand((B, T, X, Y), generator=gen) < 0.1).to(torch.int8)
                    d["agents"] = torch.stack([torch.randint(0, X, (B, Nteam), generator=gen),
                                               torch.randint(0, Y, (B, Nteam), generator=gen)], 2).to(torch.int8)
                    d["destination"] = (torch.rand((B, X, Y), generator=gen) < 0.5).to(torch.int8)
            if kind.endswith("_noise"):  # noisy stored crop; the clean "relative" crop's area channel follows the table
                for d, key, pkey in zip(ex, ("obs", "next_obs"), ("pos", "next_pos")):
                    clean = batches[-1][key].clone()
                    clean[:, -1] = torch.stack([torch.from_numpy(fix["los"])[int(p[0]), int(p[1])].float() - 1
                                                for p in batches[-1][pkey]])
                    d["relative"] = clean
                    batches[-1][key] = clean + 0.3 * torch.randn(clean.shape, generator=gen)
            batches[-1]["extras"] = ex
    batch = {"obs": pack(g, torch.stack([b["obs"] for b in batches])),
             "next_obs": pack(g, torch.stack([b["next_obs"] for b in batches])),
             "actions": torch.stack([b["actions"] for b in batches]).to(torch.int32).cuda(),
             "rewards": torch.stack([b["rewards"] for b in batches]).float().cuda(),
             "dones": torch.stack([b["dones"] for b in batches]).to(torch.uint8).cuda()}
    if base(kind) in MASKED:
        batch["pos"] = torch.stack([b["pos"] for b in batches]).to(torch.int8).cuda()
        batch["next_pos"] = torch.stack([b["next_pos"] for b in batches]).to(torch.int8).cuda()
        stack = lambda j: {k: torch.stack([b["extras"][j][k] for b in batches]) for k in batches[0]["extras"][j] if k != "relative"}
        batch.update(canvas_batch(stack(0), stack(1)))
    if fix["los"] is not None:
        g.set_los(fix["los"])
    taus = None
    if base(kind) in IQN_KINDS:
        taus = [(torch.rand((B, 4), generator=gen), torch.rand((B, 4), generator=gen), torch.rand((1, 4), generator=gen))
                for _ in range(G)]
        batch.update(taus=torch.stack([t[0] for t in taus]).cuda(), tau_dashes=torch.stack([t[1] for t in taus]).cuda(),
                     tau_sel=torch.cat([t[2] for t in taus]).cuda())
        for a in range(G):
            oracles[a].taus = taus[a]
        q = g.q_values(batch["obs"], tau_sel=batch["tau_sel"], pos=batch.get("pos"), canvas=batch).cpu()
    else:
        q = g.q_values(batch["obs"], pos=batch.get("pos"), canvas=batch).cpu()
    ins = [oracle_inputs(fix, b) for b in batches]
    for a in range(G):
        want_q = oracles[a].q_values_iqn(ins[a][0], taus[a][2]) if base(kind) in IQN_KINDS else oracles[a].q_values(ins[a][0])
        assert rel_err(q[a], want_q) < RTOL
    # gradients of the first update, then two full updates
    loss = g.grads(batch).cpu()
    grads = g.params(2).cpu()
    for a in range(G):
        o, b = oracles[a], batches[a]
        o.opt.zero_grad()
        want = o.loss(ins[a][0], ins[a][1], b["actions"], b["rewards"], b["dones"])
        want.backward()
        assert abs(float(loss[a]) - float(want.detach())) <= RTOL * abs(float(want.detach()))
        for name, off, cnt in g.table:
            got = grads[a, off:off + cnt]
            if o.sd[name].grad is None:  # the merged models' unused parent parameters (customs/conv_mlp.py:71)
                assert float(got.abs().max()) == 0.0, name
                continue
            ref = o.sd[name].grad.reshape(-1)
            # gradients: summation order differs from torch's, and a ReLU unit within rounding of zero can flip;
            # bound the tensor's relative L2 error and the worst element loosely
            assert float((got - ref).norm()) <= 1e-4 * float(ref.norm()) + 1e-9, (a, name)
            assert float((got - ref).abs().max()) <= 1e-3 * float(ref.abs().max()) + 1e-8, (a, name)
    for step in range(2):
        loss = g.learn(batch).cpu()
        for a in range(G):
            b = batches[a]
            want, norm = oracles[a].learn(ins[a][0], ins[a][1], b["actions"], b["rewards"], b["dones"])
            assert abs(float(loss[a]) - float(want)) <= 2e-5 * abs(float(want)), (step, a, float(loss[a]), float(want))
            assert abs(float(g.grad_norm.cpu()[a]) - float(norm)) <= 1e-4 * float(norm)


@pytest.mark.parametrize("B,G", [(32, 2), (256, 3), (640, 2), (2304, 2)])
def test_dqn_mlp_learner_matches_torch_oracle(B, G):
    """BASELINE configs[0] (map=simple, 2 agents, brain=dqn, model=mlp, visible_range=5): MAEXP_LEARNER_DQN_MLP.
    The reference's model=mlp path raises a TypeError, so there is no live fixture: the network is
    core/agents/models/mlp.py:14-34 on the flattened C*V*V observation (hidden 120, configs/model/mlp.yaml) under the
    DQN update of da3_brain.py:69-87, restated in oracle/torch_learner.py (`dqn_mlp`).  Q-values, loss and every gradient
    of G independently initialised networks, then two full updates; the larger B take the acting-batch inference path."""
    from maexp_b200.core.agents.brains import LearnerGroup
    cfg = baseline_config(1).merged_with(batch_size=B)
    C, V = 3, int(cfg.visible_range)
    assert cfg.brain == "dqn" and cfg.model.name == "mlp" and V == 5
    torch.manual_seed(40 + B)
    g = LearnerGroup(cfg, [C, V, V], 4, G, B)
    assert g.kind == "dqn_mlp"
    gen = torch.Generator().manual_seed(B)
    oracles = []
    for a in range(G):
        tg = {k: v + 0.02 * torch.randn(v.shape, generator=gen) for k, v in g.state_dict(a).items()}
        g.load_state_dict(a, tg, which=1)
        o = TorchLearner("dqn_mlp", g.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8)
        o.target = {k: v.clone() for k, v in tg.items()}
        oracles.append(o)
    obs = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    nxt = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    obs[:, :, -1] -= 1
    nxt[:, :, -1] -= 1
    batch = {"obs": pack(g, obs), "next_obs": pack(g, nxt),
             "actions": torch.randint(0, 4, (G, B), generator=gen).to(torch.int32).cuda(),
             "rewards": torch.randint(-1, 2, (G, B), generator=gen).float().cuda(),
             "dones": (torch.rand((G, B), generator=gen) < 0.1).to(torch.uint8).cuda()}
    q = g.q_values(batch["obs"]).cpu()
    for a in range(G):
        assert rel_err(q[a], oracles[a].q_values(obs[a])) < RTOL
    args = lambda a: (obs[a], nxt[a], batch["actions"][a].cpu().long(), batch["rewards"][a].cpu().double(),
                      batch["dones"][a].cpu().bool())
    loss = g.grads(batch).cpu()
    grads = g.params(2).cpu()
    for a in range(G):
        o = oracles[a]
        o.opt.zero_grad()
        want = o.loss(*args(a))
        want.backward()
        assert abs(float(loss[a]) - float(want.detach())) <= RTOL * abs(float(want.detach()))
        for name, off, cnt in g.table:
            got, ref = grads[a, off:off + cnt], o.sd[name].grad.reshape(-1)
            assert float((got - ref).norm()) <= 1e-4 * float(ref.norm()) + 1e-9, (a, name)
    for step in range(2):
        loss = g.learn(batch).cpu()
        for a in range(G):
            want, norm = oracles[a].learn(*args(a))
            assert abs(float(loss[a]) - float(want)) <= 2e-5 * abs(float(want)), (step, a)
            assert abs(float(g.grad_norm.cpu()[a]) - float(norm)) <= 1e-4 * float(norm)
    for a in range(G):  # RMSprop's first steps amplify last-bit differences of near-zero gradients: count outliers
        got, bad, total = g.state_dict(a), 0, 0
        for k, v in oracles[a].sd.items():
            v = v.detach()
            bad += int(((got[k] - v).abs() > 1e-4 + 1e-4 * v.abs()).sum())
            total += v.numel()
        assert bad <= total * 1e-3, (a, bad, total)


def test_apply_scale_and_inactive_agents():
    """grad_scale (1/world after an all-reduce) and the agent_tasks "-1" mask of abstract_trainer.py:49-51."""
    kind, B, G = "dqn_conv_mlp", 32, 2
    fix = load(kind)
    b = fix["batch"]
    g = make_group(kind, list(b["obs"].shape[1:]), G, B)
    for a in range(G):
        g.load_state_dict(a, fix["init"])
    g.synchronize()
    batch = {"obs": pack(g, torch.stack([b["obs"]] * G)), "next_obs": pack(g, torch.stack([b["next_obs"]] * G)),
             "actions": torch.stack([b["actions"]] * G).to(torch.int32).cuda(),
             "rewards": torch.stack([b["rewards"]] * G).float().cuda(),
             "dones": torch.stack([b["dones"]] * G).to(torch.uint8).cuda()}
    before = g.params().clone()
    g.grads(batch)
    g.params(2).mul_(4.0)  # as if four ranks had summed identical gradients
    g.apply(grad_scale=0.25, active=torch.tensor([1, 0], dtype=torch.uint8, device="cuda"))
    after = g.params()
    assert torch.equal(after[1], before[1])  # inactive agent untouched
    g2 = make_group(kind, list(b["obs"].shape[1:]), 1, B)
    g2.load_state_dict(0, fix["init"])
    g2.synchronize()
    g2.learn({k: v[:1].contiguous() for k, v in batch.items()})
    assert torch.allclose(after[0], g2.params()[0], rtol=0, atol=1e-7)


def test_replay_ring_sampling():
    from maexp_b200.core.utils.buffer import DeviceReplayRing
    from tests.cuda_backend import make_env
    cfg = baseline_config(5).merged_with(max_episode_length=7)
    E, N, S, B = 48, 8, 16, 64
    env = make_env(cfg, E, auto_reset=True)
    assert DeviceReplayRing(env, 2).positions is None  # local view: the models never read the coordinates
    ring = DeviceReplayRing(env, S, keep_positions=True)
    env.reset(ring.next_outputs(), tick=0)
    ring.commit()
    for t in range(1, 23):  # wraps the ring, crosses two auto-resets
        env.step(ring.next_outputs(), tick=t)
        ring.commit()
    assert ring.count == S and len(ring) == (S - 1) * E
    batch = ring.new_batch(B)
    # (1) injected picks: exact gather
    rng = np.random.default_rng(0)
    picks = np.stack([(ring.head - 1 - rng.integers(0, S - 1, B)) % S, rng.integers(0, E, B)], 1).astype(np.int32)
    ring.sample(batch, indices=torch.from_numpy(picks).cuda())
    torch.cuda.synchronize()
    obs, act, rew, done = ring.obs.cpu(), ring.actions.cpu(), ring.rewards.cpu(), ring.dones.cpu()
    positions = ring.positions.cpu()
    for j, (s, e) in enumerate(picks):
        n = (s + 1) % S
        assert torch.equal(batch["obs"][:, j].cpu(), obs[s, :, e])
        assert torch.equal(batch["next_obs"][:, j].cpu(), obs[n, :, e])
        assert torch.equal(batch["actions"][:, j].cpu(), act[n, e].to(torch.int32).clamp(min=0))  # -1 (reset slot) is flagged and clamped
        assert torch.equal(batch["rewards"][:, j].cpu(), rew[n, e]) and torch.equal(batch["dones"][:, j].cpu(), done[n, e])
        assert torch.equal(batch["pos"][:, j].cpu(), positions[s, e]) and torch.equal(batch["next_pos"][:, j].cpu(), positions[n, e])
    # (2) Philox picks: in range, never a transition into a reset slot, reproducible, tick-dependent
    ring.sample(batch, seed=921, tick=5)
    p1 = batch["picked"].cpu().clone()
    was_reset = ring.was_reset.cpu()
    ages = (ring.head - 1 - p1[:, 0]) % S
    assert bool(((ages >= 0) & (ages <= S - 2)).all()) and bool(((p1[:, 1] >= 0) & (p1[:, 1] < E)).all())
    assert int(was_reset[(p1[:, 0] + 1) % S, p1[:, 1]].sum()) == 0
    assert bool((batch["actions"].cpu() >= 0).all())  # actions of reset slots are -1 and must never be sampled
    ring.sample(batch, seed=921, tick=5)
    assert torch.equal(batch["picked"].cpu(), p1)
    ring.sample(batch, seed=921, tick=6)
    assert not torch.equal(batch["picked"].cpu(), p1)
    # (3) without replacement, as np.random.choice(len, B, replace=False) (buffer.py:61): 64 draws from <= 720 transitions repeat
    # one in ~94 % of the batches when drawn with replacement; every batch here is distinct, and the picks stay uniform
    seen = torch.zeros((S, E), dtype=torch.int64)
    batch["error_flag"].zero_()  # (the injected picks of (1) included transitions into reset slots)
    for tick in range(200):
        ring.sample(batch, seed=4, tick=tick)
        p = batch["picked"].cpu().long()
        assert len({(int(a), int(b)) for a, b in p}) == B, tick
        seen[p[:, 0], p[:, 1]] += 1
    ring.check(batch)
    valid = (was_reset[(torch.arange(S) + 1) % S] == 0)
    valid[ring.head] = False  # the newest slot has no successor yet
    assert int(seen[~valid].sum()) == 0
    freq = seen[valid].double() / (200 * B / int(valid.sum()))
    assert float(freq.min()) > 0.15 and float(freq.max()) < 2.5 and abs(float(freq.mean()) - 1.0) < 1e-9


def test_replay_sampler_flags_a_ring_without_valid_transitions():
    """count == 2 with every environment reset by the newest slot: no stored step is a transition.  The sampler must say so
    (error flag, DeviceReplayRing.check raises) and must not hand the loss kernels the reset slot's action -1."""
    from maexp_b200 import _lib
    from maexp_b200.core.utils.buffer import DeviceReplayRing
    from tests.cuda_backend import make_env
    cfg = baseline_config(5).merged_with(max_episode_length=3)
    E, B = 32, 16
    env = make_env(cfg, E, auto_reset=True)
    ring = DeviceReplayRing(env, 2)
    env.reset(ring.next_outputs(), tick=0)
    ring.commit()
    for t in range(1, 4):  # steps 0..2: done at step L-1 = 2 ...
        env.step(ring.next_outputs(), tick=t)
        ring.commit()
    env.step(ring.next_outputs(), tick=4)  # ... so this call resets every environment in lock-step
    ring.commit()
    assert ring.count == 2 and bool(ring.was_reset[ring.head].all())
    batch = ring.new_batch(B)
    ring.sample(batch, seed=921, tick=0)
    torch.cuda.synchronize()
    assert int(batch["error_flag"].item()) == _lib.REPLAY_ERR_NO_VALID
    assert bool((batch["actions"] >= 0).all())
    with pytest.raises(_lib.MaexpError):
        ring.check(batch)
    ring.check(batch)  # the flag was cleared by the raise
    # one more step: the transition reset-slot -> step is valid again, found by the redraws or by the fixed walk
    env.step(ring.next_outputs(), tick=5)
    ring.commit()
    ring.sample(batch, seed=921, tick=1)
    ring.check(batch)
    # more samples than distinct transitions (32 here): np.random.choice(replace=False) raises, the sampler flags the batch
    big = ring.new_batch(48)
    ring.sample(big, seed=921, tick=2)
    torch.cuda.synchronize()
    assert int(big["error_flag"].item()) == _lib.REPLAY_ERR_NOT_UNIQUE
    with pytest.raises(_lib.MaexpError):
        ring.check(big)
    # injected picks: out of range (clamped, flagged) and a pick into a reset slot (flagged, action clamped to 0)
    bad = torch.tensor([[7, 0]] + [[(ring.head - 1) % 2, 1]] * (B - 1), dtype=torch.int32).cuda()
    ring.sample(batch, indices=bad)
    torch.cuda.synchronize()
    assert int(batch["error_flag"].item()) & _lib.REPLAY_ERR_INDEX_RANGE
    with pytest.raises(_lib.MaexpError):
        ring.check(batch)
    env2 = make_env(cfg, E, auto_reset=True)
    ring2 = DeviceReplayRing(env2, 2)
    env2.reset(ring2.next_outputs(), tick=0)
    ring2.commit()
    env2.reset(ring2.next_outputs(), tick=1)
    ring2.commit()
    into_reset = torch.tensor([[(ring2.head - 1) % 2, 3]] * B, dtype=torch.int32).cuda()
    ring2.sample(batch, indices=into_reset)
    torch.cuda.synchronize()
    assert int(batch["error_flag"].item()) == _lib.REPLAY_ERR_INDEX_INVALID
    assert bool((batch["actions"] == 0).all())  # the reset slot's -1 never reaches the loss kernels


def test_reference_style_brain_interface():
    """generate_brain(...).learn / get_action with the reference's argument structure (E=1 style)."""
    from maexp_b200.core.agents.brains import generate_brain
    fix = load("c51")
    b = fix["batch"]
    cfg, shape = golden_configs()["c51"]
    brain = generate_brain(cfg, shape, 4)
    brain.network.load_state_dict(fix["init"])
    brain.target_network.load_state_dict(fix["target"])
    assert brain.get_action({"local": b["obs"][0].clone()}) == int(fix["q0"][0].argmax())
    out = brain.learn({"local": b["obs"].clone()}, b["actions"].clone(), b["rewards"].clone(), b["dones"].clone(),
                      {"local": b["next_obs"].clone()})
    assert abs(float(out["total_loss"]) - fix["losses"][0]) <= RTOL * fix["losses"][0]
    with pytest.raises(ValueError):
        generate_brain(cfg.merged_with(brain="nope"), shape, 4)


def test_attention_maps_match_torch_oracle():
    """forward_attn of DA6 (da6.py:128-165): local and relative attention weights of the acting forward pass."""
    from oracle.torch_learner import da6_forward
    kind = "da6"
    fix = load(kind)
    b = fix["batch"]
    B = b["obs"].shape[0]
    g = make_group(kind, fix["shape"], 1, B)
    g.load_state_dict(0, fix["init"])
    pos = b["pos"].to(torch.int8)[None].contiguous().cuda()
    q = g.q_values(pack(g, b["obs"][None]), pos=pos).cpu()[0]
    attns = [[], []]
    want_q = da6_forward(fix["init"], (b["obs"], b["pos"]), fix["shape"][1:], 2, attns=attns)
    assert rel_err(q, want_q) < RTOL
    for stack in (0, 1):
        got = g.attention(B, stack=stack, layer=0).cpu()[0]
        assert got.shape == attns[stack][0].shape
        assert float((got - attns[stack][0]).abs().max()) < 1e-6
    with pytest.raises(Exception):
        g.attention(B, stack=0, layer=3)
    dqn = make_group("dqn_conv_mlp", load("dqn_conv_mlp")["shape"], 1, 4)
    with pytest.raises(ValueError):
        dqn.attention(4)


def test_hard_shrink_attention_maps_match_torch_oracle():
    """HardShrinkAttention.forward_attn (hard_shrink_attention.py:82-105) returns the weights AFTER nn.Hardshrink:
    the exported maps of a model.attention == "hard" DA6 hold exact zeros where softmax <= af_lambd."""
    import oracle.torch_learner as tl
    kind = "da6_hard"
    fix = load(kind)
    b = fix["batch"]
    B = b["obs"].shape[0]
    g = make_group(kind, fix["shape"], 1, B)
    g.load_state_dict(0, fix["init"])
    pos = b["pos"].to(torch.int8)[None].contiguous().cuda()
    q = g.q_values(pack(g, b["obs"][None]), pos=pos).cpu()[0]
    assert rel_err(q, fix["q0"]) < RTOL
    attns = [[], []]
    tl._ATTN_SHRINK = fix["attn_shrink"]
    try:
        tl.da6_forward(fix["init"], (b["obs"], b["pos"]), fix["shape"][1:], 2, attns=attns)
    finally:
        tl._ATTN_SHRINK = None
    for stack in (0, 1):
        got, want = g.attention(B, stack=stack, layer=0).cpu()[0], attns[stack][0]
        cut = want == 0
        assert 0.005 < float(cut.float().mean()) < 0.95  # the threshold does cut weights (66 % local, 0.8 % relative stack)
        # a weight within rounding of af_lambd may fall on the other side; everything else agrees
        near = (want - fix["attn_shrink"]).abs() < 1e-6
        assert float(((got - want).abs() * (~near)).max()) < 1e-6
        assert bool(((got == 0) == cut)[~near].all())


def test_adam_optimizer_matches_torch():
    """configs/opt/adam.yaml through generate_optimizer (optimizers/__init__.py:18-24): five updates against torch.optim.Adam,
    with one agent masked out (its step counter and moments must not advance)."""
    from maexp_b200.core.agents.brains import LearnerGroup
    kind, B, G = "dqn_conv_mlp", 32, 2
    fix = load(kind)
    b = fix["batch"]
    cfg, _ = golden_configs()[kind]
    cfg = cfg.merged_with(batch_size=B, opt=Config(dict(name="adam", learning_rate=1e-3, betas=[0.9, 0.999], eps=1e-8)))
    g = LearnerGroup(cfg, fix["shape"], 4, G, B)
    for a in range(G):
        g.load_state_dict(a, fix["init"])
    g.synchronize()
    o = TorchLearner(kind, fix["init"], 4, fix["gamma"], 1e-3, 0.99, 1e-8, adam_betas=(0.9, 0.999))
    batch = {"obs": pack(g, torch.stack([b["obs"]] * G)), "next_obs": pack(g, torch.stack([b["next_obs"]] * G)),
             "actions": torch.stack([b["actions"]] * G).to(torch.int32).cuda(),
             "rewards": torch.stack([b["rewards"]] * G).float().cuda(),
             "dones": torch.stack([b["dones"]] * G).to(torch.uint8).cuda()}
    before = g.params().clone()
    active = torch.tensor([1, 0], dtype=torch.uint8, device="cuda")
    for step in range(5):
        loss = g.learn(batch, active=active).cpu()
        want, _ = o.learn(b["obs"], b["next_obs"], b["actions"], b["rewards"], b["dones"])
        assert abs(float(loss[0]) - float(want)) <= 5e-5 * abs(float(want)), (step, float(loss[0]), float(want))
    got = g.state_dict(0)
    for k, v in o.sd.items():
        assert float((got[k] - v.detach()).abs().max()) <= 2e-5 + 1e-4 * float(v.detach().abs().max()), k
    assert torch.equal(g.params()[1], before[1])  # the masked agent did not move
    with pytest.raises(ValueError):
        LearnerGroup(cfg.merged_with(opt=Config(dict(name="sgd", learning_rate=1e-3))), fix["shape"], 4, 1, B)


@pytest.mark.parametrize("kind,V,C,B", [("dqn_conv_mlp", 7, 7, 640), ("c51", 7, 3, 640), ("iqn", 7, 3, 640), ("c51", 5, 3, 640),
                                        ("qr", 9, 4, 640), ("c51", 7, 3, 2304), ("dqn_conv_mlp", 7, 7, 2100), ("da3", 7, 3, 2100), ("da3", 7, 3, 48), ("da3", 9, 3, 300),
                                        # tc_trunk_kernel: a partial last tile of 32 observations, and the generic-channel instantiation
                                        ("c51", 7, 3, 2100), ("c51", 7, 5, 2304), ("qr", 7, 3, 2304)])
def test_large_batch_inference_path_matches_oracle(kind, V, C, B):
    """Acting / evaluation batches: from 512 rows the MLP heads run on the TMA-fed tcgen05 kernel (csrc/tc_heads.cu, with the
    fused dueling / expectation epilogue for C51; per-layer tcgen05 GEMMs where it does not apply, e.g. QR-DQN's 400-wide head),
    from 2048 rows the conv trunk runs conv2 on the tensor cores (csrc/tc_trunk.cu); three TF32 terms per product: the
    Q-values stay within 1e-5 of the fp32 oracle."""
    from maexp_b200.core.agents.brains import LearnerGroup
    cfg, _ = golden_configs()[kind]
    G = 2
    cfg = cfg.merged_with(batch_size=B, visible_range=V)
    torch.manual_seed(11)  # fixed initialisation: the three-term TF32 error (3-6e-6 over random nets, tools/exp_tc_accuracy.py) is then the same every run
    g = LearnerGroup(cfg, [C, V, V], 4, G, B)
    gen = torch.Generator().manual_seed(3)
    obs = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    obs[:, :, -1] -= 1
    packed = pack(g, obs)
    tau = torch.rand((G, 4), generator=gen)
    q = g.q_values(packed, tau_sel=tau.cuda() if kind == "iqn" else None).cpu()
    for a in range(G):
        o = TorchLearner(kind, g.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8, num_atoms=ATOMS[kind])
        want = o.q_values_iqn(obs[a], tau[a:a + 1]) if kind == "iqn" else o.q_values(obs[a])
        assert rel_err(q[a], want) < RTOL, (kind, a, rel_err(q[a], want))
    # a smaller batch through the same handle takes the layer-by-layer path and must agree
    q_small = g.q_values(packed[:, :100].contiguous(), tau_sel=tau.cuda() if kind == "iqn" else None).cpu()
    assert float((q_small - q[:, :100]).abs().max()) <= 1e-5 * float(q.abs().max())


@pytest.mark.parametrize("B", [300, 2100])
def test_hard_shrink_large_batch_inference_path(B):
    """The acting-batch attention kernel (scores in shared memory, no probability buffer) with model.attention == "hard":
    same Q-values as the training-size kernel on the same rows, and the torch oracle's on all but the rare rows where a
    softmax weight lies within rounding of af_lambd (the shrink is discontinuous there)."""
    from maexp_b200.core.agents.brains import LearnerGroup
    cfg, _ = golden_configs()["da3_hard"]
    G, V, C = 2, 7, 3
    cfg = cfg.merged_with(batch_size=B)
    g = LearnerGroup(cfg, [C, V, V], 4, G, B)
    gen = torch.Generator().manual_seed(5)
    for a in range(G):  # the embeddings are zero-initialised (da3.py:41-44): perturb them so the weights spread out
        sd = g.state_dict(a)
        for k in ("saliency_vector", "pos_embed"):
            sd[k] = sd[k] + 0.5 * torch.randn(sd[k].shape, generator=gen)
        g.load_state_dict(a, sd)
    obs = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    obs[:, :, -1] -= 1
    packed = pack(g, obs)
    q = g.q_values(packed).cpu()
    for a in range(G):
        o = TorchLearner("da3", g.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8, attn_shrink=float(cfg.model["af_lambd"]))
        want = o.q_values(obs[a])
        soft = TorchLearner("da3", g.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8).q_values(obs[a])
        err = (q[a] - want).abs().max(dim=1)[0] / want.abs().max()
        assert float((err < RTOL).float().mean()) > 0.99, float((err < RTOL).float().mean())
        assert rel_err(soft, want) > 1e-3  # the shrink does change the result
    q_small = g.q_values(packed[:, :100].contiguous()).cpu()
    assert float((q_small - q[:, :100]).abs().max()) <= 1e-5 * float(q.abs().max())


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 64, 72), (4096, 128, 256), (4096, 51, 128), (2500, 204, 128), (1000, 32, 96),
                                   (513, 256, 40), (640, 700, 64), (300, 257, 32)])
def test_tcgen05_gemm_against_fp64(M, N, K):
    """maexp_tc_gemm (tcgen05.mma kind::tf32): three TF32 terms per product stay within 1e-5 of the exact product sum,
    single-pass TF32 within 2e-3; ragged M / N / K, misaligned weight rows, bias and ReLU."""
    import ctypes as C
    from maexp_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((M, K), generator=gen).cuda()
    flat = torch.randn((N * K + 3,), generator=gen).cuda() * 0.1
    W = flat[1:1 + N * K].view(N, K) if K % 4 else flat[:N * K].view(N, K)   # odd offsets exercise the scalar weight loads
    bias = torch.randn((N,), generator=gen).cuda()
    want = torch.relu(A.double() @ W.double().T + bias.double())
    for split, tol in ((3, 1e-5), (1, 2e-3)):
        out = torch.full((M, N), float("nan"), device="cuda")
        _lib.check(lib.maexp_tc_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(bias.data_ptr()),
                                     C.c_void_p(out.data_ptr()), M, N, K, K, K, N, 1, split, None))
        torch.cuda.synchronize()
        assert not torch.isnan(out).any()
        assert float((out.double() - want).abs().max() / want.abs().max()) < tol, (split,)
    with pytest.raises(Exception):
        _lib.check(lib.maexp_tc_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()), None, C.c_void_p(out.data_ptr()),
                                     M, N, K, K, K, N, 0, 2, None))  # split must be 1 or 3



@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (204, 128, 256), (51, 128, 300), (64, 128, 2304), (128, 576, 1100), (37, 70, 64)])
def test_tcgen05_weight_gradient_gemm_against_fp64(M, N, K):
    """maexp_tc_gemm_tn (C = A^T B with the batch as K, operands transposed while staging, three TF32 terms): within 1e-5 of the
    exact product sum; ragged M / N / K, rows that are not 16-byte aligned (lda = 51, 37), split-K partial tiles for K >= 1024."""
    import ctypes as C
    from maexp_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M), generator=gen).cuda()
    B = torch.randn((K, N), generator=gen).cuda() * 0.1
    want = A.double().T @ B.double()
    out = torch.full((M, N), float("nan"), device="cuda")
    part = torch.full((8 * M * N,), float("nan"), device="cuda")
    ks = C.c_int32(0)
    _lib.check(lib.maexp_tc_gemm_tn(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(out.data_ptr()),
                                    C.c_void_p(part.data_ptr()), part.numel(), M, N, K, M, N, N, C.byref(ks), None))
    torch.cuda.synchronize()
    assert (ks.value > 1) == (K >= 1024)
    got = out if ks.value == 1 else part[: ks.value * M * N].view(ks.value, M, N).sum(0)
    assert not torch.isnan(got).any()
    assert float((got.double() - want).abs().max() / want.abs().max()) < 1e-5
    with pytest.raises(Exception):
        _lib.check(lib.maexp_tc_gemm_tn(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(out.data_ptr()), None, 0,
                                        8, N, K, M, N, N, None, None))  # M < 16


def test_attention_agent_returns_reference_structures():
    """generate_agents(agent_type=attention): get_action_attns -> (action, [stacked maps per group]) as attention_agent.py:14-22;
    DA6 gives [local, relative], the da3_baseline brain the zero placeholder of da3_brain.py:107-123."""
    from maexp_b200.core.agents import generate_agents
    from oracle.make_golden_learner import ref_state
    fix = load("da6")
    cfg, shape = golden_configs()["da6"]
    agents = generate_agents(cfg.merged_with(agent_type="attention"), [shape], [4])
    assert len(agents) == 1
    agent = agents[0]
    agent.brain.network.load_state_dict(fix["init"])
    b = fix["batch"]
    st = ref_state(b["obs"][:1], b["pos"][:1])
    state = {"local": st["local"][0], "relative": st["relative"][0], "coordinates": {k: int(v[0]) for k, v in st["coordinates"].items()}}
    action, attns = agent.get_action_attns(state, 0.0)
    assert action == int(fix["q0"][0].argmax())
    heads, blocks = int(cfg["model"]["num_heads"]), int(cfg["model"]["block_loop"])
    assert len(attns) == 2 and attns[0].shape == (blocks, 1, heads, 26, 26) and attns[1].shape == (blocks, 1, heads, 13, 13)
    assert float((attns[0].sum(-1) - 1).abs().max()) < 1e-5  # softmax rows
    assert isinstance(agent.get_action(state, 0.0), int)
    base_cfg, base_shape = golden_configs()["dqn_conv_mlp"]
    base = generate_agents(base_cfg.merged_with(agent_type="attention"), [base_shape], [4])[0]
    a, placeholder = base.get_action_attns({"local": load("dqn_conv_mlp")["batch"]["obs"][0].clone()}, 0.0)
    assert isinstance(a, int) and float(placeholder[0].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        generate_agents(cfg.merged_with(agent_type="nope"), [shape], [4])


@pytest.mark.parametrize("kind,V,C,B", [("c51", 7, 3, 256), ("c51", 5, 3, 37), ("qr", 9, 4, 64), ("dqn_conv_mlp", 7, 11, 32),
                                        ("iqn", 7, 3, 50), ("dqn_conv_mlp", 9, 7, 300), ("c51", 7, 3, 1100)])
def test_fused_trunk_matches_layer_by_layer_and_torch(kind, V, C, B, monkeypatch):
    """The fused kernels of the update (conv trunk forward keeping pooled map + argmax, its backward through conv2 / pooling
    into conv1's gradients, one-kernel MLP heads) and the layer-by-layer path (MAEXP_FUSED_TRUNK=0, MAEXP_FUSED_MLP=0), both
    against torch autograd: loss and every gradient."""
    from maexp_b200.core.agents.brains import LearnerGroup
    cfg, _ = golden_configs()[kind]
    G = 2
    cfg = cfg.merged_with(batch_size=B, visible_range=V)
    torch.manual_seed(1000 + B)  # the networks' default initialisation: fixed, so that a run is reproducible (a sample whose
    gen = torch.Generator().manual_seed(5)  # TD error or greedy action sits within rounding of a kink would differ from CPU torch)
    obs = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    nxt = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    obs[:, :, -1] -= 1
    nxt[:, :, -1] -= 1
    groups = []
    for fused in ("1", "0"):
        monkeypatch.setenv("MAEXP_FUSED_TRUNK", fused)
        monkeypatch.setenv("MAEXP_FUSED_MLP", fused)
        groups.append(LearnerGroup(cfg, [C, V, V], 4, G, B))
    g, ref = groups
    for a in range(G):
        ref.load_state_dict(a, g.state_dict(a))
        ref.load_state_dict(a, g.state_dict(a, which=1), which=1)
    batch = {"obs": pack(g, obs), "next_obs": pack(g, nxt),
             "actions": torch.randint(0, 4, (G, B), generator=gen).to(torch.int32).cuda(),
             "rewards": torch.randint(-1, 2, (G, B), generator=gen).float().cuda(),
             "dones": (torch.rand((G, B), generator=gen) < 0.1).to(torch.uint8).cuda()}
    taus = None
    if kind == "iqn":
        taus = (torch.rand((G, B, 4), generator=gen), torch.rand((G, B, 4), generator=gen), torch.rand((G, 4), generator=gen))
        batch.update(taus=taus[0].cuda(), tau_dashes=taus[1].cuda(), tau_sel=taus[2].cuda())
    loss, loss_ref = g.grads(batch).cpu(), ref.grads(batch).cpu()
    grads, grads_ref = g.params(2).cpu(), ref.params(2).cpu()
    assert float((loss - loss_ref).abs().max()) <= 1e-5 * float(loss_ref.abs().max())
    for a in range(G):
        o = TorchLearner(kind, g.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8, num_atoms=ATOMS[kind])
        o.target = {k: v.clone() for k, v in g.state_dict(a, which=1).items()}
        if kind == "iqn":
            o.taus = (taus[0][a], taus[1][a], taus[2][a:a + 1])
        o.opt.zero_grad()
        want = o.loss(obs[a], nxt[a], batch["actions"][a].cpu().long(), batch["rewards"][a].cpu().double(), batch["dones"][a].cpu().bool())
        want.backward()
        assert abs(float(loss[a]) - float(want.detach())) <= 2e-5 * abs(float(want.detach()))
        for name, off, cnt in g.table:
            refg = o.sd[name].grad.reshape(-1)
            e_fused = float((grads[a, off:off + cnt] - refg).norm()) / (float(refg.norm()) + 1e-12)
            e_layer = float((grads_ref[a, off:off + cnt] - refg).norm()) / (float(refg.norm()) + 1e-12)
            assert e_fused <= 1e-4 and e_layer <= 1e-4, (a, name, e_fused, e_layer)


@pytest.mark.parametrize("kind,V,C,B", [("c51", 7, 3, 256), ("dqn_conv_mlp", 7, 7, 32), ("qr", 5, 3, 70)])
def test_batched_forward_passes_match_streamed(kind, V, C, B, monkeypatch):
    """MAEXP_BATCHED_FWD=1 (all forward passes of an update as grid slices of one launch per fused kernel) computes the
    same loss and gradients as the default one-stream-per-pass issue order (same kernels, same summation order: bit-equal
    up to the C51 projection's atomics)."""
    from maexp_b200.core.agents.brains import LearnerGroup
    cfg, _ = golden_configs()[kind]
    G = 2
    cfg = cfg.merged_with(batch_size=B, visible_range=V)
    torch.manual_seed(77)
    gen = torch.Generator().manual_seed(6)
    obs = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    nxt = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    groups = []
    for batched in ("1", "0"):
        monkeypatch.setenv("MAEXP_BATCHED_FWD", batched)
        groups.append(LearnerGroup(cfg, [C, V, V], 4, G, B))
    g, ref = groups
    for a in range(G):
        ref.load_state_dict(a, g.state_dict(a))
        ref.load_state_dict(a, g.state_dict(a, which=1), which=1)
    batch = {"obs": pack(g, obs), "next_obs": pack(g, nxt),
             "actions": torch.randint(0, 4, (G, B), generator=gen).to(torch.int32).cuda(),
             "rewards": torch.randint(-1, 2, (G, B), generator=gen).float().cuda(),
             "dones": (torch.rand((G, B), generator=gen) < 0.1).to(torch.uint8).cuda()}
    loss, loss_ref = g.grads(batch).cpu(), ref.grads(batch).cpu()
    grads, grads_ref = g.params(2).cpu(), ref.params(2).cpu()
    assert float((loss - loss_ref).abs().max()) <= 1e-6 * float(loss_ref.abs().max())
    assert float((grads - grads_ref).norm()) <= 1e-6 * float(grads_ref.norm())



@pytest.mark.parametrize("kind,V,C,B", [("c51", 7, 3, 256), ("dqn_conv_mlp", 7, 3, 300), ("c51", 7, 3, 1100)])
def test_opt_in_backward_kernels_match_default(kind, V, C, B, monkeypatch):
    """The two measured-slower, opt-in forms of the backward pass stay correct: MAEXP_MLP_BWD_SPLIT=1 (dh kernel + dx kernel: the
    same fmaf chains, so the same bits up to the C51 projection's atomics) and MAEXP_TC_WGRAD=1 (weight-gradient GEMMs on tcgen05,
    operands transposed while staging, three TF32 terms, split-K slices reduced in a fixed order: within 2e-5 per tensor)."""
    from maexp_b200.core.agents.brains import LearnerGroup
    cfg, _ = golden_configs()[kind]
    G = 2
    cfg = cfg.merged_with(batch_size=B, visible_range=V)
    torch.manual_seed(31)
    gen = torch.Generator().manual_seed(9)
    obs = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    nxt = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
    groups = []
    for split, tcw in (("0", "0"), ("1", "0"), ("0", "1")):
        monkeypatch.setenv("MAEXP_MLP_BWD_SPLIT", split)
        monkeypatch.setenv("MAEXP_TC_WGRAD", tcw)
        groups.append(LearnerGroup(cfg, [C, V, V], 4, G, B))
    ref = groups[0]
    for g in groups[1:]:
        for a in range(G):
            g.load_state_dict(a, ref.state_dict(a))
            g.load_state_dict(a, ref.state_dict(a, which=1), which=1)
    batch = {"obs": pack(ref, obs), "next_obs": pack(ref, nxt),
             "actions": torch.randint(0, 4, (G, B), generator=gen).to(torch.int32).cuda(),
             "rewards": torch.randint(-1, 2, (G, B), generator=gen).float().cuda(),
             "dones": (torch.rand((G, B), generator=gen) < 0.1).to(torch.uint8).cuda()}
    ref.grads(batch)
    want = ref.params(2).cpu()
    for g, tol in zip(groups[1:], (1e-6, 2e-5)):
        g.grads(batch)
        got = g.params(2).cpu()
        for a in range(G):
            for name, off, cnt in g.table:
                e = float((got[a, off:off + cnt] - want[a, off:off + cnt]).norm()) / (float(want[a, off:off + cnt].norm()) + 1e-12)
                assert e <= tol, (a, name, e, tol)


def test_graphed_update_equals_eager_updates():
    """GraphedUpdate (eager first step, then captured CUDA graphs) walks the parameters exactly like
    LearnerGroup.learn on a twin group (QR-DQN: no atomics anywhere, so bit-equal)."""
    from maexp_b200.core.agents.brains import GraphedUpdate, LearnerGroup
    kind, V, C, B, G = "qr", 7, 3, 64, 3
    cfg, _ = golden_configs()[kind]
    cfg = cfg.merged_with(batch_size=B, visible_range=V)
    torch.manual_seed(5)
    g, ref = LearnerGroup(cfg, [C, V, V], 4, G, B), LearnerGroup(cfg, [C, V, V], 4, G, B)
    for a in range(G):
        ref.load_state_dict(a, g.state_dict(a))
        ref.load_state_dict(a, g.state_dict(a, which=1), which=1)
    gen = torch.Generator().manual_seed(8)
    batch = {"obs": torch.zeros((G, B, g.obs_stride), device="cuda"), "next_obs": torch.zeros((G, B, g.obs_stride), device="cuda"),
             "actions": torch.zeros((G, B), dtype=torch.int32, device="cuda"), "rewards": torch.zeros((G, B), device="cuda"),
             "dones": torch.zeros((G, B), dtype=torch.uint8, device="cuda")}
    graphed = GraphedUpdate(g, batch)
    for step in range(5):
        obs = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
        nxt = (torch.rand((G, B, C, V, V), generator=gen) < 0.25).float()
        batch["obs"].copy_(pack(g, obs))  # the replay sampler refills the same buffers in the product
        batch["next_obs"].copy_(pack(g, nxt))
        batch["actions"].copy_(torch.randint(0, 4, (G, B), generator=gen).to(torch.int32))
        batch["rewards"].copy_(torch.randint(-1, 2, (G, B), generator=gen).float())
        loss = graphed.step().cpu().clone()
        want = ref.learn(batch).cpu()
        assert torch.equal(loss, want), step
        assert torch.equal(g.params().cpu(), ref.params().cpu()), step
    assert graphed.g_grads is not None
