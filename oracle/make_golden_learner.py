"""Writes tests/golden/learner_<brain>.npz from the LIVE reference brains (build container only).

    python -m oracle.make_golden_learner

For each brain: initial state_dict, a synthetic batch, the loss `brain.learn` returned, the
network's Q-values before the update and the online weights after 2 updates.  TEST INFRASTRUCTURE.
"""
import os
import sys

import numpy as np
import torch

from maexp_b200.config import baseline_config, make_config
from oracle import ref_runner as rr

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
B = 32
SIZE = (12, 10)  # map of the relative / merged fixtures
VIEW = 5


def configs():
    c5, c3, c2 = baseline_config(5), baseline_config(3), baseline_config(2)
    # the DQN that runs in shipped experiments is DA3BaselineBrain + ConvMLP; model.num_heads is
    # only read to size a constant attention placeholder (da3_brain.py:107-113)
    dqn = c2.merged_with(brain="da3_baseline", batch_size=B, observation_noise=False)
    dqn["model"] = dict(dqn.model, name="da3_baseline", patch_size=1, num_heads=4)
    iqn = c3.merged_with(batch_size=B, brain="iqn")
    iqn["model"] = dict(iqn.model, name="iqn", num_quantiles=4, num_quantiles_dash=4, num_cosines=4)
    da3 = c5.merged_with(batch_size=B, brain="da3", agent_type="attention")
    da3["model"] = dict(name="da3", embed_dim=64, patch_size=1, block_loop=1, num_heads=4, mlp_ratio=4.0,
                        attention="soft", af_lambd=0.5)
    # relative / merged views on a small rectangular map (the models only read map.SIZE_X / SIZE_Y)
    small = dict(hidden_channels=[8], output_channel=16, hidden_kernels=[2, 2], hidden_strides=[1, 1],
                 conv_afs=["max_pool2d", "max_pool2d"], af_kernels=[2, 2], af_strides=[2, 2], hidden_layer_sizes=[32], fc_afs=["relu"])
    def view(mask, name, brain, **extra):
        c = c2.merged_with(brain=brain, batch_size=B, observation_noise=False, observation_area_mask=mask, visible_range=5)
        c["map"] = dict(c.map, SIZE_X=SIZE[0], SIZE_Y=SIZE[1])
        c["model"] = dict(small, name=name, patch_size=1, num_heads=2, **extra)
        return c
    rel = view("relative", "da3_baseline", "da3_baseline")
    mconv = view("merged", "da3_baseline", "da3_baseline")
    mda3 = view("merged", "da3", "da3", embed_dim=32, block_loop=1, mlp_ratio=4.0, attention="soft", af_lambd=0.5)
    mda3["agent_type"] = "attention"
    da6 = view("merged", "da6", "da6", embed_dim=32, block_loop=1, mlp_ratio=4.0, attention="soft", af_lambd=0.5,
               relative_patch_size=3, local_patch_size=1)
    da6["agent_type"] = "attention"
    da6["objects_channel"], da6["agents_channel"] = False, False
    iq = dict(embed_dim=32, block_loop=1, mlp_ratio=4.0, attention="soft", af_lambd=0.5, num_quantiles=4, num_quantiles_dash=4,
              num_cosines=4, relative_patch_size=5, local_patch_size=1)
    def attn_iqn(mask, name):
        c = view(mask, name, "da3_iqn", **iq)
        c["agent_type"] = "attention"
        c["objects_channel"], c["agents_channel"], c["destination_channel"] = False, False, False
        return c
    rda3 = view("relative", "da3", "da3", embed_dim=32, block_loop=1, mlp_ratio=4.0, attention="soft", af_lambd=0.5)
    rda3["model"]["patch_size"] = 3
    rda3["agent_type"] = "attention"
    rda3_iqn = attn_iqn("relative", "da3_iqn")
    rda3_iqn["model"]["patch_size"] = 3
    def conv_iqn(mask):
        c = view(mask, "iqn", "iqn", num_quantiles=4, num_quantiles_dash=4, num_cosines=4)
        c["objects_channel"], c["agents_channel"] = False, False
        return c
    # HardShrinkBlock variants (model.attention == "hard", hard_shrink_attention.py): af_lambd near the mean attention
    # weight of the 26- and 21-token stacks so that a good share of the weights is cut
    da6_hard = da6.merged_with()
    da6_hard["model"] = dict(da6.model, attention="hard", af_lambd=0.04)
    mda3_iqn_hard = attn_iqn("merged", "da3_iqn")
    mda3_iqn_hard["model"] = dict(mda3_iqn_hard.model, attention="hard", af_lambd=0.04)
    da3_iqn_hard = attn_iqn("local", "da3_iqn")
    da3_iqn_hard["model"] = dict(da3_iqn_hard.model, attention="hard", af_lambd=0.04)
    da3_hard = da3.merged_with()
    da3_hard["model"] = dict(da3.model, attention="hard", af_lambd=0.02)  # 50 tokens: mean weight 0.02
    da6_iqn_hard = attn_iqn("merged", "da6_iqn")
    da6_iqn_hard["model"] = dict(da6_iqn_hard.model, attention="hard", af_lambd=0.04)
    # extra canvas channels (objects_channel / agents_channel / destination_channel; T = 2 object types, 4 agents)
    def channels(c, o, a, d):
        c = c.merged_with(objects_channel=o, agents_channel=a, destination_channel=d, type_objects=2, num_agents=4)
        return c
    ch = {"relative_conv_mlp_ch": (channels(rel, True, True, True), [4, *SIZE]),     # ConvMLP reads objects only
          "relative_da3_ch": (channels(rda3, True, False, False), [4, *SIZE]),
          "relative_iqn_ch": (channels(conv_iqn("relative"), True, True, False), [4, *SIZE]),
          "relative_da3_iqn_ch": (channels(rda3_iqn, True, True, True), [4, *SIZE]),
          "da6_ch": (channels(da6, True, True, False), [4, *SIZE]),
          "da6_iqn_ch": (channels(attn_iqn("merged", "da6_iqn"), True, True, True), [4, *SIZE]),
          "merged_da3_iqn_ch": (channels(attn_iqn("merged", "da3_iqn"), True, True, True), [4, *SIZE])}  # unused parent only
    # merged view under observation noise: the models read the noise-free "relative" key (merged_observtion_handler.py:27)
    noise = {"merged_conv_mlp_noise": (mconv.merged_with(observation_noise="flip"), [4, *SIZE]),
             "merged_da3_noise": (mda3.merged_with(observation_noise="flip"), [4, *SIZE]),
             "da6_noise": (da6.merged_with(observation_noise="sensing_dist"), [4, *SIZE])}
    hard = {**ch, **noise, "da3_hard": (da3_hard, [3, 7, 7]), "da6_iqn_hard": (da6_iqn_hard, [4, *SIZE]), "da6_hard": (da6_hard, [4, *SIZE]), "merged_da3_iqn_hard": (mda3_iqn_hard, [4, *SIZE]),
            "da3_iqn_hard": (da3_iqn_hard, [4, VIEW, VIEW])}
    return {**hard, "relative_iqn": (conv_iqn("relative"), [4, *SIZE]), "merged_iqn": (conv_iqn("merged"), [4, *SIZE]),
            "relative_da3": (rda3, [4, *SIZE]), "relative_da3_iqn": (rda3_iqn, [4, *SIZE]),
            "da3_iqn": (attn_iqn("local", "da3_iqn"), [4, VIEW, VIEW]), "merged_da3_iqn": (attn_iqn("merged", "da3_iqn"), [4, *SIZE]),
            "da6_iqn": (attn_iqn("merged", "da6_iqn"), [4, *SIZE]),
            "da6": (da6, [4, *SIZE]),"relative_conv_mlp": (rel, [4, *SIZE]), "merged_conv_mlp": (mconv, [4, *SIZE]), "merged_da3": (mda3, [4, *SIZE]),
            "da3": (da3, [3, 7, 7]), "dqn_conv_mlp": (dqn, [7, 7, 7]), "c51": (c5.merged_with(batch_size=B), [3, 7, 7]),
            "qr": (c3.merged_with(batch_size=B), [3, 7, 7]), "iqn": (iqn, [3, 7, 7])}


def synthetic_batch(shape, seed, V=None):
    g = torch.Generator().manual_seed(seed)
    C = shape[0]
    V = V or shape[1]
    def obs():
        x = (torch.rand((B, C, V, V), generator=g) < 0.15).float()
        x[:, -1] = -(torch.rand((B, V, V), generator=g) < 0.3).float()
        return x
    return {"obs": obs(), "next_obs": obs(), "actions": torch.randint(0, 4, (B,), generator=g),
            "rewards": torch.tensor(np.random.default_rng(seed).choice([-1.0, 0.0, 1.0], size=B, p=[0.2, 0.6, 0.2])),
            "dones": torch.rand((B,), generator=g) < 0.1}


def ref_state(obs, pos, extras=None):
    """The dict the reference's relative / merged observation handlers hand to a brain (observation_handler.py:95-105);
    extras: the state's "objects" / "agents" / "destination" keys and, under observation noise, the noise-free
    "relative" crop (merged_observtion_handler.py:41-47)."""
    from oracle.torch_learner import coordinates
    c = coordinates(pos, obs.shape[2], *SIZE)
    st = {"local": obs.clone(), "relative": obs.clone(), "coordinates": {k: torch.tensor(v) for k, v in c.items()}}
    for k, v in (extras or {}).items():
        st[k] = v.clone()
    return st


def synthetic_extras(cfg, fixture, batch, seed):
    """Synthetic "objects" int8 [B,T,X,Y], "agents" int8 [B,N,2], "destination" int8 [B,X,Y] for the state and the new state,
    and for the *_noise fixtures a random line-of-sight table with the clean "relative" crops it implies."""
    g = torch.Generator().manual_seed(seed)
    X, Y = SIZE
    T, N = int(cfg.type_objects), int(cfg.num_agents)
    out = {"": {}, "next_": {}}
    for pre in out:
        if fixture.endswith("_ch"):
            out[pre]["objects"] = (torch.rand((B, T, X, Y), generator=g) < 0.1).to(torch.int8)
            out[pre]["agents"] = torch.stack([torch.randint(0, X, (B, N), generator=g), torch.randint(0, Y, (B, N), generator=g)],
                                             dim=2).to(torch.int8)
            dest = torch.zeros((B, X, Y), dtype=torch.int8)
            half = torch.randint(0, 4, (B,), generator=g)
            for b in range(B):  # half-plane destination areas, as the central_room_destination maps use
                if half[b] == 0: dest[b, : X // 2] = 1
                elif half[b] == 1: dest[b, X // 2:] = 1
                elif half[b] == 2: dest[b, :, : Y // 2] = 1
                else: dest[b] = 1
            out[pre]["destination"] = dest
    los = None
    if fixture.endswith("_noise"):
        V = VIEW
        los = (torch.rand((X, Y, V, V), generator=g) < 0.7).to(torch.int8)
        los[:, :, V // 2, V // 2] = 1
        for pre, key, pkey in (("", "obs", "pos"), ("next_", "next_obs", "next_pos")):
            noisy = batch[key]
            clean = (torch.rand(noisy.shape, generator=g) < 0.15).float()
            clean[:, -1] = torch.stack([los[int(p[0]), int(p[1])].float() - 1 for p in batch[pkey]])
            # the local crop is the clean one plus noise on every channel (flip noise leaves integers, sensing noise reals)
            batch[key] = clean + 0.3 * torch.randn(noisy.shape, generator=g)
            out[pre]["relative"] = clean
    return out, los


def main():
    rr._import_reference()
    from core.agents.brains import generate_brain
    os.makedirs(GOLDEN, exist_ok=True)
    only = set(sys.argv[1:])
    for fixture, (cfg, shape) in configs().items():
        if only and fixture not in only:
            continue
        kind = fixture  # the fixture's network family
        for suffix in ("_hard", "_ch", "_noise"):
            if kind.endswith(suffix):
                kind = kind[:-len(suffix)]
        torch.manual_seed(921)
        brain = generate_brain(config=rr.to_dictconfig(cfg), obs_shape=shape, act_size=4)
        masked = kind in ("relative_conv_mlp", "merged_conv_mlp", "merged_da3", "da6", "merged_da3_iqn", "da6_iqn",
                          "relative_da3", "relative_da3_iqn", "relative_iqn", "merged_iqn")
        is_iqn = kind in ("iqn", "da3_iqn", "merged_da3_iqn", "da6_iqn", "relative_da3_iqn", "relative_iqn", "merged_iqn")
        with torch.no_grad():  # zero-initialised embeddings: perturb so the fixtures exercise them
            for name in {"da3_iqn": ("saliency_vector", "pos_embed"), "relative_da3": ("saliency_vector", "pos_embed"),
                         "relative_da3_iqn": ("saliency_vector", "pos_embed"), "merged_da3_iqn": ("local_saliency_vector", "local_pos_embed"),
                         "da6_iqn": ("saliency_vector", "relative_pos_embed", "local_pos_embed")}.get(kind, ()):
                p_ = getattr(brain.network, name)
                p_.add_(0.02 * torch.randn_like(p_))
        if kind == "da6":  # zero-initialised embeddings (da6.py:52-66)
            with torch.no_grad():
                for name in ("saliency_vector", "relative_pos_embed", "local_pos_embed"):
                    p_ = getattr(brain.network, name)
                    p_.add_(0.02 * torch.randn_like(p_))
        if kind == "merged_da3":
            with torch.no_grad():
                brain.network.local_saliency_vector.add_(0.02 * torch.randn_like(brain.network.local_saliency_vector))
                brain.network.local_pos_embed.add_(0.02 * torch.randn_like(brain.network.local_pos_embed))
        if kind == "da3":  # zero-initialised embeddings (da3.py:41-44): perturb so the fixture exercises them
            with torch.no_grad():
                brain.network.saliency_vector.add_(0.02 * torch.randn_like(brain.network.saliency_vector))
                brain.network.pos_embed.add_(0.02 * torch.randn_like(brain.network.pos_embed))
        # make the target differ from the online net, as it does between synchronisations
        with torch.no_grad():
            for p_ in brain.target_network.parameters():
                p_.add_(0.01 * torch.randn_like(p_))
        init = {k: v.detach().clone() for k, v in brain.network.state_dict().items()}
        tgt = {k: v.detach().clone() for k, v in brain.target_network.state_dict().items()}
        batch = synthetic_batch(shape, 7, VIEW if masked else None)
        if masked:
            gp = torch.Generator().manual_seed(11)
            for key in ("pos", "next_pos"):
                batch[key] = torch.stack([torch.randint(0, SIZE[0], (B,), generator=gp),
                                          torch.randint(0, SIZE[1], (B,), generator=gp)], dim=1)
            batch["pos"][0], batch["pos"][1] = torch.tensor([0, 0]), torch.tensor([SIZE[0] - 1, SIZE[1] - 1])
        extras, los = ({"": {}, "next_": {}}, None)
        if masked and (fixture.endswith("_ch") or fixture.endswith("_noise")):
            extras, los = synthetic_extras(cfg, fixture, batch, 13)
        state = (lambda: ref_state(batch["obs"], batch["pos"], extras[""])) if masked else (lambda: {"local": batch["obs"].clone()})
        next_state = (lambda: ref_state(batch["next_obs"], batch["next_pos"], extras["next_"])) if masked else \
            (lambda: {"local": batch["next_obs"].clone()})
        with torch.no_grad():
            if is_iqn and kind != "iqn":
                tau0 = torch.rand(1, 4)
                q0 = brain.network(state(), external_taus=tau0)
            elif masked:
                q0 = brain.network(state())
            elif kind == "c51":
                from core.agents.models.customs.categorical_dqn import ApplySoftmax
                q0 = torch.sum(brain.network({"local": batch["obs"].clone()}, ApplySoftmax.NORMAL) * brain.support, dim=2)
            elif kind == "qr":
                q0 = brain.network({"local": batch["obs"].clone()}).mean(dim=2)
            elif kind == "da3":
                q0 = brain.network({"local": batch["obs"].clone()})
            elif kind == "iqn":
                tau0 = torch.rand(1, 4)
                q0 = brain.network({"local": batch["obs"].clone()}, external_taus=tau0)
            else:
                q0 = brain.network({"local": batch["obs"].clone()})
        losses, drawn = [], []
        real_rand = torch.rand

        def recording_rand(*a, **kw):  # iqn_brain.py:74,87 and customs/iqn.py:106 draw their fractions with torch.rand
            kw.pop("device", None)
            t = real_rand(*a, **kw)
            drawn.append(t.clone())
            return t
        for _ in range(2):
            torch.rand = recording_rand if is_iqn else real_rand
            out = brain.learn(state(), batch["actions"].clone(), batch["rewards"].clone(),
                              batch["dones"].clone(), next_state())
            torch.rand = real_rand
            losses.append(float(out["total_loss"] if isinstance(out, dict) else out))
        final = {k: v.detach().clone() for k, v in brain.network.state_dict().items()}
        arrays = {"init/" + k: v.numpy() for k, v in init.items()}
        arrays.update({"target/" + k: v.numpy() for k, v in tgt.items()})
        arrays.update({"final/" + k: v.numpy() for k, v in final.items()})
        arrays.update({"batch/" + k: v.numpy() for k, v in batch.items()})
        for pre, d in extras.items():
            arrays.update({"extras/" + pre + k: v.numpy() for k, v in d.items()})
        if los is not None:
            arrays["los"] = los.numpy()
        if is_iqn:  # per learn call: taus [B,N], tau_sel [1,N], tau_dashes [B,N']
            assert len(drawn) == 6
            for u in range(2):
                arrays[f"taus/{u}/taus"], arrays[f"taus/{u}/sel"], arrays[f"taus/{u}/dash"] = (
                    drawn[3 * u].numpy(), drawn[3 * u + 1].numpy(), drawn[3 * u + 2].numpy())
            arrays["tau0"] = tau0.numpy()
        path = os.path.join(GOLDEN, f"learner_{fixture}.npz")
        np.savez_compressed(path, losses=np.array(losses, np.float64), q0=q0.numpy(), shape=np.array(shape),
                            gamma=float(cfg.gamma), area_mask=cfg.observation_area_mask,
                            attn_shrink=float(cfg.model["af_lambd"]) if fixture.endswith("_hard") else -1.0, **arrays)
        print(fixture, losses, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
