"""Pins the torch fp32 learner oracle (oracle/torch_learner.py) against the live reference brains."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.torch_learner import TorchLearner

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KINDS = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLDEN, "learner_*.npz")))
ATOMS = {"dqn_conv_mlp": 1, "c51": 51, "qr": 100, "iqn": 4, "da3": 1, "da6": 1}
IQN_KINDS = ("iqn", "da3_iqn", "merged_da3_iqn", "da6_iqn", "relative_da3_iqn", "relative_iqn", "merged_iqn",
             "da3_iqn_hard", "merged_da3_iqn_hard", "da6_iqn_hard")  # *_hard: model.attention == "hard" (HardShrinkBlock), af_lambd in the fixture
TRUNKS = {"relative_da3_iqn": "da3_iqn", "relative_iqn": "iqn", "da3_iqn_hard": "da3_iqn", "merged_da3_iqn_hard": "merged_da3_iqn",
          "da6_iqn_hard": "da6_iqn"}  # fixture name -> state-embedding trunk of the IQN head (default: the name)
# fixtures of the relative / merged views -> (network kind, attention heads)
MASKED = {"relative_conv_mlp": ("dqn_conv_mlp", 4), "merged_conv_mlp": ("dqn_conv_mlp", 4), "merged_da3": ("da3", 2),
          "da6": ("da6", 2), "merged_da3_iqn": ("iqn", 2), "da6_iqn": ("iqn", 2), "relative_da3": ("da3", 2),
          "relative_da3_iqn": ("iqn", 2), "relative_iqn": ("iqn", 2), "merged_iqn": ("iqn", 2),
          "da6_hard": ("da6", 2), "merged_da3_iqn_hard": ("iqn", 2), "da6_iqn_hard": ("iqn", 2)}


def base(kind):
    """Fixture name without the `_ch` (extra canvas channels on) / `_noise` (merged view under observation noise) suffix:
    the network family the lookup tables above are keyed by."""
    for suffix in ("_ch", "_noise"):
        if kind.endswith(suffix):
            return kind[:-len(suffix)]
    return kind


def load(kind):
    z = np.load(os.path.join(GOLDEN, f"learner_{kind}.npz"))
    part = lambda pre: {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}
    fix = {"init": part("init/"), "target": part("target/"), "final": part("final/"), "batch": part("batch/"),
           "losses": z["losses"], "q0": torch.from_numpy(z["q0"]), "gamma": float(z["gamma"]),
           "shape": [int(v) for v in z["shape"]], "area_mask": str(z["area_mask"]),
           "attn_shrink": float(z["attn_shrink"]) if "attn_shrink" in z.files and float(z["attn_shrink"]) >= 0 else None}
    # the state's / new state's "objects" [B,T,X,Y], "agents" [B,N,2], "destination" [B,X,Y] keys and the noise-free
    # "relative" crop; `los`: the [X,Y,V,V] table the *_noise fixtures' clean area channels were built from
    ex = part("extras/")
    fix["extras"] = ({k: v for k, v in ex.items() if not k.startswith("next_")},
                     {k[5:]: v for k, v in ex.items() if k.startswith("next_")})
    fix["los"] = z["los"] if "los" in z.files else None
    if base(kind) in IQN_KINDS:  # the fractions the reference drew: per update (taus, tau_dashes, tau_sel), and q0's tau
        fix["taus"] = [(torch.from_numpy(z[f"taus/{u}/taus"]), torch.from_numpy(z[f"taus/{u}/dash"]),
                        torch.from_numpy(z[f"taus/{u}/sel"])) for u in range(2)]
        fix["tau0"] = torch.from_numpy(z["tau0"])
    return fix


def make_oracle(kind, fix, sd=None, target=None):
    kind = base(kind)
    net, heads = MASKED.get(kind, ("iqn", 2) if kind in ("da3_iqn", "da3_iqn_hard") else
                            ("da3", 4) if kind == "da3_hard" else (kind, 4))
    o = TorchLearner(net, sd or fix["init"], 4, fix["gamma"], 1e-3, 0.99, 1e-8, num_atoms=ATOMS[net],
                     area_mask=fix["area_mask"], size=fix["shape"][1:], num_heads=heads,
                     trunk=TRUNKS.get(kind, kind) if kind in IQN_KINDS else "iqn", attn_shrink=fix["attn_shrink"])
    o.target = {k: v.clone() for k, v in (target or fix["target"]).items()}
    return o


def oracle_inputs(fix, b):
    """(obs, next_obs) as the oracle takes them: tensors, or (obs, pos) pairs for the relative / merged views."""
    if fix["area_mask"] == "local":
        return b["obs"], b["next_obs"]
    ex, next_ex = b.get("extras") or fix["extras"]
    return (b["obs"], b["pos"], ex), (b["next_obs"], b["next_pos"], next_ex)


@pytest.mark.parametrize("kind", KINDS)
def test_torch_oracle_matches_reference_brain(kind):
    fix = load(kind)
    o, b = make_oracle(kind, fix), fix["batch"]
    obs, next_obs = oracle_inputs(fix, b)
    q0 = o.q_values_iqn(obs, fix["tau0"]) if base(kind) in IQN_KINDS else o.q_values(obs)
    assert torch.allclose(q0, fix["q0"], rtol=1e-6, atol=1e-6)
    for u, want in enumerate(fix["losses"]):
        loss, _ = o.learn(obs, next_obs, b["actions"], b["rewards"], b["dones"],
                          taus=fix["taus"][u] if base(kind) in IQN_KINDS else None)
        assert abs(float(loss) - want) <= 1e-6 * max(1.0, abs(want)), (float(loss), want)
    for k, v in fix["final"].items():
        assert torch.allclose(o.sd[k].detach(), v, rtol=1e-5, atol=1e-6), k
