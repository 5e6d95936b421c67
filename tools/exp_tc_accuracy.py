"""GPU experiment: relative error of the acting-batch Q-values against the fp32 torch oracle over random initialisations,
for the TMA-fed heads / tensor-core trunk kernels and the per-layer GEMM path (MAEXP_TC_HEADS=0 MAEXP_TC_TRUNK=0)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_learner_gpu import golden_configs, pack, TorchLearner, rel_err, ATOMS
from maexp_b200.core.agents.brains import LearnerGroup
kind, V, C = os.environ.get("KIND", "c51"), 7, 3
B = int(os.environ.get("EXP_B", "640"))
cfg, _ = golden_configs()[kind]
cfg = cfg.merged_with(batch_size=B, visible_range=V)
errs = []
for seed in range(int(os.environ.get("SEEDS", "12"))):
    torch.manual_seed(seed)
    g = LearnerGroup(cfg, [C, V, V], 4, 2, B)
    gen = torch.Generator().manual_seed(3 + seed)
    obs = (torch.rand((2, B, C, V, V), generator=gen) < 0.25).float()
    obs[:, :, -1] -= 1
    q = g.q_values(pack(g, obs)).cpu()
    for a in range(2):
        o = TorchLearner(kind, g.state_dict(a), 4, float(cfg.gamma), 1e-3, 0.99, 1e-8, num_atoms=ATOMS[kind])
        errs.append(rel_err(q[a], o.q_values(obs[a])))
print(kind, B, "rel_err max %.2e mean %.2e" % (max(errs), sum(errs) / len(errs)), " ".join("%.1e" % e for e in errs))
