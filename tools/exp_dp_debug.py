import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maexp_b200 import _lib
from maexp_b200.core.agents.brains import LearnerGroup
from oracle.make_golden_learner import configs as golden_configs
kind = "c51"
cfg, shape = golden_configs()[kind]
G, B, world = 2, int(os.environ.get("BB", "128")), 2
C, V = shape[0], shape[1]
torch.manual_seed(3)
ranks = [LearnerGroup(cfg.merged_with(batch_size=B), shape, 4, G, B) for _ in range(world)]
one = LearnerGroup(cfg.merged_with(batch_size=B * world), shape, 4, G, B * world)
for g in ranks[1:] + [one]:
    g.params().copy_(ranks[0].params())
    g.params(_lib.BUF_TARGET).copy_(ranks[0].params(_lib.BUF_TARGET))
gen = torch.Generator().manual_seed(4)
stride = ranks[0].obs_stride
def rows(n):
    x = torch.zeros((G, n, stride))
    x[:, :, :C * V * V] = (torch.rand((G, n, C * V * V), generator=gen) < 0.25).float()
    return x.cuda()
full = {"obs": rows(B * world), "next_obs": rows(B * world),
        "actions": torch.randint(0, 4, (G, B * world), generator=gen).to(torch.int32).cuda(),
        "rewards": torch.randint(-1, 2, (G, B * world), generator=gen).float().cuda(),
        "dones": (torch.rand((G, B * world), generator=gen) < 0.1).to(torch.uint8).cuda()}
total = torch.zeros_like(one.params())
for r, g in enumerate(ranks):
    g.grads({k: v[:, r * B:(r + 1) * B].contiguous() for k, v in full.items()})
    total += g.params(_lib.BUF_GRADS)
one.grads(full)
want = one.params(_lib.BUF_GRADS)
torch.cuda.synchronize()
print("overall", float((total / world - want).norm() / want.norm()))
for name, off, cnt in one.table:
    a, b = (total / world)[:, off:off + cnt], want[:, off:off + cnt]
    print(f"{name:28s} {float((a - b).norm() / (b.norm() + 1e-30)):.3e}")
