"""GPU experiment: phase timeline of the graph-replayed C51 team update WITHOUT a profiler attached (%globaltimer stamps written
by one-thread kernels at the phase boundaries; MAEXP_LEARNER_STAMPS=1).  CUPTI timelines serialise the update's branches."""
import os, sys
if os.environ.get("STAMPS", "1") != "0":  # STAMPS=0: the plain timing, no stamp kernels in the graph
    os.environ["MAEXP_LEARNER_STAMPS"] = os.environ.get("STAMPS", "1")  # "1": all slots; else a bit mask of slots (0x641: start, loss, joined, done)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import GraphedUpdate, LearnerGroup
cfg = baseline_config(int(os.environ.get("CFG", "5")))
G, B = int(cfg.num_agents), int(cfg.batch_size)
g = LearnerGroup(cfg, [3, 7, 7], 4, G, B)
batch = {"obs": torch.rand(G, B, g.obs_stride, device="cuda").round(), "next_obs": torch.rand(G, B, g.obs_stride, device="cuda").round(),
         "actions": torch.randint(0, 4, (G, B), dtype=torch.int32, device="cuda"), "rewards": torch.zeros(G, B, device="cuda"),
         "dones": torch.zeros(G, B, dtype=torch.uint8, device="cuda")}
gu = GraphedUpdate(g, batch)
for _ in range(30):
    gu.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    gu.step()
b.record(); torch.cuda.synchronize()
stamps = "MAEXP_LEARNER_STAMPS" in os.environ
print("graph-replayed update: %.1f us (%s the stamp kernels in the graph)" % (a.elapsed_time(b) * 10, "with" if stamps else "without"))
if stamps:
    g.apply(1.0)  # eager: prints the stamps of the last replay
torch.cuda.synchronize()
