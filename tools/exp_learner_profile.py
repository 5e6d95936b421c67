"""GPU experiment: per-kernel time inside one graph-replayed team update (torch.profiler / CUPTI; not part of the product path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from torch.profiler import ProfilerActivity, profile
from maexp_b200.config import baseline_config
from maexp_b200.core.agents.brains import LearnerGroup

if os.environ.get("KIND"):  # a learner kind of oracle/make_golden_learner.py (local 3 x 7 x 7 view), e.g. KIND=iqn B=256 G=8
    from oracle.make_golden_learner import configs as golden_configs
    cfg = golden_configs()[os.environ["KIND"]][0]
    G, B = int(os.environ.get("G", "8")), int(os.environ.get("B", "256"))
    cfg = cfg.merged_with(batch_size=B)
else:
    cfg = baseline_config(int(os.environ.get("CFG", "5")))
    G, B = int(cfg.num_agents), int(cfg.batch_size)
g = LearnerGroup(cfg, [3, 7, 7], 4, G, B)
batch = {"obs": torch.rand(G, B, g.obs_stride, device="cuda").round(), "next_obs": torch.rand(G, B, g.obs_stride, device="cuda").round(),
         "actions": torch.randint(0, 4, (G, B), dtype=torch.int32, device="cuda"), "rewards": torch.zeros(G, B, device="cuda"),
         "dones": torch.zeros(G, B, dtype=torch.uint8, device="cuda")}
if g.kind == "iqn":  # fixed fraction buffers (a captured torch.rand would replay one draw)
    batch.update(taus=torch.rand(G, B, g.Z, device="cuda"), tau_dashes=torch.rand(G, B, g.Nd, device="cuda"),
                 tau_sel=torch.rand(G, g.Z, device="cuda"))
for _ in range(3):
    g.learn(batch)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    g.learn(batch)
for _ in range(5):
    gr.replay()
torch.cuda.synchronize()
R = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(R):
        gr.replay()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
t0, t1 = None, None
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name[:60]][0] += 1
        agg[e.name[:60]][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        s, f = e.time_range.start, e.time_range.end
        t0 = s if t0 is None else min(t0, s)
        t1 = f if t1 is None else max(t1, f)
tot = sum(v[1] for v in agg.values())
print(f"span per update {(t1 - t0) / R:.1f} us, sum of kernel times per update {tot / R:.1f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:62s} n/update {n / R:6.1f}  us/update {t / R:8.1f}  avg {t / n:6.2f}")

# kernel sequence of ONE replay in start order (side-stream passes interleave)
with profile(activities=[ProfilerActivity.CUDA]) as prof1:
    gr.replay()
    torch.cuda.synchronize()
ev = sorted([e for e in prof1.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print(f"{e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:6.1f}  {e.name[:70]}")
