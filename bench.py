#!/usr/bin/env python
"""Benchmark of the MAEXP hot path on B200: batched gridworld step + egocentric observation.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `maexp_env_step` over the whole per-GPU batch (BASELINE.json configs[4]:
four_rooms, 65 536 envs x 8 agents, V=7, C=3, random policy, auto-reset), observations written
as fp32 into a rotating device-resident replay-ring slot.  Prints ONE JSON line (rank 0).

`--impl reference` times the CPU implementation of the same path on the host cores.  The
reference is pure Python and cannot travel to the GPU box, so this arm runs the C oracle port
(oracle/maexp_oracle.c, pinned bit-exactly against the live reference in the build container)
with all host threads — a much stronger baseline than the reference's own ~5.7 k agent-steps/s
per core.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 65536
RING_SLOTS = 8
B_ALG = 744.0  # algorithmic bytes per agent-step for cfg5 (588 obs + 144 grid r/w + 12)
METRIC, UNIT = "agent_steps_per_sec", "agent-steps/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel at this workload, from profiles/ (ncu --set full):
# env_step_lanes_kernel<7,8> 8.8 MB read + 274.2 MB written (profiles/r2_env_step_lanes_kernel_ncu_full.csv);
# env_step_kernel<7,16,0,8> (MAEXP_LANES=0) 8.81 + 273.64 MB (profiles/r2_env_step_kernel_ncu_full.csv)
NCU_TRAFFIC_BYTES_PER_LAUNCH = {True: 282.96e6, False: 282.45e6}


def workload_config():
    from maexp_b200.config import baseline_config
    return baseline_config(5)


def config_block(n_envs, extra=None):
    c = {"workload": "BASELINE configs[4]: four_rooms 24x24, %d envs/GPU x 8 agents, V=7, C=3 (default views), "
                     "25 objects kept, L=200 auto-reset, uniform random policy + move order from on-chip Philox"
                     % n_envs,
         "envs_per_gpu": n_envs, "agents": 8, "map": "four_rooms", "visible_range": 7, "channels": 3,
         "ring_slots": RING_SLOTS,
         "l2": "each step streams a fresh ring slot (310 MB at 65536 envs) > 126 MB L2; slots rotate"}
    if extra:
        c.update(extra)
    return c


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the GPU phases run."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt, self.window = index, [], threading.Event(), [None, None]
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        nv = self.nv
        t0, t1 = self.window
        inside = [s for s in self.samples if t0 is not None and t0 <= s[0] <= t1] or self.samples[-3:]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        seen = sorted({n for _, _, r in inside for n, bit in names.items() if r & bit})
        return {"sm_mhz": statistics.median(m for _, m, _ in inside), "sm_max_mhz": self.max_mhz, "reasons": seen,
                "samples": len(inside)}


def cpu_oracle_rate(n_envs, steps, warmup, threads):
    """agent-steps/s of the C oracle port on the host cores (random policy, auto-reset)."""
    from oracle.oracle import Oracle
    cfg = workload_config()
    o = Oracle(cfg, n_envs, auto_reset=True)
    o.reset(0)
    tick = 1
    for _ in range(warmup):
        o.step_nocopy(tick, threads=threads)
        tick += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step_nocopy(tick, threads=threads)
        tick += 1
    dt = time.perf_counter() - t0
    return n_envs * 8 * steps / dt, dt


PY_REFERENCE = {"value": 5733.0, "unit": "agent-steps/s per core", "cores": 1,
                "source": "BASELINE.md §2: the unmodified Python reference (default_environment.py:153), four_rooms, 8 agents, "
                          "V=7, C=3, measured in the build container (it cannot travel to the GPU box)"}


def run_reference(args, rank, world):
    """The reference's CPU implementation of the path on the host cores: every step walks ALL envs of the configuration."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = args.envs
    steps = max(1, min(args.steps, 400))
    rate, dt = cpu_oracle_rate(n_envs, steps, max(3, min(args.warmup, 10)), threads)
    sample = "all %d envs x 8 agents per step, %d steps, %.1f s (C oracle port, %d pthreads)" % (n_envs, steps, dt, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_block(n_envs),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "python_reference": PY_REFERENCE},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def learner_cpu_rate(batch_size, steps=8):
    """updates/s of the torch fp32 restatement of CategoricalDQNBrain.learn on the host cores."""
    import torch

    from maexp_b200.core.agents.models import build_reference_modules
    from oracle.torch_learner import TorchLearner
    cfg = workload_config()
    torch.manual_seed(0)
    sd = build_reference_modules(cfg, [3, 7, 7], 4).state_dict()
    o = TorchLearner("c51", sd, 4, float(cfg.gamma), 1e-3, 0.99, 1e-8, num_atoms=51)
    B = batch_size
    obs, nxt = torch.rand(B, 3, 7, 7).round(), torch.rand(B, 3, 7, 7).round()
    act, rew, done = torch.randint(0, 4, (B,)), torch.randint(-1, 2, (B,)).double(), torch.rand(B) < 0.05
    for _ in range(2):
        o.learn(obs, nxt, act, rew, done)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.learn(obs, nxt, act, rew, done)
    return steps / (time.perf_counter() - t0)


# fp32 FLOPs of one C51 sample-update (forward of s, two forwards of s', backward), BASELINE.md §3 / SURVEY.md §8d
C51_FLOP_PER_SAMPLE = 1.45e6
SIMT_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FMA lanes x 2 FLOP at the 1965 MHz maximum clock


def learner_roofline(flop, seconds):
    """The update's contractions run as three TF32 tensor-core terms per product (fp32-level accuracy): the bound is the
    measured dense bf16 rate / 2 (TF32 runs at half the bf16 rate) / 3 (terms); the fp32 SIMT peak is printed beside it."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16, src = 1398.0, "fallback (B200_PROFILING.md sustained)"
    if os.path.exists(path):
        with open(path) as f:
            bf16, src = float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained / 2 / 3)"
    peak = bf16 / 6.0
    achieved = flop / seconds / 1e12
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": src, "flop_per_team_update": flop, "frac_of_fp32_simt_peak": achieved / SIMT_FP32_TFLOPS,
            "note": "B = 256 per agent-network: the update is a chain of small dependent layers, latency-bound by design of "
                    "the workload (2.97 GFLOP per team update)"}


def data_parallel_check(torch, dist, group, ring, batch, cfg, env, world, dev):
    """SURVEY.md §4-iv on hardware: (1) after the timed updates every rank holds the same parameters; (2) the all-reduced,
    1/world-scaled gradient of one more update equals the gradient a single GPU computes on the concatenated global batch."""
    from maexp_b200 import _lib
    from maexp_b200.core.agents.brains import LearnerGroup
    spread = group.replica_spread(dist)
    ring.sample(batch, seed=921, tick=9999)
    group.grads(batch)
    g_dp = group.params(_lib.BUF_GRADS).clone()
    dist.all_reduce(g_dp, op=dist.ReduceOp.SUM)
    g_dp /= world
    B = batch["actions"].shape[1]
    big = LearnerGroup(cfg.merged_with(batch_size=B * world), [3, 7, 7], 4, env.N, B * world, obs_stride=env.slot_floats,
                       device=dev)
    big.params(_lib.BUF_ONLINE).copy_(group.params(_lib.BUF_ONLINE))
    big.params(_lib.BUF_TARGET).copy_(group.params(_lib.BUF_TARGET))
    full = {}
    for k in ("obs", "next_obs", "actions", "rewards", "dones"):
        parts = [torch.empty_like(batch[k]) for _ in range(world)]
        dist.all_gather(parts, batch[k].contiguous())
        full[k] = torch.cat(parts, dim=1).contiguous()  # [N, world*B, ...]
    big.grads(full)
    g_one = big.params(_lib.BUF_GRADS)
    err = float((g_dp - g_one).norm() / g_one.norm())
    ok = spread == 0.0 and err <= 1e-6
    out = {"replicas_in_sync": spread == 0.0, "replica_max_abs_diff": spread,
           "dp_grad_vs_single_gpu_global_batch_rel_l2": err}
    if not ok:
        raise SystemExit("bench.py: data-parallel check failed: %s" % json.dumps(out))
    return out


def run_ours(args, rank, local_rank, world):
    import torch

    from maexp_b200 import _lib
    from maexp_b200 import distributed as mdist
    from maexp_b200.core.agents.brains import GraphedUpdate, LearnerGroup
    from maexp_b200.core.environments.batched import BatchedEnvironment
    from maexp_b200.core.maps import generate_map
    from maexp_b200.core.utils.buffer import DeviceReplayRing
    from maexp_b200.core.worlds import generate_world

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout while the communicator is created: send it to stderr so that stdout
        # carries the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    E, N, K, W = args.envs, 8, args.steps, args.warmup
    cfg = workload_config()
    world_obj = generate_world(cfg, generate_map(cfg))
    # envs are independent units: contiguous global ids per rank, no data-path collective
    env = BatchedEnvironment(cfg, world_obj, E, device=dev, auto_reset=True,
                             first_global_env_id=mdist.first_global_env_id(rank, E))
    lanes = env.uses_lanes_kernel()  # one lane per (env, agent): csrc/env_lanes.cu
    ring = DeviceReplayRing(env, RING_SLOTS)  # obs + actions + rewards + dones + was_reset rings in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()

    def step(tick):
        env.step(ring.next_outputs(), tick=tick)
        ring.commit()

    tick = 0
    env.reset(ring.next_outputs(), tick=tick)
    ring.commit()
    tick += 1
    sampler.window[0] = time.perf_counter()
    for _ in range(W):
        step(tick)
        tick += 1
    before = env.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        step(tick)
        tick += 1
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    after = env.stats()
    env_steps = after["env_steps"] - before["env_steps"]  # resets (1 in L+1 calls per env) are not counted

    # ---- end to end through the host-buffer C-ABI entry point: every step copies its pinned actions in and its
    #      rewards / done flags / was_reset out (compact encoding: int8 rewards in {-1,0,1}, one done flag per env —
    #      lossless for the default reward mode); copies and kernels of consecutive steps overlap (3 steps in flight)
    depth = 4
    hb = [{"actions": torch.randint(0, 4, (E, N), dtype=torch.int8).pin_memory(),
           "rewards": torch.empty((E, N), dtype=torch.int8).pin_memory(),
           "done": torch.empty((E,), dtype=torch.uint8).pin_memory(),
           "was_reset": torch.empty((E,), dtype=torch.uint8).pin_memory()} for _ in range(depth)]

    def host_step(i, tick):
        b = hb[i % depth]
        return env.step_host_compact_async(b["actions"], ring.obs[tick % RING_SLOTS], b["rewards"], b["done"], b["was_reset"],
                                           tick=tick)
    for i in range(3):
        host_step(i, tick)
        tick += 1
    env.host_wait()
    k2 = max(1, min(K, 200))
    b2 = env.stats()
    barrier()
    t0 = time.perf_counter()
    for i in range(k2):
        host_step(i, tick)
        tick += 1
    env.host_wait()  # every step's results are in host memory
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_steps = env.stats()["env_steps"] - b2["env_steps"]

    # ---- learner leg: fused replay sample + C51 update of the 8 per-agent networks (B = 256 per GPU),
    #      forward/backward and clip+RMSprop replayed from CUDA graphs, NCCL gradient all-reduce when N > 1
    for _ in range(3):  # refill the ring's meta arrays with ordinary steps after the host-entry steps
        step(tick)
        tick += 1
    B = int(cfg.batch_size)
    group = LearnerGroup(cfg, [3, 7, 7], 4, N, B, obs_stride=env.slot_floats, device=dev)
    if dist is not None:
        group.broadcast(dist)  # data-parallel replicas start from rank 0's weights (each rank initialised its own)
    batch = ring.new_batch(B)
    ring.sample(batch, seed=921, tick=0)
    graphed = GraphedUpdate(group, batch, dist=dist)  # first step eager (warm-up), second captures, then replays

    def update(t):
        ring.sample(batch, seed=921, tick=t)
        graphed.step()

    for t in range(5):
        update(t)
    ku = max(10, min(K, 200))
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    u0.record()
    for t in range(ku):
        update(100 + t)
    u1.record()
    barrier()
    upd_ms = u0.elapsed_time(u1)
    loss = float(group.loss.mean())
    ring.check(batch)  # the replay sampler's error flag
    dp_check = None
    if dist is not None:
        dp_check = data_parallel_check(torch, dist, group, ring, batch, cfg, env, world, dev)
    # ---- training-loop leg: what a real run of DefaultTrainer does per tick (abstract_trainer.py:63-89) -- replay sample + C51
    #      update, refresh of the acting weights, epsilon-greedy policy inference for ALL environments (conv trunk with conv2 on
    #      tcgen05 + TMA-fed MLP heads with the fused expectation epilogue), env step with those actions into the next ring slot
    from maexp_b200.core.agents import BatchedActor
    actor = BatchedActor(cfg, env, group)
    loop_tick = [tick]

    def loop_step(t):
        ring.sample(batch, seed=921, tick=t)
        graphed.step()
        actor.refresh()
        actions = actor.act(ring.obs[ring.head], 0.1, loop_tick[0])
        env.step(ring.next_outputs(), actions=actions, tick=loop_tick[0])
        ring.commit()
        loop_tick[0] += 1

    for t in range(3):
        loop_step(500 + t)
    kl = max(5, min(K, 50))
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0.record()
    for t in range(kl):
        loop_step(600 + t)
    l1.record()
    barrier()
    loop_ms = l0.elapsed_time(l1)
    sampler.window[1] = time.perf_counter()
    sampler.stop()

    vals = [ms, e2e_s, upd_ms, loop_ms]
    sums = [float(env_steps), float(e2e_steps)]
    if dist is not None:
        vals = mdist.max_over_ranks(dist, vals, dev)
        sums = mdist.sum_over_ranks(dist, sums, dev)
    ms, e2e_s, upd_ms, loop_ms = vals
    env_steps, e2e_steps = sums
    if rank == 0:
        value = env_steps * N / (ms * 1e-3)
        e2e = e2e_steps * N / e2e_s
        peak, peak_src = peaks()
        # dominant (only) kernel of the step: env_step_kernel; K launches back to back in the timed region
        launch_s = ms * 1e-3 / K
        per_launch_agent_steps = env_steps * N / K / world
        achieved = B_ALG * per_launch_agent_steps / launch_s / 1e9
        upd_s = upd_ms * 1e-3 / ku
        learner_flop = C51_FLOP_PER_SAMPLE * B * N
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic", "config": config_block(E),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": E * N, "d2h_bytes_per_step": E * N + 2 * E,
                    "steps": k2, "note": "maexp_env_step_host_compact_async: pinned actions H2D, int8 rewards + per-env done + "
                                         "was_reset D2H every step on copy streams, 3 steps in flight, wall clock until the "
                                         "last step's results are in host memory; observations stay in the device replay ring"},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH[lanes], "peak_source": peak_src,
                         "kernel": "env_step_lanes_kernel<7,8>" if lanes else "env_step_kernel<7,16,0,8>",
                         "algorithmic_bytes_per_agent_step": B_ALG,
                         "launch_us": launch_s * 1e6,
                         # the bit-packed state never moves the 144 B / agent-step of int8 grid traffic B_ALG counts:
                         # the same launch time against the DRAM bytes ncu saw
                         "real_bytes_frac": NCU_TRAFFIC_BYTES_PER_LAUNCH[lanes] * (E / ENVS_PER_GPU) / launch_s / 1e9 / peak},
            "clocks": sampler.summary(),
            "learner": {"metric": "learner_updates_per_sec", "unit": "agent-network updates/s",
                        "value": ku * N / (upd_ms * 1e-3), "ms_per_team_update": upd_ms / ku, "brain": "categorical_dqn",
                        "agents": N, "batch_per_gpu": B, "global_batch": B * world, "steps": ku, "loss": loss,
                        "note": "replay sample (1 launch) + forward/backward "
                                + ("and clip+RMSprop in ONE CUDA graph; " if graphed.single_graph else "(CUDA graph) + ")
                                + ("" if world == 1 else
                                   "ONE cooperative kernel: gradient reduce-scatter / all-gather over NVLink peer memory + 1/world + "
                                   "clip + RMSprop (maexp_learner_apply_data_parallel); " if graphed.peers is not None else
                                   "NCCL all-reduce of the flat gradient buffer + ")
                                + ("clip+RMSprop (CUDA graph); " if not graphed.single_graph and graphed.peers is None else "")
                                + "all 8 agents' networks in grouped launches",
                        "collective": None if world == 1 else ("fused NVLink peer kernel" if graphed.peers is not None else "nccl"),
                        "roofline": learner_roofline(learner_flop, upd_s)},
            "training_loop": {"metric": "agent_steps_per_sec_in_training_loop", "unit": UNIT,
                              "value": world * E * N * kl / (loop_ms * 1e-3), "ms_per_loop_step": loop_ms / kl, "steps": kl,
                              "note": "per tick: replay sample + C51 team update + acting-weight refresh + epsilon-greedy policy "
                                      "inference for all envs (tc_trunk_kernel + tc_heads_kernel: tcgen05 / TMA, 3 x TF32) + env step "
                                      "with those actions; CUDA events, max over ranks"},
        }
        if dp_check is not None:
            line["learner"].update(dp_check)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            k_cpu = 150  # the whole configuration per step, as the --impl reference arm
            rate, dt = cpu_oracle_rate(E, k_cpu, 5, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "all %d envs x 8 agents per step, %d steps, %.1f s (C oracle port, %d pthreads)"
                                              % (E, k_cpu, dt, threads),
                                    "python_reference": PY_REFERENCE}
            line["learner"]["cpu_baseline"] = {"value": learner_cpu_rate(B), "unit": "agent-network updates/s",
                                               "cores": threads, "kind": "port",
                                               "sample": "torch fp32 restatement of CategoricalDQNBrain.learn, B=%d, "
                                                         "8 updates of one network" % B}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (default: the BASELINE workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
