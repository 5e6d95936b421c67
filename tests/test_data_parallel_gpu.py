"""Data-parallel learner on hardware: replicas start equal, stay equal, and an all-reduced update equals
the single-GPU update on the concatenated batch.  Two ranks share cuda:0 over the gloo backend (NCCL refuses two ranks on
one device; gloo stages CUDA tensors through the host), so the test runs on a one-GPU box; bench.py repeats the checks over
NCCL at N > 1."""
import os
import socket

import pytest
import torch

from maexp_b200.config import baseline_config

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _trainer_worker(rank, world, port, brain, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from maexp_b200 import _lib
        from maexp_b200.core.handlers.trainers import DefaultTrainer
        from maexp_b200.core.maps import generate_map
        from maexp_b200.core.worlds import generate_world
        from maexp_b200.distributed import first_global_env_id
        E = 64
        cfg = baseline_config(5 if brain == "c51" else 3).merged_with(max_episode_length=6, batch_size=32, populate_steps=4 * E,
                                                                      capacity=16 * E, max_epochs=2)
        if brain == "iqn":
            cfg = cfg.merged_with(brain="iqn", model=dict(cfg.model, name="iqn", num_quantiles=8, num_quantiles_dash=8,
                                                          num_cosines=16))
        torch.manual_seed(100 + rank)  # every rank initialises different weights: the trainer must broadcast rank 0's

        class Env:  # the trainer only reads .world of a non-batched environment argument
            world = generate_world(cfg, generate_map(cfg))
        t = DefaultTrainer(cfg, Env, num_envs=E, first_global_env_id=first_global_env_id(rank, E), dist=dist)
        assert t.graphed is not None and t.graphed.dist is dist  # the graph-replayed update, also for IQN and N > 1
        start = t.learner.params().cpu().clone()
        hist = t.run()
        assert t.graphed.g_grads is not None
        spread = t.learner.replica_spread(dist)
        out[rank] = (start, t.learner.params().cpu().clone(), spread, hist[-1]["loss"],
                     t.learner.params(_lib.BUF_SQUARE_AVG).cpu().clone())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("brain", ["c51", "iqn"])
def test_two_rank_trainer_replicas_stay_in_sync(brain):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_trainer_worker, args=(world, port, brain, out), nprocs=world, join=True)
        res = dict(out)
    assert torch.equal(res[0][0], res[1][0])          # broadcast at construction
    assert not torch.equal(res[0][0], res[0][1])      # the updates moved the weights ...
    assert torch.equal(res[0][1], res[1][1])          # ... identically on both ranks (12 updates on different local batches)
    assert torch.equal(res[0][4], res[1][4])          # optimizer state too
    assert res[0][2] == res[1][2] == 0.0
    assert res[0][3] != res[1][3]                     # the ranks did see different batches


@pytest.mark.parametrize("kind", ["c51", "qr", "dqn_conv_mlp"])
def test_all_reduced_update_equals_single_gpu_update_on_global_batch(kind):
    """Two ranks with B samples each, gradients summed and scaled by 1/2 (what the NCCL all-reduce + maexp_learner_apply do),
    against ONE learner on the 2B concatenated batch: same gradient, same clipped RMSprop step."""
    from maexp_b200 import _lib
    from maexp_b200.core.agents.brains import LearnerGroup
    from oracle.make_golden_learner import configs as golden_configs
    cfg, shape = golden_configs()[kind]
    # B = 128 per rank: the ranks and the global-batch learner then take the same kernels (update batches of >= 128 rows run the
    # conv2 / MLP-head contractions on the tensor cores with three TF32 terms; 64 + 64 against 128 would compare two arithmetic paths)
    G, B, world = 2, 128, 2
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
    err = float((total / world - want).norm() / want.norm())
    assert err <= 1e-6, err
    for g in ranks:  # every rank applies the summed gradient with grad_scale = 1/world
        g.params(_lib.BUF_GRADS).copy_(total)
        g.apply(grad_scale=1.0 / world)
    one.apply()
    assert torch.equal(ranks[0].params(), ranks[1].params())
    assert abs(float(ranks[0].grad_norm[0]) - float(one.grad_norm[0])) <= 1e-5 * float(one.grad_norm[0])
    moved = (one.params() - ranks[0].params()).abs()
    # RMSprop's first step is +-lr/sqrt(1-alpha) whatever the gradient's size: compare where the gradient is not noise
    big = want.abs() > 1e-3 * want.abs().max()
    assert float(moved[big].max()) <= 1e-5


def test_fill_uniform_is_a_unit_uniform_stream():
    """maexp_fill_uniform (the IQN fractions of a graph-replayed update): [0,1), reproducible per (seed, tick), moments of
    U[0,1), and the same 24-bit convention as torch.rand (multiples of 2^-24)."""
    import ctypes as C
    from maexp_b200 import _lib
    lib = _lib.load()
    n = 1_000_003
    a, b, c = (torch.full((n,), -1.0, device="cuda") for _ in range(3))
    for t, (seed, tick) in zip((a, b, c), ((921, 5), (921, 5), (921, 6))):
        _lib.check(lib.maexp_fill_uniform(C.c_void_p(t.data_ptr()), n, seed, tick, None))
    torch.cuda.synchronize()
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert float(a.min()) >= 0.0 and float(a.max()) < 1.0
    assert abs(float(a.double().mean()) - 0.5) < 2e-3 and abs(float(a.double().var()) - 1 / 12) < 1e-3
    assert torch.equal((a.double() * 2 ** 24).round(), a.double() * 2 ** 24)
    hist = torch.histc(a, bins=64, min=0, max=1)
    assert float((hist - n / 64).abs().max()) < 6 * (n / 64) ** 0.5
    assert abs(float(torch.corrcoef(torch.stack([a[:-1], a[1:]]))[0, 1])) < 5e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_fused_nvlink_apply_equals_nccl_path_on_two_gpus():
    """maexp_learner_apply_data_parallel (one cooperative kernel: reduce-scatter / all-gather over NVLink peer memory + clip +
    RMSprop) against NCCL all-reduce + maexp_learner_apply on twin learners, one process per GPU: identical parameters on every
    rank and between the two paths (tools/dp_fused_check.py asserts it and prints the timings)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(root, "tools", "dp_fused_check.py")],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, B="64"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"replica_spread_fused": 0.0' in r.stdout
