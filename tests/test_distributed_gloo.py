"""World-size-2 gloo run (CPU) of the multi-rank host logic: env sharding and gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maexp_b200 import distributed as mdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, lr, w = mdist.rank_info()
        assert (r, lr, w) == (rank, rank, world)
        E = 8192
        ids = range(mdist.first_global_env_id(rank, E), mdist.first_global_env_id(rank, E) + E)
        # flat per-agent gradient buffers [G, P]; each rank holds the gradient of its local batch
        torch.manual_seed(rank)
        grads = torch.randn(8, 1000)
        local = grads.clone()
        all_reduce = mdist.make_grad_all_reduce(dist)
        n = all_reduce(grads)
        mx = mdist.max_over_ranks(dist, [1.0 + rank, 5.0 - rank], "cpu")
        sm = mdist.sum_over_ranks(dist, [float(len(ids))], "cpu")
        # replicas: every rank initialises its own weights; rank 0's are broadcast, and K identical "updates" (all-reduced
        # gradient, scaled by 1/world) keep them bit-equal
        params, state = torch.randn(8, 1000), torch.rand(8, 1000)
        spread_before = mdist.replica_spread(dist, params)
        mdist.broadcast_tensors(dist, [params, state])
        for k in range(3):
            g = torch.randn(8, 1000)  # differs per rank (seeded by rank above)
            w = all_reduce(g)
            state.mul_(0.99).addcmul_(g / w, g / w, value=0.01)
            params.sub_(1e-3 * (g / w) / (state.sqrt() + 1e-8))
        spread_after = mdist.replica_spread(dist, params)
        out[rank] = (n, grads.clone(), local, mx, sm, (ids[0], ids[-1]), spread_before, spread_after, params.clone())
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_all_reduce_and_sharding():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][0] == res[1][0] == 2
    want = res[0][2] + res[1][2]
    assert torch.equal(res[0][1], want) and torch.equal(res[1][1], want)  # sum, identical on both ranks
    assert res[0][3] == res[1][3] == [2.0, 5.0]                           # max over ranks
    assert res[0][4] == [16384.0]
    assert res[0][5] == (0, 8191) and res[1][5] == (8192, 16383)          # disjoint, contiguous global env ids
    assert res[0][6] == res[1][6] and res[0][6] > 0.1                      # unsynchronised initialisations differ ...
    assert res[0][7] == res[1][7] == 0.0                                  # ... broadcast + all-reduced updates do not
    assert torch.equal(res[0][8], res[1][8])
