"""world_size-2 gloo test of the multi-GPU host logic (alphazero_risk_b200/dist.py): sharding, weight-blob broadcast,
counter reduction, max-over-ranks timing — everything bench.py does across ranks, minus the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphazero_risk_b200 import dist as azd


def test_shard_covers_all_games_once():
    for total, world in [(65536, 1), (65536, 8), (4096, 3), (7, 4), (131072, 8)]:
        owned = []
        for r in range(world):
            first, n = azd.shard(total, r, world)
            owned.extend(range(first, first + n))
        assert owned == list(range(total))
        sizes = [azd.shard(total, r, world)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


class FakeNet:
    """stands in for api.Net (which needs a GPU): same export_blob / import_blob / num_params surface"""
    def __init__(self, n, seed):
        self.blob = np.random.default_rng(seed).standard_normal(n).astype(np.float32)

    def num_params(self):
        return self.blob.size

    def export_blob(self):
        return self.blob.copy()

    def import_blob(self, b):
        self.blob = np.array(b, np.float32)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = FakeNet(5949 + 13, seed=100 + rank)          # different "random-init" weights per rank before the broadcast
        sent = azd.broadcast_weights(net, dist, src=0)
        first, n = azd.shard(4096, rank, world)
        counters = dict(steps=n * 10, games=rank + 1, wins=[rank, 1], draws=0, illegal=0, sims=n * 64, evals=n * 65, errors=0)
        total = azd.reduce_counters(counters, dist)
        tmax = azd.max_over_ranks([10.0 + rank, 5.0 - rank], dist)
        q.put((rank, net.blob.tobytes(), sent.tobytes(), total, tmax, first, n))
    finally:
        dist.destroy_process_group()


def test_two_rank_broadcast_and_reduce():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = FakeNet(5949 + 13, seed=100).blob.tobytes()
    assert res[0][1] == want and res[1][1] == want and res[1][2] == want      # rank 1 now holds rank 0's weights
    for r in res:
        assert r[3] == dict(steps=4096 * 10, games=3, draws=0, illegal=0, sims=4096 * 64, evals=4096 * 65, errors=0, path_nodes=0, wins=[1, 2])
        assert r[4] == [11.0, 5.0]
    assert (res[0][5], res[0][6], res[1][5], res[1][6]) == (0, 2048, 2048, 2048)
