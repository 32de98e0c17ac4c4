"""Multi-GPU plumbing of the self-play path: one process per GPU, games sharded by contiguous global id.

The product's own entry points for the two exchanges are az_dist_* in libaz_b200.so (csrc/az_dist.cu, NCCL; api.Dist) — that is
what bench.py uses on the GPU box.  This module holds the sharding rule, the max-over-ranks reduction of timings, and the same two
exchanges over torch.distributed, which also runs on gloo so that the N > 1 host logic is testable without a GPU
(tests/test_dist_cpu.py).  Both exchanges are OUTSIDE the search loop:

  * broadcast_weights — rank 0's fp32 weight blob to every rank; replaces the reference's checkpoint-file hand-off
    (nn[0] saves temp.bin, the other GPUs restore it: neural_network/alphazero_gpu_cluster.cpp:221-231);
  * reduce_counters   — sum of the per-rank game / simulation counters; replaces GameResults::add after thread::join
    (game/game.cpp:298-309).

There is no collective on the data path (the reference has none either): games never migrate and every game's random
stream is keyed by its GLOBAL id, so results do not depend on the number of ranks.
"""
import numpy as np

COUNTER_KEYS = ("steps", "games", "wins0", "wins1", "draws", "illegal", "sims", "evals", "errors", "path_nodes")


def shard(n_total, rank, world):
    """contiguous block of global game ids owned by `rank`: (first_game_id, n_games); blocks differ by at most one game"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def _device_for(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_blob(blob, dist, src=0):
    """broadcast a flat fp32 numpy array from `src`; every rank passes an array of the same length (contents ignored off-src)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(blob, np.float32).copy()).to(_device_for(dist))
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def broadcast_weights(net, dist, src=0):
    """rank `src` exports its weights (az_nn_export_blob), every other rank imports them (az_nn_import_blob)"""
    rank = dist.get_rank()
    blob = net.export_blob() if rank == src else np.zeros(net.num_params(), np.float32)
    out = broadcast_blob(blob, dist, src)
    if rank != src:
        net.import_blob(out)
    return out


def flatten_counters(c):
    """az_counters dict (api.AzCounters.as_dict + 'errors') -> fixed-order float64 vector"""
    w = c.get("wins", [0, 0])
    vals = dict(c, wins0=w[0], wins1=w[1])
    return np.array([float(vals.get(k, 0)) for k in COUNTER_KEYS], np.float64)


def reduce_counters(c, dist):
    """sum the counters of all ranks (every rank gets the total)"""
    import torch
    t = torch.from_numpy(flatten_counters(c)).to(_device_for(dist))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    v = t.cpu().numpy()
    out = {k: int(v[i]) for i, k in enumerate(COUNTER_KEYS) if not k.startswith("wins")}
    out["wins"] = [int(v[COUNTER_KEYS.index("wins0")]), int(v[COUNTER_KEYS.index("wins1")])]
    return out


def max_over_ranks(values, dist):
    """element-wise max of a list of floats over all ranks (device timings are reported as the max over ranks)"""
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=_device_for(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]
