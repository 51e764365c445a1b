"""Multi-GPU partition of a batch of document pairs.

Document pairs are closed computations (seg_align/align.py:206-230 aligns them one by one), so
the path shards by pair with NO data-path collective: each rank aligns its shard on its own GPU
and the per-pair results (a few KB) are gathered on the host.  The reference's own sharding helper
(utils/mp_utils.py:7-16 get_shard_range) cuts contiguous ranges; here pairs are length-balanced
with the LPT rule on an estimate of the per-pair work (bench.py's config-4 arm, the seg_align driver
and align_sharded below all partition with it).

Host RNG: with ``seeds`` (one per pair) every pair draws from its own np.random stream, so results
do not depend on the partition; without seeds the caller must draw in input order on one rank.
"""
import numpy as np


def estimate_work(n0, n1, alignment_max_size, search_buffer_size=5, dim=1024, max_size_full_dp=300):
    """~FLOPs + bytes proxy per pair: banded dots T*(n0+n1+3)*B*D at level 0, (n0+n1)*B*D for the
    coarser levels together, the prologue's K*(n0+n1)*D row traffic and the dense level."""
    n0 = np.asarray(n0, dtype=np.float64)
    n1 = np.asarray(n1, dtype=np.float64)
    k = alignment_max_size - 1
    t = alignment_max_size * (alignment_max_size - 1) / 2
    band = 2 * (np.ceil(k / 2) + search_buffer_size)
    return (t + 1) * (n0 + n1 + 3) * band * dim + 8 * k * (n0 + n1) * dim + float(max_size_full_dp) ** 2 * dim


def lpt_partition(work, nranks):
    """Longest-processing-time-first: returns a list of index arrays, one per rank (each sorted
    ascending so that a rank keeps input order inside its shard)."""
    work = np.asarray(work, dtype=np.float64)
    order = np.argsort(-work, kind="stable")
    load = np.zeros(nranks)
    shards = [[] for _ in range(nranks)]
    for i in order:
        r = int(np.argmin(load))
        shards[r].append(int(i))
        load[r] += work[i]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def gather_in_order(local_results, local_indices, total, group=None, dst=0):
    """Host-side gather of per-pair results to rank `dst`, restored to input order.
    Uses torch.distributed object gather (NCCL or gloo); returns None on the other ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        out = [None] * total
        for i, r in zip(local_indices, local_results):
            out[int(i)] = r
        return out
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    payload = (list(map(int, local_indices)), local_results)
    bucket = [None] * world if rank == dst else None
    dist.gather_object(payload, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    out = [None] * total
    for idx, res in bucket:
        for i, r in zip(idx, res):
            out[i] = r
    return out


def align_sharded(pairs_meta, load_pair, align_fn, alignment_max_size, seeds, group=None):
    """Generic driver: `pairs_meta` = list of (n0, n1); `load_pair(i)` yields the pair's tensors on
    this rank; `align_fn(list_of_pairs, seeds)` aligns a shard.  Returns the ordered result list on
    rank 0 (None elsewhere)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n0 = [m[0] for m in pairs_meta]
    n1 = [m[1] for m in pairs_meta]
    shards = lpt_partition(estimate_work(n0, n1, alignment_max_size), world)
    mine = shards[rank]
    res = align_fn([load_pair(int(i)) for i in mine], [seeds[int(i)] for i in mine]) if len(mine) else []
    return gather_in_order(res, mine, len(pairs_meta), group=group)
