"""CPU, world_size 2 over gloo: the multi-GPU path shards by document pair with no data-path
collective.  The shard driver (speech_vecalign_b200.sharding) is exercised with the
oracle standing in for the per-rank aligner: sharded result == serial result, input order
restored, independent of the partition thanks to per-pair RNG seeds."""
import math
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


SHAPES = [(60, 66), (210, 190), (35, 40), (120, 118), (320, 300), (15, 9), (88, 95)]
A = 4


def _align_with_oracle(pairs, seeds):
    from oracle import vecalign_oracle as vo
    k = A - 1
    out = []
    for (v0, v1), s in zip(pairs, seeds):
        np.random.seed(int(s))
        st = vo.vecalign(v0, v1, vo.alignment_types(A), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100, fast_host=True)
        out.append({"alignments": [(list(x), list(y)) for x, y in st[0]["final_alignments"]],
                    "scores": st[0]["alignment_scores"].tolist()})
    return out


def _load(i):
    from speech_vecalign_b200 import synth
    return synth.synth_pair(SHAPES[i][0], SHAPES[i][1], A - 1, dim=128, seed=500 + i)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from speech_vecalign_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = [9000 + i for i in range(len(SHAPES))]
    res = sharding.align_sharded(SHAPES, _load, _align_with_oracle, A, seeds)
    shards = sharding.lpt_partition(sharding.estimate_work([s[0] for s in SHAPES], [s[1] for s in SHAPES], A), world)
    q.put((rank, res, [s.tolist() for s in shards]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_serial():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, res, shards = q.get(timeout=300)
        got[rank] = (res, shards)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1][0] is None                        # results live on rank 0 only
    res, shards = got[0]
    assert sorted(shards[0] + shards[1]) == list(range(len(SHAPES)))
    assert len(shards[0]) and len(shards[1])
    serial = _align_with_oracle([_load(i) for i in range(len(SHAPES))], [9000 + i for i in range(len(SHAPES))])
    assert len(res) == len(serial)
    for r, s in zip(res, serial):
        assert r["alignments"] == s["alignments"] and r["scores"] == s["scores"]


def test_lpt_partition_is_balanced():
    from speech_vecalign_b200 import sharding, synth
    n0, n1 = synth.batch_sizes(8192)
    work = sharding.estimate_work(n0, n1, 6)
    for world in (2, 4, 8):
        shards = sharding.lpt_partition(work, world)
        loads = np.array([work[s].sum() for s in shards])
        assert np.concatenate(shards).size == 8192 and len(set(np.concatenate(shards).tolist())) == 8192
        assert loads.max() / loads.mean() < 1.001
        for s in shards:
            assert np.all(np.diff(s) > 0)          # input order kept inside a shard


def test_gather_without_process_group_restores_order():
    from speech_vecalign_b200 import sharding
    out = sharding.gather_in_order(["c", "a"], [2, 0], 3)
    assert out == ["a", None, "c"]
