"""Multi-GPU partition of BASELINE configs[3] through the product path (sharding.lpt_partition + vecalign_batch(seeds) +
sharding.gather_in_order): the sharded run reproduces the single-batch records bit for bit, and a 512-pair sample of
the corpus equals the oracle (all host cores).  The two-process test needs two GPUs and is skipped on a one-GPU box;
the same partition logic is exercised there by aligning the shards one after the other on the single GPU."""
import math
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, same_alignments

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

A, K, DIMS = 6, 5, 1024
ARGS = (0.2, math.ceil(K / 2) + 5, 300, 20000, 100)


def _corpus(npairs, dim=DIMS):
    """config-4 lengths (synth.batch_sizes, seed 1234) and fp16-rounded embeddings, pair id -> data seed"""
    from speech_vecalign_b200 import synth
    n0, n1 = synth.batch_sizes(npairs, seed=1234)
    return n0, n1


def _pair(g, n0, n1, dim=DIMS):
    from speech_vecalign_b200 import synth
    v0, v1 = synth.synth_pair(int(n0[g]), int(n1[g]), K, dim=dim, seed=7_000_000 + g)
    return v0.astype(np.float16).astype(np.float32), v1.astype(np.float16).astype(np.float32)


def _align(svb, ids, n0, n1, dim=DIMS):
    types = svb.make_alignment_types(A)
    pairs = [_pair(int(g), n0, n1, dim) for g in ids]
    out = svb.vecalign_batch(pairs, types, *ARGS, output="records", seeds=[int(g) for g in ids])
    return [(o["recs"].tobytes(), o["nrecs"], o["status"]) for o in out]


def test_sharded_equals_single_batch_on_one_gpu(svb):
    """256 pairs of the corpus: one batch vs the LPT shards of 2, 4 and 8 ranks aligned shard by shard and gathered
    in input order - identical bytes.  (Per-pair seeds make every pair independent of the partition.)"""
    from speech_vecalign_b200.sharding import estimate_work, gather_in_order, lpt_partition
    npairs = 256
    n0, n1 = _corpus(npairs)
    whole = _align(svb, range(npairs), n0, n1, dim=256)
    assert all(w[2] == 0 for w in whole)
    for world in (2, 8):
        shards = lpt_partition(estimate_work(n0, n1, A), world)
        assert sorted(int(i) for s in shards for i in s) == list(range(npairs))
        out = [None] * npairs
        for ids in shards:
            for g, r in zip(ids, _align(svb, ids, n0, n1, dim=256)):
                out[int(g)] = r
        assert out == whole
    load = [estimate_work(n0, n1, A)[s].sum() for s in lpt_partition(estimate_work(n0, n1, A), 8)]
    assert max(load) / min(load) < 1.05                         # length-balanced


def _oracle_worker(job):
    g, n0, n1 = job
    sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    from threadpoolctl import threadpool_limits
    from oracle import vecalign_oracle as vo
    with threadpool_limits(1):
        v0, v1 = _pair(g, n0, n1)
        np.random.seed(g)
        st = vo.vecalign(v0, v1, vo.alignment_types(A), *ARGS, fast_host=True)
    return g, [(list(x), list(y)) for x, y in st[0]["final_alignments"]], np.asarray(st[0]["alignment_scores"]), \
        [float(st[d]["del_penalty"]) for d in sorted(st)]


def test_config4_512_pairs_equal_the_oracle(svb, oracle):
    """512 pairs of the config-4 corpus (full dimension), one vecalign_batch call; the oracle runs on all host cores."""
    from speech_vecalign_b200.engine import records_to_alignments
    npairs = 512
    n0, n1 = _corpus(npairs)
    types = svb.make_alignment_types(A)
    got = {}
    for lo in range(0, npairs, 128):                           # host memory: 128 pairs = 2.7 GB of fp32 at a time
        ids = range(lo, lo + 128)
        pairs = [_pair(g, n0, n1) for g in ids]
        for g, o in zip(ids, svb.vecalign_batch(pairs, types, *ARGS, output="records", seeds=list(ids))):
            got[g] = o
    ctx = mp.get_context("fork")
    with ctx.Pool(len(os.sched_getaffinity(0))) as pool:
        refs = pool.map(_oracle_worker, [(g, n0, n1) for g in range(npairs)], chunksize=4)
    ties = 0
    for g, al_ref, sc_ref, pens in refs:
        o = got[g]
        assert o["status"] == 0
        al, sc = records_to_alignments(o["recs"])
        if not same_alignments(al, al_ref):
            # only a demonstrated DeletionKnob tie may move an alignment (tests/test_gpu_random_configs.py)
            mine = {d: float(p) for d, p in enumerate(o["del_penalty"])}
            assert any(abs(mine[d] - pens[d]) > 1e-6 * max(1.0, abs(pens[d])) for d in mine), g
            v0, v1 = _pair(g, n0, n1)
            np.random.seed(g)
            st = oracle.vecalign(v0, v1, oracle.alignment_types(A), *ARGS, fast_host=True, penalties=mine)
            assert same_alignments(al, st[0]["final_alignments"]), g
            ties += 1
            continue
        assert np.max(np.abs(sc - sc_ref), initial=0) <= 1e-4, g
    assert ties <= 2, ties


def _rank_main(rank, world, port, npairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      LOCAL_WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import speech_vecalign_b200 as svb
    from speech_vecalign_b200.sharding import align_sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    n0, n1 = _corpus(npairs)
    types = svb.make_alignment_types(A)

    def align_fn(pairs, seeds):
        out = svb.vecalign_batch(pairs, types, *ARGS, output="records", seeds=seeds)
        return [(o["recs"].tobytes(), o["nrecs"], o["status"]) for o in out]

    res = align_sharded(list(zip(n0, n1)), lambda g: _pair(g, n0, n1, 256), align_fn, A, list(range(npairs)))
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_run_equals_single_gpu(svb):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    npairs = 256
    n0, n1 = _corpus(npairs)
    whole = _align(svb, range(npairs), n0, n1, dim=256)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, 29533, npairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == whole
