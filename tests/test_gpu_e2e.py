"""GPU parity, end to end: speech_vecalign_b200.vecalign / vecalign_batch against the oracle's
vecalign on the same seeded inputs and the same np.random stream.

Bars: final alignments and every level's search path bit-exact; level vectors
bit-exact (numpy-order normalise/downsample); norms <= 2.4e-7 abs (sgemm order is CPU dependent);
del_penalty <= 1e-6 relative; alignment scores <= 1e-4 abs; backpointers identical on the path and
>= 99.9 % identical over the whole band (cost noise of ~1e-7 can flip near-ties off the path).
"""
import math

import numpy as np
import pytest

from conftest import same_alignments

pytestmark = pytest.mark.gpu


def _run_both(svb, oracle, n0, n1, a, seed, rng_seed=0, mode="exact", **kw):
    from speech_vecalign_b200 import synth
    k = a - 1
    v0, v1 = synth.synth_pair(n0, n1, k, dim=kw.get("dim", 1024), seed=seed)
    types = oracle.alignment_types(a)
    w = math.ceil(k / 2) + kw.get("search_buffer_size", 5)
    args = (types, 0.2, w, kw.get("max_size_full_dp", 300), kw.get("costs_sample_size", 20000), 100)
    np.random.seed(rng_seed)
    ref = oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True)
    ref_state = np.random.get_state()[1].copy()
    np.random.seed(rng_seed)
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), *args, cost_mode=mode, debug=True)
    got_state = np.random.get_state()[1].copy()
    assert np.array_equal(ref_state, got_state), "np.random stream consumed differently from the reference"
    return ref, got


def _compare(ref, got, exact_costs_tol=5e-6):
    assert set(ref.keys()) == set(got.keys())
    for d in sorted(ref.keys(), reverse=True):
        r, g = ref[d], got[d]
        assert g["size0"] == r["size0"] and g["size1"] == r["size1"]
        assert np.array_equal(g["v0"], r["v0"]) and np.array_equal(g["v1"], r["v1"]), f"level {d} vectors"
        assert np.max(np.abs(g["n0"].astype(np.float64) - r["n0"]), initial=0) <= 2.4e-7, f"level {d} n0"
        assert np.max(np.abs(g["n1"].astype(np.float64) - r["n1"]), initial=0) <= 2.4e-7, f"level {d} n1"
        assert abs(g["del_penalty"] - r["del_penalty"]) <= 1e-6 * max(1.0, abs(r["del_penalty"])), f"level {d} del_penalty"
        if "costs_1to1" in r:
            assert np.allclose(g["costs_1to1"], r["costs_1to1"], rtol=0, atol=exact_costs_tol)
            assert same_alignments(g["alignments"], r["alignments"]), f"level {d} dense alignments"
        if "searchpath" in r:
            assert g["searchpath"] == [tuple(p) for p in r["searchpath"]], f"level {d} search path"
            assert np.array_equal(g["b_offset"], r["b_offset"]) and np.array_equal(g["new_b_offset"], r["new_b_offset"])
            fin = np.isfinite(r["a_b_costs"])
            assert np.array_equal(np.isfinite(g["a_b_costs"]), fin)
            assert np.max(np.abs(g["a_b_costs"][fin] - r["a_b_costs"][fin]), initial=0) <= exact_costs_tol * 40
            same = (g["a_b_xp"] == r["a_b_xp"]) & (g["a_b_yp"] == r["a_b_yp"])
            assert same.mean() >= 0.999, f"level {d}: {100 * (1 - same.mean()):.3f}% backpointers differ"
            key = "final_alignments" if d == 0 else "alignments"
            assert same_alignments(g[key], r[key]), f"level {d} alignments"
            assert np.max(np.abs(g["alignment_scores"] - r["alignment_scores"]), initial=0) <= 1e-4


@pytest.mark.parametrize("n0,n1,a,seed", [
    (237, 217, 4, 1), (150, 160, 6, 2), (310, 305, 5, 3), (700, 650, 5, 4), (800, 817, 6, 5),
    (489, 500, 6, 6), (1200, 1100, 8, 7), (2000, 2000, 5, 8)])
def test_vecalign_matches_oracle(svb, oracle, n0, n1, a, seed):
    ref, got = _run_both(svb, oracle, n0, n1, a, seed)
    _compare(ref, got)


@pytest.mark.parametrize("n0,n1", [(1, 1), (2, 1), (1, 5), (0, 5), (5, 0), (0, 0), (3, 2000), (5, 301), (299, 302),
                                   (1, 100000), (2, 50000)])
def test_edge_sizes(svb, oracle, n0, n1):
    """none of these raise in the reference."""
    ref, got = _run_both(svb, oracle, n0, n1, 4, seed=n0 * 7 + n1, dim=128)
    assert same_alignments(got[0]["final_alignments"], ref[0]["final_alignments"])
    assert np.allclose(got[0]["alignment_scores"], ref[0]["alignment_scores"], rtol=0, atol=1e-4)
    for d in ref:
        assert abs(got[d]["del_penalty"] - ref[d]["del_penalty"]) <= 1e-6


def test_fast_mode_same_path(svb, oracle):
    ref, got = _run_both(svb, oracle, 600, 640, 6, seed=21, mode="fast")
    assert same_alignments(got[0]["final_alignments"], ref[0]["final_alignments"])
    assert np.max(np.abs(got[0]["alignment_scores"] - ref[0]["alignment_scores"])) <= 1e-4


def test_small_parameters(svb, oracle):
    """non-default knobs: search_buffer_size, max_size_full_dp, costs_sample_size (full-grid branch)."""
    ref, got = _run_both(svb, oracle, 130, 120, 4, seed=33, search_buffer_size=2, max_size_full_dp=40,
                         costs_sample_size=50000)
    _compare(ref, got)


def test_batch_equals_serial_loop(svb, oracle):
    """vecalign_batch consumes np.random pair by pair in input order: same results as a serial loop
    of the reference (seg_align/align.py:206), whatever the sizes in the batch."""
    from speech_vecalign_b200 import synth
    shapes = [(220, 230), (640, 600), (90, 100), (301, 322), (1300, 1250)]
    a, k = 6, 5
    types = oracle.alignment_types(a)
    args = (types, 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    pairs = [synth.synth_pair(n0, n1, k, seed=100 + i) for i, (n0, n1) in enumerate(shapes)]
    np.random.seed(5)
    refs = [oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True) for v0, v1 in pairs]
    np.random.seed(5)
    gots = svb.vecalign_batch([(v0.copy(), v1.copy()) for v0, v1 in pairs], *args)
    for r, g in zip(refs, gots):
        assert same_alignments(g[0]["final_alignments"], r[0]["final_alignments"])
        assert np.max(np.abs(g[0]["alignment_scores"] - r[0]["alignment_scores"])) <= 1e-4
        assert abs(g[0]["del_penalty"] - r[0]["del_penalty"]) <= 1e-6


def test_user_norms_at_depth0(svb, oracle):
    """norms0/norms1 passed by the caller skip compute_norms (and its RNG draws) at depth 0
    (dp_utils.py:428-444) and make every level-0 cost bit-exact."""
    from speech_vecalign_b200 import synth
    v0, v1 = synth.synth_pair(260, 250, 3, seed=9)
    types = oracle.alignment_types(4)
    n0 = np.random.default_rng(1).uniform(0.8, 1.0, (3, 260)).astype(np.float32)
    n1 = np.random.default_rng(2).uniform(0.8, 1.0, (3, 250)).astype(np.float32)
    np.random.seed(2)
    ref = oracle.vecalign(v0.copy(), v1.copy(), types, 0.2, 7, 300, 20000, 100, norms0=n0, norms1=n1, fast_host=True)
    np.random.seed(2)
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), types, 0.2, 7, 300, 20000, 100, norms0=n0, norms1=n1, debug=True)
    assert np.array_equal(got[0]["a_b_costs"], ref[0]["a_b_costs"])
    assert got[0]["del_penalty"] == ref[0]["del_penalty"]
    assert np.array_equal(got[0]["a_b_csum"], ref[0]["a_b_csum"])
    assert np.array_equal(got[0]["a_b_xp"], ref[0]["a_b_xp"]) and np.array_equal(got[0]["a_b_yp"], ref[0]["a_b_yp"])
    assert same_alignments(got[0]["final_alignments"], ref[0]["final_alignments"])
    assert np.array_equal(got[0]["alignment_scores"], ref[0]["alignment_scores"])


def test_device_tensors_in_place(svb, oracle):
    """torch CUDA inputs are normalised in place, as the reference mutates its numpy inputs."""
    import torch
    from speech_vecalign_b200 import synth
    v0, v1 = synth.synth_pair(120, 125, 3, seed=4)
    t0, t1 = torch.from_numpy(v0).cuda(), torch.from_numpy(v1).cuda()
    np.random.seed(1)
    got = svb.dp_utils.vecalign(t0, t1, oracle.alignment_types(4), 0.2, 7, 300, 20000, 100)
    r0 = v0.copy()
    oracle.unit_rows(r0, fast=True)
    assert np.array_equal(t0.cpu().numpy(), r0)
    np.random.seed(1)
    ref = oracle.vecalign(v0.copy(), v1.copy(), oracle.alignment_types(4), 0.2, 7, 300, 20000, 100, fast_host=True)
    assert same_alignments(got[0]["final_alignments"], ref[0]["final_alignments"])


def test_errors_mirror_reference(svb):
    v = np.zeros((2, 5, 1024), np.float32)
    with pytest.raises(Exception, match="overlaps requrested"):
        svb.dp_utils.vecalign(v, v.copy(), [(1, 1), (3, 1)], 0.2, 7, 300, 20000, 100)
    with pytest.raises(Exception, match="norms0 wrong shape"):
        svb.dp_utils.vecalign(v, v.copy(), [(1, 1)], 0.2, 7, 300, 20000, 100, norms0=np.ones((2, 4), np.float32))
    with pytest.raises(ValueError):
        svb.dp_utils.vecalign(v.astype(np.float64), v.copy(), [(1, 1)], 0.2, 7, 300, 20000, 100)


def test_stream_groups_do_not_change_results(svb, oracle):
    """Pair groups on separate CUDA streams (engine.BatchRun.run(ngroups)) are a scheduling choice:
    bit-identical records, and still the oracle's alignments."""
    from speech_vecalign_b200 import synth
    shapes = [(220, 230), (640, 600), (90, 100), (301, 322), (700, 650), (33, 30), (410, 400), (150, 170), (820, 800)]
    a, k = 5, 4
    types = oracle.alignment_types(a)
    args = (types, 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    pairs = [synth.synth_pair(n0, n1, k, seed=300 + i) for i, (n0, n1) in enumerate(shapes)]
    seeds = [40 + i for i in range(len(shapes))]
    outs = []
    for streams in (1, 3, 9):
        outs.append(svb.vecalign_batch([(v0.copy(), v1.copy()) for v0, v1 in pairs], *args, output="records",
                                       seeds=seeds, streams=streams))
    for other in outs[1:]:
        for r0, r1 in zip(outs[0], other):
            assert np.array_equal(r0["recs"], r1["recs"]) and r0["del_penalty"] == r1["del_penalty"]
    from speech_vecalign_b200.engine import records_to_alignments
    for (v0, v1), s, r in zip(pairs, seeds, outs[1]):
        np.random.seed(s)
        ref = oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True)
        al, sc = records_to_alignments(r["recs"])
        assert same_alignments(al, ref[0]["final_alignments"])
        assert np.max(np.abs(sc - ref[0]["alignment_scores"])) <= 1e-4


def test_chunked_host_pipeline_equals_serial_loop(svb, oracle):
    """>= 16 host pairs take the chunk-pipelined path of vecalign_batch (copies queued on a copy
    stream, chunks planned in input order): the global np.random stream must still be consumed exactly
    as a serial loop of the reference consumes it."""
    from speech_vecalign_b200 import synth
    rng = np.random.default_rng(3)
    shapes = [(int(a), int(b)) for a, b in zip(rng.integers(40, 420, 19), rng.integers(40, 420, 19))]
    a, k = 4, 3
    types = oracle.alignment_types(a)
    args = (types, 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    pairs = [synth.synth_pair(n0, n1, k, dim=256, seed=900 + i) for i, (n0, n1) in enumerate(shapes)]
    np.random.seed(11)
    refs = [oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True) for v0, v1 in pairs]
    ref_state = np.random.get_state()[1].copy()
    np.random.seed(11)
    gots = svb.vecalign_batch([(v0.copy(), v1.copy()) for v0, v1 in pairs], *args)
    assert np.array_equal(ref_state, np.random.get_state()[1])
    assert len(gots) == len(refs)
    for r, g in zip(refs, gots):
        assert same_alignments(g[0]["final_alignments"], r[0]["final_alignments"])
        assert np.max(np.abs(g[0]["alignment_scores"] - r[0]["alignment_scores"]), initial=0) <= 1e-4


def test_fp16_inputs_equal_widened_fp32(svb, oracle):
    """Embeddings kept in their on-disk dtype: fp16 (K, N, D) inputs are widened on the device and must
    give exactly what the same values passed as fp32 give — single pair, small batch and the
    chunk-pipelined host path."""
    from speech_vecalign_b200 import synth
    a, k = 5, 4
    types = oracle.alignment_types(a)
    args = (types, 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    shapes = [(300 + 13 * i, 280 + 17 * i) for i in range(17)]
    p16 = [tuple(v.astype(np.float16) for v in synth.synth_pair(n0, n1, k, dim=256, seed=50 + i)) for i, (n0, n1) in enumerate(shapes)]
    p32 = [(v0.astype(np.float32), v1.astype(np.float32)) for v0, v1 in p16]
    seeds = list(range(len(shapes)))
    r16 = svb.vecalign_batch(p16, *args, output="records", seeds=seeds)
    r32 = svb.vecalign_batch(p32, *args, output="records", seeds=seeds)
    for x, y in zip(r16, r32):
        assert np.array_equal(x["recs"], y["recs"]) and x["del_penalty"] == y["del_penalty"]
    np.random.seed(3)
    one16 = svb.dp_utils.vecalign(p16[0][0], p16[0][1], *args)
    np.random.seed(3)
    ref = oracle.vecalign(p32[0][0].copy(), p32[0][1].copy(), *args, fast_host=True)
    assert same_alignments(one16[0]["final_alignments"], ref[0]["final_alignments"])


@pytest.mark.parametrize("sbs,a", [(15, 4), (30, 3)])
def test_wide_search_buffer(svb, oracle, sbs, a):
    """search_buffer_size large enough that the band (2 * width_over2) exceeds a warp: the generic
    wavefront kernel takes several band slots per lane."""
    ref, got = _run_both(svb, oracle, 340, 330, a, seed=60 + sbs, search_buffer_size=sbs, dim=256)
    _compare(ref, got)


def test_many_to_one_types(svb, oracle):
    """vecalign.py:165-171 make_many_to_one_alignment_types: a non-standard type list (generic kernels)."""
    from speech_vecalign_b200 import synth
    types = svb.make_many_to_one_alignment_types(4)
    assert types == [(1, 1), (2, 1), (3, 1), (4, 1)]
    v0, v1 = synth.synth_pair(420, 380, 4, dim=256, seed=77)
    v1 = v1[:1].copy()                                   # the target side only provides overlap 0
    w = math.ceil(4 / 2) + 5
    np.random.seed(4)
    ref = oracle.vecalign(v0.copy(), v1.copy(), types, 0.2, w, 300, 20000, 100, fast_host=True)
    np.random.seed(4)
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), types, 0.2, w, 300, 20000, 100, debug=True)
    _compare(ref, got)


def test_numpy_inputs_are_normalised_in_place_by_default(svb, oracle):
    """dp_utils.py:396-397: the reference leaves its inputs normalised; vecalign() copies the rows back
    into writable numpy arrays (also through the pinned ring for large arrays) unless writeback=False."""
    from speech_vecalign_b200 import synth
    for n0, n1, k in ((120, 125, 3), (1500, 1400, 4)):        # below / above the staging threshold
        v0, v1 = synth.synth_pair(n0, n1, k, seed=4)
        a0, a1 = v0.copy(), v1.copy()
        keep = v0.copy()
        np.random.seed(1)
        svb.dp_utils.vecalign(a0, a1, oracle.alignment_types(k + 1), 0.2, 7, 300, 20000, 100)
        r0, r1 = v0.copy(), v1.copy()
        oracle.unit_rows(r0, fast=True)
        oracle.unit_rows(r1, fast=True)
        assert np.array_equal(a0, r0) and np.array_equal(a1, r1)
        b0 = v0.copy()
        np.random.seed(1)
        svb.dp_utils.vecalign(b0, v1.copy(), oracle.alignment_types(k + 1), 0.2, 7, 300, 20000, 100, writeback=False)
        assert np.array_equal(b0, keep)


def test_production_path_levels_match_oracle(svb, oracle):
    """The non-debug path stores only overlap 0 of the coarser levels (SvxLevelJob.keep = 1) - a different prologue
    code path from the debug stack's keep = k.  Its overlap-0 rows and norms of every level are fetched from the
    arena of a production run and compared with the oracle: rows bit-exact, norms <= 2.4e-7."""
    import torch
    from speech_vecalign_b200 import synth
    a, k = 5, 4
    v0, v1 = synth.synth_pair(1300, 1210, k, seed=9)
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    np.random.seed(3)
    ref = oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True)
    np.random.seed(3)
    run = svb.vecalign_batch([(v0.copy(), v1.copy())], *args, sync=False)
    torch.cuda.synchronize()
    assert not run.keep_dense_csum and len(ref) == int(run.nlev[0]) == 4
    for d in range(len(ref)):
        r = run.level_record(0, d)
        for side, (kv, kn) in enumerate((("v0", "n0"), ("v1", "n1"))):
            s = ref[d][kv].shape[1]
            if d > 0:
                assert np.array_equal(run.fetch_vecs(r, side)[0], ref[d][kv][0]), (d, side)
            norms = run.fetch("norms0" if side == 0 else "norms1", r, (k, s), np.float32)
            rows = slice(0, k) if d == 0 else slice(0, 1)
            assert np.max(np.abs(norms[rows] - ref[d][kn][rows])) <= 2.4e-7, (d, side)
        assert abs(run.results()[0]["del_penalty"][d] - ref[d]["del_penalty"]) <= 1e-6 * max(1.0, abs(ref[d]["del_penalty"]))
    from speech_vecalign_b200.engine import records_to_alignments
    al, sc = records_to_alignments(run.results()[0]["recs"])
    assert same_alignments(al, ref[0]["final_alignments"])


@pytest.mark.parametrize("sbs", [6, 7, 8])
def test_alignment_max_size_9_wide_buffers(svb, oracle, sbs):
    """a = 9 (K = 8) with search buffers 6-8: band + K <= 32 selects the warp-per-band DP, whose chunk plan exceeds
    shared memory at these sizes - the launcher must fall through to the generic kernel (round-1 advisor finding)."""
    ref, got = _run_both(svb, oracle, 260, 250, 9, seed=90 + sbs, search_buffer_size=sbs, dim=128)
    _compare(ref, got)


def test_many_to_one_50_default_of_the_cli(svb, oracle):
    """The reference CLI's --many_to_one default (const=50: 50 types, width_over2 = 30, band 60) runs through the
    generic cost and DP kernels (round-1 advisor finding: bands >= 46 were rejected)."""
    from speech_vecalign_b200 import synth
    m = 50
    types = svb.make_many_to_one_alignment_types(m)
    v0, v1 = synth.synth_pair(150, 40, m, dim=128, seed=78)
    v1 = v1[:1].copy()
    w = math.ceil(m / 2) + 5
    np.random.seed(4)
    ref = oracle.vecalign(v0.copy(), v1.copy(), types, 0.2, w, 300, 20000, 100, fast_host=True)
    np.random.seed(4)
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), types, 0.2, w, 300, 20000, 100, debug=True)
    _compare(ref, got)


def test_num_samps_for_norm_above_fused_limit(svb, oracle):
    """num_samps_for_norm > 2048 exceeds the fused prologue's scratch: the planner switches to the unfused launchers."""
    from speech_vecalign_b200 import synth
    v0, v1 = synth.synth_pair(400, 390, 3, dim=128, seed=5)
    args = (oracle.alignment_types(4), 0.2, 7, 300, 20000, 2500)
    np.random.seed(2)
    ref = oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True)
    np.random.seed(2)
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), *args)
    assert same_alignments(got[0]["final_alignments"], ref[0]["final_alignments"])
    assert np.max(np.abs(got[0]["alignment_scores"] - ref[0]["alignment_scores"])) <= 1e-4
