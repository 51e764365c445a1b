"""CPU: the C-ABI library loads and exports everything include/svx.h declares; the host twins of
the DP kernels (identical source, compiled for the CPU) reproduce the reference's known answers;
the batch planner replays the reference's RNG stream; the product refuses to run without CUDA.
No compute call touches a GPU here."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, same_alignments


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "functions.npz"))


def test_library_exports_every_declared_symbol(svb):
    hdr = open(os.path.join(ROOT, "include", "svx.h")).read()
    declared = re.findall(r"SVX_API\s+[\w\s\*]+?\b(svx_\w+)\s*\(", hdr)
    assert len(declared) >= 17
    L = ctypes.CDLL(svb.capi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/svx.h but not exported by libsvx.so"
    assert sorted(declared) == sorted(svb.capi.EXPORTED_SYMBOLS)
    assert svb.capi.lib().svx_version() == 100


def test_struct_layouts_match(svb):
    L = svb.capi.lib()
    for i, dt in enumerate([svb.capi.ROWS, svb.capi.DOWN, svb.capi.NORM, svb.capi.SCORE, svb.capi.DENSE,
                            svb.capi.BAND, svb.capi.REC, svb.capi.LEVEL]):
        assert L.svx_sizeof_job(i) == dt.itemsize


def test_no_cpu_fallback(svb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    v = np.zeros((2, 5, 1024), np.float32)
    with pytest.raises(svb.capi.SvxError, match="no CUDA device"):
        svb.dp_utils.vecalign(v, v.copy(), [(1, 1)], 0.2, 7, 300, 20000, 100)
    with pytest.raises(svb.capi.SvxError, match="no CUDA device"):
        svb.dp_core.make_norm1(v)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "speech-vecalign_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f


# ------------------------------------------------------------------------------------------------
def test_host_del_knob_bit_exact(svb, gold):
    L = svb.capi.lib()
    sc = np.ascontiguousarray(gold["score_out"])
    for frac, want in zip(gold["knob_fracs"], gold["knob_pens"]):
        out = np.zeros(1, np.float64)
        assert L.svx_host_del_knob(sc.ctypes.data, sc.shape[0], float(frac), out.ctypes.data) == 0
        assert out[0] == want, (frac, out[0], want)


@pytest.mark.parametrize("seed,n", [(0, 20000), (1, 777), (2, 3), (3, 1)])
def test_host_del_knob_random(svb, oracle, seed, n):
    rng = np.random.default_rng(seed)
    sc = np.abs(rng.normal(0.9, 0.3, n)).astype(np.float32)
    knob = oracle.PercentileKnob(sc, 0, max(sc))
    for frac in (0.0, 0.02, 0.2, 0.33, 0.8, 1.0):
        out = np.zeros(1, np.float64)
        svb.capi.lib().svx_host_del_knob(sc.ctypes.data, n, frac, out.ctypes.data)
        assert out[0] == knob.percentile_frac_to_del_penalty(frac)


def _dense_job(svb, costs, pen, t0, t1, ups):
    s0, s1 = costs.shape
    keep = {}
    keep["costs"] = np.ascontiguousarray(costs, dtype=np.float32)
    keep["pen"] = np.array([pen], np.float64)
    keep["bp"] = np.zeros((s0 + 1, s1 + 1), np.uint8)
    keep["csum"] = np.zeros((s0 + 1, s1 + 1), np.float64)
    plen = svb.capi.lib().svx_path_len(s0, s1, t0, t1, ups)
    keep["ypath"] = np.full(max(plen, 1), -1, np.int32)
    keep["status"] = np.zeros(1, np.int32)
    job = np.zeros(1, dtype=svb.capi.DENSE)
    job["costs"], job["del_penalty"], job["bp"], job["csum"] = (keep[k].ctypes.data for k in ("costs", "pen", "bp", "csum"))
    job["ypath"], job["status_d"] = keep["ypath"].ctypes.data, keep["status"].ctypes.data
    job["s0"], job["s1"], job["t0"], job["t1"], job["upsample"], job["path_len"] = s0, s1, t0, t1, ups, plen
    return job, keep, plen


def test_host_dense_dp_known_answer(svb, gold):
    s0, s1 = gold["dense_costs"].shape
    job, keep, plen = _dense_job(svb, gold["dense_costs"], float(gold["pen"][0]), s0, s1, 0)
    assert svb.capi.lib().svx_host_dense_dp(job.ctypes.data) == 0
    assert keep["status"][0] == 0
    assert np.array_equal(keep["csum"], gold["dense_csum"])
    assert np.array_equal(keep["bp"].astype(np.int32), gold["dense_bp"])
    path = gold["path_same"]
    assert plen == path.shape[0]
    assert np.array_equal(keep["ypath"], path[:, 1]) and np.array_equal(np.arange(plen) - keep["ypath"], path[:, 0])


def _band_job(svb, costs_tab, boff, types, pen, s0, s1, target=None):
    T, A, B = costs_tab.shape
    w = B // 2
    keep = {"costs": np.ascontiguousarray(costs_tab.transpose(1, 0, 2)), "ypath": (boff + w).astype(np.int32),
            "pen": np.array([pen], np.float64), "bp": np.zeros((A + 2, B), np.uint8),
            "csum": np.zeros((A + 2, B), np.float64), "recs": np.zeros(s0 + s1 + 2, dtype=svb.capi.REC),
            "nrecs": np.zeros(1, np.int32), "status": np.zeros(1, np.int32)}
    job = np.zeros(1, dtype=svb.capi.BAND)
    for f, k in (("costs", "costs"), ("ypath", "ypath"), ("del_penalty", "pen"), ("bp", "bp"), ("csum", "csum"),
                 ("recs", "recs"), ("nrecs", "nrecs"), ("status_d", "status")):
        job[f] = keep[k].ctypes.data
    job["s0"], job["s1"], job["a_len"], job["band"], job["width_over2"], job["ntypes"] = s0, s1, A, B, w, T
    job["rec_cap"] = s0 + s1 + 2
    for t, (x, y) in enumerate(types):
        job["xo"][0, t], job["yo"][0, t] = x, y
    job["amax"] = max([2] + [x + y for x, y in types])
    if target is not None:
        nlen = svb.capi.lib().svx_path_len(s0, s1, target[0], target[1], 1)
        keep["next"] = np.full(nlen, -1, np.int32)
        job["next_ypath"], job["t0"], job["t1"], job["next_len"] = keep["next"].ctypes.data, target[0], target[1], nlen
    return job, keep


def _xy_of(svb, bp, types):
    tx = np.array([x for x, _ in types] + [0, 1], np.int32)
    ty = np.array([y for _, y in types] + [1, 0], np.int32)
    xp = np.full(bp.shape, -42, np.int32)
    yp = np.full(bp.shape, -42, np.int32)
    ok = bp != svb.capi.SVX_BP_NONE
    xp[ok], yp[ok] = tx[bp[ok]], ty[bp[ok]]
    return xp, yp


@pytest.mark.parametrize("tag", [None, "even", "odd"])
def test_host_banded_dp_known_answer(svb, oracle, gold, tag):
    types = oracle.alignment_types(4)
    s0, s1 = gold["unit0"].shape[1], gold["unit1"].shape[1]
    target = None if tag is None else ((2 * s0, 2 * s1) if tag == "even" else (2 * s0 + 1, 2 * s1 + 1))
    job, keep = _band_job(svb, gold["sparse_costs"], gold["b_offset"], types, float(gold["sparse_pen"][0]), s0, s1, target)
    assert svb.capi.lib().svx_host_banded_dp(job.ctypes.data) == 0
    assert keep["status"][0] == 0
    assert np.array_equal(keep["csum"], gold["sparse_csum"])
    xp, yp = _xy_of(svb, keep["bp"], types)
    assert np.array_equal(xp, gold["sparse_xp"]) and np.array_equal(yp, gold["sparse_yp"])
    n = int(keep["nrecs"][0])
    recs = keep["recs"][len(keep["recs"]) - n:]
    assert np.array_equal(recs["nx"], gold["trace_nx"]) and np.array_equal(recs["ny"], gold["trace_ny"])
    assert np.array_equal(recs["score"], gold["trace_scores"])
    has_x = gold["trace_nx"] > 0
    assert np.array_equal(recs["x_end"][has_x], gold["trace_x_end"][has_x])
    if tag is not None:
        path = gold[f"path_up_{tag}"]
        assert np.array_equal(keep["next"], path[:, 1])


@pytest.mark.parametrize("seed,s0,s1,a,w", [(0, 40, 37, 4, 7), (1, 25, 60, 6, 8), (2, 64, 64, 8, 9), (3, 3, 9, 3, 3), (4, 1, 1, 2, 3)])
def test_host_twins_vs_oracle_random(svb, oracle, ocore, seed, s0, s1, a, w):
    """Dense twin -> same-level path -> banded twin -> upsampled path, against the oracle chain."""
    rng = np.random.default_rng(seed)
    k = a - 1
    v0 = rng.standard_normal((k, s0, 128)).astype(np.float32)
    v1 = rng.standard_normal((k, s1, 128)).astype(np.float32)
    oracle.unit_rows(v0, fast=True)
    oracle.unit_rows(v1, fast=True)
    n0 = rng.uniform(0.7, 1.1, (k, s0)).astype(np.float32)
    n1 = rng.uniform(0.7, 1.1, (k, s1)).astype(np.float32)
    pen = 0.41
    costs = ocore.make_dense_costs(v0, v1, n0, n1)
    cs, bp = ocore.dense_dp(costs, pen)
    job, keep, plen = _dense_job(svb, costs, pen, s0, s1, 0)
    assert svb.capi.lib().svx_host_dense_dp(job.ctypes.data) == 0
    assert np.array_equal(keep["csum"], cs) and np.array_equal(keep["bp"].astype(np.int32), bp)
    path = oracle.search_path(oracle.dense_backtrace(bp))
    assert [(i - int(y), int(y)) for i, y in enumerate(keep["ypath"])] == path
    types = oracle.alignment_types(a)
    feats, boff = ocore.make_sparse_costs(v0, v1, n0, n1, path, types, w)
    csum, xp, yp, nbo = ocore.sparse_dp(feats, boff, types, pen, s0, s1)
    al, sc = oracle.banded_backtrace(csum, xp, yp, nbo, s0, s1)
    t0, t1 = 2 * s0 + 1, 2 * s1
    job, keep = _band_job(svb, feats, boff, types, pen, s0, s1, (t0, t1))
    assert svb.capi.lib().svx_host_banded_dp(job.ctypes.data) == 0
    assert keep["status"][0] == 0
    assert np.array_equal(keep["csum"], csum)
    gx, gy = _xy_of(svb, keep["bp"], types)
    assert np.array_equal(gx, xp) and np.array_equal(gy, yp)
    n = int(keep["nrecs"][0])
    from speech_vecalign_b200.engine import records_to_alignments
    gal, gsc = records_to_alignments(keep["recs"][len(keep["recs"]) - n:])
    assert same_alignments(gal, al) and np.array_equal(gsc, sc)
    up = oracle.double_resolution(al)
    oracle.extend_to(up, t0, t1)
    want = oracle.search_path(up)
    assert [(i - int(y), int(y)) for i, y in enumerate(keep["next"])] == want


@pytest.mark.parametrize("c0,c1,t0,t1", [(5, 7, 10, 14), (5, 7, 11, 15), (0, 3, 1, 7), (0, 0, 1, 1), (150, 151, 301, 303)])
def test_path_len_matches_reference_glue(svb, oracle, c0, c1, t0, t1):
    al = [([i], [i]) for i in range(min(c0, c1))]
    al += [([i], []) for i in range(min(c0, c1), c0)] + [([], [i]) for i in range(min(c0, c1), c1)]
    up = oracle.double_resolution(al)
    oracle.extend_to(up, t0, t1)
    assert svb.capi.lib().svx_path_len(c0, c1, t0, t1, 1) == len(oracle.search_path(up))
    assert svb.capi.lib().svx_path_len(c0, c1, c0, c1, 0) == len(oracle.search_path(al))


# ------------------------------------------------------------------------------------------------
# The planner lives in libsvx.so (csrc/plan.cu, svx_plan_*): it is host code until the plan is bound to
# device memory, so everything below runs without a GPU.
# ------------------------------------------------------------------------------------------------
def _plan(svb, n0, n1, k0, k1, a_types, sample_size=20000, nsfn=100, full=300, w=7, dim=128):
    from speech_vecalign_b200 import engine
    prm = engine.make_params(k0, k1, dim, a_types, 0.2, w, full, sample_size, nsfn)
    return engine.Plan(prm, n0, n1)


def test_planner_level_sizes(svb, oracle):
    """dp_utils.py:403-408 level sizes and the search-path lengths that follow from the path glue."""
    pl = _plan(svb, [237, 2000, 20000, 0, 5, 299], [217, 2000, 20000, 5, 301, 302], 4, 4, oracle.alignment_types(5))
    depth, first, rs0, rs1, A = (pl.array(k) for k in ("depth", "first", "rs0", "rs1", "A"))
    assert list(depth) == [0, 3, 7, 0, 0, 1]
    assert list(rs0[first[1]:first[1] + 4]) == [2000, 1000, 500, 250] and rs0[first[2] + 7] == 156
    assert rs0[first[5] + 1] == 149 and rs1[first[5] + 1] == 151
    # banded levels: A = len(search path built from the next coarser level); the coarsest level has none
    L = svb.capi.lib()
    assert A[first[0]] == L.svx_path_len(237, 217, 237, 217, 0)
    assert A[first[1]] == L.svx_path_len(1000, 1000, 2000, 2000, 1) == 4003 and A[first[1] + 3] == 0
    assert A[first[5]] == L.svx_path_len(149, 151, 299, 302, 1)
    assert int(pl.info["band"]) == 14 and int(pl.info["per0"]) == 25


def _python_draws(rs0, rs1, first, nlev, k0, k1, nsfn, ss, rs):
    """The reference's call order, spelled out: per pair, for each level the n0 draws (K1 calls
    over range(size1), dp_utils.py:346), then the n1 draws; then per level the knob draws x, y (:301-302)."""
    per1, per0 = -(-nsfn // k1), -(-nsfn // k0)
    out = {}
    for p in range(len(first)):
        f, n = int(first[p]), int(nlev[p])
        r_ = rs(p)
        for r in range(f, f + n):
            a, b = int(rs0[r]), int(rs1[r])
            if b:
                out[("idx0", r)] = np.stack([r_.randint(0, b, per1) for _ in range(k1)])
            if a:
                out[("idx1", r)] = np.stack([r_.randint(0, a, per0) for _ in range(k0)])
        for r in range(f, f + n):
            a, b = int(rs0[r]), int(rs1[r])
            if a > 0 and b > 0 and a * b >= ss:
                out[("xi", r)] = r_.randint(0, a, ss)
                out[("yi", r)] = r_.randint(0, b, ss)
    return out, per0, per1


@pytest.mark.parametrize("n0,n1,a", [(237, 217, 4), (700, 650, 5), (90, 80, 6), (0, 5, 4), (3, 2000, 4)])
def test_planner_replays_reference_rng_stream(svb, oracle, n0, n1, a):
    """svx_plan_draw_stream must leave np.random in exactly the state the reference's vecalign leaves it in
     - checked against the oracle driver, which is pinned to the reference."""
    from speech_vecalign_b200 import synth
    k = a - 1
    v0, v1 = synth.synth_pair(n0, n1, k, dim=128, seed=1)
    np.random.seed(77)
    oracle.vecalign(v0, v1, oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100, fast_host=True)
    want = np.random.get_state()
    pl = _plan(svb, [n0], [n1], k, k, oracle.alignment_types(a), w=math.ceil(k / 2) + 5)
    stage = np.zeros(max(int(pl.info["host_bytes"]), 16), dtype=np.uint8)
    ptrs = np.zeros(1, dtype=np.uint64)
    pl.bind(4096, stage.ctypes.data, ptrs, ptrs)          # the arena address only enters the descriptors
    np.random.seed(77)
    pl.draw(None)
    got = np.random.get_state()
    assert np.array_equal(got[1], want[1]) and got[2] == want[2]


@pytest.mark.parametrize("seeded", [False, True])
def test_draws_into_staging_match_reference_order(svb, oracle, seeded):
    """The C planner's draws (global stream continued in C, or one MT19937 stream per pair seed) land in the staging
    block exactly where the descriptors point and equal numpy's numbers in the reference's call order."""
    n0 = [237, 700, 90, 0, 3, 2000, 1]
    n1 = [217, 650, 80, 5, 2000, 2000, 1]
    k0, k1, nsfn, ss = 3, 4, 100, 20000
    types = [(1, 1), (1, 2), (2, 1)]
    pl = _plan(svb, n0, n1, k0, k1, types, sample_size=ss, nsfn=nsfn)
    first, nlev, rs0, rs1 = (pl.array(k) for k in ("first", "nlev", "rs0", "rs1"))
    seeds = [11 + 3 * i for i in range(len(n0))] if seeded else None
    np.random.seed(5)
    want, per0, per1 = _python_draws(rs0, rs1, first, nlev, k0, k1, nsfn, ss,
                                     (lambda p: np.random.RandomState(seeds[p])) if seeded else (lambda p: np.random))
    state_ref = np.random.get_state()
    stage = np.zeros(int(pl.info["host_bytes"]), dtype=np.uint8)
    ptrs = np.zeros(len(n0), dtype=np.uint64)
    pl.bind(4096, stage.ctypes.data, ptrs, ptrs)
    np.random.seed(5)
    pl.draw(seeds)
    if not seeded:
        got = np.random.get_state()
        assert np.array_equal(got[1], state_ref[1]) and got[2] == state_ref[2]
    off = {k: pl.array(k) for k in ("idx0", "idx1", "xi", "yi")}
    assert (int(pl.info["per0"]), int(pl.info["per1"])) == (per0, per1)
    has_draw = pl.array("has_draw")
    for r in range(len(rs0)):
        for key, rows, per in (("idx0", k1, per1), ("idx1", k0, per0)):
            if (key, r) in want:
                got = stage[off[key][r]:off[key][r] + rows * per * 4].view(np.int32).reshape(rows, per)
                assert np.array_equal(got, want[(key, r)]), (key, r)
        assert bool(has_draw[r]) == (("xi", r) in want)
        for key in ("xi", "yi"):
            if (key, r) in want:
                assert np.array_equal(stage[off[key][r]:off[key][r] + 4 * ss].view(np.int32), want[(key, r)]), (key, r)


def test_plan_matches_workspace_query_and_rejects_bad_types(svb, oracle):
    """svx_workspace_bytes == the plan's own sizes; the reference's overlap check (dp_core.pyx:204-209) is kept."""
    from speech_vecalign_b200 import engine
    L = svb.capi.lib()
    n0 = np.array([237, 2000], dtype=np.int32)
    n1 = np.array([217, 1900], dtype=np.int32)
    prm = engine.make_params(4, 4, 1024, oracle.alignment_types(5), 0.2, 7, 300, 20000, 100)
    pl = engine.Plan(prm, n0, n1)
    a, h = np.zeros(1, np.int64), np.zeros(1, np.int64)
    assert L.svx_workspace_bytes(prm.ctypes.data, 2, n0.ctypes.data, n1.ctypes.data, a.ctypes.data, h.ctypes.data) == 0
    assert (int(a[0]), int(h[0])) == (int(pl.info["arena_bytes"]), int(pl.info["host_bytes"]))
    bad = engine.make_params(2, 4, 1024, oracle.alignment_types(5), 0.2, 7, 300, 20000, 100)
    with pytest.raises(svb.capi.SvxError, match="x overlaps requrested"):
        engine.Plan(bad, n0, n1)
    # the default penalty of an empty pair is the reference's fallback knob (dp_utils.py:315-321)
    assert float(pl.info["fallback_del_penalty"]) == engine.fallback_del_penalty(0.2)


@pytest.mark.parametrize("B", [2, 6, 14, 16, 18, 24])
def test_packed_cost_kernel_block_cover(B):
    """banded_p2.cu maps band/2 + 1 warps onto the 2x2 position blocks of a block-diagonal: enumerate every band
    offset pattern (the path moves 0 or 1 in y per anti-diagonal) and check that Ymin and the block count hold."""
    for b0 in range(-5, 6):
        for s1 in (0, 1):
            for s2 in (0, 1):
                b1, b2 = b0 + s1, b0 + s1 + s2
                for vA in (0, 1):
                    for vB in (0, 1):
                        for vC in (0, 1):
                            ys = set()
                            for ok, b, par in ((vA, b0, (0,)), (vB, b1, (0, 1)), (vC, b2, (1,))):
                                if ok:
                                    ys |= {(yy - yy % 2) // 2 for yy in range(b, b + B) if yy % 2 in par}
                            if not ys:
                                continue
                            cand = ([(b0 + 1) >> 1] if vA else []) + ([b1 >> 1] if vB else []) + ([b2 >> 1] if vC else [])
                            assert min(cand) == min(ys)
                            assert max(ys) - min(ys) + 1 <= B // 2 + 1


def test_c_randint_replay_equals_numpy(svb):
    L = svb.capi.lib()
    np.random.seed(99)
    np.random.randint(0, 5, 333)
    st = np.random.get_state()
    key, pos = st[1].copy(), np.array([st[2]], dtype=np.int32)
    highs = np.array([1, 2, 3, 255, 256, 257, 20000, 2 ** 31 - 1], dtype=np.int32)
    counts = np.array([4, 9, 100, 1000, 1000, 1000, 5000, 50], dtype=np.int64)
    outs = [np.zeros(int(c), np.int32) for c in counts]
    ptrs = np.array([o.ctypes.data for o in outs], dtype=np.uint64)
    assert L.svx_host_randint_stream(key.ctypes.data, pos.ctypes.data, len(outs), highs.ctypes.data, counts.ctypes.data, ptrs.ctypes.data) == 0
    for h, c, o in zip(highs, counts, outs):
        assert np.array_equal(np.random.randint(0, int(h), int(c)), o)
    st2 = np.random.get_state()
    assert np.array_equal(st2[1], key) and st2[2] == pos[0]


@pytest.mark.parametrize("nbytes,nthreads", [(0, 4), (1, 4), (4095, 2), (1 << 20, 1), (5_000_003, 3), (37_000_001, 8), (37_000_001, 64)])
def test_host_memcpy_threads(svb, nbytes, nthreads):
    """svx_host_memcpy (pinned staging of pageable inputs): byte-exact for any size / thread count."""
    rng = np.random.default_rng(nbytes % 97)
    src = rng.integers(0, 256, size=nbytes + 64, dtype=np.uint8)
    dst = np.zeros(nbytes + 64, dtype=np.uint8)
    assert svb.capi.lib().svx_host_memcpy(dst.ctypes.data + 32, src.ctypes.data + 32, nbytes, nthreads) == 0
    assert np.array_equal(dst[32:32 + nbytes], src[32:32 + nbytes])
    assert not dst[:32].any() and not dst[32 + nbytes:].any()


def test_plan_sources_argument_checks(svb):
    """svx_plan_set_sources (host only): both sides or none, every pair needs rows, and the unfused prologue refuses."""
    from speech_vecalign_b200 import engine
    prm = engine.make_params(4, 4, 1024, [(1, 1)], 0.2, 7, 300, 20000, 100)
    pl = engine.Plan(prm, [10], [12])
    src = np.zeros(1, dtype=svb.capi.ROW_SOURCE)
    rc = svb.capi.lib().svx_plan_set_sources(pl.handle, svb.capi.hptr(src), None)
    assert rc != 0 and "both sides" in svb.capi.lib().svx_last_error_string().decode()
    rc = svb.capi.lib().svx_plan_set_sources(pl.handle, svb.capi.hptr(src), svb.capi.hptr(src))
    assert rc != 0 and "no source rows" in svb.capi.lib().svx_last_error_string().decode()
    big = engine.Plan(engine.make_params(4, 4, 1024, [(1, 1)], 0.2, 7, 300, 20000, 5000), [10], [12])
    src["rows"] = 16
    rc = svb.capi.lib().svx_plan_set_sources(big.handle, svb.capi.hptr(src), svb.capi.hptr(src))
    assert rc != 0 and "unfused" in svb.capi.lib().svx_last_error_string().decode()


def test_row_sources_reject_host_tensors(svb):
    """engine.row_sources describes DEVICE row matrices; a host tensor or another dtype is an error, not a silent copy."""
    import torch
    from speech_vecalign_b200 import engine
    with pytest.raises(ValueError):
        engine.row_sources([torch.zeros(4, 128, dtype=torch.float16)])
    with pytest.raises(ValueError):
        engine.row_sources([torch.zeros(4, 128, dtype=torch.float64)])
