"""GPU parity, function by function: each kernel of libsvx.so against the oracle's restatement of
the reference function it replaces, on the same inputs.  Bar: bit-exact for everything except the
sample norms (the reference's sgemm order is CPU-dependent): <= 2 ulp there, stated below.
Wrappers under test: speech_vecalign_b200.dp_core (B1-shaped, reference argument order)."""
import math

import numpy as np
import pytest

from conftest import same_alignments, ulp_diff

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(20240611)


def _vecs(k, n, d, zero_rows=True):
    v = RNG.standard_normal((k, n, d)).astype(np.float32)
    if zero_rows and n:
        for j in range(1, k):
            v[j, :min(j, n)] = 0.0        # PAD rows of make_doc_embedding
    return v


@pytest.mark.parametrize("d", [128, 256, 512, 1024])
def test_make_norm1_bit_exact(svb, oracle, d):
    from speech_vecalign_b200 import dp_core
    v = _vecs(3, 67, d)
    v[0, 5] *= 1e-3
    v[1, 7] *= 1e3
    ref = v.copy()
    oracle.unit_rows(ref)
    got = dp_core.make_norm1(v.copy())
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 64, 301])
def test_downsample_bit_exact(svb, oracle, n):
    from speech_vecalign_b200 import dp_core
    v = _vecs(4, n, 1024)
    oracle.unit_rows(v, fast=True)
    ref = oracle.halve(v)
    got = dp_core.downsample_vectors(v)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_sample_norms_within_2ulp(svb, oracle):
    from speech_vecalign_b200 import dp_core
    v0, v1 = _vecs(3, 150, 1024), _vecs(4, 170, 1024)
    oracle.unit_rows(v0, fast=True)
    oracle.unit_rows(v1, fast=True)
    np.random.seed(11)
    ref = oracle.sampled_norms(v0, v1, 100)
    np.random.seed(11)
    per = math.ceil(100 / 4)
    idx = np.stack([np.random.randint(0, 170, per) for _ in range(4)]).astype(np.int32)
    got = dp_core.compute_norms_from_samples(v0, v1, idx)
    # tolerance: 2 ulp of fp32 at 1.0 (2.4e-7); reference = OpenBLAS sgemm + fp32 mean
    assert np.max(np.abs(got.astype(np.float64) - ref)) <= 2.4e-7
    assert got[1, 0] == 1.0 and got[2, 1] == 1.0      # zero rows -> exactly 1


def test_score_path_bit_exact(svb, ocore, oracle):
    from speech_vecalign_b200 import dp_core
    e, f = _vecs(1, 90, 1024)[0], _vecs(1, 110, 1024)[0]
    ne = RNG.uniform(0.7, 1.1, 90).astype(np.float32)
    nf = RNG.uniform(0.7, 1.1, 110).astype(np.float32)
    xi = RNG.integers(0, 90, 5000).astype(np.int32)
    yi = RNG.integers(0, 110, 5000).astype(np.int32)
    ref = np.empty(5000, np.float32)
    ocore.score_path(xi, yi, ne, nf, e, f, ref)
    got = np.empty(5000, np.float32)
    dp_core.score_path(xi, yi, ne, nf, e, f, got)
    assert np.array_equal(got, ref)


def test_del_knob_bit_exact(svb, oracle):
    from speech_vecalign_b200 import dp_core
    for n, frac in [(20000, 0.2), (777, 0.2), (20000, 0.05), (3, 0.9), (20000, 0.5)]:
        s = RNG.gamma(4.0, 0.25, n).astype(np.float32)
        ref = oracle.PercentileKnob(s, 0, max(s)).percentile_frac_to_del_penalty(frac)
        assert dp_core.del_penalty_from_scores(s, frac) == ref


@pytest.mark.parametrize("s0,s1", [(40, 50), (33, 1), (1, 77), (250, 250), (2, 700)])
def test_dense_costs_and_dp(svb, ocore, oracle, s0, s1):
    from speech_vecalign_b200 import dp_core
    v0, v1 = _vecs(2, s0, 1024), _vecs(2, s1, 1024)
    oracle.unit_rows(v0, fast=True)
    oracle.unit_rows(v1, fast=True)
    n0 = RNG.uniform(0.7, 1.1, (2, s0)).astype(np.float32)
    n1 = RNG.uniform(0.7, 1.1, (2, s1)).astype(np.float32)
    ref = ocore.make_dense_costs(v0, v1, n0, n1)
    got = dp_core.make_dense_costs(v0, v1, n0, n1)
    assert np.array_equal(got, ref)
    fast = dp_core.make_dense_costs(v0, v1, n0, n1, cost_mode=1)
    assert np.max(np.abs(fast - ref)) <= 2e-6            # FMA contraction only
    pen = 0.31415926
    rcs, rbp = ocore.dense_dp(ref, pen)
    cs, bp, path = dp_core.dense_dp(ref, pen, want_path=True)
    assert np.array_equal(bp, rbp) and np.array_equal(cs, rcs)
    assert path == [tuple(p) for p in oracle.search_path(oracle.dense_backtrace(rbp))]
    # upsampled + extended path of the next finer level (odd and even target sizes)
    for t0, t1 in [(2 * s0, 2 * s1), (2 * s0 + 1, 2 * s1 + 1), (2 * s0 + 1, 2 * s1)]:
        _, _, up = dp_core.dense_dp(ref, pen, target_sizes=(t0, t1), want_path=True)
        coarse = oracle.double_resolution(oracle.dense_backtrace(rbp))
        oracle.extend_to(coarse, t0, t1)
        assert up == [tuple(p) for p in oracle.search_path(coarse)]


@pytest.mark.parametrize("a,n0,n1", [(2, 60, 70), (4, 80, 75), (5, 120, 131), (6, 90, 95), (8, 70, 66), (10, 50, 55)])
def test_banded_costs_and_dp(svb, ocore, oracle, a, n0, n1):
    from speech_vecalign_b200 import dp_core
    k = a - 1
    types = oracle.alignment_types(a)
    w = math.ceil(k / 2) + 5
    v0, v1 = _vecs(k, n0, 1024), _vecs(k, n1, 1024)
    oracle.unit_rows(v0, fast=True)
    oracle.unit_rows(v1, fast=True)
    nn0 = RNG.uniform(0.7, 1.1, (k, n0)).astype(np.float32)
    nn1 = RNG.uniform(0.7, 1.1, (k, n1)).astype(np.float32)
    # a wavy but legal search path: random deletions/insertions around the diagonal
    al = []
    x = y = 0
    while x < n0 or y < n1:
        r = RNG.random()
        if (r < 0.15 and x < n0) or y >= n1:
            al.append(([x], [])); x += 1
        elif r < 0.3 or x >= n0:
            al.append(([], [y])); y += 1
        else:
            al.append(([x], [y])); x += 1; y += 1
    path = oracle.search_path(al)
    ref, rboff = ocore.make_sparse_costs(v0, v1, nn0, nn1, path, types, w)
    got, boff = dp_core.make_sparse_costs(v0, v1, nn0, nn1, path, types, w)
    assert np.array_equal(boff, rboff)
    assert np.array_equal(got, ref)                       # +inf outside the documents included
    fast, _ = dp_core.make_sparse_costs(v0, v1, nn0, nn1, path, types, w, cost_mode=1)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(fast), fin) and np.max(np.abs(fast[fin] - ref[fin])) <= 2e-6 * a * a
    pen = 0.27182818
    rcs, rxp, ryp, rbo = ocore.sparse_dp(ref, rboff, types, pen, n0, n1)
    cs, xp, yp, bo, al_d, sc_d = dp_core.sparse_dp(ref, rboff, types, pen, n0, n1, want_traceback=True)
    assert np.array_equal(bo, rbo) and np.array_equal(xp, rxp) and np.array_equal(yp, ryp)
    assert np.array_equal(cs, rcs)
    ral, rsc = oracle.banded_backtrace(rcs, rxp, ryp, rbo, n0, n1)
    assert same_alignments(al_d, ral) and np.array_equal(sc_d, rsc)


def test_banded_many_to_one_types(svb, ocore, oracle):
    """vecalign.py:165-171 type lists go through the generic (non-triangular) cost kernel."""
    from speech_vecalign_b200 import dp_core
    types = oracle.many_to_one_types(6)
    v0, v1 = _vecs(6, 50, 1024), _vecs(1, 40, 1024)
    oracle.unit_rows(v0, fast=True)
    oracle.unit_rows(v1, fast=True)
    nn0 = RNG.uniform(0.7, 1.1, (6, 50)).astype(np.float32)
    nn1 = RNG.uniform(0.7, 1.1, (1, 40)).astype(np.float32)
    al = [([i], [i]) for i in range(40)] + [([i], []) for i in range(40, 50)]
    path = oracle.search_path(al)
    w = 3 + 5
    ref, rboff = ocore.make_sparse_costs(v0, v1, nn0, nn1, path, types, w)
    got, boff = dp_core.make_sparse_costs(v0, v1, nn0, nn1, path, types, w)
    assert np.array_equal(got, ref) and np.array_equal(boff, rboff)
    rcs, rxp, ryp, rbo = ocore.sparse_dp(ref, rboff, types, 0.2, 50, 40)
    cs, xp, yp, bo = dp_core.sparse_dp(ref, rboff, types, 0.2, 50, 40)
    assert np.array_equal(cs, rcs) and np.array_equal(xp, rxp) and np.array_equal(yp, ryp)


def test_sparse_dp_next_path(svb, ocore, oracle):
    """Traceback of a (1,1)-only level lays the next finer level's search path on the device."""
    from speech_vecalign_b200 import dp_core
    n0, n1, w = 100, 104, 7
    v0, v1 = _vecs(1, n0, 1024), _vecs(1, n1, 1024)
    oracle.unit_rows(v0, fast=True)
    oracle.unit_rows(v1, fast=True)
    ones0, ones1 = np.ones((1, n0), np.float32), np.ones((1, n1), np.float32)
    al = [([i], [i]) for i in range(n0)] + [([], [i]) for i in range(n0, n1)]
    path = oracle.search_path(al)
    ref, rboff = ocore.make_sparse_costs(v0, v1, ones0, ones1, path, [(1, 1)], w)
    for t0, t1 in [(2 * n0, 2 * n1), (2 * n0 + 1, 2 * n1 + 1)]:
        out = dp_core.sparse_dp(ref, rboff, [(1, 1)], 0.9, n0, n1, target_sizes=(t0, t1), want_traceback=True)
        rcs, rxp, ryp, rbo = ocore.sparse_dp(ref, rboff, [(1, 1)], 0.9, n0, n1)
        ral, _ = oracle.banded_backtrace(rcs, rxp, ryp, rbo, n0, n1)
        coarse = oracle.double_resolution(ral)
        oracle.extend_to(coarse, t0, t1)
        assert out[6] == [tuple(p) for p in oracle.search_path(coarse)]
