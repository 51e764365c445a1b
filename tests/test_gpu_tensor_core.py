"""GPU: the tcgen05 / TMA dense cost kernel (cost_mode 'tc', csrc/dense_tc.cu) against the oracle's
make_dense_costs (dp_core.pyx:36-77).

Tolerance, stated: the tensor-core path computes every dot product as a 3xTF32 split GEMM with fp32
accumulation in TMEM instead of the reference's sequential fp32 sum.  On unit-norm 1024-d rows
|dot_tc - dot_ref| <= 4e-6, so costs = 2(1-dot)/(1e-6+n0+n1) with n0+n1 >= 1 agree to <= 1e-5 absolute.
The coarse level only positions the search band (the finer level re-searches +-search_buffer_size
around it), and the end-to-end tests below check that final alignments stay identical."""
import math

import numpy as np
import pytest

from conftest import same_alignments

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("s0,s1,d", [(250, 250, 1024), (156, 156, 1024), (300, 37, 1024), (129, 257, 1024),
                                     (1, 1, 1024), (128, 128, 256), (5, 300, 128)])
def test_dense_costs_tc_within_tolerance(svb, oracle, ocore, s0, s1, d):
    from speech_vecalign_b200 import capi, dp_core
    rng = np.random.default_rng(s0 * 1000 + s1)
    # correlated rows so that dots span [-0.2, 1], like real coarse-level documents
    base = rng.standard_normal((max(s0, s1), d)).astype(np.float32)
    v0 = (base[:s0] + 0.7 * rng.standard_normal((s0, d))).astype(np.float32)[None]
    v1 = (base[:s1] + 0.7 * rng.standard_normal((s1, d))).astype(np.float32)[None]
    oracle.unit_rows(v0, fast=True)
    oracle.unit_rows(v1, fast=True)
    n0 = rng.uniform(0.6, 1.1, (1, s0)).astype(np.float32)
    n1 = rng.uniform(0.6, 1.1, (1, s1)).astype(np.float32)
    ref = ocore.make_dense_costs(v0, v1, n0, n1)
    got = dp_core.make_dense_costs(v0, v1, n0, n1, cost_mode=capi.SVX_COST_TC)
    assert got.shape == ref.shape
    err = float(np.max(np.abs(got.astype(np.float64) - ref)))
    assert err <= 1e-5, f"max |cost_tc - cost_ref| = {err:.3e}"
    exact = dp_core.make_dense_costs(v0, v1, n0, n1)
    assert np.array_equal(exact, ref)


@pytest.mark.parametrize("n0,n1,a,seed", [(237, 217, 4, 1), (700, 650, 5, 4), (800, 817, 6, 5), (2000, 2000, 5, 8)])
def test_tc_mode_keeps_the_alignment(svb, oracle, n0, n1, a, seed):
    from speech_vecalign_b200 import synth
    k = a - 1
    v0, v1 = synth.synth_pair(n0, n1, k, seed=seed)
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    np.random.seed(0)
    ref = oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True)
    np.random.seed(0)
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), *args, cost_mode="tc", debug=True)
    top = max(ref)
    assert np.max(np.abs(got[top]["costs_1to1"].astype(np.float64) - ref[top]["costs_1to1"])) <= 1e-5
    assert same_alignments(got[top]["alignments"], ref[top]["alignments"])
    assert same_alignments(got[0]["final_alignments"], ref[0]["final_alignments"])
    assert np.max(np.abs(got[0]["alignment_scores"] - ref[0]["alignment_scores"])) <= 1e-4


def test_tc_batch(svb, oracle):
    from speech_vecalign_b200 import synth
    shapes = [(420, 400), (237, 217), (1300, 1250), (90, 100)]
    a, k = 6, 5
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    pairs = [synth.synth_pair(n0, n1, k, seed=700 + i) for i, (n0, n1) in enumerate(shapes)]
    np.random.seed(9)
    refs = [oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True) for v0, v1 in pairs]
    np.random.seed(9)
    gots = svb.vecalign_batch([(v0.copy(), v1.copy()) for v0, v1 in pairs], *args, cost_mode="tc")
    for r, g in zip(refs, gots):
        assert same_alignments(g[0]["final_alignments"], r[0]["final_alignments"])
