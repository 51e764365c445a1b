"""GPU parity on randomly drawn configurations: document sizes, alignment_max_size, search buffer,
max_size_full_dp and sample sizes are drawn from a fixed seed so that tile boundaries of the kernels
(30 / 16 / 24 anti-diagonal tiles, 32-diagonal DP chunks, 64x64 dense tiles, odd sizes, several levels
on small documents) are crossed in many combinations.  Bars as in test_gpu_e2e.py.

One documented exception (parity): the reference's DeletionKnob quantises its 29 percentile
points to histogram bins whose float32 edges depend on max(sampled scores).  When a cumulative count lands
exactly on a percentile point (e.g. a full n0*n1 grid whose size is a multiple of 14), the side it falls on is
decided by ulp-level noise of the scores - which differ between the reference's BLAS norms and any other
summation order.  Such a level shows a del_penalty difference of a fraction of ONE bin (<= max/1000); it is
accepted here, the levels after it are not compared, and at most 2 of the drawn cases may hit it."""
import math

import numpy as np
import pytest

from conftest import same_alignments

pytestmark = pytest.mark.gpu


def _draw_cases(n, seed):
    rng = np.random.default_rng(seed)
    cases = []
    for i in range(n):
        a = int(rng.integers(2, 9))
        n0 = int(rng.integers(1, 700))
        n1 = max(1, int(n0 * rng.uniform(0.6, 1.5)))
        cases.append(dict(n0=n0, n1=n1, a=a, sbs=int(rng.integers(1, 9)), full=int(rng.choice([20, 37, 64, 150, 300])),
                          css=int(rng.choice([500, 5000, 20000])), nsn=int(rng.choice([7, 100])), seed=1000 + i,
                          frac=float(rng.choice([0.05, 0.2, 0.5])), dim=1024 if i % 6 == 0 else 128))
    return cases


_bin_events = []


@pytest.mark.parametrize("case", _draw_cases(36, 2024), ids=lambda c: f"{c['n0']}x{c['n1']}-a{c['a']}-b{c['sbs']}-f{c['full']}")
def test_random_configuration(svb, oracle, case):
    from speech_vecalign_b200 import synth
    a, k = case["a"], case["a"] - 1
    v0, v1 = synth.synth_pair(case["n0"], case["n1"], k, dim=case["dim"], seed=case["seed"])
    types = oracle.alignment_types(a)
    w = math.ceil(k / 2) + case["sbs"]
    args = (types, case["frac"], w, case["full"], case["css"], case["nsn"])
    np.random.seed(case["seed"])
    ref = oracle.vecalign(v0.copy(), v1.copy(), *args, fast_host=True)
    state = np.random.get_state()[1].copy()
    np.random.seed(case["seed"])
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), *args, debug=True)
    assert np.array_equal(state, np.random.get_state()[1])
    assert set(ref) == set(got)
    for d in sorted(ref, reverse=True):
        r, g = ref[d], got[d]
        assert np.array_equal(g["v0"], r["v0"]) and np.array_equal(g["v1"], r["v1"]), d
        diff = abs(g["del_penalty"] - r["del_penalty"])
        if diff > 1e-6 * max(1.0, abs(r["del_penalty"])):
            one_bin = float(np.max(r["sample_scores"])) / 1000.0
            assert diff <= one_bin * (1 + 1e-6), (d, diff, one_bin)
            _bin_events.append(case)
            assert len(_bin_events) <= 2, _bin_events
            return
        if "searchpath" in r:
            assert g["searchpath"] == [tuple(p) for p in r["searchpath"]], d
            fin = np.isfinite(r["a_b_costs"])
            assert np.array_equal(np.isfinite(g["a_b_costs"]), fin), d
            assert np.max(np.abs(g["a_b_costs"][fin] - r["a_b_costs"][fin]), initial=0) <= 2e-4, d
            key = "final_alignments" if d == 0 else "alignments"
            assert same_alignments(g[key], r[key]), d
            assert np.max(np.abs(g["alignment_scores"] - r["alignment_scores"]), initial=0) <= 1e-4, d
        if "costs_1to1" in r:
            assert same_alignments(g["alignments"], r["alignments"]), d
