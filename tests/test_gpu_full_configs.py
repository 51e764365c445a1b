"""GPU parity at BASELINE.json's full sizes (configs 3, 4 (a sample of the batch) and 5): final
alignments bit-exact against the oracle on the same seeded inputs, scores within 1e-4, plus the
size-independent properties of an alignment (a monotone partition of both documents)."""
import math

import numpy as np
import pytest

from conftest import same_alignments

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def _check_partition(al, n0, n1):
    """every segment of both documents is consumed exactly once, in order"""
    xs = [i for x, _ in al for i in x]
    ys = [j for _, y in al for j in y]
    assert xs == list(range(n0)) and ys == list(range(n1))


@pytest.mark.parametrize("n0,n1,a,seed", [(20000, 20000, 5, 31), (5000, 5000, 8, 32)])
def test_full_size_pair(svb, oracle, n0, n1, a, seed):
    from speech_vecalign_b200 import synth
    k = a - 1
    v0, v1 = synth.synth_pair(n0, n1, k, seed=seed)
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    np.random.seed(seed)
    got = svb.dp_utils.vecalign(v0.copy(), v1.copy(), *args)
    _check_partition(got[0]["final_alignments"], n0, n1)
    np.random.seed(seed)
    ref = oracle.vecalign(v0, v1, *args, fast_host=True)
    assert len(ref) == len(got)                      # same number of levels
    assert same_alignments(got[0]["final_alignments"], ref[0]["final_alignments"])
    assert np.max(np.abs(got[0]["alignment_scores"] - ref[0]["alignment_scores"])) <= 1e-4
    for d in ref:
        assert abs(got[d]["del_penalty"] - ref[d]["del_penalty"]) <= 1e-6 * max(1.0, abs(ref[d]["del_penalty"]))


def test_config4_sample_batch(svb, oracle):
    """48 pairs drawn from config 4's length distribution (200-800 segments, a=6), one batch with
    per-pair seeds: every pair equals the oracle run with that seed."""
    from speech_vecalign_b200 import synth
    from speech_vecalign_b200.engine import records_to_alignments
    n0s, n1s = synth.batch_sizes(48, seed=1234)
    a, k = 6, 5
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    pairs = [synth.synth_pair(int(n0), int(n1), k, seed=4000 + i) for i, (n0, n1) in enumerate(zip(n0s, n1s))]
    seeds = [77 + i for i in range(len(pairs))]
    outs = svb.vecalign_batch([(v0.copy(), v1.copy()) for v0, v1 in pairs], *args, seeds=seeds, output="records")
    for (v0, v1), s, r in zip(pairs, seeds, outs):
        np.random.seed(s)
        ref = oracle.vecalign(v0, v1, *args, fast_host=True)
        al, sc = records_to_alignments(r["recs"])
        assert same_alignments(al, ref[0]["final_alignments"])
        assert np.max(np.abs(sc - ref[0]["alignment_scores"])) <= 1e-4
