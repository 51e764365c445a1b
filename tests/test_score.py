"""Known answers of the reference's scorer (README.md:289-296 of the reference): the shipped en-de
alignment against the human gold alignment."""
import os

from conftest import GOLDEN


def test_readme_table():
    from speech_vecalign_b200 import score
    from speech_vecalign_b200.vecalign import read_alignments
    ex = os.path.join(GOLDEN, "example")
    res = score.score_multiple([read_alignments(f"{ex}/human.gold")], [read_alignments(f"{ex}/shipped_alignment_a6.txt")])
    want = dict(precision_strict=0.558, precision_lax=0.942, recall_strict=0.632, recall_lax=0.993, f1_strict=0.593, f1_lax=0.967)
    for k, v in want.items():
        assert abs(res[k] - v) < 5e-4, (k, res[k], v)


def test_against_reference_scorer_if_present():
    import random
    import pytest
    from oracle import ref_loader
    ref = ref_loader.ref_package("svecalign.vecalign.score") if ref_loader.have_reference() else None
    if ref is None:
        pytest.skip("/root/reference not present")
    from speech_vecalign_b200 import score
    rnd = random.Random(3)

    def rand_alignment(n):
        out, x, y = [], 0, 0
        while x < n and y < n:
            dx, dy = rnd.choice([(1, 1), (1, 1), (1, 2), (2, 1), (0, 1), (1, 0), (2, 2)])
            out.append((list(range(x, min(n, x + dx))), list(range(y, min(n, y + dy)))))
            x, y = x + dx, y + dy
        return out
    for _ in range(20):
        g, t = rand_alignment(60), rand_alignment(60)
        a, b = ref.score_multiple([g], [t]), score.score_multiple([g], [t])
        for k in a:
            assert abs(a[k] - b[k]) < 1e-12, k
