"""Alignment scorer (reference: svecalign/vecalign/score.py:35-134): strict / lax precision, recall
and F1 of a test alignment against a gold alignment, as `align(gold_alignment=...)` reports them
(vecalign.py:290-293).  Evaluation only — nothing here is on the GPU path.

Definitions (Varga-style, as the reference applies them):
  strict hit  a test alignment that appears verbatim in the gold set;
  lax hit     a test alignment that shares at least one (source id, target id) link with the gold set;
  precision   hits over all non-empty test alignments;
  recall      the same counts with the roles of gold and test swapped, after dropping insertions and
              deletions from both.
"""
import sys
from typing import Dict, Iterable, List, Sequence, Tuple

Alignment = Tuple[Sequence[int], Sequence[int]]


def _links(alignments: Iterable[Alignment]):
    """all (source id, target id) links an alignment set asserts"""
    return {(i, j) for xs, ys in alignments for i in xs for j in ys}


def _hits(reference: List[Alignment], candidate: List[Alignment]):
    """(strict hits, lax hits, total) of `candidate` measured against `reference`"""
    ref = {(tuple(x), tuple(y)) for x, y in reference if len(x) or len(y)}
    cand = {(tuple(x), tuple(y)) for x, y in candidate if len(x) or len(y)}
    links = _links(ref)
    strict = lax = 0
    for xs, ys in cand:
        if (xs, ys) in ref:
            strict += 1
            lax += 1
        elif any((i, j) in links for i in xs for j in ys):
            lax += 1
    return strict, lax, len(cand)


def _ratio(num, den, default):
    return num / float(den) if den else default


def score_multiple(gold_list, test_list, value_for_div_by_0=0.0) -> Dict[str, float]:
    """Counts accumulated over all (gold, test) pairs, then the six figures of the reference's table."""
    p_strict = p_lax = p_all = r_strict = r_lax = r_all = 0
    for gold, test in zip(gold_list, test_list):
        s, l, n = _hits(gold, test)
        p_strict, p_lax, p_all = p_strict + s, p_lax + l, p_all + n
        full = lambda al: [(x, y) for x, y in al if len(x) and len(y)]
        s, l, n = _hits(full(test), full(gold))
        r_strict, r_lax, r_all = r_strict + s, r_lax + l, r_all + n
    ps, pl = _ratio(p_strict, p_all, value_for_div_by_0), _ratio(p_lax, p_all, value_for_div_by_0)
    rs, rl = _ratio(r_strict, r_all, value_for_div_by_0), _ratio(r_lax, r_all, value_for_div_by_0)
    f = lambda p, r: 2 * p * r / (p + r) if (p + r) else value_for_div_by_0
    return dict(recall_strict=rs, recall_lax=rl, precision_strict=ps, precision_lax=pl, f1_strict=f(ps, rs), f1_lax=f(pl, rl))


def log_final_scores(res, file=sys.stderr):
    bar = ' ' + '-' * 33
    rows = [('Precision', 'precision'), ('Recall', 'recall'), ('F1', 'f1')]
    print(bar, file=file)
    print('|             |  Strict |    Lax  |', file=file)
    for label, key in rows:
        print('| %-11s |   %.3f |   %.3f |' % (label, res[key + '_strict'], res[key + '_lax']), file=file)
    print(bar, file=file)


def main(argv=None):
    import argparse
    from .vecalign import read_alignments
    ap = argparse.ArgumentParser('strict/lax precision and recall for pairs of gold/test alignment files')
    ap.add_argument('-t', '--test', nargs='+', required=True)
    ap.add_argument('-g', '--gold', nargs='+', required=True)
    args = ap.parse_args(argv)
    if len(args.test) != len(args.gold):
        raise Exception('number of gold/test files must be the same')
    res = score_multiple([read_alignments(g) for g in args.gold], [read_alignments(t) for t in args.test])
    log_final_scores(res)
    return res


if __name__ == '__main__':
    main()
