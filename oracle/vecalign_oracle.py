"""oracle/vecalign_oracle.py — TEST INFRASTRUCTURE (checker), not product code.

CPU restatement (numpy + the C core in ``oracle/core.py``) of the reference's coarse-to-fine
aligner ``svecalign/vecalign/dp_utils.py``.  Every function cites the reference lines it
follows.  It exists so that the CUDA path can be checked, and a CPU baseline timed, on the GPU
box where /root/reference does not exist.  The product package never imports it.

Parity pin: ``tests/test_oracle_vs_reference.py`` runs this file and the real reference
(``/root/reference`` + ``oracle/_ref``) on the same seeded inputs and requires identical
results for every stack entry (bit-for-bit for paths/backpointers/costs produced by the C core;
bit-for-bit for the numpy steps because the same numpy executes them); ``tests/golden/`` holds
reference outputs for the GPU box.

Numerics that matter: numpy pairwise fp32 sums, fp32 ``+1e-5`` under NEP 50,
sgemm + fp32 mean in the sample norms, numpy histogram/cumsum/searchsorted/interp for the
deletion knob, Python ``round`` (half-to-even) in the slant, and the global ``np.random``
stream consumed in the reference's order (a14).
"""
from math import ceil
from time import perf_counter

import numpy as np

from . import core as _port_core


# ----------------------------------------------------------------------------- a3, a4
def unit_rows(vecs, fast=False):
    """dp_utils.py:32-40 make_norm1 — in place, v /= sqrt(sum(v*v)) + 1e-5 per (overlap,row)."""
    if fast:  # same pairwise reduction per contiguous row, one numpy call (bit-identical; tested)
        nrm = np.sqrt(np.square(vecs).sum(axis=2, keepdims=True))
        np.divide(vecs, nrm + np.float32(1e-5), out=vecs)
        return
    k, n, _ = vecs.shape
    for o in range(k):
        for r in range(n):
            row = vecs[o, r, :]
            nrm = np.sqrt(np.square(row).sum())
            vecs[o, r, :] = row / (nrm + 1e-5)


def halve(vecs, fast=False):
    """dp_utils.py:362-378 downsample_vectors — pair sums, minus per-overlap mean row, unit rows."""
    k, n, d = vecs.shape
    m = n // 2
    out = np.empty((k, m, d), dtype=np.float32)
    for o in range(k):
        if fast:
            out[o] = vecs[o, 0:2 * m:2, :] + vecs[o, 1:2 * m:2, :]
            out[o] -= np.mean(out[o], axis=0)
        else:
            for j in range(m):
                out[o, j, :] = vecs[o, 2 * j, :] + vecs[o, 2 * j + 1, :]
            centre = np.mean(out[o, :, :], axis=0)
            for j in range(m):
                out[o, j, :] = out[o, j, :] - centre
    unit_rows(out, fast=fast)
    return out


# ----------------------------------------------------------------------------- a5
def sampled_norms(vecs_a, vecs_b, num_samples):
    """dp_utils.py:326-359 compute_norms: norms_a[o,i] = 1 - mean_s(vecs_a[o,i] . sample_s), the
    samples being ceil(num_samples/K_b) random rows of every overlap of the OTHER side, drawn
    with np.random.choice(range(n_b), size) (one call per overlap of b, :346)."""
    kb, nb, d = vecs_b.shape
    ka, na, da = vecs_a.shape
    assert d == da
    per = ceil(num_samples / kb)
    if not (nb and per):
        return np.ones((ka, na)).astype(np.float32)
    bag = np.empty((per * kb, d), dtype=np.float32)
    for o in range(kb):
        pick = np.random.choice(range(nb), size=per, replace=True)
        bag[o * per:(o + 1) * per, :] = vecs_b[o, pick, :]
    norms = np.empty((ka, na), dtype=np.float32)
    for o in range(ka):
        sim = np.matmul(vecs_a[o, :, :], bag.T)
        norms[o, :] = 1.0 - sim.mean(axis=1)
    return norms


# ----------------------------------------------------------------------------- a6, a7
class PercentileKnob:
    """dp_utils.py:43-79 DeletionKnob: 1000-bin density histogram of sampled costs over
    [lo, hi], running cdf, 29 interpolation points at k/28."""

    def __init__(self, samples, lo, hi):
        self.res_min, self.res_max = lo, hi
        if self.res_min >= self.res_max:
            self.res_max = self.res_min + 1e-4
        nbins, npts = 1000, 30
        self.hist, self.bin_edges = np.histogram(samples, bins=nbins,
                                                 range=[self.res_min, self.res_max], density=True)
        width = self.bin_edges[1] - self.bin_edges[0]
        self.cdf = np.cumsum(self.hist) * width
        xs, ys = [0], [self.res_min]
        for q in np.linspace(0, 1, npts - 1)[1:-1]:
            idx = np.searchsorted(self.cdf, q)
            xs.append(q)
            ys.append(self.res_min + idx / float(nbins) * (self.res_max - self.res_min))
        xs.append(1)
        ys.append(self.res_max)
        self.x, self.y = tuple(xs), tuple(ys)

    def percentile_frac_to_del_penalty(self, frac):
        return np.interp([frac], self.x, self.y)[0]


def sample_cost_knob(core, e_vecs, f_vecs, e_norms, f_norms, sample_size, keep=None):
    """dp_utils.py:278-323 make_del_knob on overlap-0 vectors: full e x f grid (row-major, no RNG)
    when e*f < sample_size, otherwise two np.random.choice draws (x then y, :301-302)."""
    ne, nf = e_vecs.shape[0], f_vecs.shape[0]
    if ne > 0 and nf > 0 and sample_size > 0:
        if ne * nf < sample_size:
            sample_size = ne * nf
            xi = np.repeat(np.arange(ne, dtype=np.int32), nf)
            yi = np.tile(np.arange(nf, dtype=np.int32), ne)
        else:
            xi = np.random.choice(range(ne), size=sample_size, replace=True).astype(np.int32)
            yi = np.random.choice(range(nf), size=sample_size, replace=True).astype(np.int32)
        scores = np.empty(sample_size, dtype=np.float32)
        core.score_path(xi, yi, e_norms, f_norms, e_vecs, f_vecs, scores)
        lo, hi = 0, max(scores)
    else:
        xi = yi = None
        scores = np.array([0.0, 0.5, 1.0])
        lo, hi = 0, 1
    if keep is not None:
        keep.update(sample_x=xi, sample_y=yi, sample_scores=scores)
    return PercentileKnob(scores, lo, hi)


# ----------------------------------------------------------------------------- a9, a12
def dense_backtrace(bp):
    """dp_utils.py:146-174 dense_traceback."""
    x, y = bp.shape[0] - 1, bp.shape[1] - 1
    out = []
    while not (x == y == 0):
        code = bp[x, y]
        if code == 0:
            out.append(([x - 1], [y - 1])); x -= 1; y -= 1
        elif code == 1:
            out.append(([], [y - 1])); y -= 1
        elif code == 2:
            out.append(([x - 1], [])); x -= 1
        else:
            raise Exception('got unknown value')
    out.reverse()
    return out


def finish_scores(scores, alignments):
    """dp_utils.py:89-102 process_scores."""
    scores = np.clip(scores, a_min=0, a_max=None)
    for i, (xs, ys) in enumerate(alignments):
        if len(xs) == 0 or len(ys) == 0:
            scores[i] = 0.0
        else:
            scores[i] = scores[i] / len(xs) / len(ys)
    return scores


def banded_backtrace(csum, xp, yp, b_offset, xsize, ysize):
    """dp_utils.py:105-143 sparse_traceback (+ xy2ab_w_offset :82-86)."""
    x, y = xsize, ysize
    out, cum = [], []
    while True:
        a = x + y
        b = y - b_offset[a]
        cum.append(csum[a, b])
        dx, dy = xp[a, b], yp[a, b]
        if x == y == 0:
            break
        if x < 0 or y < 0:
            raise Exception('traceback bug')
        out.append((list(range(x - dx, x)), list(range(y - dy, y))))
        x, y = x - dx, y - dy
    out.reverse()
    cum.reverse()
    steps = np.array(cum[1:]) - np.array(cum[:-1])
    return out, finish_scores(steps, out)


# ----------------------------------------------------------------------------- a13
def _slant(path, xw, yw):
    """dp_utils.py:177-196 append_slant (Python round = half-to-even, then force a unit step)."""
    total = xw + yw
    x0, y0 = path[-1]
    for i in range(1, total + 1):
        x = x0 + round(xw * i / total)
        y = y0 + round(yw * i / total)
        px, py = path[-1]
        jump = x + y - px - py
        if jump == 1:
            path.append((x, y))
        elif jump == 2:
            path.append((x - 1, y))
        elif jump == 0:
            path.append((x + 1, y))


def search_path(alignments):
    """dp_utils.py:199-225 alignment_to_search_path."""
    path = [(0, 0)]
    dx = dy = 0
    for xs, ys in alignments:
        if len(xs) and len(ys):
            _slant(path, dx, dy)
            dx = dy = 0
            _slant(path, len(xs), len(ys))
        elif len(xs):
            dx += len(xs)
        elif len(ys):
            dy += len(ys)
    _slant(path, dx, dy)
    return path


def extend_to(alignments, size0, size1):
    """dp_utils.py:228-258 extend_alignments (in place)."""
    xmax = ymax = 0
    for xs, ys in alignments:
        for v in xs:
            xmax = max(xmax, v)
        for v in ys:
            ymax = max(ymax, v)
    if xmax > size0 or ymax > size1:
        raise Exception('asked to extend alignments but already bigger than requested')
    more_x = list(range(xmax + 1, size0 + 1))
    more_y = list(range(ymax + 1, size1 + 1))
    if len(more_x) == 0:
        alignments.extend(([], [v]) for v in more_y)
    elif len(more_y) == 0:
        alignments.extend(([v], []) for v in more_x)
    else:
        alignments.append((more_x, more_y))


def double_resolution(alignments):
    """dp_utils.py:261-275 upsample_alignment."""
    def grow(ids):
        return list(range(min(ids) * 2, (max(ids) + 1) * 2))
    out = []
    for xs, ys in alignments:
        if len(xs) == 0:
            out.extend(([], [v]) for v in grow(ys))
        elif len(ys) == 0:
            out.extend(([v], []) for v in grow(xs))
        else:
            out.append((grow(xs), grow(ys)))
    return out


# ----------------------------------------------------------------------------- a15
def vecalign(vecs0, vecs1, final_alignment_types, del_percentile_frac, width_over2,
             max_size_full_dp, costs_sample_size, num_samps_for_norm, norms0=None, norms1=None,
             core=None, fast_host=False, timings=None, penalties=None):
    """dp_utils.py:381-537 vecalign.  Same arguments and same ``stack`` result; ``core`` selects
    the native module (default: the C port; tests also pass the reference's compiled dp_core),
    ``fast_host`` swaps the per-row Python loops of a3/a4 for bit-identical vectorised numpy,
    ``timings`` (dict) receives the reference's phase names (:419-529), ``penalties`` ({depth: value},
    tests only) replaces the deletion penalty of those levels after the knob has been sampled (the RNG
    stream is consumed as always) - how a parity test follows a case past a DeletionKnob tie."""
    core = core or _port_core
    tm = timings if timings is not None else {}
    if width_over2 < 3:
        width_over2 = 3

    t0 = perf_counter()
    unit_rows(vecs0, fast=fast_host)
    unit_rows(vecs1, fast=fast_host)
    tm['make_norm1'] = perf_counter() - t0

    s0, s1 = vecs0.shape[1], vecs1.shape[1]
    depth_max = 0
    while s0 * s1 > max_size_full_dp ** 2:
        depth_max += 1
        s0, s1 = s0 // 2, s1 // 2

    stack = {0: {'v0': vecs0, 'v1': vecs1}}
    t0 = perf_counter()
    for d in range(1, depth_max + 1):
        stack[d] = {'v0': halve(stack[d - 1]['v0'], fast=fast_host),
                    'v1': halve(stack[d - 1]['v1'], fast=fast_host)}
    tm['Downsample embeddings'] = perf_counter() - t0

    t0 = perf_counter()
    for d in stack:
        lv = stack[d]
        lv['size0'], lv['size1'] = lv['v0'].shape[1], lv['v1'].shape[1]
        lv['alignment_types'] = final_alignment_types if d == 0 else [(1, 1)]
        if d == 0 and norms0 is not None:
            if norms0.shape != vecs0.shape[:2]:
                raise Exception('norms0 wrong shape')
            lv['n0'] = norms0
        else:
            lv['n0'] = sampled_norms(lv['v0'], lv['v1'], num_samps_for_norm)
        if d == 0 and norms1 is not None:
            if norms1.shape != vecs1.shape[:2]:
                raise Exception('norms1 wrong shape')
            lv['n1'] = norms1
        else:
            lv['n1'] = sampled_norms(lv['v1'], lv['v0'], num_samps_for_norm)
    tm['Normalize embeddings'] = perf_counter() - t0

    t0 = perf_counter()
    for d in stack:
        lv = stack[d]
        lv['del_knob'] = sample_cost_knob(core, lv['v0'][0, :, :], lv['v1'][0, :, :],
                                          lv['n0'][0, :], lv['n1'][0, :], costs_sample_size, keep=lv)
        lv['del_penalty'] = lv['del_knob'].percentile_frac_to_del_penalty(del_percentile_frac)
        if penalties is not None and d in penalties:
            lv['del_penalty'] = type(lv['del_penalty'])(penalties[d])
    tm['Compute deletion penalties'] = perf_counter() - t0

    top = stack[depth_max]
    t0 = perf_counter()
    top['costs_1to1'] = core.make_dense_costs(top['v0'], top['v1'], top['n0'], top['n1'])
    tm['Full DP make features'] = perf_counter() - t0
    t0 = perf_counter()
    _, top['x_y_tb'] = core.dense_dp(top['costs_1to1'], top['del_penalty'])
    top['alignments'] = dense_backtrace(top['x_y_tb'])
    tm['Full DP'] = perf_counter() - t0

    cost_t, dp_t = [], []
    for d in ([0] if depth_max == 0 else range(depth_max - 1, -1, -1)):
        lv = stack[d]
        if depth_max > 0:
            coarse = double_resolution(stack[d + 1]['alignments'])
            extend_to(coarse, lv['size0'], lv['size1'])
        else:
            coarse = stack[0]['alignments']
        lv['searchpath'] = search_path(coarse)

        t0 = perf_counter()
        lv['a_b_costs'], lv['b_offset'] = core.make_sparse_costs(
            lv['v0'], lv['v1'], lv['n0'], lv['n1'], lv['searchpath'], lv['alignment_types'], width_over2)
        cost_t.append(perf_counter() - t0)

        t0 = perf_counter()
        lv['a_b_csum'], lv['a_b_xp'], lv['a_b_yp'], lv['new_b_offset'] = core.sparse_dp(
            lv['a_b_costs'], lv['b_offset'], lv['alignment_types'], lv['del_penalty'],
            lv['size0'], lv['size1'])
        key = 'final_alignments' if d == 0 else 'alignments'
        lv[key], lv['alignment_scores'] = banded_backtrace(
            lv['a_b_csum'], lv['a_b_xp'], lv['a_b_yp'], lv['new_b_offset'], lv['size0'], lv['size1'])
        dp_t.append(perf_counter() - t0)

    tm['Upsample DP compute costs'] = sum(cost_t[:-1])
    tm['Upsample DP'] = sum(dp_t[:-1])
    tm['Final DP compute costs'] = cost_t[-1]
    tm['Final DP'] = dp_t[-1]
    return stack


# ----------------------------------------------------------------------------- a1 (vecalign.py)
def alignment_types(max_alignment_size):
    """vecalign.py:154-162 make_alignment_types: (x,y), x outer, y inner, x+y <= max."""
    return [(x, y) for x in range(1, max_alignment_size) for y in range(1, max_alignment_size)
            if x + y <= max_alignment_size]


def many_to_one_types(m_max):
    """vecalign.py:165-171 make_many_to_one_alignment_types."""
    return [(m, 1) for m in range(1, m_max + 1)]


def dp_cells(size0, size1, width_over2, max_size_full_dp=300):
    """DP-cell count of one pair as defined in BASELINE.md §3 / SURVEY.md §8d."""
    sizes = [(size0, size1)]
    while sizes[-1][0] * sizes[-1][1] > max_size_full_dp ** 2:
        sizes.append((sizes[-1][0] // 2, sizes[-1][1] // 2))
    depth = len(sizes) - 1
    band = 2 * max(3, width_over2)
    cells = (sizes[-1][0] + 1) * (sizes[-1][1] + 1)
    for a, b in (sizes[:1] if depth == 0 else sizes[:-1]):
        a_len = a + b + (1 if depth == 0 else 3)
        cells += (a_len + 2) * band
    return cells
