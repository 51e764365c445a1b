"""Host-side driver of the B200 alignment path: plans a batch of document pairs, replays the
reference's host RNG draws, lays out one device arena, builds the job descriptors of
``include/svx.h`` and enqueues the kernels level by level.  One call = one batch; a single pair is
a batch of one.  Mirrors the control flow of ``svecalign/vecalign/dp_utils.py:381-537`` (vecalign)
with every numeric step executed by libsvx.so on the GPU.

Nothing here computes alignment arithmetic on the CPU: numpy is used for shapes, offsets, RNG
draws (the reference draws from the global ``np.random`` stream, dp_utils.py:301-302,346) and for
unpacking results.
"""
import os
from concurrent.futures import ThreadPoolExecutor
from math import ceil

import numpy as np
import torch

from . import capi

_ALIGN = 256


def level_sizes(n0, n1, max_size_full_dp):
    """dp_utils.py:403-408: halve both sides until s0*s1 <= max_size_full_dp**2.
    Returns (depth[P], s0[P, Lmax+1], s1[P, Lmax+1]) with sizes beyond a pair's depth set to -1."""
    n0 = np.asarray(n0, dtype=np.int64)
    n1 = np.asarray(n1, dtype=np.int64)
    lim = int(max_size_full_dp) ** 2
    depth = np.zeros(n0.shape[0], dtype=np.int64)
    s0, s1 = n0.copy(), n1.copy()
    while True:
        more = s0 * s1 > lim
        if not more.any():
            break
        depth += more
        s0 = np.where(more, s0 // 2, s0)
        s1 = np.where(more, s1 // 2, s1)
    lmax = int(depth.max()) if depth.size else 0
    lv = np.arange(lmax + 1)[None, :]
    a0 = n0[:, None] >> lv
    a1 = n1[:, None] >> lv
    valid = lv <= depth[:, None]
    return depth, np.where(valid, a0, -1), np.where(valid, a1, -1)


def path_len(c0, c1, t0, t1, upsample):
    """Search-path length A of a target level (vectorised twin of svx_path_len; consequence of
    dp_utils.py:228-258 extend_alignments)."""
    c0, c1, t0, t1 = (np.asarray(v, dtype=np.int64) for v in (c0, c1, t0, t1))
    xmax = np.where(c0 > 0, 2 * c0 - 1, 0)
    ymax = np.where(c1 > 0, 2 * c1 - 1, 0)
    lenx = np.maximum(t0 - xmax, 0)
    leny = np.maximum(t1 - ymax, 0)
    up = 1 + 2 * c0 + lenx + 2 * c1 + leny
    same = 1 + c0 + c1
    return np.where(np.asarray(upsample) != 0, up, same)


class _Arena:
    """Bump allocator over byte offsets (vectorised)."""

    def __init__(self):
        self.top = 0

    def take(self, nbytes):
        nbytes = np.asarray(nbytes, dtype=np.int64)
        padded = (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        ends = np.cumsum(padded.ravel())
        offs = (self.top + ends - padded.ravel()).reshape(nbytes.shape)
        if ends.size:
            self.top = int(self.top + ends[-1])
        return offs


def draw_samples(rec_s0, rec_s1, pair_first, pair_nlev, k0, k1, num_samps_for_norm, costs_sample_size,
                 skip_norm0, skip_norm1, seeds=None):
    """Replays the reference's np.random consumption for every pair, in input order
    (SURVEY.md §8a a14): per pair, for each depth ascending the n0 draws (K1 calls over range(size1))
    then the n1 draws (K0 calls over range(size0)); then for each depth ascending the knob draws
    (x then y) iff size0*size1 >= costs_sample_size.  np.random.choice(range(n), size=k) consumes
    the stream exactly like np.random.randint(0, n, k) (tested)."""
    per1 = ceil(num_samps_for_norm / k1) if k1 else 0   # samples per overlap of side 1 (for n0)
    per0 = ceil(num_samps_for_norm / k0) if k0 else 0
    nrec = rec_s0.shape[0]
    idx0 = np.zeros((nrec, k1, per1), dtype=np.int32)
    idx1 = np.zeros((nrec, k0, per0), dtype=np.int32)
    knob = [None] * nrec

    def draw_pair(p, rs):
        """All draws of pair p from the RandomState-like `rs`, in the reference's call order."""
        first, nlev = int(pair_first[p]), int(pair_nlev[p])
        for r in range(first, first + nlev):
            a, b = int(rec_s0[r]), int(rec_s1[r])
            lvl0 = r == first
            if not (lvl0 and skip_norm0) and b and per1:
                for o in range(k1):
                    idx0[r, o] = rs.randint(0, b, per1)
            if not (lvl0 and skip_norm1) and a and per0:
                for o in range(k0):
                    idx1[r, o] = rs.randint(0, a, per0)
        for r in range(first, first + nlev):
            a, b = int(rec_s0[r]), int(rec_s1[r])
            if a > 0 and b > 0 and costs_sample_size > 0 and a * b >= costs_sample_size:
                xi = rs.randint(0, a, costs_sample_size).astype(np.int32)
                yi = rs.randint(0, b, costs_sample_size).astype(np.int32)
                knob[r] = (xi, yi)

    npairs = pair_first.shape[0]
    if seeds is None:
        for p in range(npairs):                          # the global stream: inherently sequential
            draw_pair(p, np.random)
    elif npairs < 4:
        for p in range(npairs):
            draw_pair(p, np.random.RandomState(int(seeds[p])))
    else:
        # per-pair streams are independent (RandomState(seed) == np.random after np.random.seed(seed)) and
        # randint releases the GIL: draw the pairs on a thread pool
        with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
            list(pool.map(lambda p: draw_pair(p, np.random.RandomState(int(seeds[p]))), range(npairs)))
    return idx0, idx1, knob, per0, per1


def draw_samples_into(stage, off, rs0, rs1, rec_pair, rec_level, k0, k1, per0, per1, sample_size, has_draw,
                      skip_norm0, skip_norm1, seeds):
    """The reference's RNG draws of a batch, written at their final place in the (pinned) staging
    buffer.  Call order per pair: for each level ascending the n0 draws (K1 calls
    over range(size1), dp_utils.py:346), then the n1 draws (K0 calls over range(size0)); then for each
    level ascending the knob draws x, y (dp_utils.py:301-302) iff size0*size1 >= sample_size.
    seeds is None: the global np.random stream, pair after pair (what a serial loop of the reference
    consumes).  seeds given: pair p draws from RandomState(seeds[p]); the streams are independent, so
    they are generated by libsvx's bit-exact MT19937/randint replay on all host cores.
    Returns knob[r] = (xi, yi) int32 views into `stage` (or None)."""
    R = rs0.shape[0]
    rec = np.arange(R)
    l0 = rec_level == 0
    c0 = ~(l0 & skip_norm0) & (rs1 > 0) & (per1 > 0)          # n0 draws of record r happen
    c1 = ~(l0 & skip_norm1) & (rs0 > 0) & (per0 > 0)
    parts = []      # (pair, phase, rec, sub, high, count, dst offset)
    for cond, k, per, high, key in ((c0, k1, per1, rs1, "idx0"), (c1, k0, per0, rs0, "idx1")):
        r = np.repeat(rec[cond], k)
        sub = np.tile(np.arange(k), int(cond.sum())) + (0 if key == "idx0" else 1000)
        parts.append((rec_pair[r], np.zeros_like(r), r, sub, high[r], np.full(r.shape, per, dtype=np.int64),
                      off[key][r] + (sub % 1000) * per * 4))
    r = rec[has_draw]
    for sub, high, key in ((0, rs0, "xi"), (1, rs1, "yi")):
        parts.append((rec_pair[r], np.ones_like(r), r, np.full(r.shape, sub), high[r],
                      np.full(r.shape, sample_size, dtype=np.int64), off[key][r]))
    pair, phase, recs, sub, high, count, dst = (np.concatenate([p[i] for p in parts]) for i in range(7))
    order = np.lexsort((sub, recs, phase, pair))
    pair, high, count, dst = pair[order], high[order].astype(np.int32), count[order].astype(np.int64), dst[order].astype(np.int64)
    ncalls = pair.shape[0]
    if ncalls:
        if seeds is None:
            for c in range(ncalls):
                n = int(count[c])
                stage[dst[c]:dst[c] + 4 * n].view(np.int32)[:] = np.random.randint(0, int(high[c]), n)
        else:
            npairs = int(rec_pair.max()) + 1 if R else 0
            begin = np.searchsorted(pair, np.arange(npairs + 1)).astype(np.int64)
            sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.int64)[:npairs] & 0xFFFFFFFF, dtype=np.uint32)
            ptrs = (stage.ctypes.data + dst).astype(np.uint64)
            capi.check(capi.lib().svx_host_randint_seeded(npairs, capi.hptr(sd), capi.hptr(begin), capi.hptr(high), capi.hptr(count),
                                                          capi.hptr(ptrs), min(32, os.cpu_count() or 1)), "svx_host_randint_seeded")
    knob = [None] * R
    for r in np.nonzero(has_draw)[0]:
        n = 4 * sample_size
        knob[r] = (stage[off["xi"][r]:off["xi"][r] + n].view(np.int32), stage[off["yi"][r]:off["yi"][r] + n].view(np.int32))
    return knob


def fallback_del_penalty(frac):
    """dp_utils.py:315-321: with an empty side the knob is built from [0, .5, 1] on [0, 1]."""
    samp = np.array([0.0, 0.5, 1.0])
    hist, edges = np.histogram(samp, bins=1000, range=[0, 1], density=True)
    cdf = np.cumsum(hist) * (edges[1] - edges[0])
    xs, ys = [0], [0]
    for q in np.linspace(0, 1, 29)[1:-1]:
        xs.append(q)
        ys.append(0 + np.searchsorted(cdf, q) / 1000.0 * (1 - 0))
    xs.append(1)
    ys.append(1)
    return float(np.interp([frac], xs, ys)[0])


class BatchRun:
    """One batch of document pairs on one GPU."""

    def __init__(self, vec_ptrs0, vec_ptrs1, n0, n1, k0, k1, dim, alignment_types, del_percentile_frac,
                 width_over2, max_size_full_dp, costs_sample_size, num_samps_for_norm, device,
                 cost_mode=capi.SVX_COST_EXACT, norms0=None, norms1=None, keep_dense_csum=False, seeds=None):
        self.P = P = len(n0)
        self.dev = device
        self.dim = dim
        self.k0, self.k1 = int(k0), int(k1)
        self.types = [(int(x), int(y)) for x, y in alignment_types]
        for x, y in self.types:
            assert x > 0 and y > 0                       # dp_core.pyx:28-30
        mx = max([0] + [x for x, _ in self.types])
        my = max([0] + [y for _, y in self.types])
        if mx > self.k0:                                 # dp_core.pyx:204-209
            raise Exception('%d x overlaps requrested (via alignment_types), but vecs0 only has %d' % (mx, self.k0))
        if my > self.k1:
            raise Exception('%d y overlaps requrested (via alignment_types), but vecs1 only has %d' % (my, self.k1))
        if len(self.types) + 2 > capi.SVX_MAX_TYPES:
            raise capi.SvxError("too many alignment types for this build")
        self.frac = float(del_percentile_frac)
        self.w = max(3, int(width_over2))                # dp_utils.py:391-393
        self.band = 2 * self.w
        self.cost_mode = cost_mode
        self.sample_size = int(costs_sample_size)
        n0 = np.asarray(n0, dtype=np.int64)
        n1 = np.asarray(n1, dtype=np.int64)
        self.depth, S0, S1 = level_sizes(n0, n1, max_size_full_dp)
        self.nlev = self.depth + 1
        self.first = np.concatenate([[0], np.cumsum(self.nlev)[:-1]]).astype(np.int64)
        R = int(self.nlev.sum())
        self.R = R
        rp = np.repeat(np.arange(P), self.nlev)
        rl = np.arange(R) - self.first[rp]
        self.rec_pair, self.rec_level = rp, rl
        rs0, rs1 = S0[rp, rl], S1[rp, rl]
        self.rs0, self.rs1 = rs0, rs1
        is_top = rl == self.depth[rp]                    # coarsest level of its pair
        is_l0 = rl == 0
        # banded levels: every level below the top, or level 0 itself when the pair has one level
        banded = (~is_top) | (self.depth[rp] == 0)
        self.banded = banded
        # search path length of each banded level
        coarse = np.minimum(np.arange(R) + 1, R - 1)     # record of the next coarser level
        c0 = np.where(self.depth[rp] == 0, rs0, rs0[coarse])
        c1 = np.where(self.depth[rp] == 0, rs1, rs1[coarse])
        A = path_len(c0, c1, rs0, rs1, (self.depth[rp] > 0).astype(np.int64))
        A = np.where(banded, A, 0)
        self.A = A
        T = np.where(is_l0, len(self.types), 1)
        self.T = T

        # ---- sampling plan (sizes only; the draws themselves go straight into the staging buffer) ------
        per1 = ceil(int(num_samps_for_norm) / self.k1) if self.k1 else 0   # samples per overlap of side 1 (for n0)
        per0 = ceil(int(num_samps_for_norm) / self.k0) if self.k0 else 0
        self.per0, self.per1 = per0, per1
        prod = rs0 * rs1
        nsamp = np.where((rs0 > 0) & (rs1 > 0) & (self.sample_size > 0), np.minimum(prod, self.sample_size), 0).astype(np.int64)
        self.nsamp = nsamp
        # dp_utils.py:288-302: the full grid when e*f < sample_size, two RNG draws otherwise
        has_draw = (rs0 > 0) & (rs1 > 0) & (self.sample_size > 0) & (prod >= self.sample_size)

        # ---- arena layout ---------------------------------------------------------------------
        ar = _Arena()
        D = dim
        o = {}
        # host-initialised region first (one H2D copy)
        o["idx0"] = ar.take(np.full(R, self.k1 * per1 * 4))
        o["idx1"] = ar.take(np.full(R, self.k0 * per0 * 4))
        o["xi"] = ar.take(np.where(has_draw, nsamp * 4, 0))
        o["yi"] = ar.take(np.where(has_draw, nsamp * 4, 0))
        o["delpen"] = ar.take(np.full(R, 8))
        # TMA descriptors of the coarsest level's operands (tensor-core mode), encoded on the host
        o["tmaps"] = ar.take(np.where(is_top & (cost_mode == capi.SVX_COST_TC), 256, 0))
        # every job descriptor of the batch (upper bound incl. 16-byte alignment slack per array)
        jobs_bytes = (2 * P * capi.ROWS.itemsize + 2 * R * capi.DOWN.itemsize + 2 * R * capi.NORM.itemsize +
                      2 * R * capi.LEVEL.itemsize + 64 * 16 +
                      R * capi.SCORE.itemsize + P * capi.DENSE.itemsize + R * capi.BAND.itemsize + 4096)
        o["jobs"] = ar.take(np.array([jobs_bytes]))
        self._jobs_off = int(o["jobs"][0])
        self._jobs_cap = jobs_bytes
        host_end = ar.top
        self.host_init_bytes = host_end
        # device-only region
        o["norms0"] = ar.take(self.k0 * rs0 * 4)
        o["norms1"] = ar.take(self.k1 * rs1 * 4)
        norms_lo, norms_hi = int(o["norms0"].min()) if R else ar.top, ar.top
        o["vec0"] = ar.take(np.where(is_l0, 0, self.k0 * rs0 * D * 4))
        o["vec1"] = ar.take(np.where(is_l0, 0, self.k1 * rs1 * D * 4))
        o["mean0"] = ar.take(np.where(is_l0, 0, self.k0 * D * 4))
        o["mean1"] = ar.take(np.where(is_l0, 0, self.k1 * D * 4))
        o["mbar0"] = ar.take(np.full(R, (D + 1024) * 8))     # + denominators of the sampled rows (SvxLevelJob.mbar)
        o["mbar1"] = ar.take(np.full(R, (D + 1024) * 8))
        o["scores"] = ar.take(nsamp * 4)
        o["perm"] = ar.take(np.where(has_draw & ~is_top, nsamp * 4, 0))
        o["dcost"] = ar.take(np.where(is_top, rs0 * rs1 * 4, 0))
        o["ddots"] = ar.take(np.where(is_top, rs0 * rs1 * 4, 0))
        o["dbp"] = ar.take(np.where(is_top, (rs0 + 1) * (rs1 + 1), 0))
        o["dcsum"] = ar.take(np.where(is_top & bool(keep_dense_csum), (rs0 + 1) * (rs1 + 1) * 8, 0))
        o["ypath"] = ar.take(A * 4)
        o["bcost"] = ar.take(A * T * self.band * 4)
        o["bbp"] = ar.take(np.where(banded, (A + 2) * self.band, 0))
        o["bcsum"] = ar.take(np.where(banded, (A + 2) * self.band * 8, 0))
        self.rec_cap = np.where(banded, rs0 + rs1 + 2, 0)
        o["recs"] = ar.take(self.rec_cap * capi.REC.itemsize)
        o["nrecs"] = ar.take(np.full(R, 4))
        o["status"] = ar.take(np.full(R, 8))       # [0] banded status, [1] dense status
        self.off = o
        self.nbytes = ar.top
        self.keep_dense_csum = bool(keep_dense_csum)

        self.arena = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.base = self.arena.data_ptr()
        assert self.base % 16 == 0
        # norms default to 1.0 (dp_utils.py:356-357), status/nrecs to 0
        if norms_hi > norms_lo:
            self.arena[norms_lo:norms_hi].view(torch.float32).fill_(1.0)
        st_lo = int(o["nrecs"].min()) if R else self.nbytes
        self.arena[st_lo:self.nbytes].zero_()

        # ---- host staging of the host-initialised region ---------------------------------------
        # pinned staging (torch's caching host allocator recycles it): the single H2D copy below is then
        # asynchronous, so planning the next batch never waits for this batch's kernels
        self._stage = torch.zeros(host_end, dtype=torch.uint8, pin_memory=True)
        stage = self._stage.numpy()
        fb = fallback_del_penalty(self.frac)
        if R:
            dp_view = np.full(R, fb, dtype=np.float64)
            for r in range(R):
                stage[o["delpen"][r]:o["delpen"][r] + 8] = dp_view[r:r + 1].view(np.uint8)
        self.knob = draw_samples_into(stage, o, rs0, rs1, rp, rl, self.k0, self.k1, per0, per1, self.sample_size,
                                      has_draw, norms0 is not None, norms1 is not None, seeds)

        # ---- job descriptors ------------------------------------------------------------------
        b = self.base
        v0p = np.asarray(vec_ptrs0, dtype=np.uint64)
        v1p = np.asarray(vec_ptrs1, dtype=np.uint64)
        vec0 = np.where(is_l0, v0p[rp], (b + o["vec0"]).astype(np.uint64))
        vec1 = np.where(is_l0, v1p[rp], (b + o["vec1"]).astype(np.uint64))
        self.vec0_ptr, self.vec1_ptr = vec0, vec1
        ptr = lambda key: (b + o[key]).astype(np.uint64)

        rows = np.zeros(2 * P, dtype=capi.ROWS)
        rows["ptr"][0::2], rows["nrows"][0::2] = v0p, self.k0 * n0
        rows["ptr"][1::2], rows["nrows"][1::2] = v1p, self.k1 * n1

        lmax = int(self.depth.max()) if P else 0
        self.down_jobs = []
        for lvl in range(1, lmax + 1):
            sel = np.nonzero(rl == lvl)[0]
            dj = np.zeros(2 * sel.size, dtype=capi.DOWN)
            dj["in"][0::2], dj["out"][0::2], dj["mean"][0::2] = vec0[sel - 1], vec0[sel], ptr("mean0")[sel]
            dj["k"][0::2], dj["n"][0::2] = self.k0, rs0[sel - 1]
            dj["in"][1::2], dj["out"][1::2], dj["mean"][1::2] = vec1[sel - 1], vec1[sel], ptr("mean1")[sel]
            dj["k"][1::2], dj["n"][1::2] = self.k1, rs1[sel - 1]
            self.down_jobs.append((dj, np.repeat(rp[sel], 2)))

        skip0 = is_l0 & (norms0 is not None)
        skip1 = is_l0 & (norms1 is not None)
        sel0 = np.nonzero((rs1 > 0) & (per1 > 0) & (self.k1 > 0) & ~skip0 & (rs0 > 0))[0]
        sel1 = np.nonzero((rs0 > 0) & (per0 > 0) & (self.k0 > 0) & ~skip1 & (rs1 > 0))[0]
        nj = np.zeros(sel0.size + sel1.size, dtype=capi.NORM)
        a_, b_ = nj[:sel0.size], nj[sel0.size:]
        a_["vecs"], a_["other"], a_["idx"], a_["mbar"], a_["norms"] = vec0[sel0], vec1[sel0], ptr("idx0")[sel0], ptr("mbar0")[sel0], ptr("norms0")[sel0]
        a_["k"], a_["n"], a_["ko"], a_["no"], a_["per"] = self.k0, rs0[sel0], self.k1, rs1[sel0], per1
        b_["vecs"], b_["other"], b_["idx"], b_["mbar"], b_["norms"] = vec1[sel1], vec0[sel1], ptr("idx1")[sel1], ptr("mbar1")[sel1], ptr("norms1")[sel1]
        b_["k"], b_["n"], b_["ko"], b_["no"], b_["per"] = self.k1, rs1[sel1], self.k0, rs0[sel1], per0
        # fused prologue: one job per (record, side), grouped by level
        want0 = np.zeros(R, dtype=bool); want0[sel0] = True      # norms0 computed from samples
        want1 = np.zeros(R, dtype=bool); want1[sel1] = True
        nxt = np.minimum(np.arange(R) + 1, max(R - 1, 0))
        has_next = rl < self.depth[rp]
        self.level_jobs = []
        for lvl in range(0, lmax + 1):
            sel = np.nonzero(rl == lvl)[0]
            lj = np.zeros(2 * sel.size, dtype=capi.LEVEL)
            for side, (vec_a, vec_b, mean_a, mean_b, idx_k, mbar_k, norms_k, want, ka, kb, sa, sb, per) in enumerate([
                    (vec0, vec1, "mean0", "mean1", "idx0", "mbar0", "norms0", want0, self.k0, self.k1, rs0, rs1, per1),
                    (vec1, vec0, "mean1", "mean0", "idx1", "mbar1", "norms1", want1, self.k1, self.k0, rs1, rs0, per0)]):
                v = lj[side::2]
                v["vecs"], v["other"] = vec_a[sel], vec_b[sel]
                if lvl > 0:
                    v["mean"], v["other_mean"] = ptr(mean_a)[sel], ptr(mean_b)[sel]
                v["next"] = np.where(has_next[sel], vec_a[nxt[sel]], 0)
                v["idx"] = np.where(want[sel], ptr(idx_k)[sel], 0)
                v["norms"] = np.where(want[sel], ptr(norms_k)[sel], 0)
                v["mbar"] = ptr(mbar_k)[sel]
                v["k"], v["n"], v["ko"], v["no"], v["per"] = ka, sa[sel], kb, sb[sel], per
                # levels >= 1 align 1-1 only: later kernels read overlap 0 (debug keeps everything)
                v["keep"] = ka if (lvl == 0 or keep_dense_csum) else min(1, ka)
            self.level_jobs.append((lj, np.repeat(rp[sel], 2)))
        nj_pair = np.concatenate([rp[sel0], rp[sel1]])
        order = np.argsort(nj_pair, kind="stable")       # pair-major so that a pair range is a job range
        self.norm_jobs = nj[order]
        nj_pair = nj_pair[order]

        ssel = np.nonzero(nsamp > 0)[0]
        sj = np.zeros(ssel.size, dtype=capi.SCORE)
        sj["e"], sj["f"], sj["norm_e"], sj["norm_f"] = vec0[ssel], vec1[ssel], ptr("norms0")[ssel], ptr("norms1")[ssel]
        sj["xi"] = np.where(has_draw[ssel], ptr("xi")[ssel], 0)
        sj["yi"] = np.where(has_draw[ssel], ptr("yi")[ssel], 0)
        sj["scores"], sj["del_penalty"] = ptr("scores")[ssel], ptr("delpen")[ssel]
        # coarsest level: the dense cost kernel has already produced every dot product of the level
        sj["dots"] = np.where(is_top[ssel], ptr("ddots")[ssel], 0)
        sj["perm"] = np.where(has_draw[ssel] & ~is_top[ssel], ptr("perm")[ssel], 0)
        sj["ne"], sj["nf"], sj["nsamp"] = rs0[ssel], rs1[ssel], nsamp[ssel]
        self.score_jobs = sj

        top = np.nonzero(is_top)[0]                      # one per pair, in pair order
        tgt = np.where(self.depth > 0, top - 1, top)     # record whose search path the dense DP lays
        dj = np.zeros(P, dtype=capi.DENSE)
        dj["v0"], dj["v1"], dj["n0"], dj["n1"] = vec0[top], vec1[top], ptr("norms0")[top], ptr("norms1")[top]
        dj["costs"], dj["del_penalty"], dj["bp"] = ptr("dcost")[top], ptr("delpen")[top], ptr("dbp")[top]
        dj["dots"] = ptr("ddots")[top]
        dj["csum"] = ptr("dcsum")[top] if keep_dense_csum else 0
        dj["ypath"] = ptr("ypath")[tgt]
        dj["status_d"] = ptr("status")[top] + np.uint64(4)
        dj["s0"], dj["s1"], dj["t0"], dj["t1"] = rs0[top], rs1[top], rs0[tgt], rs1[tgt]
        dj["upsample"], dj["path_len"] = (self.depth > 0), A[tgt]
        if cost_mode == capi.SVX_COST_TC and P:
            dj["tmap0"] = ptr("tmaps")[top]
            dj["tmap1"] = ptr("tmaps")[top] + np.uint64(128)
            blobs = np.zeros((P, 2, 128), dtype=np.uint8)
            capi.check(capi.lib().svx_dense_tmaps_encode(capi.hptr(dj), P, D, capi.hptr(blobs)), "svx_dense_tmaps_encode")
            for i, r in enumerate(top):
                stage[o["tmaps"][r]:o["tmaps"][r] + 256] = blobs[i].ravel()
        self.dense_jobs = dj
        self.top_rec, self.tgt_rec = top, tgt

        # banded stages: stage s (1-based) handles level max(depth,1) - s of every pair that has it
        xo = np.zeros(capi.SVX_MAX_TYPES, dtype=np.int8)
        yo = np.zeros(capi.SVX_MAX_TYPES, dtype=np.int8)
        for t, (x, y) in enumerate(self.types):
            xo[t], yo[t] = x, y
        xo1 = np.zeros_like(xo); yo1 = np.zeros_like(yo)
        xo1[0] = yo1[0] = 1
        self.band_stages = []
        nstage = max(1, lmax)
        for s in range(1, nstage + 1):
            lvl = np.maximum(self.depth, 1) - s
            ps = np.nonzero(lvl >= 0)[0]
            recs_ = self.first[ps] + lvl[ps]
            groups = []
            for want_l0 in (False, True):
                sel = recs_[(rl[recs_] == 0) == want_l0]
                if sel.size == 0:
                    continue
                bj = np.zeros(sel.size, dtype=capi.BAND)
                bj["v0"], bj["v1"], bj["n0"], bj["n1"] = vec0[sel], vec1[sel], ptr("norms0")[sel], ptr("norms1")[sel]
                bj["ypath"], bj["costs"], bj["del_penalty"] = ptr("ypath")[sel], ptr("bcost")[sel], ptr("delpen")[sel]
                bj["bp"], bj["csum"], bj["recs"], bj["nrecs"] = ptr("bbp")[sel], ptr("bcsum")[sel], ptr("recs")[sel], ptr("nrecs")[sel]
                bj["status_d"] = ptr("status")[sel]
                bj["s0"], bj["s1"], bj["k0"], bj["k1"] = rs0[sel], rs1[sel], self.k0, self.k1
                bj["a_len"], bj["band"], bj["width_over2"] = A[sel], self.band, self.w
                bj["rec_cap"] = self.rec_cap[sel]
                if want_l0:
                    bj["ntypes"], bj["xo"], bj["yo"] = len(self.types), xo, yo
                    bj["amax"] = max([2] + [x + y for x, y in self.types])
                    bj["next_ypath"] = 0
                else:
                    bj["ntypes"], bj["xo"], bj["yo"], bj["amax"] = 1, xo1, yo1, 2
                    bj["next_ypath"] = ptr("ypath")[sel - 1]
                    bj["t0"], bj["t1"], bj["next_len"] = rs0[sel - 1], rs1[sel - 1], A[sel - 1]
                groups.append((bj, rp[sel]))
            self.band_stages.append(groups)

        # pack descriptors into the staging buffer
        self._job_views = {}
        cur = self._jobs_off

        def pack(name, arr, pairs):
            nonlocal cur
            nb = arr.nbytes
            cur = (cur + 15) // 16 * 16
            assert cur + nb <= self._jobs_off + self._jobs_cap, "descriptor region too small"
            stage[cur:cur + nb] = arr.view(np.uint8).ravel()
            self._job_views[name] = (self.base + cur, arr, np.asarray(pairs, dtype=np.int64))
            cur += nb

        pack("rows", rows, np.repeat(np.arange(P), 2))
        for i, (dj_, pr_) in enumerate(self.down_jobs):
            pack(("down", i), dj_, pr_)
        pack("norm", self.norm_jobs, nj_pair)
        for i, (lj_, pr_) in enumerate(self.level_jobs):
            pack(("level", i), lj_, pr_)
        pack("score", self.score_jobs, rp[ssel])
        pack("dense", self.dense_jobs, np.arange(P))
        for s, groups in enumerate(self.band_stages):
            for g, (bj, pr_) in enumerate(groups):
                pack(("band", s, g), bj, pr_)
        # descriptor block -> device by a kernel reading the pinned buffer (not the DMA queue, which the bulk
        # embedding copies of the following batches occupy)
        if host_end:
            capi.check(capi.lib().svx_upload_pinned(self.base, self._stage.data_ptr(), host_end,
                                                    torch.cuda.current_stream(device).cuda_stream), "svx_upload_pinned")
        if norms0 is not None:
            self._put(o["norms0"][self.first[0]], np.ascontiguousarray(norms0, dtype=np.float32))
        if norms1 is not None:
            self._put(o["norms1"][self.first[0]], np.ascontiguousarray(norms1, dtype=np.float32))

    # ------------------------------------------------------------------------------------------
    def _put(self, off, arr):
        t = torch.from_numpy(arr.view(np.uint8).ravel())
        self.arena[int(off):int(off) + t.numel()].copy_(t)

    def _get(self, off, nbytes, dtype):
        return self.arena[int(off):int(off) + int(nbytes)].cpu().numpy().view(dtype)

    def _call(self, fn, name, key, *extra):
        dptr, arr, pairs = self._job_views[key]
        lo, hi = 0, arr.shape[0]
        if self._pair_range is not None:
            lo, hi = (int(v) for v in np.searchsorted(pairs, self._pair_range))
        if hi <= lo:
            return
        dptr += lo * arr.dtype.itemsize
        hp = capi.hptr(arr) + lo * arr.dtype.itemsize
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        if self._events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.check(fn(dptr, hp, hi - lo, *extra, stream), name)
            e1.record()
            self._events.append((name if not isinstance(key, tuple) or key[0] != "band" else
                                 name + ("_level0" if self._band_is_l0(key) else "_coarse"), e0, e1))
            return
        capi.check(fn(dptr, hp, hi - lo, *extra, stream), name)

    def _band_is_l0(self, key):
        return bool(self._job_views[key][1]["ntypes"][0] == len(self.types) and
                    self._job_views[key][1]["next_ypath"][0] == 0)

    def kernel_times(self):
        """Per-launcher device time (ms) of the last run(timing=True), summed by launcher name;
        CUDA events on the launching stream.  Synchronises."""
        torch.cuda.synchronize(self.dev)
        out = {}
        for name, e0, e1 in self._events or []:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out

    _events = None
    fused_prologue = True        # False: the three separate prologue launchers (A/B measurements)
    _pair_range = None
    _streams = None

    def _enqueue_chain(self):
        """The whole path for the pairs in self._pair_range (all pairs if None) on the current stream."""
        L = capi.lib()
        D, mode = self.dim, self.cost_mode
        if self.fused_prologue:
            for i in range(len(self.level_jobs)):
                self._call(L.svx_level_prologue, "svx_level_prologue", ("level", i), D)
        else:
            self._call(L.svx_normalize_rows, "svx_normalize_rows", "rows", D)
            for i in range(len(self.down_jobs)):
                self._call(L.svx_downsample, "svx_downsample", ("down", i), D)
            self._call(L.svx_sample_norms, "svx_sample_norms", "norm", D)
        self._call(L.svx_dense_costs, "svx_dense_costs", "dense", D, mode)
        self._call(L.svx_score_pairs, "svx_score_pairs", "score", D, mode)
        self._call(L.svx_del_knob, "svx_del_knob", "score", self.frac)
        self._call(L.svx_dense_dp, "svx_dense_dp", "dense")
        for s, groups in enumerate(self.band_stages):
            for g in range(len(groups)):
                self._call(L.svx_banded_costs, "svx_banded_costs", ("band", s, g), D, mode)
            for g in range(len(groups)):
                self._call(L.svx_banded_dp, "svx_banded_dp", ("band", s, g))

    def pair_groups(self, ngroups):
        """Contiguous pair ranges of roughly equal work (bytes of level-0 rows)."""
        ngroups = max(1, min(int(ngroups), self.P))
        l0 = self.first
        work = np.cumsum((self.rs0[l0] + self.rs1[l0]).astype(np.float64) + 1.0)
        cuts = np.searchsorted(work, work[-1] * np.arange(1, ngroups) / ngroups, side="left") + 1
        b = np.unique(np.concatenate([[0], np.minimum(cuts, self.P), [self.P]]))
        return [(int(b[i]), int(b[i + 1])) for i in range(len(b) - 1)]

    def run(self, timing=False, ngroups=1):
        """Enqueue the whole batch (asynchronous).  ngroups > 1 splits the pairs into contiguous
        groups, each running its own kernel chain on its own CUDA stream (forked from / joined to the
        current stream with events): the latency-bound kernels of one group (wavefront DPs, knob)
        then overlap the bandwidth- and FP32-bound kernels of the others.  Pairs are independent, so
        the result does not depend on the grouping."""
        self._events = [] if timing else None
        if ngroups <= 1 or self.P <= 1:
            self._pair_range = None
            self._enqueue_chain()
            return
        groups = self.pair_groups(ngroups)
        if self._streams is None or len(self._streams) < len(groups):
            self._streams = [torch.cuda.Stream(device=self.dev) for _ in range(len(groups))]
        cur = torch.cuda.current_stream(self.dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        try:
            for st, pr in zip(self._streams, groups):
                st.wait_event(fork)
                self._pair_range = pr
                with torch.cuda.stream(st):
                    self._enqueue_chain()
                    done = torch.cuda.Event()
                    done.record(st)
                cur.wait_event(done)
        finally:
            self._pair_range = None

    # ------------------------------------------------------------------------------------------
    def algorithmic_bytes(self):
        """Compulsory HBM bytes per launcher for this batch (SURVEY.md §8d formulae):
        every needed embedding row read once, every output written once."""
        D, K0, K1, B = self.dim, self.k0, self.k1, self.band
        s0, s1, A, T = self.rs0, self.rs1, self.A, self.T
        l0 = self.rec_level == 0
        top = self.rec_level == self.depth[self.rec_pair]
        band_l0, band_co = self.banded & l0, self.banded & ~l0
        nj = self.norm_jobs
        lvl_bytes = 0
        for lj, _ in self.level_jobs:
            kk, nn = lj["k"].astype(np.int64), lj["n"].astype(np.int64)
            rows = kk * nn * D * 4
            kept = lj["keep"].astype(np.int64) * nn * D * 4
            lvl_bytes += int((rows * (1 + (lj["mean"] != 0)) + kept + (lj["next"] != 0) * kk * (nn // 2) * D * 4 + kk * nn * 4 +
                              (lj["idx"] != 0) * lj["ko"].astype(np.int64) * lj["per"] * D * 4 * 2).sum())
        out = {
            "svx_level_prologue": lvl_bytes,
            "svx_normalize_rows": 2 * 4 * D * int(K0 * s0[l0].sum() + K1 * s1[l0].sum()),
            "svx_downsample": 4 * D * int((K0 * (s0[~top] + 3 * (s0[~top] // 2)) + K1 * (s1[~top] + 3 * (s1[~top] // 2))).sum()),
            "svx_sample_norms": int((nj["k"].astype(np.int64) * nj["n"] * (D * 4 + 4) +
                                     nj["ko"].astype(np.int64) * nj["per"] * D * 4).sum()),
            "svx_score_pairs": int((self.nsamp * (2 * D * 4 + 4)).sum()),
            "svx_del_knob": int((self.nsamp * 8).sum()),
            "svx_dense_costs": int(((s0[top] + s1[top]) * D * 4 + s0[top] * s1[top] * 4).sum()),
            "svx_dense_dp": int(((s0[top] + 1) * (s1[top] + 1) * 5).sum()),
            "svx_banded_costs_level0": int((K0 * s0[band_l0] * D * 4 + K1 * s1[band_l0] * D * 4 + 4 * T[band_l0] * A[band_l0] * B).sum()),
            "svx_banded_costs_coarse": int(((s0[band_co] + s1[band_co]) * D * 4 + 4 * A[band_co] * B).sum()),
            "svx_banded_dp_level0": int(((A[band_l0] + 2) * B * (4 * T[band_l0] + 9)).sum()),
            "svx_banded_dp_coarse": int(((A[band_co] + 2) * B * (4 + 9)).sum()),
        }
        return out

    def cost_flops(self):
        """Algorithmic FLOPs of the cost launchers (2*D per dot product)."""
        D, B = self.dim, self.band
        l0 = self.rec_level == 0
        top = self.rec_level == self.depth[self.rec_pair]
        band_l0, band_co = self.banded & l0, self.banded & ~l0
        return {
            "svx_banded_costs_level0": int(2 * D * (self.T[band_l0] * self.A[band_l0] * B).sum()),
            "svx_banded_costs_coarse": int(2 * D * (self.A[band_co] * B).sum()),
            "svx_dense_costs": int(2 * D * (self.rs0[top] * self.rs1[top]).sum()),
        }

    def dp_cells(self):
        """DP cells of the batch: banded nodes (A+2)*B per banded level plus the
        (s0+1)(s1+1) dense nodes of the coarsest level."""
        top = self.rec_level == self.depth[self.rec_pair]
        return int(((self.A[self.banded] + 2) * self.band).sum() + ((self.rs0[top] + 1) * (self.rs1[top] + 1)).sum())

    def results(self):
        """Device -> host read of the level-0 alignment records, scores, penalties and status.
        Returns a list (one per pair) of dicts with 'recs' (structured array, forward order),
        'del_penalty' (per level), 'status'."""
        o = self.off
        lo = int(o["recs"].min()) if self.R else 0
        blob = self.arena[lo:self.nbytes].cpu().numpy()
        dp_lo = int(o["delpen"].min()) if self.R else 0
        dp_blob = self.arena[dp_lo:dp_lo + self.R * _ALIGN].cpu().numpy()
        out = []
        for p in range(self.P):
            r0 = int(self.first[p])
            cap = int(self.rec_cap[r0])
            n = int(blob[o["nrecs"][r0] - lo:o["nrecs"][r0] - lo + 4].view(np.int32)[0])
            st = blob[o["status"][r0] - lo:o["status"][r0] - lo + 8].view(np.int32).copy()
            top = int(self.top_rec[p])
            st_dense = int(blob[o["status"][top] - lo + 4:o["status"][top] - lo + 8].view(np.int32)[0])
            recs = blob[o["recs"][r0] - lo:o["recs"][r0] - lo + cap * capi.REC.itemsize].view(capi.REC)
            recs = recs[cap - min(n, cap):cap].copy()
            pens = [float(dp_blob[o["delpen"][r] - dp_lo:o["delpen"][r] - dp_lo + 8].view(np.float64)[0])
                    for r in range(r0, r0 + int(self.nlev[p]))]
            status = int(st[0]) | st_dense
            for r in range(r0 + 1, r0 + int(self.nlev[p])):
                status |= int(blob[o["status"][r] - lo:o["status"][r] - lo + 4].view(np.int32)[0])
            out.append({"recs": recs, "nrecs": n, "del_penalty": pens, "status": status})
        return out

    # ---- debug / parity accessors (device -> host copies of intermediates) -----------------------
    def level_record(self, p, level):
        return int(self.first[p]) + level

    def fetch(self, key, r, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        return self._get(self.off[key][r], n, dtype).reshape(shape).copy()

    def fetch_vecs(self, r, side):
        k = self.k0 if side == 0 else self.k1
        s = int(self.rs0[r] if side == 0 else self.rs1[r])
        ptr = int(self.vec0_ptr[r] if side == 0 else self.vec1_ptr[r])
        n = k * s * self.dim
        if n == 0:
            return np.zeros((k, s, self.dim), dtype=np.float32)
        if self.rec_level[r] == 0:
            raise ValueError("level-0 vectors live in the caller's tensors")
        off = ptr - self.base
        return self._get(off, n * 4, np.float32).reshape(k, s, self.dim).copy()


def records_to_alignments(recs):
    """SvxAlignRec array -> the reference's list of (list[int], list[int]) + float64 scores
    (dp_utils.py:126-143)."""
    al = [(list(range(int(r["x_end"] - r["nx"]), int(r["x_end"]))),
           list(range(int(r["y_end"] - r["ny"]), int(r["y_end"])))) for r in recs]
    return al, recs["score"].astype(np.float64).copy()
