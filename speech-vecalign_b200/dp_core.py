"""Drop-in for the reference's native module ``svecalign.vecalign.dp_core`` (dp_core.pyx): the same
six callables with the same argument order, dtypes and return shapes, each executed by the
sm_100a kernels of libsvx.so on one job.  These per-function entry points exist for parity work
(device -> host copies on every call); the production path is ``dp_utils.vecalign_batch``.

  make_x_y_offsets   dp_core.pyx:24-34
  make_dense_costs   dp_core.pyx:36-77
  dense_dp           dp_core.pyx:79-141
  score_path         dp_core.pyx:143-161
  make_sparse_costs  dp_core.pyx:165-267
  sparse_dp          dp_core.pyx:269-404
plus the numeric helpers of dp_utils.py that became kernels: make_norm1 (:32-40),
downsample_vectors (:362-378), compute_norms given the sampled indices (:326-359),
del_penalty_from_scores (DeletionKnob :43-79), dense_path / sparse_path (traceback + path glue).
"""
import numpy as np
import torch

from . import capi
from .engine import records_to_alignments


def _dev():
    if not torch.cuda.is_available():
        raise capi.SvxError("no CUDA device: speech_vecalign_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _up(a, dtype=None):
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype))
    if a.size == 0:
        return torch.empty(max(a.size, 4), dtype=torch.uint8, device=_dev())[:0]
    return torch.from_numpy(a).to(_dev())


def _f32(a, ndim):
    a = np.asarray(a)
    if a.dtype != np.float32 or a.ndim != ndim:
        raise ValueError("Buffer dtype mismatch, expected 'float' with ndim=%d" % ndim)
    return np.ascontiguousarray(a)


def _ptr(t):
    return t.data_ptr() if t.numel() else 0


def _launch(fn, name, job, *extra):
    jd = torch.from_numpy(job.view(np.uint8).reshape(-1).copy()).to(_dev())
    stream = torch.cuda.current_stream().cuda_stream
    capi.check(fn(jd.data_ptr(), capi.hptr(job), job.shape[0], *extra, stream), name)
    torch.cuda.synchronize()


def _empty(n, dtype):
    return torch.empty(max(int(n), 1), dtype=dtype, device=_dev())


def make_x_y_offsets(alignment_types):
    for x, y in alignment_types:
        assert (x > 0)
        assert (y > 0)
    return (np.array([x for x, _ in alignment_types], dtype=np.int32),
            np.array([y for _, y in alignment_types], dtype=np.int32))


# ------------------------------------------------------------------------------------------------
def make_norm1(vecs0, out=None):
    """dp_utils.py:32-40 — in place on `vecs0` (numpy, like the reference) via the GPU."""
    v = _f32(vecs0, 3)
    t = _up(v)
    job = np.zeros(1, dtype=capi.ROWS)
    job["ptr"], job["nrows"] = _ptr(t), v.shape[0] * v.shape[1]
    _launch(capi.lib().svx_normalize_rows, "svx_normalize_rows", job, v.shape[2])
    res = t.cpu().numpy().reshape(v.shape) if v.size else v
    vecs0[...] = res
    return vecs0


def downsample_vectors(vecs1):
    """dp_utils.py:362-378."""
    v = _f32(vecs1, 3)
    k, n, d = v.shape
    t = _up(v)
    out = _empty(k * (n // 2) * d, torch.float32)
    mean = _empty(k * d, torch.float32)
    job = np.zeros(1, dtype=capi.DOWN)
    job["in"], job["out"], job["mean"], job["k"], job["n"] = _ptr(t), out.data_ptr(), mean.data_ptr(), k, n
    _launch(capi.lib().svx_downsample, "svx_downsample", job, d)
    return out[:k * (n // 2) * d].cpu().numpy().reshape(k, n // 2, d)


def compute_norms_from_samples(vecs0, vecs1, idx):
    """dp_utils.py:326-359 with the sampled row ids `idx` (K1, per) given (host RNG)."""
    v0, v1 = _f32(vecs0, 3), _f32(vecs1, 3)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    k, n, d = v0.shape
    ko, no, _ = v1.shape
    t0, t1, ti = _up(v0), _up(v1), _up(idx)
    mbar = _empty(d, torch.float64)
    norms = _empty(k * n, torch.float32)
    job = np.zeros(1, dtype=capi.NORM)
    job["vecs"], job["other"], job["idx"], job["mbar"], job["norms"] = _ptr(t0), _ptr(t1), _ptr(ti), mbar.data_ptr(), norms.data_ptr()
    job["k"], job["n"], job["ko"], job["no"], job["per"] = k, n, ko, no, idx.shape[1]
    _launch(capi.lib().svx_sample_norms, "svx_sample_norms", job, d)
    return norms[:k * n].cpu().numpy().reshape(k, n)


def score_path(xx, yy, norm1, norm2, vecs1, vecs2, out, cost_mode=capi.SVX_COST_EXACT):
    """dp_core.pyx:143-161 — fills `out`."""
    for a in (xx, yy):
        if np.asarray(a).dtype != np.int32:
            raise ValueError("Buffer dtype mismatch, expected 'int'")
    n1, n2 = _f32(norm1, 1), _f32(norm2, 1)
    v1, v2 = _f32(vecs1, 2), _f32(vecs2, 2)
    assert out.dtype == np.float32
    n = int(np.asarray(xx).shape[0])
    tx, ty, tn1, tn2, tv1, tv2 = _up(xx), _up(yy), _up(n1), _up(n2), _up(v1), _up(v2)
    sc = _empty(n, torch.float32)
    job = np.zeros(1, dtype=capi.SCORE)
    job["e"], job["f"], job["norm_e"], job["norm_f"] = _ptr(tv1), _ptr(tv2), _ptr(tn1), _ptr(tn2)
    job["xi"], job["yi"], job["scores"] = _ptr(tx), _ptr(ty), sc.data_ptr()
    job["ne"], job["nf"], job["nsamp"] = v1.shape[0], v2.shape[0], n
    _launch(capi.lib().svx_score_pairs, "svx_score_pairs", job, v1.shape[1], cost_mode)
    out[...] = sc[:n].cpu().numpy()


def del_penalty_from_scores(scores, frac):
    """DeletionKnob(samp, 0, max(samp)).percentile_frac_to_del_penalty(frac) (dp_utils.py:43-79)."""
    s = _f32(scores, 1)
    ts = _up(s)
    dp = _empty(1, torch.float64)
    job = np.zeros(1, dtype=capi.SCORE)
    job["scores"], job["del_penalty"], job["nsamp"] = _ptr(ts), dp.data_ptr(), s.shape[0]
    _launch(capi.lib().svx_del_knob, "svx_del_knob", job, float(frac))
    return np.float64(dp.cpu().numpy()[0])


# ------------------------------------------------------------------------------------------------
def _dense_job(vecs0, vecs1, norm0, norm1, offset0, offset1):
    v0, v1 = _f32(vecs0, 3), _f32(vecs1, 3)
    n0, n1 = _f32(norm0, 2), _f32(norm1, 2)
    assert v0.shape[0] > offset0 and v1.shape[0] > offset1
    assert n0.shape[0] > offset0 and n1.shape[0] > offset1
    s0, s1, d = v0.shape[1], v1.shape[1], v0.shape[2]
    assert n0.shape[1] == s0 and n1.shape[1] == s1 and v1.shape[2] == d
    if offset0 or offset1:
        raise capi.SvxError("offset0/offset1 != 0 are never used by the reference's callers (dp_utils.py:465-468)")
    return v0[offset0], v1[offset1], n0[offset0], n1[offset1], s0, s1, d


def make_dense_costs(vecs0, vecs1, norm0, norm1, offset0=0, offset1=0, cost_mode=capi.SVX_COST_EXACT):
    """dp_core.pyx:36-77."""
    a, b, na, nb, s0, s1, d = _dense_job(vecs0, vecs1, norm0, norm1, offset0, offset1)
    ta, tb, tna, tnb = _up(a), _up(b), _up(na), _up(nb)
    costs = _empty(s0 * s1, torch.float32)
    job = np.zeros(1, dtype=capi.DENSE)
    job["v0"], job["v1"], job["n0"], job["n1"], job["costs"] = _ptr(ta), _ptr(tb), _ptr(tna), _ptr(tnb), costs.data_ptr()
    job["s0"], job["s1"] = s0, s1
    if cost_mode == capi.SVX_COST_TC:      # tcgen05 path: residual-plane scratch; TMA descriptors encoded on the host
        lo0, lo1 = _empty(max(1, s0 * d), torch.float32), _empty(max(1, s1 * d), torch.float32)
        job["lo0"], job["lo1"] = lo0.data_ptr(), lo1.data_ptr()
        blobs = np.zeros((1, 4, 128), dtype=np.uint8)
        capi.check(capi.lib().svx_dense_tmaps_encode(capi.hptr(job), 1, d, capi.hptr(blobs)), "svx_dense_tmaps_encode")
        tm = _up(blobs.reshape(-1))
        job["tmap0"], job["tmap1"] = tm.data_ptr(), tm.data_ptr() + 256
    _launch(capi.lib().svx_dense_costs, "svx_dense_costs", job, d, cost_mode)
    return costs[:s0 * s1].cpu().numpy().reshape(s0, s1)


def dense_dp(alignment_cost, pen, target_sizes=None, want_path=False):
    """dp_core.pyx:79-141 -> (csum float64, bp int32).  With want_path=True also returns the
    search path (list of (x,y)) that dp_utils.py:479-489 derives from the traceback: for the same
    level (target_sizes=None) or upsampled + extended to target_sizes=(t0,t1)."""
    cost = _f32(alignment_cost, 2)
    s0, s1 = cost.shape
    tc = _up(cost)
    bp = _empty((s0 + 1) * (s1 + 1), torch.uint8)
    csum = _empty((s0 + 1) * (s1 + 1), torch.float64)
    pen_t = torch.tensor([float(pen)], dtype=torch.float64, device=_dev())
    ups = target_sizes is not None
    t0, t1 = target_sizes if ups else (s0, s1)
    plen = capi.lib().svx_path_len(s0, s1, int(t0), int(t1), int(ups))
    ypath = torch.full((max(plen, 1),), -1, dtype=torch.int32, device=_dev())
    status = torch.zeros(1, dtype=torch.int32, device=_dev())
    job = np.zeros(1, dtype=capi.DENSE)
    job["costs"], job["del_penalty"], job["bp"], job["csum"] = _ptr(tc), pen_t.data_ptr(), bp.data_ptr(), csum.data_ptr()
    job["ypath"], job["status_d"] = ypath.data_ptr(), status.data_ptr()
    job["s0"], job["s1"], job["t0"], job["t1"], job["upsample"], job["path_len"] = s0, s1, t0, t1, int(ups), plen
    _launch(capi.lib().svx_dense_dp, "svx_dense_dp", job)
    if int(status.item()):
        raise Exception('got unknown value')
    n = (s0 + 1) * (s1 + 1)
    cs = csum[:n].cpu().numpy().reshape(s0 + 1, s1 + 1)
    b = bp[:n].cpu().numpy().reshape(s0 + 1, s1 + 1).astype(np.int32)
    if want_path:
        yp = ypath[:plen].cpu().numpy()
        return cs, b, [(int(a - y), int(y)) for a, y in enumerate(yp)]
    return cs, b


def _band_job(s0, s1, k0, k1, a_len, alignment_types, width_over2):
    job = np.zeros(1, dtype=capi.BAND)
    job["s0"], job["s1"], job["k0"], job["k1"] = s0, s1, k0, k1
    job["a_len"], job["band"], job["width_over2"], job["ntypes"] = a_len, 2 * width_over2, width_over2, len(alignment_types)
    for t, (x, y) in enumerate(alignment_types):
        job["xo"][0, t], job["yo"][0, t] = x, y
    job["amax"] = max([2] + [x + y for x, y in alignment_types])
    return job


def make_sparse_costs(vecs0, vecs1, norms0, norms1, x_y_path, alignment_types, width_over2,
                      cost_mode=capi.SVX_COST_EXACT):
    """dp_core.pyx:165-267 -> (a_b_feats (T, A, B) float32, b_offset (A,) int32)."""
    v0, v1 = _f32(vecs0, 3), _f32(vecs1, 3)
    n0, n1 = _f32(norms0, 2), _f32(norms1, 2)
    path = np.array(x_y_path).astype(np.int32).reshape(-1, 2)
    assert v0.shape[0] == n0.shape[0] and v1.shape[0] == n1.shape[0]
    assert v0.shape[1] == n0.shape[1] and v1.shape[1] == n1.shape[1]
    mx = max([0] + [x for x, _ in alignment_types])
    my = max([0] + [y for _, y in alignment_types])
    if mx > v0.shape[0]:
        raise Exception('%d x overlaps requrested (via alignment_types), but vecs0 only has %d' % (mx, v0.shape[0]))
    if my > v1.shape[0]:
        raise Exception('%d y overlaps requrested (via alignment_types), but vecs1 only has %d' % (my, v1.shape[0]))
    assert v0.shape[2] == v1.shape[2]
    make_x_y_offsets(alignment_types)
    A, B, T = path.shape[0], 2 * int(width_over2), len(alignment_types)
    # the kernel indexes the path by anti-diagonal (every reference path advances x+y by one)
    assert np.array_equal(path.sum(axis=1), np.arange(A)), "search path must advance x+y by one per point"
    ypath = np.ascontiguousarray(path[:, 1])
    t0, t1, tn0, tn1, typ = _up(v0), _up(v1), _up(n0), _up(n1), _up(ypath)
    costs = _empty(A * T * B, torch.float32)
    job = _band_job(v0.shape[1], v1.shape[1], v0.shape[0], v1.shape[0], A, alignment_types, int(width_over2))
    job["v0"], job["v1"], job["n0"], job["n1"], job["ypath"], job["costs"] = _ptr(t0), _ptr(t1), _ptr(tn0), _ptr(tn1), _ptr(typ), costs.data_ptr()
    _launch(capi.lib().svx_banded_costs, "svx_banded_costs", job, v0.shape[2], cost_mode)
    feats = costs[:A * T * B].cpu().numpy().reshape(A, T, B).transpose(1, 0, 2)
    return np.ascontiguousarray(feats), (ypath - int(width_over2)).astype(np.int32)


def sparse_dp(a_b_costs, b_offset_in, alignment_types, del_penalty, x_in_size, y_in_size,
              target_sizes=None, want_traceback=False):
    """dp_core.pyx:269-404 -> (a_b_csum float64, a_b_xp int32, a_b_yp int32, b_offset_out int32).
    With want_traceback=True additionally returns (alignments, scores[, next search path]) as
    produced on the device (dp_utils.py:105-143 and, when target_sizes is given, :199-275)."""
    costs = _f32(a_b_costs, 3)
    boff = np.ascontiguousarray(b_offset_in)
    assert boff.dtype == np.int32
    T, A, B = costs.shape
    w = B // 2
    ypath = (boff + w).astype(np.int32)
    tc = _up(np.ascontiguousarray(costs.transpose(1, 0, 2)))
    typ = _up(ypath)
    bp = _empty((A + 2) * B, torch.uint8)
    csum = _empty((A + 2) * B, torch.float64)
    cap = x_in_size + y_in_size + 2
    recs = _empty(cap * capi.REC.itemsize, torch.uint8)
    nrecs = torch.zeros(1, dtype=torch.int32, device=_dev())
    status = torch.zeros(1, dtype=torch.int32, device=_dev())
    pen_t = torch.tensor([float(del_penalty)], dtype=torch.float64, device=_dev())
    job = _band_job(x_in_size, y_in_size, 0, 0, A, alignment_types, w)
    job["ypath"], job["costs"], job["del_penalty"] = _ptr(typ), _ptr(tc), pen_t.data_ptr()
    job["bp"], job["csum"], job["recs"], job["nrecs"], job["status_d"] = bp.data_ptr(), csum.data_ptr(), recs.data_ptr(), nrecs.data_ptr(), status.data_ptr()
    job["rec_cap"] = cap
    nxt = None
    if target_sizes is not None:
        t0, t1 = target_sizes
        nlen = capi.lib().svx_path_len(x_in_size, y_in_size, int(t0), int(t1), 1)
        nxt = torch.full((max(nlen, 1),), -1, dtype=torch.int32, device=_dev())
        job["next_ypath"], job["t0"], job["t1"], job["next_len"] = nxt.data_ptr(), t0, t1, nlen
    _launch(capi.lib().svx_banded_dp, "svx_banded_dp", job)
    n = (A + 2) * B
    cs = csum[:n].cpu().numpy().reshape(A + 2, B)
    b = bp[:n].cpu().numpy().reshape(A + 2, B)
    tx = np.array([x for x, _ in alignment_types] + [0, 1], dtype=np.int32)
    ty = np.array([y for _, y in alignment_types] + [1, 0], dtype=np.int32)
    xp = np.full(b.shape, -42, dtype=np.int32)
    yp = np.full(b.shape, -42, dtype=np.int32)
    ok = b != capi.SVX_BP_NONE
    xp[ok], yp[ok] = tx[b[ok]], ty[b[ok]]
    boff_out = np.concatenate([[boff[0], boff[0]], boff + 1]).astype(np.int32)
    if not want_traceback:
        return cs, xp, yp, boff_out
    st = int(status.item())
    if st:
        raise Exception('traceback bug (device status %d)' % st)
    k = int(nrecs.item())
    r = recs[:cap * capi.REC.itemsize].cpu().numpy().view(capi.REC)[cap - k:]
    al, sc = records_to_alignments(r)
    extra = (al, sc)
    if nxt is not None:
        yn = nxt[:nlen].cpu().numpy()
        extra = extra + ([(int(a - y), int(y)) for a, y in enumerate(yn)],)
    return (cs, xp, yp, boff_out) + extra
