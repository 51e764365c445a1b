"""Drop-in for ``svecalign.vecalign.dp_utils`` (reference: svecalign/vecalign/dp_utils.py).

``vecalign(vecs0, vecs1, final_alignment_types, del_percentile_frac, width_over2,
max_size_full_dp, costs_sample_size, num_samps_for_norm, norms0=None, norms1=None)`` keeps the
reference signature and returns the reference's ``stack`` (dp_utils.py:381-390, 537):
``stack[0]['final_alignments']`` / ``['alignment_scores']`` are what callers read
(vecalign.py:279,292).  All arithmetic runs on the GPU through libsvx.so; there is no CPU path.

Additions: ``vecalign_batch`` (many pairs per call), ``debug=True`` to fill every other stack key
for parity work, ``cost_mode`` ('exact' | 'fast' | 'tc': fast + the coarsest-level cost matrix
as a 3xTF32 tcgen05 GEMM).

Semantics kept from the reference: the global ``np.random`` stream is consumed in the reference's
order; ``width_over2 < 3`` is raised to 3 (:391-393); torch CUDA inputs are normalised in place
(:396-397).  Deviation (documented in DESIGN.md): numpy inputs are not written back unless
``writeback=True`` (the reference's callers never reuse them).
"""
import logging

import numpy as np
import torch

from . import capi
from .engine import BatchRun, records_to_alignments

logger = logging.getLogger('vecalign')

_MODES = {"exact": capi.SVX_COST_EXACT, "fast": capi.SVX_COST_FAST, "tc": capi.SVX_COST_TC}


def _device():
    if not torch.cuda.is_available():
        raise capi.SvxError("no CUDA device: speech_vecalign_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(v, dev):
    """(K, N, D) fp32 on the device.  Returns (tensor, host_array_or_None)."""
    if isinstance(v, torch.Tensor):
        if v.dtype != torch.float32 or v.dim() != 3:
            raise ValueError("Buffer dtype mismatch, expected 'float' (K, N, D) tensor")
        if v.is_cuda:
            if not v.is_contiguous():
                raise ValueError("device tensors must be contiguous")
            return v, None
        return v.contiguous().to(dev, non_blocking=True), None
    v = np.asarray(v)
    if v.dtype != np.float32 or v.ndim != 3:
        # the reference's Cython buffers reject anything else (dp_core.pyx:168-171)
        raise ValueError("Buffer dtype mismatch, expected 'float' with ndim=3")
    t = torch.from_numpy(np.ascontiguousarray(v)).to(dev, non_blocking=True)
    return t, v


def vecalign_batch(pairs, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
                   costs_sample_size, num_samps_for_norm, cost_mode="exact", debug=False, writeback=False,
                   norms0=None, norms1=None, sync=True, output="stack", seeds=None, streams=None):
    """Align many document pairs in one pass over the GPU.

    pairs: sequence of (vecs0, vecs1), each (K, N, D) fp32 numpy array or torch tensor (host or
    device).  np.random draws are made pair by pair in input order, i.e. exactly as a serial loop of
    the reference would consume the stream.  Returns one reference-style ``stack`` dict per pair, or
    with output="records" the packed device records per pair ({'recs': SvxAlignRec array in path
    order, 'del_penalty': [per level], 'nrecs', 'status'}) without building Python lists.
    seeds: optional per-pair np.random seeds (np.random.seed(seeds[p]) before pair p's draws), which
    makes every pair's result independent of batch order and of the multi-GPU partition.
    streams: pair groups enqueued on separate CUDA streams (default 4 for batches of >= 8 pairs);
    pairs are independent, so results do not depend on it.
    """
    if width_over2 < 3:
        logger.warning('width_over2 was set to %d, which does not make sense. increasing to 3.', width_over2)
        width_over2 = 3
    dev = _device()
    dv, hosts = [], []
    for v0, v1 in pairs:
        t0, h0 = _to_device(v0, dev)
        t1, h1 = _to_device(v1, dev)
        if t0.shape[2] != t1.shape[2]:
            raise AssertionError("embedding dimensions differ")
        dv.append((t0, t1))
        hosts.append((h0, h1))
    if not dv:
        return []
    k0, k1, dim = dv[0][0].shape[0], dv[0][1].shape[0], dv[0][0].shape[2]
    for t0, t1 in dv:
        if t0.shape[0] != k0 or t1.shape[0] != k1 or t0.shape[2] != dim:
            raise ValueError("all pairs of a batch must share (K0, K1, D)")
    if norms0 is not None and tuple(norms0.shape) != tuple(dv[0][0].shape[:2]):
        raise Exception('norms0 wrong shape')          # dp_utils.py:429-432
    if norms1 is not None and tuple(norms1.shape) != tuple(dv[0][1].shape[:2]):
        raise Exception('norms1 wrong shape')
    if (norms0 is not None or norms1 is not None) and len(dv) != 1:
        raise ValueError("norms0/norms1 are a single-pair option")

    run = BatchRun([t0.data_ptr() for t0, _ in dv], [t1.data_ptr() for _, t1 in dv],
                   [t0.shape[1] for t0, _ in dv], [t1.shape[1] for _, t1 in dv],
                   k0, k1, dim, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
                   costs_sample_size, num_samps_for_norm, dev, cost_mode=_MODES[cost_mode],
                   norms0=norms0, norms1=norms1, keep_dense_csum=debug, seeds=seeds)
    run.run(ngroups=(4 if len(dv) >= 8 else 1) if streams is None else int(streams))
    if not sync:
        return run
    res = run.results()
    for p, r in enumerate(res):
        if r["status"]:
            # the reference fails here with IndexError / 'traceback bug' (dp_utils.py:123-124)
            raise Exception('traceback bug (device status %d for pair %d)' % (r["status"], p))
    if output == "records":
        return res
    stacks = []
    for p, r in enumerate(res):
        al, sc = records_to_alignments(r["recs"])
        st = {0: {"final_alignments": al, "alignment_scores": sc, "del_penalty": np.float64(r["del_penalty"][0])}}
        for lvl in range(1, len(r["del_penalty"])):
            st[lvl] = {"del_penalty": np.float64(r["del_penalty"][lvl])}
        if debug:
            _fill_debug(st, run, p, dv[p], final_alignment_types)
        stacks.append(st)
    if writeback:
        for (t0, t1), (h0, h1) in zip(dv, hosts):
            if h0 is not None and h0.flags.writeable:
                h0[...] = t0.cpu().numpy()
            if h1 is not None and h1.flags.writeable:
                h1[...] = t1.cpu().numpy()
    return stacks


def vecalign(vecs0, vecs1, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
             costs_sample_size, num_samps_for_norm, norms0=None, norms1=None, cost_mode="exact", debug=False,
             writeback=False):
    """Reference signature (dp_utils.py:381-390) + keyword-only extras; returns the ``stack``."""
    return vecalign_batch([(vecs0, vecs1)], final_alignment_types, del_percentile_frac, width_over2,
                          max_size_full_dp, costs_sample_size, num_samps_for_norm, cost_mode=cost_mode,
                          debug=debug, writeback=writeback, norms0=norms0, norms1=norms1)[0]


# ---------------------------------------------------------------------------------------------
# debug stack: every key the reference's stack carries, read back from the
# device buffers.  Only used by parity tests / --debug_save_stack.
# ---------------------------------------------------------------------------------------------
def _bp_to_xy(bp, types):
    tx = np.array([x for x, _ in types] + [0, 1], dtype=np.int32)
    ty = np.array([y for _, y in types] + [1, 0], dtype=np.int32)
    xp = np.full(bp.shape, -42, dtype=np.int32)
    yp = np.full(bp.shape, -42, dtype=np.int32)
    ok = bp != capi.SVX_BP_NONE
    xp[ok] = tx[bp[ok]]
    yp[ok] = ty[bp[ok]]
    return xp, yp


def _dense_alignments(bp):
    """dp_utils.py:146-174 on the device-produced backpointer matrix (host unpacking only)."""
    x, y = bp.shape[0] - 1, bp.shape[1] - 1
    out = []
    while not (x == 0 and y == 0):
        c = bp[x, y]
        if c == 0:
            out.append(([x - 1], [y - 1])); x -= 1; y -= 1
        elif c == 1:
            out.append(([], [y - 1])); y -= 1
        elif c == 2:
            out.append(([x - 1], [])); x -= 1
        else:
            raise Exception('got unknown value')
    out.reverse()
    return out


def _fill_debug(st, run, p, dv_pair, final_types):
    nlev = int(run.nlev[p])
    B, w = run.band, run.w
    for lvl in range(nlev):
        r = run.level_record(p, lvl)
        d = st.setdefault(lvl, {})
        s0, s1 = int(run.rs0[r]), int(run.rs1[r])
        d["size0"], d["size1"] = s0, s1
        d["alignment_types"] = list(final_types) if lvl == 0 else [(1, 1)]
        d["v0"] = dv_pair[0].cpu().numpy() if lvl == 0 else run.fetch_vecs(r, 0)
        d["v1"] = dv_pair[1].cpu().numpy() if lvl == 0 else run.fetch_vecs(r, 1)
        d["n0"] = run.fetch("norms0", r, (run.k0, s0), np.float32)
        d["n1"] = run.fetch("norms1", r, (run.k1, s1), np.float32)
        ns = int(run.nsamp[r])
        d["sample_scores"] = run.fetch("scores", r, (ns,), np.float32)
        if run.knob[r] is not None:
            d["sample_x"], d["sample_y"] = run.knob[r]
        if run.rec_level[r] == run.depth[p]:
            d["costs_1to1"] = run.fetch("dcost", r, (s0, s1), np.float32)
            d["x_y_tb"] = run.fetch("dbp", r, (s0 + 1, s1 + 1), np.uint8).astype(np.int32)
            if run.keep_dense_csum:
                d["dense_csum"] = run.fetch("dcsum", r, (s0 + 1, s1 + 1), np.float64)
            d["alignments"] = _dense_alignments(d["x_y_tb"])
        if run.banded[r]:
            A, T = int(run.A[r]), int(run.T[r])
            yp = run.fetch("ypath", r, (A,), np.int32)
            d["searchpath"] = [(int(a - y), int(y)) for a, y in enumerate(yp)]
            d["a_b_costs"] = np.ascontiguousarray(run.fetch("bcost", r, (A, T, B), np.float32).transpose(1, 0, 2))
            d["b_offset"] = (yp - w).astype(np.int32)
            d["a_b_csum"] = run.fetch("bcsum", r, (A + 2, B), np.float64)
            bp = run.fetch("bbp", r, (A + 2, B), np.uint8)
            d["a_b_xp"], d["a_b_yp"] = _bp_to_xy(bp, d["alignment_types"])
            d["new_b_offset"] = np.concatenate([[yp[0] - w, yp[0] - w], yp - w + 1]).astype(np.int32)
            cap = int(run.rec_cap[r])
            n = int(run.fetch("nrecs", r, (1,), np.int32)[0])
            recs = run.fetch("recs", r, (cap,), capi.REC)[cap - min(n, cap):]
            al, sc = records_to_alignments(recs)
            d["final_alignments" if lvl == 0 else "alignments"] = al
            d["alignment_scores"] = sc
