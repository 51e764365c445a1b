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
order; ``width_over2 < 3`` is raised to 3 (:391-393); the inputs are normalised IN PLACE (:396-397):
torch CUDA tensors directly, writable numpy arrays by copying the normalised rows back
(``writeback=True``, the default of ``vecalign``; ``writeback=False`` skips the device->host copy for
callers that do not reuse the arrays - the reference's own callers never do).  Host torch tensors are
an extension of the interface and are left untouched.

``vecalign_batch`` is not re-entrant per device: concurrent calls from several Python threads serialise
on a lock around the pinned staging ring and the side streams.
"""
import logging
import os
import threading

import numpy as np
import torch

from . import capi
from .engine import BatchRun, WidenJobs, host_threads, records_to_alignments, row_sources

logger = logging.getLogger('vecalign')

_MODES = {"exact": capi.SVX_COST_EXACT, "fast": capi.SVX_COST_FAST, "tc": capi.SVX_COST_TC}


def _device():
    if not torch.cuda.is_available():
        raise capi.SvxError("no CUDA device: speech_vecalign_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


_SIDE_STREAMS = {}


def _side_stream(dev, name):
    """Long-lived side streams per device (copy stream, chunk streams): the caching allocator keys its
    free blocks by stream, so fresh streams per call would cudaMalloc every batch."""
    key = (dev.type, dev.index, name)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


# Pageable host inputs (plain numpy arrays - what the reference's callers pass, dp_utils.py:381) go through a
# ring of pinned staging slots: a multi-threaded memcpy into a slot (svx_host_memcpy), then an asynchronous copy
# from it.  The driver's own pageable path is a single staging thread (~10 GB/s measured; the link takes 55).
_STAGE_SLOT = 64 << 20
_STAGE = {"slots": [], "events": [], "next": 0}
_STAGE_MIN = 4 << 20                      # smaller arrays are not worth a slot round trip
_LOCK = threading.RLock()                 # the staging ring and the side streams are process-wide


def _stage_threads():
    if os.environ.get("SVX_STAGE_THREADS"):
        return max(1, int(os.environ["SVX_STAGE_THREADS"]))
    return host_threads(16)                  # this rank's share of the host cores


def _stage_slot():
    """Next pinned slot of the ring, free again (its last device copy has completed)."""
    nslots = int(os.environ.get("SVX_STAGE_SLOTS", "8"))
    st = _STAGE
    if len(st["slots"]) < nslots:
        st["slots"].append(torch.empty(_STAGE_SLOT, dtype=torch.uint8, pin_memory=True))
        st["events"].append(None)
        i = len(st["slots"]) - 1
    else:
        i = st["next"] % nslots
        if st["events"][i] is not None:
            st["events"][i].synchronize()
    st["next"] = i + 1
    return i


def _staged_copy(src_ptr, nbytes, dst):
    """pageable host memory [src_ptr, +nbytes) -> device tensor dst (contiguous), on the current stream."""
    flat = dst.view(torch.uint8).view(-1)
    nthreads = _stage_threads()
    stream = torch.cuda.current_stream(dst.device)
    off = 0
    while off < nbytes:
        n = min(_STAGE_SLOT, nbytes - off)
        i = _stage_slot()
        slot = _STAGE["slots"][i]
        capi.check(capi.lib().svx_host_memcpy(slot.data_ptr(), src_ptr + off, n, nthreads), "svx_host_memcpy")
        flat[off:off + n].copy_(slot[:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        _STAGE["events"][i] = ev
        off += n


def _copy_back(t, h):
    """device tensor -> writable numpy array of the same shape (the reference normalises its inputs in place,
    dp_utils.py:396-397): through the pinned ring, multi-threaded host copy out of each slot."""
    flat = t.reshape(-1).view(torch.uint8)
    nbytes = flat.numel()
    if not h.flags.c_contiguous or nbytes < _STAGE_MIN or h.dtype != np.float32 or t.dtype != torch.float32:
        h[...] = t.cpu().numpy()
        return
    nthreads = _stage_threads()
    off, pending = 0, []
    while off < nbytes or pending:
        while off < nbytes and len(pending) < 4:
            n = min(_STAGE_SLOT, nbytes - off)
            i = _stage_slot()
            slot = _STAGE["slots"][i]
            slot[:n].copy_(flat[off:off + n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(t.device))
            _STAGE["events"][i] = ev
            pending.append((i, off, n, ev))
            off += n
        i, o, n, ev = pending.pop(0)
        ev.synchronize()
        capi.check(capi.lib().svx_host_memcpy(h.ctypes.data + o, _STAGE["slots"][i].data_ptr(), n, nthreads), "svx_host_memcpy")


def _to_device(v, dev):
    """(K, N, D) on the device, fp32 or — an extension for embeddings kept in their on-disk dtype — fp16
    (widened on the device by _widen).  Returns (tensor, host_array_or_None)."""
    if isinstance(v, torch.Tensor):
        if v.dtype not in (torch.float32, torch.float16) or v.dim() != 3:
            raise ValueError("Buffer dtype mismatch, expected 'float' (K, N, D) tensor")
        if v.is_cuda:
            if not v.is_contiguous():
                raise ValueError("device tensors must be contiguous")
            return v, None
        v = v.contiguous()
        nbytes = v.numel() * v.element_size()
        if v.is_pinned() or nbytes < _STAGE_MIN:
            return v.to(dev, non_blocking=True), None
        t = torch.empty(v.shape, dtype=v.dtype, device=dev)
        _staged_copy(v.data_ptr(), nbytes, t)
        return t, None
    v = np.asarray(v)
    if v.dtype not in (np.float32, np.float16) or v.ndim != 3:
        # the reference's Cython buffers reject anything else (dp_core.pyx:168-171)
        raise ValueError("Buffer dtype mismatch, expected 'float' with ndim=3")
    c = np.ascontiguousarray(v)
    if c.nbytes < _STAGE_MIN:
        return torch.from_numpy(c).to(dev, non_blocking=True), v
    t = torch.empty(c.shape, dtype=torch.float32 if c.dtype == np.float32 else torch.float16, device=dev)
    _staged_copy(c.ctypes.data, c.nbytes, t)
    return t, v


def _widen(pairs_dev, dev):
    """fp16 device tensors -> fp32 working tensors (svx_gather_doc_embedding with the identity table: a
    widening copy that also zeroes rows containing NaNs, as make_doc_embedding does).  One launch for
    all fp16 tensors of the list; fp32 tensors pass through."""
    todo = [(i, s) for i, pr in enumerate(pairs_dev) for s in (0, 1) if pr[s].dtype == torch.float16]
    if not todo:
        return pairs_dev, None
    out = [list(pr) for pr in pairs_dev]
    srcs, dsts = [], []
    for i, s in todo:
        src = pairs_dev[i][s]
        dst = torch.empty(src.shape, dtype=torch.float32, device=dev)
        out[i][s] = dst
        srcs.append(src)
        dsts.append(dst)
    job = WidenJobs(srcs, dsts, int(pairs_dev[0][0].shape[2]), dev)
    job.run()
    # the launches above are asynchronous: the caller keeps `job` (pinned descriptors, sources) alive until it has synchronised
    return [tuple(pr) for pr in out], job


def vecalign_batch(pairs, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
                   costs_sample_size, num_samps_for_norm, cost_mode="exact", debug=False, writeback=False,
                   norms0=None, norms1=None, sync=True, output="stack", seeds=None, streams=None):
    with _LOCK:
        return _vecalign_batch(pairs, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
                               costs_sample_size, num_samps_for_norm, cost_mode, debug, writeback, norms0, norms1, sync,
                               output, seeds, streams)


def _vecalign_batch(pairs, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
                    costs_sample_size, num_samps_for_norm, cost_mode, debug, writeback, norms0, norms1, sync, output,
                    seeds, streams):
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
    A pair whose traceback fails on the device (the reference's IndexError / 'traceback bug',
    dp_utils.py:123-124) raises in stack mode; with output="records" its entry carries the non-zero
    'status' and the other pairs of the batch are returned normally (seg_align writes those and skips
    the failing one).
    """
    if width_over2 < 3:
        logger.warning('width_over2 was set to %d, which does not make sense. increasing to 3.', width_over2)
        width_over2 = 3
    dev = _device()
    pairs = list(pairs)
    if not pairs:
        return []

    def _meta(v):
        if isinstance(v, torch.Tensor):
            if v.dtype not in (torch.float32, torch.float16) or v.dim() != 3:
                raise ValueError("Buffer dtype mismatch, expected 'float' (K, N, D) tensor")
            return tuple(v.shape), v.is_cuda
        a = np.asarray(v)
        if a.dtype not in (np.float32, np.float16) or a.ndim != 3:
            # the reference's Cython buffers reject anything else (dp_core.pyx:168-171)
            raise ValueError("Buffer dtype mismatch, expected 'float' with ndim=3")
        return tuple(a.shape), False

    metas = [(_meta(v0), _meta(v1)) for v0, v1 in pairs]
    k0, k1, dim = metas[0][0][0][0], metas[0][1][0][0], metas[0][0][0][2]
    for (sh0, _), (sh1, _) in metas:
        if sh0[2] != sh1[2]:
            raise AssertionError("embedding dimensions differ")
        if sh0[0] != k0 or sh1[0] != k1 or sh0[2] != dim:
            raise ValueError("all pairs of a batch must share (K0, K1, D)")
    if norms0 is not None and tuple(norms0.shape) != tuple(metas[0][0][0][:2]):
        raise Exception('norms0 wrong shape')          # dp_utils.py:429-432
    if norms1 is not None and tuple(norms1.shape) != tuple(metas[0][1][0][:2]):
        raise Exception('norms1 wrong shape')
    if (norms0 is not None or norms1 is not None) and len(pairs) != 1:
        raise ValueError("norms0/norms1 are a single-pair option")

    # Host inputs of a large batch are pipelined: all host->device copies are queued on a copy stream up
    # front (they run back to back at PCIe rate), and the pairs are planned and enqueued chunk by chunk,
    # a chunk's kernels starting as soon as its embeddings have landed.  The np.random draws stay in
    # input order because chunks are planned in input order.
    P = len(pairs)
    any_host = any(not c0 or not c1 for (_, c0), (_, c1) in metas)
    nchunks = 1
    if any_host and sync and not debug and P >= 16:
        # ~512 MB of embeddings per chunk (10 ms of PCIe): enough chunks to overlap copy and compute, few
        # enough that per-chunk planning (RNG replay, descriptors) stays off the critical path
        total_bytes = sum(4 * (sh0[0] * sh0[1] + sh1[0] * sh1[1]) * sh0[2] for (sh0, _), (sh1, _) in metas)
        cap = int(os.environ.get("SVX_PIPE_CHUNKS", "8"))
        nchunks = max(1, min(cap, P // 4, int(total_bytes * cap // (4096 << 20))))
    # equal chunks except the last two (3/4 and 1/2 of a chunk): what is left to compute after the last copy has landed
    # is the tail of the timeline
    wts = [1.0] * nchunks
    if nchunks >= 4:
        wts[-2], wts[-1] = 0.75, 0.5
    acc_w = np.concatenate([[0.0], np.cumsum(wts)]) / sum(wts)
    bounds = [int(round(P * a)) for a in acc_w]
    bounds[-1] = P
    cur = torch.cuda.current_stream(dev)
    piped = nchunks > 1
    copy_stream = _side_stream(dev, "copy") if piped else cur
    dv, hosts, landed = [None] * P, [None] * P, [None] * nchunks

    def queue_copies(c):
        with torch.cuda.stream(copy_stream):
            for p in range(bounds[c], bounds[c + 1]):
                t0, h0 = _to_device(pairs[p][0], dev)
                t1, h1 = _to_device(pairs[p][1], dev)
                dv[p], hosts[p] = (t0, t1), (h0, h1)
            landed[c] = torch.cuda.Event(enable_timing=True)
            landed[c].record(copy_stream)

    ngroups = (4 if P >= 8 else 1) if streams is None else int(streams)
    runs, keepalive = [], []
    import os as _os, time as _time
    _tr = _os.environ.get("SVX_TRACE")
    _t0 = _time.perf_counter()
    def _mark(what):
        if _tr:
            print(f"[trace] {1e3 * (_time.perf_counter() - _t0):8.2f} ms  {what}", flush=True)
    _ev = []
    if piped:
        fork = torch.cuda.Event(enable_timing=True)
        fork.record(cur)
        copy_stream.wait_event(fork)
    queue_copies(0)
    for c in range(nchunks):
        if c + 1 < nchunks:
            queue_copies(c + 1)        # the DMA queue stays one chunk ahead of the planner
        _mark(f"copies queued through chunk {min(c + 1, nchunks - 1)}")
        lo, hi = bounds[c], bounds[c + 1]
        # each chunk runs its whole chain on its own stream (4 in rotation): the latency-bound kernels of
        # one chunk (wavefront DPs) overlap the other chunks' work and copies
        st = _side_stream(dev, ("chunk", c % 4)) if piped else cur
        if piped and c < 4:
            st.wait_event(fork)
        with torch.cuda.stream(st):
            sources = None
            if any(t.dtype == torch.float16 for pr in dv[lo:hi] for t in pr):
                if all(t.dtype == torch.float16 for pr in dv[lo:hi] for t in pr):
                    # embeddings in their on-disk dtype: the level-0 prologue reads the fp16 rows itself (identity
                    # sources: widened exactly, rows holding a NaN zeroed as make_doc_embedding does) and writes the
                    # normalised fp32 working tensors - no widening pass over HBM
                    sources = (row_sources([t0 for t0, _ in dv[lo:hi]]), row_sources([t1 for _, t1 in dv[lo:hi]]))
                    keepalive.append(dv[lo:hi])
                    dv[lo:hi] = [(torch.empty(t0.shape, dtype=torch.float32, device=dev),
                                  torch.empty(t1.shape, dtype=torch.float32, device=dev)) for t0, t1 in dv[lo:hi]]
                else:
                    if piped:
                        st.wait_event(landed[c])
                    dv[lo:hi], alive = _widen(dv[lo:hi], dev)
                    keepalive.append(alive)
            run = BatchRun([t0.data_ptr() for t0, _ in dv[lo:hi]], [t1.data_ptr() for _, t1 in dv[lo:hi]],
                           [t0.shape[1] for t0, _ in dv[lo:hi]], [t1.shape[1] for _, t1 in dv[lo:hi]],
                           k0, k1, dim, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
                           costs_sample_size, num_samps_for_norm, dev, cost_mode=_MODES[cost_mode],
                           norms0=norms0, norms1=norms1, keep_dense_csum=debug,
                           seeds=None if seeds is None else list(seeds)[lo:hi], sources=sources)
            _mark(f"chunk {c} planned")
            if piped:
                st.wait_event(landed[c])
            if _tr and piped:
                e_a = torch.cuda.Event(enable_timing=True); e_a.record(st)
            run.run(ngroups=1 if piped else ngroups)
            if sync and not debug:
                run.start_fetch()          # results of this chunk travel while the next chunks run
            if _tr and piped:
                e_b = torch.cuda.Event(enable_timing=True); e_b.record(st); _ev.append((c, e_a, e_b))
            _mark(f"chunk {c} enqueued")
        runs.append(run)
    if piped:
        for c in range(min(4, nchunks)):
            cur.wait_stream(_side_stream(dev, ("chunk", c)))
    run = runs[0]
    if not sync:
        run._keepalive = (keepalive, dv)        # the kernels are still running: inputs and descriptors stay alive with the run
        return run
    res = []
    if _tr:
        torch.cuda.synchronize(dev)
        _mark("device idle")
        for c, e_a, e_b in _ev:
            print(f"[trace] chunk {c}: copies landed {fork.elapsed_time(landed[c]):7.2f}  kernels start {fork.elapsed_time(e_a):7.2f}  end {fork.elapsed_time(e_b):7.2f} ms", flush=True)
    for r_ in runs:
        res.extend(r_.results())
    _mark("results read")
    if writeback:
        for (t0, t1), (h0, h1) in zip(dv, hosts):
            if h0 is not None and h0.flags.writeable:
                _copy_back(t0, h0)
            if h1 is not None and h1.flags.writeable:
                _copy_back(t1, h1)
    if output == "records":
        return res
    for p, r in enumerate(res):
        if r["status"]:
            # the reference fails here with IndexError / 'traceback bug' (dp_utils.py:123-124)
            raise Exception('traceback bug (device status %d for pair %d)' % (r["status"], p))
    stacks = []
    for p, r in enumerate(res):
        al, sc = records_to_alignments(r["recs"])
        st = {0: {"final_alignments": al, "alignment_scores": sc, "del_penalty": np.float64(r["del_penalty"][0])}}
        for lvl in range(1, len(r["del_penalty"])):
            st[lvl] = {"del_penalty": np.float64(r["del_penalty"][lvl])}
        if debug:
            _fill_debug(st, run, p, dv[p], final_alignment_types)
        stacks.append(st)
    return stacks


def vecalign(vecs0, vecs1, final_alignment_types, del_percentile_frac, width_over2, max_size_full_dp,
             costs_sample_size, num_samps_for_norm, norms0=None, norms1=None, cost_mode="exact", debug=False,
             writeback=True):
    """Reference signature (dp_utils.py:381-390) + keyword-only extras; returns the ``stack``.  Like the
    reference (:396-397) it leaves vecs0 / vecs1 normalised: numpy arrays get the normalised rows copied
    back unless writeback=False."""
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
