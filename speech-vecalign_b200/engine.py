"""Host-side driver of the B200 alignment path.  The planning itself - level sizes, arena layout, the
reference's RNG call order, job descriptors, launch chain: the control flow of
``svecalign/vecalign/dp_utils.py:381-537`` (vecalign) for many pairs at once - lives in libsvx.so
(csrc/plan.cu, ``svx_plan_*`` in include/svx.h); this module owns the torch tensors the plan is bound
to (device arena, pinned staging block), hands the global ``np.random`` state to the C replay and back
(dp_utils.py:301-302,346 draw from it), and unpacks results.  One call = one batch; a single pair is a
batch of one.

Nothing here computes alignment arithmetic on the CPU.
"""
import ctypes
import os

import numpy as np
import torch

from . import capi

_ALIGN = 256


def host_threads(cap=16):
    """Host threads this process may use for the RNG replay / staging copies: its CPU affinity divided by
    the ranks that share the node (torchrun sets LOCAL_WORLD_SIZE), not os.cpu_count() - eight ranks
    each starting cpu_count() threads oversubscribe the node 8x."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, min(cap, n // local))


_OFF_KEYS = ["idx0", "idx1", "xi", "yi", "delpen", "tmaps", "norms0", "norms1", "vec0", "vec1", "mean0", "mean1", "mbar0",
             "mbar1", "scores", "perm", "dcost", "ddots", "dbp", "dcsum", "ypath", "bcost", "bbp", "bcsum", "recs", "nrecs",
             "status", "dlo0", "dlo1"]       # include/svx.h SVX_PO_*
_PA = {name: i for i, name in enumerate(
    ["first", "nlev", "depth", "rec_pair", "rec_level", "rs0", "rs1", "A", "T", "banded", "rec_cap", "nsamp", "has_draw",
     "top_rec", "tgt_rec", "draw_pair", "draw_high", "draw_count", "draw_off", "draw_begin"])}      # SVX_PA_*
_PA_OFFSETS = 64


def make_params(k0, k1, dim, alignment_types, del_percentile_frac, width_over2, max_size_full_dp, costs_sample_size,
                num_samps_for_norm, cost_mode=capi.SVX_COST_EXACT, keep_all=False, unfused_prologue=False,
                skip_norms0=False, skip_norms1=False):
    """SvxAlignParams (include/svx.h) as a one-element numpy record."""
    types = [(int(x), int(y)) for x, y in alignment_types]
    if len(types) + 2 > capi.SVX_MAX_TYPES:
        raise capi.SvxError("too many alignment types for this build")
    prm = np.zeros(1, dtype=capi.PARAMS)
    prm["k0"], prm["k1"], prm["dim"], prm["ntypes"] = int(k0), int(k1), int(dim), len(types)
    for t, (x, y) in enumerate(types):
        prm["xo"][0, t], prm["yo"][0, t] = x, y
    prm["del_percentile_frac"] = float(del_percentile_frac)
    prm["width_over2"], prm["max_size_full_dp"] = int(width_over2), int(max_size_full_dp)
    prm["costs_sample_size"], prm["num_samps_for_norm"] = int(costs_sample_size), int(num_samps_for_norm)
    prm["cost_mode"], prm["keep_all"], prm["unfused_prologue"] = int(cost_mode), int(bool(keep_all)), int(bool(unfused_prologue))
    prm["skip_norms0"], prm["skip_norms1"] = int(bool(skip_norms0)), int(bool(skip_norms1))
    return prm


class Plan:
    """Owner of one SvxPlan (csrc/plan.cu): level sizes, arena layout, RNG call list, launch chain.
    Host-only until bind(); usable without a GPU (the not-gpu tests check it against the oracle)."""

    def __init__(self, params, n0, n1):
        L = capi.lib()
        self.params = params
        self.n0 = np.ascontiguousarray(n0, dtype=np.int32)
        self.n1 = np.ascontiguousarray(n1, dtype=np.int32)
        handle = ctypes.c_void_p()
        capi.check(L.svx_plan_create(capi.hptr(params), int(self.n0.shape[0]), capi.hptr(self.n0), capi.hptr(self.n1),
                                     ctypes.byref(handle)), "svx_plan_create")
        self.handle = handle
        info = np.zeros(1, dtype=capi.PLAN_INFO)
        capi.check(L.svx_plan_info(handle, capi.hptr(info)), "svx_plan_info")
        self.info = info[0]

    def array(self, which):
        """Copy of one of the plan's int64 arrays (name from SVX_PA_*, or an arena offset key)."""
        idx = _PA[which] if which in _PA else _PA_OFFSETS + _OFF_KEYS.index(which)
        ptr, cnt = ctypes.POINTER(ctypes.c_int64)(), ctypes.c_int64()
        capi.check(capi.lib().svx_plan_array(self.handle, idx, ctypes.byref(ptr), ctypes.byref(cnt)), "svx_plan_array")
        if cnt.value == 0:
            return np.zeros(0, dtype=np.int64)
        return np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy()

    def bind(self, arena_ptr, stage_ptr, v0_ptrs, v1_ptrs):
        self._v0 = np.ascontiguousarray(v0_ptrs, dtype=np.uint64)
        self._v1 = np.ascontiguousarray(v1_ptrs, dtype=np.uint64)
        capi.check(capi.lib().svx_plan_bind(self.handle, int(arena_ptr), int(stage_ptr), capi.hptr(self._v0), capi.hptr(self._v1)),
                   "svx_plan_bind")
        info = np.zeros(1, dtype=capi.PLAN_INFO)
        capi.check(capi.lib().svx_plan_info(self.handle, capi.hptr(info)), "svx_plan_info")
        self.info = info[0]

    def set_sources(self, src0, src1):
        """svx_plan_set_sources: per-pair capi.ROW_SOURCE arrays (or None, None to clear); before bind()."""
        if src0 is None:
            capi.check(capi.lib().svx_plan_set_sources(self.handle, None, None), "svx_plan_set_sources")
            return
        self._src = (np.ascontiguousarray(src0, dtype=capi.ROW_SOURCE), np.ascontiguousarray(src1, dtype=capi.ROW_SOURCE))
        capi.check(capi.lib().svx_plan_set_sources(self.handle, capi.hptr(self._src[0]), capi.hptr(self._src[1])),
                   "svx_plan_set_sources")

    def draw(self, seeds=None):
        """The reference's RNG draws into the bound staging block: per-pair np.random.seed(seed) streams (C replay on
        the host cores this rank owns), or the global np.random stream continued and handed back."""
        L = capi.lib()
        if int(self.info["ndraw_calls"]) == 0:
            return
        if seeds is not None:
            sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.int64)[:int(self.info["npairs"])] & 0xFFFFFFFF, dtype=np.uint32)
            capi.check(L.svx_plan_draw_seeded(self.handle, capi.hptr(sd), host_threads()), "svx_plan_draw_seeded")
            return
        st = np.random.get_state()
        key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
        pos = np.array([st[2]], dtype=np.int32)
        capi.check(L.svx_plan_draw_stream(self.handle, capi.hptr(key), capi.hptr(pos)), "svx_plan_draw_stream")
        np.random.set_state((st[0], key, int(pos[0]), st[3], st[4]))

    def launcher_names(self):
        buf = ctypes.create_string_buffer(64)
        out = []
        for i in range(int(self.info["nlaunchers"])):
            capi.check(capi.lib().svx_plan_launcher_name(self.handle, i, buf, 64), "svx_plan_launcher_name")
            out.append(buf.value.decode())
        return out

    def __del__(self):
        try:
            if self.handle:
                capi.lib().svx_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def workspace_bytes(params, n0, n1):
    """(device arena bytes, host staging bytes) of a batch - svx_workspace_bytes."""
    n0 = np.ascontiguousarray(n0, dtype=np.int32)
    n1 = np.ascontiguousarray(n1, dtype=np.int32)
    a, h = np.zeros(1, np.int64), np.zeros(1, np.int64)
    capi.check(capi.lib().svx_workspace_bytes(capi.hptr(params), int(n0.shape[0]), capi.hptr(n0), capi.hptr(n1), capi.hptr(a), capi.hptr(h)),
               "svx_workspace_bytes")
    return int(a[0]), int(h[0])


class _SourceGather:
    """Row sources materialised by svx_gather_doc_embedding (BatchRun with sources on the unfused prologue)."""

    def __init__(self, sources, out0, out1, n0, n1, k0, k1, dim, device):
        P = len(n0)
        jobs = np.zeros(2 * P, dtype=capi.GATHER)
        for p in range(P):
            for side, (src, out, n, k) in enumerate(((sources[0][p], out0[p], n0[p], k0), (sources[1][p], out1[p], n1[p], k1))):
                j = jobs[2 * p + side]
                j["rows"], j["table"], j["nan_rows"], j["out"] = src["rows"], src["table"], src["nan_rows"], int(out)
                j["k"], j["n"], j["nrows"], j["is_fp16"] = k, n, src["nrows"], src["is_fp16"]
        self.jobs, self.dim, self.dev = jobs, int(dim), device
        self._dev = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).to(device)

    def run(self):
        capi.check(capi.lib().svx_gather_doc_embedding(self._dev.data_ptr(), capi.hptr(self.jobs), len(self.jobs), self.dim,
                                                       torch.cuda.current_stream(self.dev).cuda_stream), "svx_gather_doc_embedding")


class WidenJobs:
    """fp16 (K, N, D) device tensors -> fp32 working tensors with one launch of svx_gather_doc_embedding (identity
    table: a widening copy that also zeroes rows containing NaNs, as make_doc_embedding does,
    utils/embedding_utils.py:196-200).  The descriptors are uploaded once; run() can be repeated."""

    def __init__(self, srcs, dsts, dim, device):
        self.n, self.dim, self.dev = len(srcs), int(dim), device
        jobs = np.zeros(self.n, dtype=capi.GATHER)
        for j, (src, dst) in enumerate(zip(srcs, dsts)):
            jobs[j]["rows"], jobs[j]["out"] = src.data_ptr(), dst.data_ptr()
            jobs[j]["k"], jobs[j]["n"], jobs[j]["nrows"], jobs[j]["is_fp16"] = src.shape[0], src.shape[1], src.shape[0] * src.shape[1], 1
        self.jobs = jobs
        self._stage = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).pin_memory() if self.n else None
        self._dev = torch.empty(max(jobs.nbytes, 16) + 16, dtype=torch.uint8, device=device)
        self._keep = (srcs, dsts)
        if self.n:
            capi.check(capi.lib().svx_upload_pinned(self._dev.data_ptr(), self._stage.data_ptr(), self._stage.numel(),
                                                    torch.cuda.current_stream(device).cuda_stream), "svx_upload_pinned")

    def run(self):
        if self.n:
            capi.check(capi.lib().svx_gather_doc_embedding(self._dev.data_ptr(), capi.hptr(self.jobs), self.n, self.dim,
                                                           torch.cuda.current_stream(self.dev).cuda_stream), "svx_gather_doc_embedding")


def row_sources(rows, tables=None, nan_counts=None):
    """capi.ROW_SOURCE array for a list of device row matrices: (nrows, D) tensors with a (K, N) int32 table each, or
    (K, N, D) tensors with tables=None (identity: the tensor is the overlap tensor in its stored dtype).  fp16 or fp32.
    nan_counts: optional int32 device tensor, one counter per entry."""
    out = np.zeros(len(rows), dtype=capi.ROW_SOURCE)
    for j, t in enumerate(rows):
        if t.dtype not in (torch.float16, torch.float32) or not t.is_cuda or not t.is_contiguous():
            raise ValueError("row sources must be contiguous fp16 / fp32 CUDA tensors")
        out[j]["rows"] = t.data_ptr()
        out[j]["nrows"] = t.shape[0] if t.dim() == 2 else t.shape[0] * t.shape[1]
        out[j]["is_fp16"] = int(t.dtype == torch.float16)
        if tables is not None and tables[j] is not None:
            out[j]["table"] = tables[j].data_ptr()
        if nan_counts is not None:
            out[j]["nan_rows"] = nan_counts.data_ptr() + 4 * j
    return out


def fallback_del_penalty(frac):
    """dp_utils.py:315-321: with an empty side the knob is built from [0, .5, 1] on [0, 1] (host twin of
    svx_del_knob; what svx_plan_bind presets every level's penalty to)."""
    out = np.zeros(1, dtype=np.float64)
    samp = np.array([0.0, 0.5, 1.0], dtype=np.float32)
    capi.check(capi.lib().svx_host_del_knob(capi.hptr(samp), 3, float(frac), capi.hptr(out)), "svx_host_del_knob")
    return float(out[0])


class BatchRun:
    """One batch of document pairs on one GPU: a bound SvxPlan plus the torch tensors it lives in."""

    def __init__(self, vec_ptrs0, vec_ptrs1, n0, n1, k0, k1, dim, alignment_types, del_percentile_frac,
                 width_over2, max_size_full_dp, costs_sample_size, num_samps_for_norm, device,
                 cost_mode=capi.SVX_COST_EXACT, norms0=None, norms1=None, keep_dense_csum=False, seeds=None,
                 arena=None, fused_prologue=True, sources=None):
        """vec_ptrs0 / vec_ptrs1: device (K, N, D) fp32 tensors, normalised in place.  With sources=(src0, src1)
        (row_sources() arrays) they are pure outputs: the raw rows are read through the sources by the level-0 prologue
        (fp16 rows and row tables never become an fp32 tensor first); a plan that needs the unfused prologue
        materialises them with svx_gather_doc_embedding at the start of every run() instead."""
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
        self.frac = float(del_percentile_frac)
        self.cost_mode = cost_mode
        self.sample_size = int(costs_sample_size)
        self.keep_dense_csum = bool(keep_dense_csum)
        prm = make_params(k0, k1, dim, self.types, del_percentile_frac, width_over2, max_size_full_dp, costs_sample_size,
                          num_samps_for_norm, cost_mode, keep_dense_csum, not fused_prologue, norms0 is not None,
                          norms1 is not None)
        self.plan = plan = Plan(prm, n0, n1)
        info = plan.info
        self.w, self.band = int(info["width_over2"]), int(info["band"])
        self.per0, self.per1 = int(info["per0"]), int(info["per1"])
        self.R = int(info["nrecords"])
        self.nbytes, self.host_init_bytes = int(info["arena_bytes"]), int(info["host_bytes"])
        for name in ("first", "nlev", "depth", "rec_pair", "rec_level", "rs0", "rs1", "A", "T", "rec_cap", "nsamp", "top_rec",
                     "tgt_rec"):
            setattr(self, name, plan.array(name))
        self.banded = plan.array("banded").astype(bool)
        self.has_draw = plan.array("has_draw").astype(bool)
        self.off = {key: plan.array(key) for key in _OFF_KEYS}

        if arena is None:
            arena = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        elif arena.numel() < self.nbytes:
            raise capi.SvxError("arena of %d bytes given, the batch needs %d" % (arena.numel(), self.nbytes))
        self.arena = arena
        self.base = self.arena.data_ptr()
        assert self.base % 16 == 0
        # pinned staging (torch's caching host allocator recycles it): the upload below is then asynchronous,
        # so planning the next batch never waits for this batch's kernels
        self._stage = torch.empty(max(self.host_init_bytes, 16), dtype=torch.uint8, pin_memory=True)
        self._gather = None
        self._src_elt = 4               # bytes per element of the level-0 rows the prologue reads
        if sources is not None:
            if bool(info["fused_prologue"]):
                plan.set_sources(sources[0], sources[1])
                if len(sources[0]) and all(int(x) for x in sources[0]["is_fp16"]) and all(int(x) for x in sources[1]["is_fp16"]):
                    self._src_elt = 2
            else:
                self._gather = _SourceGather(sources, vec_ptrs0, vec_ptrs1, n0, n1, k0, k1, dim, device)
        plan.bind(self.base, self._stage.data_ptr(), vec_ptrs0, vec_ptrs1)
        plan.draw(seeds)
        self._names = plan.launcher_names()
        self.fused_prologue = bool(plan.info["fused_prologue"])
        self.upload()
        if norms0 is not None:
            self._put(self.off["norms0"][self.first[0]], np.ascontiguousarray(norms0, dtype=np.float32))
        if norms1 is not None:
            self._put(self.off["norms1"][self.first[0]], np.ascontiguousarray(norms1, dtype=np.float32))

    def upload(self):
        """Staging block -> arena on the current stream (descriptors, draws, default penalties; norms preset to 1.0,
        status cleared).  Needed again only when another batch has used the same arena in between."""
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        if self._init_copy is not None:
            capi.check(capi.lib().svx_plan_restore(self.plan.handle, self._init_copy.data_ptr(), stream), "svx_plan_restore")
        else:
            capi.check(capi.lib().svx_plan_upload(self.plan.handle, 1, stream), "svx_plan_upload")

    _init_copy = None

    def keep_init_on_device(self):
        """Keeps a device copy of the staging block: later upload() calls restore the arena's host-initialised prefix
        from it (several batches taking turns in one arena, bench.py's config-4 corpus) instead of crossing PCIe."""
        self._init_copy = torch.empty(max(self.host_init_bytes, 16), dtype=torch.uint8, device=self.dev)
        self._init_copy.copy_(self._stage[:self._init_copy.numel()], non_blocking=True)

    @property
    def knob(self):
        """knob[r] = (xi, yi) int32 views of the drawn knob sample ids in the staging block, or None (debug stack)."""
        stage = self._stage.numpy()
        out = [None] * self.R
        n = 4 * self.sample_size
        for r in np.nonzero(self.has_draw)[0]:
            out[r] = (stage[self.off["xi"][r]:self.off["xi"][r] + n].view(np.int32), stage[self.off["yi"][r]:self.off["yi"][r] + n].view(np.int32))
        return out

    # ------------------------------------------------------------------------------------------
    def _put(self, off, arr):
        t = torch.from_numpy(arr.view(np.uint8).ravel())
        self.arena[int(off):int(off) + t.numel()].copy_(t)

    def _get(self, off, nbytes, dtype):
        return self.arena[int(off):int(off) + int(nbytes)].cpu().numpy().view(dtype)

    def kernel_times(self):
        """Per-launcher device time (ms) of the last run(timing=True), summed by launcher name;
        CUDA events on the launching stream.  Synchronises."""
        torch.cuda.synchronize(self.dev)
        out = {}
        for name, e0, e1 in self._events or []:
            out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out

    _events = None
    _pair_range = None
    _streams = None

    def _enqueue_chain(self):
        """The whole path for the pairs in self._pair_range (all pairs if None) on the current stream: one C call
        (svx_plan_enqueue walks the launch chain), or one call per launcher with CUDA events around it."""
        L = capi.lib()
        lo, hi = self._pair_range if self._pair_range is not None else (0, self.P)
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        if self._events is None:
            capi.check(L.svx_plan_enqueue(self.plan.handle, -1, lo, hi, stream), "svx_plan_enqueue")
            return
        for i, name in enumerate(self._names):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.check(L.svx_plan_enqueue(self.plan.handle, i, lo, hi, stream), name)
            e1.record()
            self._events.append((name, e0, e1))

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
        if self._gather is not None:
            self._gather.run()
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
        pair = self.rec_pair
        has_next = self.rec_level < self.depth[pair]
        keep_all = self.keep_dense_csum
        sk0, sk1 = bool(self.plan.params["skip_norms0"][0]), bool(self.plan.params["skip_norms1"][0])
        want0 = (s1 > 0) & (self.per1 > 0) & (K1 > 0) & ~(l0 & sk0) & (s0 > 0)      # norms0 come from samples of side 1
        want1 = (s0 > 0) & (self.per0 > 0) & (K0 > 0) & ~(l0 & sk1) & (s1 > 0)
        lvl_bytes = 0
        for kk, nn, ko, per, want in ((K0, s0, K1, self.per1, want0), (K1, s1, K0, self.per0, want1)):
            rows = kk * nn * D * 4
            keep = np.where(l0 | keep_all, kk, min(1, kk))
            lvl_bytes += int((rows * np.where(l0, self._src_elt / 4.0, 2.0) + keep * nn * D * 4 + has_next * kk * (nn // 2) * D * 4 + kk * nn * 4 +
                              want * ko * per * D * 4 * 2).sum())
        norm_bytes = int((want0 * (K0 * s0 * (D * 4 + 4) + K1 * self.per1 * D * 4) +
                          want1 * (K1 * s1 * (D * 4 + 4) + K0 * self.per0 * D * 4)).sum())
        out = {
            "svx_level_prologue": lvl_bytes,
            "svx_normalize_rows": 2 * 4 * D * int(K0 * s0[l0].sum() + K1 * s1[l0].sum()),
            "svx_downsample": 4 * D * int((K0 * (s0[~top] + 3 * (s0[~top] // 2)) + K1 * (s1[~top] + 3 * (s1[~top] // 2))).sum()),
            "svx_sample_norms": norm_bytes,
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

    def start_fetch(self):
        """Queues the device -> host copies of the results on the current stream, into pinned buffers: the level-0
        alignment records, the record counts + status words, the penalties.  results() waits for them; calling this right
        after run() lets the copies of one batch overlap the kernels of the next."""
        info = self.plan.info
        lo, n = int(info["result_offset"]), int(info["result_bytes"])
        clo = int(info["counts_offset"])
        dp_lo = int(self.off["delpen"].min()) if self.R else 0
        spans = [(lo, n), (clo, self.nbytes - clo), (dp_lo, self.R * _ALIGN)]
        bufs = []
        for off, nb in spans:
            h = torch.empty(max(nb, 1), dtype=torch.uint8, pin_memory=True)
            if nb > 0:
                h[:nb].copy_(self.arena[off:off + nb], non_blocking=True)
            bufs.append(h)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self._fetch = (bufs, ev, lo, clo, dp_lo)

    _fetch = None

    def results(self):
        """Level-0 alignment records, scores, penalties and status of every pair (device -> host).
        Returns a list (one per pair) of dicts with 'recs' (structured array, forward order),
        'del_penalty' (per level), 'status'."""
        if self._fetch is None:
            self.start_fetch()
        (rb, cb, pb), ev, lo, clo, dp_lo = self._fetch
        self._fetch = None
        ev.synchronize()
        blob, cnt, dp_blob = rb.numpy(), cb.numpy(), pb.numpy()
        o = self.off
        out = []
        for p in range(self.P):
            r0 = int(self.first[p])
            cap = int(self.rec_cap[r0])
            n = int(cnt[o["nrecs"][r0] - clo:o["nrecs"][r0] - clo + 4].view(np.int32)[0])
            recs = blob[o["recs"][r0] - lo:o["recs"][r0] - lo + cap * capi.REC.itemsize].view(capi.REC)
            recs = recs[cap - min(n, cap):cap].copy()
            pens = [float(dp_blob[o["delpen"][r] - dp_lo:o["delpen"][r] - dp_lo + 8].view(np.float64)[0])
                    for r in range(r0, r0 + int(self.nlev[p]))]
            status = 0
            for r in range(r0, r0 + int(self.nlev[p])):
                st = cnt[o["status"][r] - clo:o["status"][r] - clo + 8].view(np.int32)
                status |= int(st[0]) | int(st[1])
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
        n = k * s * self.dim
        if n == 0:
            return np.zeros((k, s, self.dim), dtype=np.float32)
        if self.rec_level[r] == 0:
            raise ValueError("level-0 vectors live in the caller's tensors")
        off = self.off["vec0" if side == 0 else "vec1"][r]
        return self._get(off, n * 4, np.float32).reshape(k, s, self.dim).copy()


def records_to_alignments(recs):
    """SvxAlignRec array -> the reference's list of (list[int], list[int]) + float64 scores
    (dp_utils.py:126-143)."""
    al = [(list(range(int(r["x_end"] - r["nx"]), int(r["x_end"]))),
           list(range(int(r["y_end"] - r["ny"]), int(r["y_end"])))) for r in recs]
    return al, recs["score"].astype(np.float64).copy()
