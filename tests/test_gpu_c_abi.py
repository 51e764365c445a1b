"""The whole path through the C ABI alone (include/svx.h: svx_workspace_bytes + svx_align_batch): no Python planner,
no engine module - torch only provides device memory.  What a cgo / JNI / N-API host would do."""
import ctypes
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, same_alignments

pytestmark = pytest.mark.gpu


def _params(svb, k, dim, types, w, frac=0.2, full=300, css=20000, nsn=100):
    prm = np.zeros(1, dtype=svb.capi.PARAMS)
    prm["k0"] = prm["k1"] = k
    prm["dim"], prm["ntypes"] = dim, len(types)
    for t, (x, y) in enumerate(types):
        prm["xo"][0, t], prm["yo"][0, t] = x, y
    prm["del_percentile_frac"], prm["width_over2"], prm["max_size_full_dp"] = frac, w, full
    prm["costs_sample_size"], prm["num_samps_for_norm"] = css, nsn
    return prm


def _align_batch_c(svb, pairs, prm, seeds):
    """svx_workspace_bytes -> buffers -> svx_align_batch; returns per pair (records, nrecs, penalty, status)."""
    import torch
    L = svb.capi.lib()
    P = len(pairs)
    n0 = np.array([p[0].shape[1] for p in pairs], dtype=np.int32)
    n1 = np.array([p[1].shape[1] for p in pairs], dtype=np.int32)
    arena_b, stage_b = np.zeros(1, np.int64), np.zeros(1, np.int64)
    assert L.svx_workspace_bytes(prm.ctypes.data, P, n0.ctypes.data, n1.ctypes.data, arena_b.ctypes.data, stage_b.ctypes.data) == 0
    dev = torch.device("cuda", torch.cuda.current_device())
    d0 = [torch.from_numpy(p[0]).to(dev) for p in pairs]
    d1 = [torch.from_numpy(p[1]).to(dev) for p in pairs]
    arena = torch.empty(int(arena_b[0]), dtype=torch.uint8, device=dev)
    stage = torch.empty(int(stage_b[0]), dtype=torch.uint8, pin_memory=True)
    v0 = np.array([t.data_ptr() for t in d0], dtype=np.uint64)
    v1 = np.array([t.data_ptr() for t in d1], dtype=np.uint64)
    begin = np.concatenate([[0], np.cumsum(n0.astype(np.int64) + n1 + 2)]).astype(np.int64)
    recs = np.zeros(int(begin[-1]), dtype=svb.capi.REC)
    nrecs, status, pen = np.zeros(P, np.int32), np.zeros(P, np.int32), np.zeros(P, np.float64)
    sd = np.asarray(seeds, dtype=np.uint32)
    rc = L.svx_align_batch(prm.ctypes.data, P, n0.ctypes.data, n1.ctypes.data, v0.ctypes.data, v1.ctypes.data, sd.ctypes.data,
                           arena.data_ptr(), int(arena_b[0]), stage.data_ptr(), int(stage_b[0]), 1,
                           recs.ctypes.data, begin.ctypes.data, nrecs.ctypes.data, pen.ctypes.data, status.ctypes.data,
                           torch.cuda.current_stream(dev).cuda_stream)
    assert rc == 0, L.svx_last_error_string().decode()
    out = []
    for p in range(P):
        out.append((recs[begin[p]:begin[p] + nrecs[p]].copy(), int(nrecs[p]), float(pen[p]), int(status[p])))
    return out, (d0, d1)


def test_c_abi_aligns_the_shipped_example(svb):
    """BASELINE configs[0] and the shipped a=6 alignment through svx_align_batch; the reference ran with
    np.random.seed(0), i.e. pair seed 0 (tests/golden/make_golden.py)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import example_pair
    from speech_vecalign_b200.engine import records_to_alignments
    ref = json.load(open(os.path.join(GOLDEN, "example_reference.json")))
    for a in (4, 6):
        v0, v1 = example_pair(a)
        k = a - 1
        prm = _params(svb, k, v0.shape[2], svb.make_alignment_types(a), math.ceil(k / 2) + 5)
        out, _ = _align_batch_c(svb, [(v0, v1)], prm, [0])
        recs, n, pen, status = out[0]
        assert status == 0
        al, sc = records_to_alignments(recs)
        assert same_alignments(al, ref[f"a{a}"]["alignments"])
        assert np.max(np.abs(sc - np.array(ref[f"a{a}"]["scores"]))) <= 2e-5


def test_c_abi_batch_matches_oracle_and_normalises_in_place(svb, oracle):
    from speech_vecalign_b200 import synth
    from speech_vecalign_b200.engine import records_to_alignments
    a, k = 5, 4
    shapes = [(700, 650), (90, 80), (0, 5), (310, 305), (3, 2000)]
    pairs = [synth.synth_pair(n0, n1, k, dim=128, seed=40 + i) for i, (n0, n1) in enumerate(shapes)]
    seeds = [500 + i for i in range(len(pairs))]
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    prm = _params(svb, k, 128, args[0], args[2])
    out, (d0, d1) = _align_batch_c(svb, [(v0.copy(), v1.copy()) for v0, v1 in pairs], prm, seeds)
    for (v0, v1), s, (recs, n, pen, status), t0 in zip(pairs, seeds, out, d0):
        np.random.seed(s)
        r0 = v0.copy()
        ref = oracle.vecalign(r0, v1.copy(), *args, fast_host=True)
        al, sc = records_to_alignments(recs)
        assert status == 0 and same_alignments(al, ref[0]["final_alignments"])
        assert np.max(np.abs(sc - ref[0]["alignment_scores"]), initial=0) <= 1e-4
        assert abs(pen - ref[0]["del_penalty"]) <= 1e-6 * max(1.0, abs(ref[0]["del_penalty"]))
        assert np.array_equal(t0.cpu().numpy(), r0)            # dp_utils.py:396-397: inputs normalised in place


def test_c_abi_rejects_small_workspace_and_missing_seeds(svb):
    import torch
    from speech_vecalign_b200 import synth
    L = svb.capi.lib()
    v0, v1 = synth.synth_pair(50, 40, 2, dim=128, seed=1)
    prm = _params(svb, 2, 128, [(1, 1), (1, 2), (2, 1)], 6)
    n0, n1 = np.array([50], np.int32), np.array([40], np.int32)
    dev = torch.device("cuda", torch.cuda.current_device())
    t0, t1 = torch.from_numpy(v0).to(dev), torch.from_numpy(v1).to(dev)
    p0, p1 = np.array([t0.data_ptr()], np.uint64), np.array([t1.data_ptr()], np.uint64)
    arena = torch.empty(1024, dtype=torch.uint8, device=dev)
    stage = torch.empty(1024, dtype=torch.uint8, pin_memory=True)
    seeds = np.array([1], np.uint32)
    rc = L.svx_align_batch(prm.ctypes.data, 1, n0.ctypes.data, n1.ctypes.data, p0.ctypes.data, p1.ctypes.data, seeds.ctypes.data,
                           arena.data_ptr(), 1024, stage.data_ptr(), 1024, 1, None, None, None, None, None, None)
    assert rc == 2 and b"workspace too small" in L.svx_last_error_string()
    a_b, s_b = np.zeros(1, np.int64), np.zeros(1, np.int64)
    assert L.svx_workspace_bytes(prm.ctypes.data, 1, n0.ctypes.data, n1.ctypes.data, a_b.ctypes.data, s_b.ctypes.data) == 0
    arena = torch.empty(int(a_b[0]), dtype=torch.uint8, device=dev)
    stage = torch.empty(int(s_b[0]), dtype=torch.uint8, pin_memory=True)
    rc = L.svx_align_batch(prm.ctypes.data, 1, n0.ctypes.data, n1.ctypes.data, p0.ctypes.data, p1.ctypes.data, None,
                           arena.data_ptr(), int(a_b[0]), stage.data_ptr(), int(s_b[0]), 1, None, None, None, None, None, None)
    assert rc == 2 and b"seeds are required" in L.svx_last_error_string()
