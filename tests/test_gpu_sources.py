"""Row sources of the level-0 prologue (include/svx.h SvxRowSource, svx_plan_set_sources): the raw rows are read THROUGH
the source - fp16 / fp32 row matrix + optional (K, N) row table, zeros for missing rows and for rows holding a NaN, i.e.
what make_doc_embedding (utils/embedding_utils.py:135-203) builds on the host - instead of from a materialised fp32
(K, N, D) tensor.  Every result must be bit-identical to gathering first (svx_gather_doc_embedding) and running the plain
path: normalised rows, norms of every level, records, penalties."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stack_equal(a, b):
    assert sorted(a) == sorted(b)
    for lvl in a:
        for key in a[lvl]:
            x, y = a[lvl][key], b[lvl][key]
            if isinstance(x, np.ndarray):
                assert x.shape == y.shape and x.dtype == y.dtype, (lvl, key)
                assert np.array_equal(x, y, equal_nan=True), (lvl, key, float(np.nanmax(np.abs(x.astype(np.float64) - y))))
            elif isinstance(x, (list, tuple)):
                assert list(map(repr, x)) == list(map(repr, y)), (lvl, key)
            else:
                assert x == y or (x != x and y != y), (lvl, key, x, y)


@pytest.mark.parametrize("n0,n1,a,dim", [(237, 217, 4, 1024), (401, 388, 6, 1024), (1, 3, 3, 128), (2, 2, 2, 256),
                                         (63, 700, 5, 512), (900, 31, 8, 128)])
def test_fp16_inputs_through_sources_equal_widened_inputs(svb, oracle, n0, n1, a, dim):
    """vecalign with fp16 (K, N, D) inputs (identity sources, the on-disk dtype of .embed files) against the same
    values passed as fp32: the whole debug stack - vectors, norms, costs, paths, alignments, scores of every level."""
    from speech_vecalign_b200 import synth
    k = a - 1
    v0, v1 = synth.synth_pair(n0, n1, k, dim=dim, seed=31 + n0)
    h0, h1 = v0.astype(np.float16), v1.astype(np.float16)
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    np.random.seed(5)
    ref = svb.dp_utils.vecalign(h0.astype(np.float32), h1.astype(np.float32), *args, debug=True)
    np.random.seed(5)
    got = svb.dp_utils.vecalign(h0, h1, *args, debug=True)
    _stack_equal(got, ref)


def test_fp16_rows_holding_nans_are_zeroed_like_the_loader(svb, oracle):
    """embedding_utils.py:196-200 resets an overlap row that holds a NaN to zeros; identity sources do the same."""
    from speech_vecalign_b200 import synth
    a, k = 5, 4
    v0, v1 = synth.synth_pair(310, 295, k, seed=77)
    h0, h1 = v0.astype(np.float16), v1.astype(np.float16)
    h0[0, 17, 5] = np.nan
    h0[3, 200, 1023] = np.nan
    h1[1, 0, 0] = np.nan
    h1[2, 294, 511] = np.nan
    z0, z1 = h0.astype(np.float32), h1.astype(np.float32)
    for z in (z0, z1):
        bad = np.isnan(z).any(axis=2)
        z[bad] = 0.0
    args = (oracle.alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
    np.random.seed(9)
    ref = svb.dp_utils.vecalign(z0, z1, *args, debug=True)
    np.random.seed(9)
    got = svb.dp_utils.vecalign(h0, h1, *args, debug=True)
    _stack_equal(got, ref)


def _tables(rng, k, n, nrows):
    """a (K, N) row table like the one the step-5.4 driver builds: overlap j of position e = some row of the file,
    -1 where the concatenation does not exist; a few entries past the end of the file (treated as missing)."""
    t = rng.integers(0, nrows, size=(k, n)).astype(np.int32)
    for j in range(1, k):
        t[j, :j] = -1
    t[rng.random((k, n)) < 0.03] = -1
    if n > 4:
        t[0, n // 2] = nrows + 3
        t[0, 1] = t[2, n - 1] = 5                # the file row that holds a NaN in the test below
    return t


@pytest.mark.parametrize("is_fp16,nsn", [(True, 100), (False, 100), (True, 3000)])
def test_table_sources_equal_gather_then_align(svb, oracle, is_fp16, nsn):
    """BatchRun with (rows, table) sources against svx_gather_doc_embedding + the plain BatchRun, three pairs of
    different lengths: records, penalties, the normalised level-0 rows left in the output tensors, the per-document NaN
    counts.  nsn = 3000 draws more norm samples than the fused prologue takes: the same call then materialises the rows
    itself (unfused plan)."""
    import torch
    from speech_vecalign_b200 import engine
    dev = torch.device("cuda", torch.cuda.current_device())
    rng = np.random.default_rng(4)
    a, k, dim = 6, 5, 1024
    sizes = [(333, 301), (58, 77), (512, 480)]
    types = oracle.alignment_types(a)
    rows, tabs = [[], []], [[], []]
    for n0, n1 in sizes:
        for side, n in enumerate((n0, n1)):
            nrows = k * n + 7
            r = rng.standard_normal((nrows, dim)).astype(np.float32)
            r[5, 100] = np.nan                       # a file row with a NaN: every (overlap, position) that uses it is zeroed
            r = r.astype(np.float16) if is_fp16 else r
            rows[side].append(torch.from_numpy(r).to(dev))
            tabs[side].append(torch.from_numpy(_tables(rng, k, n, nrows)).to(dev))
    P = len(sizes)
    n0s, n1s = [s[0] for s in sizes], [s[1] for s in sizes]
    w = math.ceil(k / 2) + 5

    def run(sources):
        out0 = [torch.empty((k, n, dim), dtype=torch.float32, device=dev) for n in n0s]
        out1 = [torch.empty((k, n, dim), dtype=torch.float32, device=dev) for n in n1s]
        nan0 = torch.zeros(P, dtype=torch.int32, device=dev)
        nan1 = torch.zeros(P, dtype=torch.int32, device=dev)
        src = None
        if sources:
            src = (engine.row_sources(rows[0], tabs[0], nan0), engine.row_sources(rows[1], tabs[1], nan1))
        else:
            jobs = np.zeros(2 * P, dtype=svb.capi.GATHER)
            for p in range(P):
                for side, (o, n, cnt) in enumerate(((out0[p], n0s[p], nan0), (out1[p], n1s[p], nan1))):
                    j = jobs[2 * p + side]
                    j["rows"], j["table"], j["out"] = rows[side][p].data_ptr(), tabs[side][p].data_ptr(), o.data_ptr()
                    j["nan_rows"] = cnt.data_ptr() + 4 * p
                    j["k"], j["n"], j["nrows"], j["is_fp16"] = k, n, rows[side][p].shape[0], int(is_fp16)
            jd = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).to(dev)
            svb.capi.check(svb.capi.lib().svx_gather_doc_embedding(jd.data_ptr(), svb.capi.hptr(jobs), 2 * P, dim,
                                                                  torch.cuda.current_stream(dev).cuda_stream), "gather")
        br = engine.BatchRun([t.data_ptr() for t in out0], [t.data_ptr() for t in out1], n0s, n1s, k, k, dim, types, 0.2, w,
                             300, 20000, nsn, dev, seeds=[11, 12, 13], sources=src)
        assert br.fused_prologue == (nsn <= 2000)
        br.run()
        res = br.results()
        torch.cuda.synchronize()
        return res, [t.cpu().numpy() for t in out0], [t.cpu().numpy() for t in out1], nan0.cpu().numpy(), nan1.cpu().numpy()

    ref = run(False)
    got = run(True)
    for p in range(P):
        assert got[0][p]["status"] == 0 and ref[0][p]["status"] == 0
        assert np.array_equal(got[0][p]["recs"], ref[0][p]["recs"])
        assert list(got[0][p]["del_penalty"]) == list(ref[0][p]["del_penalty"])
        assert np.array_equal(got[1][p], ref[1][p]) and np.array_equal(got[2][p], ref[2][p])
    assert ref[3].sum() > 0 and ref[4].sum() > 0
    assert np.array_equal(got[3], ref[3]) and np.array_equal(got[4], ref[4])
