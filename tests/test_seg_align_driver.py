"""The batch driver that replaces svecalign/seg_align/align.py: file
resolution and sharding on the CPU, and one real run over the shipped example on the GPU."""
import os
import shutil

import numpy as np
import pytest

from conftest import GOLDEN, same_alignments

NAME = "20180313-0900-PLENARY-15"


def _tree(tmp_path, with_ignore=True, break_one=False):
    ex = os.path.join(GOLDEN, "example")
    for lang in ("en", "de"):
        for sub, src, suffix in (("segments", f"{lang}.segments.txt", ".txt"), ("cat_segs", f"{lang}.cat_segs.txt", ".txt"),
                                 ("embeds", f"{lang}.embed", ".embed")):
            d = tmp_path / sub / lang
            d.mkdir(parents=True, exist_ok=True)
            shutil.copy(os.path.join(ex, src), d / f"{NAME}_{lang}{suffix}")
    if with_ignore:
        d = tmp_path / "ign" / "en-de"
        d.mkdir(parents=True)
        shutil.copy(os.path.join(ex, "ignore.src.txt"), d / f"{NAME}_en-{NAME}_de.src.txt")
        shutil.copy(os.path.join(ex, "ignore.tgt.txt"), d / f"{NAME}_en-{NAME}_de.tgt.txt")
    meta = tmp_path / "metadata.tsv"
    lines = [f"/audio/en/{NAME}_en.ogg\t/audio/de/{NAME}_de.ogg"]
    if break_one:
        lines.append("/audio/en/missing_en.ogg\t/audio/de/missing_de.ogg")
    meta.write_text("\n".join(lines) + "\n")
    argv = [str(meta), str(tmp_path / "out"), "--src_lang", "en", "--tgt_lang", "de", "--seg_dir", str(tmp_path / "segments"),
            "--concat_dir", str(tmp_path / "cat_segs"), "--embed_dir", str(tmp_path / "embeds"), "--is_stopes_embed", "-a", "6"]
    if with_ignore:
        argv += ["--ign_indices_dir", str(tmp_path / "ign")]
    return argv


def test_resolve_pairs_and_seeds(tmp_path):
    from speech_vecalign_b200 import seg_align
    argv = _tree(tmp_path, break_one=True)
    args = seg_align.build_parser().parse_args(argv)
    assert args.alignment_max_size == 6 and args.search_buffer_size == 5 and args.del_percentile_frac == 0.2
    assert args.max_size_full_dp == 300 and args.costs_sample_size == 20000 and args.num_samps_for_norm == 100
    out_dir, jobs = seg_align.resolve_pairs(open(args.metadata), args)
    assert out_dir.name == "en-de" and len(jobs) == 1                 # the pair with missing files is dropped
    j = jobs[0]
    assert j["out"].name == f"{NAME}_en-{NAME}_de.txt"
    assert j["src_emb"].name == f"{NAME}_en.embed" and j["tgt_cat"].name == f"{NAME}_de.txt"
    assert j["src_ign"] is not None and j["tgt_ign"] is not None
    assert seg_align.pair_seed(j, 0) == seg_align.pair_seed(j, 0) != seg_align.pair_seed(j, 1)
    v0, v1 = seg_align.load_pair(j, 5, args)
    assert v0.shape == (5, 237, 1024) and v1.shape == (5, 217, 1024) and v0.dtype == np.float32


@pytest.mark.gpu
def test_driver_reproduces_shipped_alignment(tmp_path):
    from speech_vecalign_b200 import seg_align
    from speech_vecalign_b200.vecalign import read_alignments
    argv = _tree(tmp_path)
    assert seg_align.main(argv) == 1
    out = tmp_path / "out" / "en-de" / f"{NAME}_en-{NAME}_de.txt"
    got = read_alignments(str(out))
    shipped = os.path.join(GOLDEN, "example", "shipped_alignment_a6.txt")
    assert same_alignments(got, read_alignments(shipped))
    scores = [float(ln.rsplit(":", 1)[1]) for ln in open(out)]
    ref = [float(ln.rsplit(":", 1)[1]) for ln in open(shipped)]
    assert np.max(np.abs(np.array(scores) - np.array(ref))) <= 0.05      # other RNG state
    assert seg_align.main(argv + ["--skip_existing"]) == 0
    # step 6.1 from the same records: equals the filter applied to the written step-5.4 file
    out.unlink()
    assert seg_align.main(argv + ["--max_cost", "0.7"]) == 1
    filt = tmp_path / "out_0.7" / "en-de" / out.name
    kept = seg_align.filter_by_cost(read_alignments(str(out)), [float(ln.rsplit(":", 1)[1]) for ln in open(out)], 0.7)
    assert filt.read_text().splitlines() == [f"{xs}:{ys}:{c}" for xs, ys, c in kept]
    assert 120 <= len(kept) <= 156


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float16, np.float32])
def test_device_gather_equals_make_doc_embedding(svb, dtype):
    """svx_gather_doc_embedding == make_doc_embedding (utils/embedding_utils.py:135-203) bit for bit:
    shipped example with its ignore indices, plus NaN rows and unknown keys."""
    from speech_vecalign_b200 import embedding_utils as eu
    from speech_vecalign_b200.vecalign import load_ignore_index_file
    ex = os.path.join(GOLDEN, "example")
    sent2id, rows = eu.read_in_embedding_rows(f"{ex}/en.cat_segs.txt", f"{ex}/en.embed", use_stopes=True)
    assert rows.dtype == np.float16
    rows = np.array(rows, dtype=dtype)
    rows[7, 100] = np.nan                                  # a NaN row must come out as zeros
    lines = open(f"{ex}/en.segments.txt").readlines()
    ign = load_ignore_index_file(f"{ex}/ignore.src.txt")
    sent2id.pop(next(iter(sent2id)))                       # an unknown key -> zero row
    for k in (1, 3, 5):
        ref = eu.make_doc_embedding(sent2id, rows.astype(np.float32), lines, k, ignore_indices=ign, overlap_segments=True)
        got = eu.make_doc_embedding_device(sent2id, rows, lines, k, ignore_indices=ign, overlap_segments=True)
        assert got.shape == ref.shape and np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.gpu
def test_driver_host_and_device_gather_agree(tmp_path):
    from speech_vecalign_b200 import seg_align
    argv = _tree(tmp_path)
    assert seg_align.main(argv) == 1
    out = tmp_path / "out" / "en-de" / f"{NAME}_en-{NAME}_de.txt"
    dev_txt = out.read_text()
    out.unlink()
    assert seg_align.main(argv + ["--host_gather"]) == 1
    assert out.read_text() == dev_txt


def test_filter_by_cost_known_answer():
    """Step 6.1 on the shipped step-5.4 file must give the shipped align_0.7 file, line for line
    (reference README: filter_by_cost --max_cost 0.7)."""
    from speech_vecalign_b200 import seg_align
    from speech_vecalign_b200.vecalign import read_alignments
    ex = os.path.join(GOLDEN, "example")
    src = os.path.join(ex, "shipped_alignment_a6.txt")
    al = read_alignments(src)
    scores = [float(ln.rsplit(":", 1)[1]) for ln in open(src)]
    kept = seg_align.filter_by_cost(al, scores, 0.7)
    want = open(os.path.join(ex, "shipped_align_0.7.txt")).read().splitlines()
    got = [f"{xs}:{ys}:{c}" for xs, ys, c in kept]
    assert got == want


def _synthetic_corpus(root, n_docs, k, rng):
    """A step-5.3 style tree for n_docs document pairs: segment lists '<start> <end>', the concatenation keys
    of every (start_i, end_{i+j}), j < k, in shuffled order, and raw fp16 .embed files (--fp16_embed)."""
    from speech_vecalign_b200 import synth
    meta = []
    for d in range(n_docs):
        n0 = int(rng.integers(3, 420))
        n1 = max(1, int(n0 * rng.uniform(0.7, 1.4)))
        vecs = synth.synth_pair(n0, n1, k, seed=900 + d)
        for lang, v in zip(("en", "de"), vecs):
            n = v.shape[1]
            bounds = np.cumsum(rng.integers(1600, 64000, size=n + 1))
            (root / "segments" / lang).mkdir(parents=True, exist_ok=True)
            (root / "cat_segs" / lang).mkdir(parents=True, exist_ok=True)
            (root / "embeds" / lang).mkdir(parents=True, exist_ok=True)
            (root / "segments" / lang / f"doc{d}_{lang}.txt").write_text(
                "".join(f"{bounds[i]} {bounds[i + 1]}\n" for i in range(n)))
            keys, rows = [], []
            for i in range(n):
                for j in range(k):
                    if i + j < n:
                        keys.append(f"{bounds[i]} {bounds[i + j + 1]}")
                        rows.append(v[j, i + j])
            order = rng.permutation(len(keys))
            (root / "cat_segs" / lang / f"doc{d}_{lang}.txt").write_text("".join(keys[o] + "\n" for o in order))
            np.stack([rows[o] for o in order]).astype(np.float16).tofile(root / "embeds" / lang / f"doc{d}_{lang}.embed")
        meta.append(f"/a/en/doc{d}_en.ogg\t/a/de/doc{d}_de.ogg")
    (root / "metadata.tsv").write_text("\n".join(meta) + "\n")
    return [str(root / "metadata.tsv"), str(root / "out"), "--src_lang", "en", "--tgt_lang", "de", "--seg_dir",
            str(root / "segments"), "--concat_dir", str(root / "cat_segs"), "--embed_dir", str(root / "embeds"),
            "--fp16_embed", "-a", str(k + 1), "--max_size_full_dp", "100", "--costs_sample_size", "4000"]


def test_native_row_tables_equal_the_python_contract(tmp_path, svb):
    """svx_host_overlap_tables (C, many documents per call, threads) == embedding_utils.overlap_row_table, the Python
    restatement of make_doc_embedding's key lookups (utils/embedding_utils.py:106-203): the shipped example with its
    ignore lists, and documents with unknown keys, blank lines and odd whitespace."""
    import ctypes
    from speech_vecalign_b200 import embedding_utils as eu, seg_align
    from speech_vecalign_b200.vecalign import load_ignore_index_file
    argv = _tree(tmp_path)
    args = seg_align.build_parser().parse_args(argv)
    _, jobs = seg_align.resolve_pairs(open(args.metadata), args)
    docs = []
    for side in ("src", "tgt"):
        docs.append({"seg": jobs[0][side + "_seg"], "cat": jobs[0][side + "_cat"], "ign": load_ignore_index_file(jobs[0][side + "_ign"])})
    odd = tmp_path / "odd"
    odd.mkdir()
    (odd / "seg.txt").write_text("0.0 1.5\n 1.5\t2.25 \r\n2.25 3.0 extra\n3.0 4.0\n4.0 5.5")
    (odd / "cat.txt").write_text("0.0 1.5\n0.0 2.25\n0.0 1.5\n1.5 3.0\n\n3.0 5.5\n2.25 3.0\n")
    docs.append({"seg": odd / "seg.txt", "cat": odd / "cat.txt", "ign": {(1, 2), (3, 3)}})
    k = 5
    lines = [open(d["seg"], "rt", encoding="utf-8").readlines() for d in docs]
    want = []
    for d, ln in zip(docs, lines):
        key_to_row = {}
        for i, line in enumerate(open(d["cat"], "rt", encoding="utf-8")):
            key_to_row.setdefault(line.strip(), i)
        want.append(eu.overlap_row_table(key_to_row, ln, k, ignore_indices=d["ign"], overlap_segments=True))
    nd = len(docs)
    nlines = np.array([len(ln) for ln in lines], dtype=np.int32)
    outs = [np.full((k, int(n)), -7, dtype=np.int32) for n in nlines]
    ign = [np.array(sorted(d["ign"]), dtype=np.int32).reshape(-1, 2) for d in docs]
    for nthreads in (1, 3):
        rc = svb.capi.lib().svx_host_overlap_tables(
            nd, (ctypes.c_char_p * nd)(*[str(d["seg"]).encode() for d in docs]), (ctypes.c_char_p * nd)(*[str(d["cat"]).encode() for d in docs]),
            (ctypes.c_void_p * nd)(*[a.ctypes.data for a in ign]), np.array([a.shape[0] for a in ign], dtype=np.int32).ctypes.data, k,
            (ctypes.c_void_p * nd)(*[o.ctypes.data for o in outs]), nlines.ctypes.data, None, nthreads)
        assert rc == 0, svb.capi.lib().svx_last_error_string()
        for got, ref in zip(outs, want):
            assert np.array_equal(got, ref)
    assert (want[2] >= 0).sum() == 3 and (want[0] >= 0).sum() > 1000


@pytest.mark.gpu
def test_driver_batches_and_shards_synthetic_corpus(tmp_path, oracle):
    """Nine synthetic document pairs through the driver: small --batch_gb (several GPU batches), two shards
    (--rank/--n_shard), --skip_existing; every output file equals the oracle run on the same host-built
    tensors with the pair's seed, whatever the batching or sharding."""
    import math
    from speech_vecalign_b200 import seg_align
    from speech_vecalign_b200.vecalign import read_alignments
    k = 3
    argv = _synthetic_corpus(tmp_path, 9, k, np.random.default_rng(5))
    assert seg_align.main(argv + ["--batch_gb", "0.004"]) == 9
    out_dir = tmp_path / "out" / "en-de"
    first = {p.name: p.read_text() for p in sorted(out_dir.iterdir())}
    assert len(first) == 9
    # the same corpus in two shards and one big batch: identical files
    for p in out_dir.iterdir():
        p.unlink()
    done = [seg_align.main(argv + ["--rank", str(r), "--n_shard", "2"]) for r in range(2)]
    assert sum(done) == 9 and min(done) >= 1
    assert {p.name: p.read_text() for p in sorted(out_dir.iterdir())} == first
    assert seg_align.main(argv + ["--skip_existing"]) == 0
    # against the oracle on host-gathered tensors, seeded like the driver seeds its pairs
    args = seg_align.build_parser().parse_args(argv)
    _, jobs = seg_align.resolve_pairs(open(args.metadata), args)
    for item in jobs[:4]:
        v0, v1 = seg_align.load_pair(item, k, args, on_device=False)
        np.random.seed(seg_align.pair_seed(item, 0))
        ref = oracle.vecalign(v0, v1, oracle.alignment_types(k + 1), 0.2, math.ceil(k / 2) + 5, 100, 4000, 100, fast_host=True)
        assert same_alignments(read_alignments(str(item["out"])), ref[0]["final_alignments"])
