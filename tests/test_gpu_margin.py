"""Margin scoring on the GPU (svx_margin_scores: tcgen05 fp16 GEMM + fused top-k) against the numpy oracle of
svecalign/postprocess/score_align.py:124-161 and against the scores the reference shipped for its example.
Tolerances: kernel vs oracle on the same fp16-stored vectors 3e-5 (fp32 accumulation order of a 1024-term dot
product, scores ~1.2); vs the shipped scores 2.5e-4 (the reference's own fp16 faiss-gpu arithmetic)."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _fixture():
    g = os.path.join(GOLDEN, "margin")
    x = np.load(os.path.join(g, "en.index_vectors.f16.npy"))
    y = np.load(os.path.join(g, "de.index_vectors.f16.npy"))
    shipped = np.array([float(ln.rsplit(":", 1)[1]) for ln in open(os.path.join(g, "shipped_margin.txt"))])
    return x, y, shipped


@pytest.mark.parametrize("margin", ["ratio", "distance"])
def test_shipped_example_scores(svb, margin):
    from oracle import margin_oracle as mo
    from speech_vecalign_b200 import score_align
    x16, y16, shipped = _fixture()
    for x, y in ((x16.astype(np.float32), y16.astype(np.float32)), (x16, y16)):          # fp32 and fp16 rows in
        got = score_align.compute_sim(x, y, 16, margin)
        ref = mo.margin_scores(x.astype(np.float32), y.astype(np.float32), 16, margin)
        assert got.dtype == np.float32 and got.shape == (347,)
        assert np.max(np.abs(got - ref)) <= 3e-5
        if margin == "ratio":
            assert np.max(np.abs(got - shipped)) <= 2.5e-4


@pytest.mark.parametrize("n,nxb,nyb,dim,k,seed", [(16, 0, 0, 128, 1, 1), (129, 300, 257, 1024, 16, 2), (1000, 0, 2500, 256, 4, 3),
                                                (2500, 700, 0, 1024, 16, 4), (5, 4000, 16, 64, 16, 5)])
def test_random_collections(svb, n, nxb, nyb, dim, k, seed):
    """Pairs scored against collections of other sizes (one file against the corpus-wide indexes, score_align.py:233-253),
    sizes that are not multiples of the 128 x 256 tiles, small k, small dimensions."""
    from oracle import margin_oracle as mo
    from speech_vecalign_b200 import score_align
    rng = np.random.default_rng(seed)
    cen = rng.standard_normal((40, dim)).astype(np.float32)

    def draw(m):            # clustered, so that neighbourhoods are not all alike
        return (cen[rng.integers(0, 40, m)] + 0.7 * rng.standard_normal((m, dim))).astype(np.float32) * rng.uniform(0.5, 3.0, (m, 1)).astype(np.float32)

    x = draw(n)
    y = (x + 0.5 * rng.standard_normal((n, dim))).astype(np.float32)
    xb = draw(nxb) if nxb else None
    yb = draw(nyb) if nyb else None
    got = score_align.compute_sim(x.copy(), y.copy(), k, "ratio", x_base=xb, y_base=yb)
    ref = mo.margin_scores(x, y, k, "ratio", x_base=xb, y_base=yb)
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) <= 3e-5


def test_reference_signature_flat_index_files_and_in_place(svb, tmp_path):
    from oracle import margin_oracle as mo
    from speech_vecalign_b200 import score_align
    x16, y16, shipped = _fixture()
    paths = []
    for name, v in (("en", x16), ("de", y16)):               # a faiss IndexFlat file: 45-byte header + raw fp32 rows
        v32 = v.astype(np.float32)
        p = tmp_path / f"{name}.Flat.populate.idx"
        with open(p, "wb") as f:
            f.write(b"IxF2" + struct.pack("<iqqq", v32.shape[1], v32.shape[0], 1 << 20, 1 << 20) + b"\x01" + struct.pack("<iq", 1, v32.size))
            f.write(v32.tobytes())
        paths.append(p)
    idx_x, idx_y = score_align.load_flat_index(paths[0]), score_align.load_flat_index(paths[1])
    assert idx_x.ntotal == 347 and np.array_equal(idx_x.vectors, x16.astype(np.float32))
    x = 3.0 * x16.astype(np.float32)
    y = y16.astype(np.float32)
    got = score_align.compute_sim_with_nonflat_idx(idx_x, idx_y, x, y, 16, "ratio")
    assert np.max(np.abs(got - shipped)) <= 2.5e-4
    assert np.allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-5)        # faiss.normalize_L2(x) of the reference (:136)
    with pytest.raises(ValueError, match="Wrong margin type"):
        score_align.compute_sim(x, y, 16, "cosine")
    with pytest.raises(ValueError, match="not a flat faiss index"):
        bad = tmp_path / "ivf.idx"
        bad.write_bytes(b"IwFl" + b"\0" * 64)
        score_align.load_flat_index(bad)


def test_driver_writes_the_reference_file_format(svb, tmp_path):
    """python -m speech_vecalign_b200.score_align on a small tree laid out like example/voxpopuli (tsvs pointing into
    .embed files, concatenated alignments, Flat indexes): same file names and line format as score_align.py:96-115."""
    from oracle import margin_oracle as mo
    from speech_vecalign_b200 import score_align
    x16, y16, shipped = _fixture()
    n = 120
    lang = tmp_path / "emb" / "en-de"
    lang.mkdir(parents=True)
    al = tmp_path / "align" / "en-de"
    al.mkdir(parents=True)
    meta = tmp_path / "metadata.tsv"
    names = []
    with open(meta, "w") as mf:
        for f_i, (lo, hi) in enumerate(((0, 70), (70, n))):
            s, t = f"a{f_i}_en", f"a{f_i}_de"
            names.append(f"{s}-{t}")
            mf.write(f"/audio/{s}.wav\t/audio/{t}.wav\n")
            for side, v in (("src", x16), ("tgt", y16)):
                emb = lang / f"{s}-{t}.{side}.embed"
                np.save(str(emb) + ".npy", v[lo:hi])            # stopes .npy framing, fp16
                os.replace(str(emb) + ".npy", emb)
                with open(lang / f"{s}-{t}.{side}.tsv", "w") as tf:
                    for r in range(hi - lo):
                        tf.write(f"{emb}\t{r}\n")
            with open(al / f"{s}-{t}.txt", "w") as af:
                for r in range(hi - lo):
                    af.write(f"[{r}]:[{r}, {r + 1}]\n")
    out = tmp_path / "margin"
    scores = score_align.main([str(meta), str(out), "--embed_dir", str(tmp_path / "emb"), "--align_dir", str(tmp_path / "align"),
                               "--src_lang", "en", "--tgt_lang", "de", "--embed_fp16", "--embed_stopes"])
    ref = mo.margin_scores(x16[:n].astype(np.float32), y16[:n].astype(np.float32), 16, "ratio")
    assert np.max(np.abs(scores - ref)) <= 3e-5
    lines = [ln.strip() for nm in names for ln in open(out / "en-de" / f"{nm}.txt")]
    assert len(lines) == n and lines[0].startswith("[0]:[0, 1]:") and float(lines[0].rsplit(":", 1)[1]) == float(scores[0])
