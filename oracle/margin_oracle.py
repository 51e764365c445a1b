"""oracle/margin_oracle.py — TEST INFRASTRUCTURE (checker only; never imported by the product).

numpy restatement of svecalign/postprocess/score_align.py:118-161 (inplace_l2_to_cosine,
compute_sim_with_nonflat_idx) with the index search spelled out as an EXACT flat search.

Third-party dependency: the reference delegates the k-nearest-neighbour search to faiss (faiss-gpu through
stopes' load_index; neither is vendored in /root/reference nor installed here; README.md:121 pins no version).
Its published algorithm for a `Flat` index - the type prep_index.py's determine_faiss_index_type picks for small
collections and the one the shipped example holds (example/voxpopuli/..._embed_indexes/en-de/*/Flat.populate.idx)
- is exhaustive search by squared L2 distance |q|^2 + |b|^2 - 2 q.b, results sorted ascending; with
`--gpu_type fp16-shard` (score_align.py:48-50) the vectors are stored in fp16.  For larger collections the reference
trains IVF/PQ indexes, whose results are approximations of this search.

Parity pin: the reference's own known answer - the 347 margin scores it shipped for the example, reproduced from
the vectors its two indexes hold (tests/golden/margin, tests/test_oracle_golden.py): max |diff| 1.2e-4 (the
reference's fp16 GPU arithmetic), bar 2e-4.
"""
import numpy as np


def normalize_L2(x):
    """faiss.normalize_L2 (score_align.py:136-137): in place, fp32, rows of zero norm are left alone."""
    nr = np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float32)).astype(np.float32)
    ok = nr > 0
    x[ok] *= (np.float32(1.0) / nr[ok])[:, None]
    return x


def flat_search(base, queries, k):
    """IndexFlatL2.search: the k smallest squared L2 distances of every query, ascending (float64 arithmetic on the
    stored values)."""
    b = base.astype(np.float64)
    q = queries.astype(np.float64)
    out = np.empty((q.shape[0], k), dtype=np.float64)
    bn = np.einsum("ij,ij->i", b, b)
    for lo in range(0, q.shape[0], 1024):
        blk = q[lo:lo + 1024]
        d2 = np.einsum("ij,ij->i", blk, blk)[:, None] + bn[None, :] - 2.0 * (blk @ b.T)
        out[lo:lo + 1024] = np.sort(np.partition(d2, k - 1, axis=1)[:, :k], axis=1)
    return out


def l2_to_cosine(x):
    """score_align.py:118-121 inplace_l2_to_cosine: cosine = (2 - L2^2) / 2"""
    return (2.0 - x) / 2.0


def margin_scores(x, y, k=16, margin="ratio", x_base=None, y_base=None, index_dtype=np.float16):
    """score_align.py:124-161.  x, y: (n, d) rows of the aligned pairs (copied, then normalised as the reference
    normalises its inputs in place); x_base / y_base: the collections the indexes were populated with (default: the
    pairs themselves); index_dtype: storage type of the flat index (queries are converted alike by faiss-gpu)."""
    x = normalize_L2(np.array(x, dtype=np.float32))
    y = normalize_L2(np.array(y, dtype=np.float32))
    assert x.shape == y.shape, f"{x.shape} {y.shape}"
    xb = x if x_base is None else normalize_L2(np.array(x_base, dtype=np.float32))
    yb = y if y_base is None else normalize_L2(np.array(y_base, dtype=np.float32))
    st = lambda v: v.astype(index_dtype).astype(np.float32)
    avg_xy = flat_search(st(yb), st(x), k).mean(axis=1)
    avg_yx = flat_search(st(xb), st(y), k).mean(axis=1)
    cxy, cyx = l2_to_cosine(avg_xy), l2_to_cosine(avg_yx)
    a = np.einsum("ij,ij->i", x.astype(np.float64), y.astype(np.float64))
    b = (cxy + cyx) / 2
    if margin == "ratio":
        return (a / b).astype(np.float32)
    if margin == "distance":
        return (a - b).astype(np.float32)
    raise ValueError(f"Wrong margin type: {margin}")
