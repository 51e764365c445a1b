#!/usr/bin/env python
"""tests/golden/make_golden.py — regenerates every fixture in this directory by running the REAL
reference (/root/reference Python driver + its own Cython core compiled into oracle/_ref by
`make -C oracle ref`).  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Fixtures
  example/                 the reference's shipped worked example (inputs of BASELINE config 1):
                           segment lists, concatenation keys, fp16 embedding rows, ignore ids and
                           the shipped golden alignment (156 lines, -a 6) — data files copied
                           verbatim from /root/reference/example/voxpopuli (no source code).
  example_reference.json   reference outputs on that pair: a=4 and a=6, np.random.seed(0).
  functions.npz            per-function known answers of dp_core.pyx (+ dp_utils helpers) on small
                           seeded inputs: every array the reference returned.
  e2e.json                 reference final alignments / scores / del_penalty for seeded synthetic
                           pairs (inputs are regenerated from the seed by speech_vecalign_b200.synth).
  margin/                  step 6.7 (postprocess/score_align.py) on the shipped example: the 347 + 347 vectors the
                           reference's two `Flat` faiss indexes hold (read out of Flat.populate.idx: 45-byte header,
                           then raw fp32 rows whose values are fp16-representable - stored here as fp16) and the
                           margin scores the reference shipped for them (align_0.7_clean_cat3_min1s_margin).
"""
import json
import math
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

du = ref_loader.ref_dp_utils()
rc = ref_loader.ref_core()
assert du is not None and rc is not None, "needs /root/reference and oracle/_ref (make -C oracle ref)"
from svecalign.utils.embedding_utils import make_doc_embedding  # noqa: E402  (reference)
from svecalign.vecalign.vecalign import make_alignment_types, load_ignore_index_file  # noqa: E402
from speech_vecalign_b200 import synth  # noqa: E402

EX = "/root/reference/example/voxpopuli"
NAME = "20180313-0900-PLENARY-15"

E2E_CASES = [  # (n0, n1, a, data seed, rng seed)
    (60, 70, 4, 11, 0), (237, 217, 4, 12, 1), (150, 160, 6, 13, 2), (310, 305, 5, 14, 3),
    (700, 650, 5, 15, 4), (489, 500, 6, 16, 5), (800, 817, 6, 17, 6), (620, 640, 8, 18, 7),
    (2000, 2000, 5, 19, 8),
]


def jsonable_alignments(al):
    return [[list(map(int, x)), list(map(int, y))] for x, y in al]


def example():
    dst = os.path.join(HERE, "example")
    os.makedirs(dst, exist_ok=True)
    for lang in ("en", "de"):
        shutil.copy(f"{EX}/segments/{lang}/{NAME}_{lang}.txt", f"{dst}/{lang}.segments.txt")
        shutil.copy(f"{EX}/cat_segs/{lang}/{NAME}_{lang}.txt", f"{dst}/{lang}.cat_segs.txt")
        shutil.copy(f"{EX}/embeds/{lang}/{NAME}_{lang}.embed", f"{dst}/{lang}.embed")
    shutil.copy(f"{EX}/untrans_cat_seg_ids/en-de/{NAME}_en-{NAME}_de.src.txt", f"{dst}/ignore.src.txt")
    shutil.copy(f"{EX}/untrans_cat_seg_ids/en-de/{NAME}_en-{NAME}_de.tgt.txt", f"{dst}/ignore.tgt.txt")
    shutil.copy(f"{EX}/alignments/en-de/{NAME}_en-{NAME}_de.txt", f"{dst}/shipped_alignment_a6.txt")
    shutil.copy(f"{EX}/align_0.7/en-de/{NAME}_en-{NAME}_de.txt", f"{dst}/shipped_align_0.7.txt")   # step 6.1 output (max_cost 0.7)
    shutil.copy(f"{EX}/{NAME}.gold", f"{dst}/human.gold")          # human alignment (README.md:289-296 known answers)

    def load(lang, ign, k):
        emb = np.load(f"{dst}/{lang}.embed").astype(np.float32)
        keys = {}
        for i, ln in enumerate(open(f"{dst}/{lang}.cat_segs.txt")):
            keys.setdefault(ln.strip(), i)
        lines = open(f"{dst}/{lang}.segments.txt").readlines()
        return make_doc_embedding(keys, emb, lines, k, ignore_indices=load_ignore_index_file(ign), overlap_segments=True)

    out = {}
    for a in (4, 6):
        k = a - 1
        v0 = load("en", f"{dst}/ignore.src.txt", k)
        v1 = load("de", f"{dst}/ignore.tgt.txt", k)
        chk = (float(np.abs(v0).sum()), float(np.abs(v1).sum()))   # before the in-place normalisation
        np.random.seed(0)
        st = du.vecalign(v0, v1, make_alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
        out[f"a{a}"] = {"alignments": jsonable_alignments(st[0]["final_alignments"]),
                        "scores": [float(s) for s in st[0]["alignment_scores"]],
                        "del_penalty": float(st[0]["del_penalty"]), "rng_seed": 0,
                        "vecs0_checksum": chk[0], "vecs1_checksum": chk[1]}
    json.dump(out, open(os.path.join(HERE, "example_reference.json"), "w"))


def functions():
    rng = np.random.default_rng(4711)
    d = 128
    out = {}
    k, s0, s1 = 3, 37, 41
    v0 = rng.standard_normal((k, s0, d)).astype(np.float32)
    v1 = rng.standard_normal((k, s1, d)).astype(np.float32)
    v0[1, 0] = 0
    v0[2, :2] = 0
    v1[1, 0] = 0
    v1[2, :2] = 0
    out["raw0"], out["raw1"] = v0.copy(), v1.copy()
    du.make_norm1(v0)
    du.make_norm1(v1)
    out["unit0"], out["unit1"] = v0.copy(), v1.copy()
    out["half0"] = du.downsample_vectors(v0)
    n0 = rng.uniform(0.7, 1.1, (k, s0)).astype(np.float32)
    n1 = rng.uniform(0.7, 1.1, (k, s1)).astype(np.float32)
    out["n0"], out["n1"] = n0, n1
    out["dense_costs"] = rc.make_dense_costs(v0, v1, n0, n1)
    out["pen"] = np.array([0.3141592653589793])
    cs, bp = rc.dense_dp(out["dense_costs"], float(out["pen"][0]))
    out["dense_csum"], out["dense_bp"] = cs, bp
    dal = du.dense_traceback(bp)
    xi = rng.integers(0, s0, 600).astype(np.int32)
    yi = rng.integers(0, s1, 600).astype(np.int32)
    sc = np.empty(600, np.float32)
    rc.score_path(xi, yi, n0[0], n1[0], v0[0], v1[0], sc)
    out["score_x"], out["score_y"], out["score_out"] = xi, yi, sc
    knob = du.DeletionKnob(sc, 0, max(sc))
    out["knob_fracs"] = np.array([0.0, 0.05, 0.2, 0.5, 0.9, 1.0])
    out["knob_pens"] = np.array([knob.percentile_frac_to_del_penalty(f) for f in out["knob_fracs"]])
    types = make_alignment_types(4)
    w = 7
    path = du.alignment_to_search_path(dal)
    out["path_same"] = np.array(path, dtype=np.int32)
    feats, boff = rc.make_sparse_costs(v0, v1, n0, n1, path, types, w)
    out["sparse_costs"], out["b_offset"] = feats, boff
    csum, xp, yp, nbo = rc.sparse_dp(feats, boff, types, 0.2718281828, s0, s1)
    out["sparse_pen"] = np.array([0.2718281828])
    out["sparse_csum"], out["sparse_xp"], out["sparse_yp"], out["new_b_offset"] = csum, xp, yp, nbo
    al, scores = du.sparse_traceback(csum, xp, yp, nbo, s0, s1)
    out["trace_x_end"] = np.array([(x[-1] + 1) if len(x) else -1 for x, _ in al], dtype=np.int32)
    out["trace_nx"] = np.array([len(x) for x, _ in al], dtype=np.int32)
    out["trace_ny"] = np.array([len(y) for _, y in al], dtype=np.int32)
    out["trace_scores"] = scores
    for t0, t1, tag in [(2 * s0, 2 * s1, "even"), (2 * s0 + 1, 2 * s1 + 1, "odd")]:
        up = du.upsample_alignment([(list(x), list(y)) for x, y in al])
        du.extend_alignments(up, t0, t1)
        out[f"path_up_{tag}"] = np.array(du.alignment_to_search_path(up), dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "functions.npz"), **out)


def e2e():
    cases = []
    for n0, n1, a, seed, rseed in E2E_CASES:
        k = a - 1
        v0, v1 = synth.synth_pair(n0, n1, k, seed=seed)
        chk = [float(np.abs(v0).sum()), float(np.abs(v1).sum())]      # before the in-place normalisation
        np.random.seed(rseed)
        st = du.vecalign(v0, v1, make_alignment_types(a), 0.2, math.ceil(k / 2) + 5, 300, 20000, 100)
        cases.append({"n0": n0, "n1": n1, "a": a, "seed": seed, "rng_seed": rseed,
                      "input_checksum": chk,
                      "alignments": jsonable_alignments(st[0]["final_alignments"]),
                      "scores": [float(s) for s in st[0]["alignment_scores"]],
                      "del_penalty": [float(st[d]["del_penalty"]) for d in sorted(st)]})
        print("e2e", n0, n1, a, len(st[0]["final_alignments"]))
    json.dump({"numpy": np.__version__, "cases": cases}, open(os.path.join(HERE, "e2e.json"), "w"))


def margin():
    dst = os.path.join(HERE, "margin")
    os.makedirs(dst, exist_ok=True)
    for lang in ("en", "de"):
        raw = open(f"{EX}/align_0.7_clean_cat3_min1s_embed_indexes/en-de/{lang}/Flat.populate.idx", "rb").read()
        n = (len(raw) - 45) // (1024 * 4)
        vec = np.frombuffer(raw[45:45 + n * 1024 * 4], dtype=np.float32).reshape(n, 1024)
        assert n == 347 and len(raw) == 45 + n * 4096 and np.array_equal(vec.astype(np.float16).astype(np.float32), vec)
        np.save(f"{dst}/{lang}.index_vectors.f16.npy", vec.astype(np.float16))
    shutil.copy(f"{EX}/align_0.7_clean_cat3_min1s_margin/en-de/{NAME}_en-{NAME}_de.txt", f"{dst}/shipped_margin.txt")
    print("margin: 2 x 347 index vectors + shipped scores")


if __name__ == "__main__":
    example()
    margin()
    functions()
    e2e()
    print("golden fixtures written to", HERE)
