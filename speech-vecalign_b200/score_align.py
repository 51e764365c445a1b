"""Step 6.7, margin scoring of aligned segment pairs — drop-in for ``svecalign.postprocess.score_align``
(reference: svecalign/postprocess/score_align.py; driven by example/voxpopuli/run.sh:162-169).

The reference searches faiss indexes for the k nearest neighbours of every aligned segment embedding in the other
language's collection and divides the pair's cosine by the mean neighbourhood cosine (ratio margin,
https://aclanthology.org/P19-1309).  Here the search is an exact flat search on the GPU's tensor cores
(``svx_margin_scores``: tcgen05 fp16 GEMM fed by TMA with the top-k fused into the epilogue); ``Flat`` index files -
what prep_index.py builds for collections the size of the shipped example - are read directly, and a trained
(IVF/PQ) index, whose answers approximate this search, is replaced by the exact answer over the same vectors.

    compute_sim_with_nonflat_idx(idx_x, idx_y, x, y, k, margin)   reference signature (:124-129); idx_* = FlatIndex
    compute_sim(x, y, k, margin, x_base=None, y_base=None)        the same without index objects
    python -m speech_vecalign_b200.score_align METADATA OUT_DIR --embed_dir ... --align_dir ... --index_dir ...

All arithmetic runs in libsvx.so on the GPU; there is no CPU fallback.
"""
import argparse
import logging
import struct
from collections import defaultdict
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import capi
from .embedding_utils import load_sent_embeddings
from .vecalign import read_alignments

logger = logging.getLogger(__name__)
_MARGINS = {"ratio": 0, "distance": 1}


class FlatIndex:
    """The vectors of a faiss ``Flat`` index (what idx.search() of the reference scans exhaustively)."""

    def __init__(self, vectors):
        self.vectors = vectors

    @property
    def ntotal(self):
        return int(self.vectors.shape[0])


def load_flat_index(path) -> FlatIndex:
    """Reads a faiss IndexFlat file as written by prep_index.py:153-183 for small collections
    (``Flat.populate.idx``): fourcc 'IxF2' / 'IxFI' / 'IxFl', d, ntotal, two dummies, is_trained, metric type,
    vector-size, then ntotal * d fp32 values."""
    raw = Path(path).read_bytes()
    four = raw[:4]
    if four not in (b"IxF2", b"IxFI", b"IxFl"):
        raise ValueError(f"{path}: not a flat faiss index (fourcc {four!r}); trained IVF/PQ indexes approximate the exact "
                         "search - pass the vectors they were populated with instead")
    d, ntotal = struct.unpack_from("<iq", raw, 4)
    nvals = struct.unpack_from("<q", raw, 37)[0]
    if nvals != d * ntotal or len(raw) != 45 + 4 * nvals:
        raise ValueError(f"{path}: unexpected flat index layout (d={d}, ntotal={ntotal}, {len(raw)} bytes)")
    return FlatIndex(np.frombuffer(raw, dtype=np.float32, offset=45, count=nvals).reshape(ntotal, d))


def inplace_l2_to_cosine(x: np.ndarray):
    """score_align.py:118-121 (kept for callers that post-process distances themselves)."""
    np.negative(x, out=x)
    np.add(x, 2, out=x)
    np.divide(x, 2.0, out=x)


def _dev(v, dev):
    if isinstance(v, torch.Tensor):
        t = v
    else:
        a = np.asarray(v)
        if a.dtype not in (np.float32, np.float16):
            a = a.astype(np.float32)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype not in (torch.float32, torch.float16):
        t = t.float()
    return t.to(dev, non_blocking=True).contiguous()


def compute_sim(x, y, k: int = 16, margin: str = "ratio", x_base=None, y_base=None, as_numpy=True):
    """Margin scores of the pairs (x[i], y[i]); x_base / y_base are the collections searched for neighbours (default:
    x and y themselves).  numpy or torch inputs, fp32 or fp16, (n, d) with d a multiple of 64."""
    if margin not in _MARGINS:
        raise ValueError(f"Wrong margin type: {margin}")           # score_align.py:159
    if not torch.cuda.is_available():
        raise capi.SvxError("no CUDA device: speech_vecalign_b200 has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    xd, yd = _dev(x, dev), _dev(y, dev)
    num_x, dim_x = xd.shape
    num_y, dim_y = yd.shape
    assert num_x == num_y and dim_x == dim_y, f"{tuple(xd.shape)} {tuple(yd.shape)}"      # score_align.py:134
    if xd.dtype != yd.dtype:
        xd, yd = xd.float(), yd.float()
    xb = None if x_base is None else _dev(x_base, dev).to(xd.dtype)
    yb = None if y_base is None else _dev(y_base, dev).to(xd.dtype)
    L = capi.lib()
    need = np.zeros(1, dtype=np.int64)
    capi.check(L.svx_margin_workspace_bytes(num_x, 0 if xb is None else xb.shape[0], 0 if yb is None else yb.shape[0], dim_x,
                                            capi.hptr(need)), "svx_margin_workspace_bytes")
    work = torch.empty(int(need[0]) + 256, dtype=torch.uint8, device=dev)
    wptr = (work.data_ptr() + 255) // 256 * 256
    scores = torch.empty(num_x, dtype=torch.float32, device=dev)
    capi.check(L.svx_margin_scores(xd.data_ptr(), yd.data_ptr(), num_x,
                                   0 if xb is None else xb.data_ptr(), 0 if xb is None else xb.shape[0],
                                   0 if yb is None else yb.data_ptr(), 0 if yb is None else yb.shape[0],
                                   dim_x, int(xd.dtype == torch.float16), int(k), _MARGINS[margin], scores.data_ptr(), wptr,
                                   int(need[0]), torch.cuda.current_stream(dev).cuda_stream), "svx_margin_scores")
    return scores.cpu().numpy() if as_numpy else scores


def compute_sim_with_nonflat_idx(idx_x, idx_y, x: np.ndarray, y: np.ndarray, k: int, margin: str) -> np.ndarray:
    """Reference signature (score_align.py:124-129).  idx_x / idx_y: FlatIndex objects (load_flat_index) or plain
    (m, d) arrays of the indexed vectors.  Like the reference (faiss.normalize_L2, :136-137) numpy inputs x, y are
    left L2-normalised in place."""
    xb = idx_x.vectors if isinstance(idx_x, FlatIndex) else idx_x
    yb = idx_y.vectors if isinstance(idx_y, FlatIndex) else idx_y
    scores = compute_sim(x, y, k, margin, x_base=xb, y_base=yb)
    for v in (x, y):
        if isinstance(v, np.ndarray) and v.flags.writeable and v.dtype == np.float32:
            nr = np.sqrt(np.einsum("ij,ij->i", v, v))
            ok = nr > 0
            v[ok] /= nr[ok][:, None]
    return scores


# ------------------------------------------------------------------------------------------------
# Driver: same files in, same files out as the reference's main() (:164-262)
# ------------------------------------------------------------------------------------------------
def load_embed_from_tsv(tsv_path: Path, fp16_embed: bool, use_stopes: bool) -> np.ndarray:
    """prep_index.py:91-126: every line of the tsv is '<embedding file>\\t<row>'; rows come back in line order."""
    by_file = defaultdict(list)
    with open(tsv_path) as fp:
        for ii, line in enumerate(fp):
            path, _id = line.strip().split("\t")
            by_file[path].append((ii, int(_id)))
    n = sum(len(v) for v in by_file.values())
    out = None
    for path, items in by_file.items():
        emb = load_sent_embeddings(path, use_stopes=use_stopes, fp16_embed=fp16_embed, stopes_mode="memory")
        if out is None:
            out = np.empty((n, emb.shape[1]), dtype=np.float32)
        for line_no, row in items:
            out[line_no] = emb[row]
    return out


def find_valid_metas(meta: List[Tuple[str, str]], embed_dir: Path) -> List[str]:
    """score_align.py:72-93"""
    res = []
    for src_aud, tgt_aud in meta:
        src_id, tgt_id = Path(src_aud).stem, Path(tgt_aud).stem
        src_tsv, tgt_tsv = embed_dir / f"{src_id}-{tgt_id}.src.tsv", embed_dir / f"{src_id}-{tgt_id}.tgt.tsv"
        if src_tsv.exists() and tgt_tsv.exists():
            res.append(f"{src_id}-{tgt_id}")
        elif not src_tsv.exists() and not tgt_tsv.exists():
            logger.warning(f"{src_tsv} and {tgt_tsv} not exist")
        else:
            raise Exception(f"{src_tsv}: {src_tsv.exists()} | {tgt_tsv}: {tgt_tsv.exists()}")
    logger.info(f"Kept {len(res)}/{len(meta)}")
    return res


def write_to_output(align_dir: Path, align_ids: List[str], margin_scores: np.ndarray, out_dir: Path):
    """score_align.py:96-115: '<src ids>:<tgt ids>:<score>' in the order of the alignment files."""
    margin_id = 0
    for ali_id in align_ids:
        alignments = read_alignments(align_dir / f"{ali_id}.txt")
        with open(out_dir / f"{ali_id}.txt", mode="w") as fp:
            for src, tgt in alignments:
                fp.write(f"{src}:{tgt}:{margin_scores[margin_id]}\n")
                margin_id += 1
    assert margin_id == margin_scores.shape[0], f"{margin_id}, {margin_scores.shape}"


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("metadata", type=str, help="the meta file that each line contains paired audio paths")
    parser.add_argument("out_dir", type=str, help="dir to store the margin-scored alignments")
    parser.add_argument("--embed_dir", type=str, required=True, help="the dir for embedding tsvs.")
    parser.add_argument("--align_dir", type=str, required=True, help="the dir for concatenated alignments.")
    parser.add_argument("--src_lang", type=str, required=True)
    parser.add_argument("--tgt_lang", type=str, required=True)
    parser.add_argument("--index_dir", type=str, default=None,
                        help="where the Flat indexes are saved; omitted: the collections are the embeddings of all files")
    parser.add_argument("--num_probe", type=int, default=128, help="accepted for compatibility (the search is exact)")
    parser.add_argument("--gpu_type", type=str, default="fp16-shard", help="accepted for compatibility")
    parser.add_argument("--embed_fp16", action="store_true", default=False)
    parser.add_argument("--embed_stopes", action="store_true", default=False)
    parser.add_argument("--margin", type=str, default="ratio")
    parser.add_argument("--k", type=int, default=16)
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    logger.info(args)
    with open(args.metadata, "rt", encoding="utf-8") as f:
        all_pairs = [tuple(ln.strip().split("\t")[:2]) for ln in f if ln.strip()]
    src_lang, tgt_lang = args.src_lang, args.tgt_lang
    embed_dir = Path(args.embed_dir) / f"{src_lang}-{tgt_lang}"
    align_dir = Path(args.align_dir) / f"{src_lang}-{tgt_lang}"
    out_dir = Path(args.out_dir) / f"{src_lang}-{tgt_lang}"
    out_dir.mkdir(parents=True, exist_ok=True)
    metas = find_valid_metas(all_pairs, embed_dir)
    embeds = [(load_embed_from_tsv(embed_dir / f"{m}.src.tsv", args.embed_fp16, args.embed_stopes),
               load_embed_from_tsv(embed_dir / f"{m}.tgt.tsv", args.embed_fp16, args.embed_stopes)) for m in metas]
    if args.index_dir is not None:
        index_dir = Path(args.index_dir) / f"{src_lang}-{tgt_lang}"
        src_index = load_flat_index(list((index_dir / src_lang).glob("*.populate.idx"))[0])
        tgt_index = load_flat_index(list((index_dir / tgt_lang).glob("*.populate.idx"))[0])
    else:                       # what prep_index.py populates the indexes with: every file's embeddings
        src_index = FlatIndex(np.concatenate([e[0] for e in embeds]) if embeds else np.zeros((0, 1024), np.float32))
        tgt_index = FlatIndex(np.concatenate([e[1] for e in embeds]) if embeds else np.zeros((0, 1024), np.float32))
    dev = torch.device("cuda", torch.cuda.current_device())
    xb, yb = _dev(src_index.vectors, dev), _dev(tgt_index.vectors, dev)        # uploaded once for all files
    scores = [compute_sim(src, tgt, args.k, args.margin, x_base=xb, y_base=yb) for src, tgt in embeds]
    margin_scores = np.concatenate(scores, axis=0) if scores else np.zeros(0, np.float32)
    logger.info(f"Writing to {out_dir}...")
    write_to_output(align_dir, metas, margin_scores, out_dir)
    return margin_scores


if __name__ == "__main__":
    logging.basicConfig(level="INFO")
    main()
