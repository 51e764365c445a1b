"""Input contract of the alignment path (reference: svecalign/utils/embedding_utils.py): loaders
of the ``.embed`` files and the (K, N, D) overlap tensor ``vecs[j, i+j] = emb(segments i..i+j)``.

Host-side (file formats, string keys -> row ids); the arithmetic starts at ``dp_utils.vecalign``.
``overlap_row_table`` exposes the (K, N) row-index table so that the tensor can also be gathered
on the device straight from the fp16/fp32 rows.
"""
import logging
from typing import Dict, List, Optional, Set, Tuple

import numpy as np

EMBED_DIM = 1024
PAD_LABEL = "PAD"
logger = logging.getLogger(__name__)


def preprocess_line(line: str) -> str:
    line = line.strip()
    if len(line) == 0:                       # embedding_utils.py:29-35
        logger.warning("Encountered empty line.")
        line = '[BLANK_LINE]'
    return line


def load_stopes_embeddings(path: str, mode: str = "mmap") -> np.ndarray:
    """embedding_utils.py:38-44.  stopes writes .npy-framed files; when stopes is not installed
    the same bytes are read with numpy (verified on the shipped example)."""
    try:
        from stopes.utils.embedding_utils import Embedding  # noqa
    except ImportError:
        return np.load(path, mmap_mode="r" if mode == "mmap" else None).astype(np.float32)
    with Embedding(path).open_for_read(mode) as e:
        return e.astype(np.float32)


def load_np_embeddings(embed_file: str, fp16_embed: bool) -> np.ndarray:
    """embedding_utils.py:47-55: raw fp16/fp32 dumps."""
    if fp16_embed:
        return np.fromfile(embed_file, dtype=np.float16, count=-1).astype(np.float32)
    return np.fromfile(embed_file, dtype=np.float32, count=-1)


def load_sent_embeddings(embed_file: str, use_stopes: bool = False, fp16_embed: bool = False,
                         stopes_mode: str = "mmap") -> np.ndarray:
    """embedding_utils.py:58-76 — always returns fp32 (rows, EMBED_DIM)."""
    if use_stopes:
        emb = load_stopes_embeddings(embed_file, mode=stopes_mode)
    else:
        emb = load_np_embeddings(embed_file, fp16_embed)
        if emb.size == 0:
            raise Exception('Got empty embedding file')
        emb = emb.reshape(emb.shape[0] // EMBED_DIM, EMBED_DIM)
    assert emb.dtype == np.float32, embed_file
    return emb


def read_in_embeddings(text_file: str, embed_file: str, use_stopes: bool = False,
                       fp16_embed: bool = False) -> Tuple[Dict[str, int], np.ndarray]:
    """embedding_utils.py:79-103: candidate string -> first row that carries it."""
    sent2line: Dict[str, int] = {}
    with open(text_file, 'rt', encoding="utf-8") as fin:
        for i, line in enumerate(fin):
            sent2line.setdefault(line.strip(), i)
    return sent2line, load_sent_embeddings(embed_file, use_stopes, fp16_embed)


def make_overlap(lines: List[str], num_overlaps: int, start_id: int,
                 ignore_indices: Optional[Set[Tuple[int, int]]] = None, comb: str = ' ',
                 overlap_segments: bool = False) -> List[str]:
    """embedding_utils.py:106-132: keys of the concatenations start_id..start_id+j, PAD from an
    ignored (start, j) onwards."""
    out: List[str] = []
    for j in range(start_id, min(start_id + num_overlaps, len(lines))):
        if ignore_indices and (start_id, j) in ignore_indices:
            out.extend([PAD_LABEL] * (min(len(lines), start_id + num_overlaps) - j))
            break
        if overlap_segments:
            out.append(f"{lines[start_id].split()[0]} {lines[j].split()[1]}")
        else:
            out.append(comb.join(lines[start_id:j + 1]))
    return out


def overlap_row_table(sent2id: dict, lines: List[str], max_overlaps: int,
                      ignore_indices: Optional[Set[Tuple[int, int]]] = None,
                      overlap_segments: bool = False) -> np.ndarray:
    """(K, N) int32: row of the embedding file that fills vecs[j, e], -1 for a zero row (PAD,
    unknown key, position before the document start)."""
    lines = [preprocess_line(x) for x in lines]
    table = np.full((max_overlaps, len(lines)), -1, dtype=np.int32)
    for i in range(len(lines)):
        keys = make_overlap(lines, max_overlaps, i, ignore_indices=ignore_indices,
                            overlap_segments=overlap_segments)
        for j, key in enumerate(keys):
            if key != PAD_LABEL:
                table[j, i + j] = sent2id.get(key, -1)
    return table


def make_doc_embedding(sent2id: dict, line_embeddings: np.ndarray, lines: List[str], max_overlaps: int,
                       ignore_indices: Optional[Set[Tuple[int, int]]] = None,
                       overlap_segments: bool = False) -> np.ndarray:
    """embedding_utils.py:135-203 -> (max_overlaps, len(lines), dim) fp32; rows with NaNs, unknown
    keys, PAD and positions before the start stay zero."""
    table = overlap_row_table(sent2id, lines, max_overlaps, ignore_indices, overlap_segments)
    dim = line_embeddings.shape[1]
    vecs = np.zeros((max_overlaps, table.shape[1], dim), dtype=np.float32)
    hit = table >= 0
    rows = np.asarray(line_embeddings[table[hit]], dtype=np.float32)
    bad = np.isnan(rows).any(axis=1)
    if bad.any():
        logger.error("loaded %d vector(s) with nan values; reset to zero", int(bad.sum()))
        rows[bad] = 0.0
    vecs[hit] = rows
    return vecs


# ---------------------------------------------------------------------------------------------
# Device path: the .embed rows travel in their on-disk dtype and the (K, N, D)
# fp32 overlap tensor is gathered on the GPU.
# ---------------------------------------------------------------------------------------------
def load_embedding_rows(embed_file: str, use_stopes: bool = False, fp16_embed: bool = False) -> np.ndarray:
    """The rows of an .embed file WITHOUT widening: (rows, EMBED_DIM) float16 or float32
    (same files as load_sent_embeddings, embedding_utils.py:38-76)."""
    if use_stopes:
        try:
            from stopes.utils.embedding_utils import Embedding  # noqa
            with Embedding(embed_file).open_for_read("mmap") as e:
                emb = np.asarray(e)
        except ImportError:
            emb = np.load(embed_file, mmap_mode="r")
    else:
        emb = np.fromfile(embed_file, dtype=np.float16 if fp16_embed else np.float32, count=-1)
        if emb.size == 0:
            raise Exception('Got empty embedding file')
        emb = emb.reshape(emb.shape[0] // EMBED_DIM, EMBED_DIM)
    if emb.dtype not in (np.float16, np.float32):
        emb = emb.astype(np.float32)
    return emb


def read_in_embedding_rows(text_file: str, embed_file: str, use_stopes: bool = False, fp16_embed: bool = False):
    """read_in_embeddings (embedding_utils.py:79-103) keeping the on-disk dtype of the rows."""
    sent2line: Dict[str, int] = {}
    with open(text_file, 'rt', encoding="utf-8") as fin:
        for i, line in enumerate(fin):
            sent2line.setdefault(line.strip(), i)
    return sent2line, load_embedding_rows(embed_file, use_stopes, fp16_embed)


def make_doc_embedding_device(sent2id: dict, rows: np.ndarray, lines: List[str], max_overlaps: int,
                              ignore_indices: Optional[Set[Tuple[int, int]]] = None, overlap_segments: bool = False,
                              device=None):
    """make_doc_embedding (embedding_utils.py:135-203) with the copy loop on the GPU: uploads `rows`
    (fp16 or fp32, as stored) and the (K, N) row table, returns the (K, N, D) fp32 CUDA tensor —
    bit-identical to the host function (fp16 -> fp32 is exact; NaN rows, PAD, unknown keys -> zeros)."""
    import torch
    from . import capi
    if not torch.cuda.is_available():
        raise capi.SvxError("no CUDA device: speech_vecalign_b200 has no CPU fallback")
    dev = device or torch.device("cuda", torch.cuda.current_device())
    table = overlap_row_table(sent2id, lines, max_overlaps, ignore_indices, overlap_segments)
    k, n = table.shape
    rows = np.ascontiguousarray(rows)
    if rows.dtype not in (np.float16, np.float32) or rows.ndim != 2:
        raise ValueError("rows must be a (nrows, dim) float16 or float32 array")
    dim = rows.shape[1]
    t_rows = torch.from_numpy(rows).to(dev, non_blocking=True)
    t_tab = torch.from_numpy(np.ascontiguousarray(table)).to(dev, non_blocking=True)
    out = torch.empty((k, n, dim), dtype=torch.float32, device=dev)
    nan_rows = torch.zeros(1, dtype=torch.int32, device=dev)
    if k * n:
        job = np.zeros(1, dtype=capi.GATHER)
        job["rows"], job["table"], job["out"], job["nan_rows"] = t_rows.data_ptr(), t_tab.data_ptr(), out.data_ptr(), nan_rows.data_ptr()
        job["k"], job["n"], job["nrows"], job["is_fp16"] = k, n, rows.shape[0], int(rows.dtype == np.float16)
        jd = torch.from_numpy(job.view(np.uint8).reshape(-1).copy()).to(dev)
        capi.check(capi.lib().svx_gather_doc_embedding(jd.data_ptr(), capi.hptr(job), 1, dim,
                                                       torch.cuda.current_stream(dev).cuda_stream), "svx_gather_doc_embedding")
        bad = int(nan_rows.item())                        # also orders the kernel before t_rows / t_tab are released
        if bad:
            logger.error("loaded %d vector(s) with nan values; reset to zero", bad)
    return out


# ---------------------------------------------------------------------------------------------
# Batched device path of the seg_align driver: many documents per call.  Host part (file reads, row tables in C on
# several threads, rows packed into ONE pinned slab) and device part (one upload, ONE gather launch, one NaN-count
# read-back per batch) are separate so that a loader thread can prepare batch b + 1 while the GPU aligns batch b.
# ---------------------------------------------------------------------------------------------
def open_embedding_rows(embed_file: str, use_stopes: bool = False, fp16_embed: bool = False) -> np.ndarray:
    """load_embedding_rows without reading the file: a memory map in the on-disk dtype."""
    if use_stopes:
        return load_embedding_rows(embed_file, True, fp16_embed)
    emb = np.memmap(embed_file, dtype=np.float16 if fp16_embed else np.float32, mode="r")
    if emb.size == 0:
        raise Exception('Got empty embedding file')
    return emb.reshape(emb.shape[0] // EMBED_DIM, EMBED_DIM)


def prepare_documents_host(docs, max_overlaps: int, use_stopes: bool, fp16_embed: bool, nthreads: int = 8):
    """docs: list of dicts {seg, cat, emb, nlines, ignore (set of (start, end) or None)}.  Returns the host half of a
    batch: pinned `slab` holding every document's embedding rows back to back (on-disk dtype), pinned `tables`
    (concatenated (K, N) int32 row tables, svx_host_overlap_tables), and per-document offsets."""
    import ctypes
    import torch
    from . import capi
    L = capi.lib()
    nd = len(docs)
    rows = [open_embedding_rows(str(d["emb"]), use_stopes, fp16_embed) for d in docs]
    dim = rows[0].shape[1] if nd else EMBED_DIM
    is16 = [r.dtype == np.float16 for r in rows]
    nbytes = np.array([r.shape[0] * r.shape[1] * r.dtype.itemsize for r in rows], dtype=np.int64)
    row_off = np.concatenate([[0], np.cumsum((nbytes + 255) // 256 * 256)]).astype(np.int64)
    nlines = np.array([d["nlines"] for d in docs], dtype=np.int32)
    tab_off = np.concatenate([[0], np.cumsum(max_overlaps * nlines.astype(np.int64))]).astype(np.int64)
    slab = torch.empty(max(int(row_off[-1]), 16), dtype=torch.uint8, pin_memory=True)
    tables = torch.empty(max(int(tab_off[-1]), 4), dtype=torch.int32, pin_memory=True)
    tnp = tables.numpy()
    # row tables: C, all documents of the batch in one call
    seg = (ctypes.c_char_p * nd)(*[str(d["seg"]).encode() for d in docs])
    cat = (ctypes.c_char_p * nd)(*[str(d["cat"]).encode() for d in docs])
    ign_arrays = [np.array(sorted(d["ignore"]), dtype=np.int32).reshape(-1, 2) if d.get("ignore") else np.zeros((0, 2), np.int32) for d in docs]
    ign_ptrs = (ctypes.c_void_p * nd)(*[a.ctypes.data if a.size else None for a in ign_arrays])
    n_ign = np.array([a.shape[0] for a in ign_arrays], dtype=np.int32)
    out_ptrs = (ctypes.c_void_p * nd)(*[tnp.ctypes.data + 4 * int(tab_off[i]) for i in range(nd)])
    nrows = np.zeros(nd, dtype=np.int32)
    capi.check(L.svx_host_overlap_tables(nd, seg, cat, ign_ptrs, capi.hptr(n_ign), max_overlaps, out_ptrs, capi.hptr(nlines),
                                         capi.hptr(nrows), nthreads), "svx_host_overlap_tables")
    # rows -> pinned slab (reads the memory-mapped files), several threads per copy
    base = slab.data_ptr()
    for i, r in enumerate(rows):
        if nbytes[i]:
            src = r if r.flags.c_contiguous else np.ascontiguousarray(r)
            capi.check(L.svx_host_memcpy(base + int(row_off[i]), src.ctypes.data, int(nbytes[i]), nthreads), "svx_host_memcpy")
    return {"slab": slab, "tables": tables, "row_off": row_off, "tab_off": tab_off, "nlines": nlines, "is16": is16,
            "nrows": np.array([r.shape[0] for r in rows], dtype=np.int32), "dim": dim, "k": max_overlaps}


def gather_documents_device(host, device=None):
    """Device half: uploads the slab and the tables (two copies), gathers every document's (K, N, D) fp32 overlap
    tensor with ONE svx_gather_doc_embedding launch.  Returns (list of tensors, nan_rows int32 device tensor [ndocs],
    keepalive) - read nan_rows after the batch has been synchronised anyway."""
    import torch
    from . import capi
    dev = device or torch.device("cuda", torch.cuda.current_device())
    nd = len(host["nlines"])
    k, dim = host["k"], host["dim"]
    slab_d = host["slab"].to(dev, non_blocking=True)
    tab_d = host["tables"].to(dev, non_blocking=True)
    out_off = np.concatenate([[0], np.cumsum(k * host["nlines"].astype(np.int64) * dim)]).astype(np.int64)
    out = torch.empty(max(int(out_off[-1]), 1), dtype=torch.float32, device=dev)
    nan_rows = torch.zeros(max(nd, 1), dtype=torch.int32, device=dev)
    jobs = np.zeros(nd, dtype=capi.GATHER)
    jobs["rows"] = slab_d.data_ptr() + host["row_off"][:-1].astype(np.uint64)
    jobs["table"] = tab_d.data_ptr() + 4 * host["tab_off"][:-1].astype(np.uint64)
    jobs["out"] = out.data_ptr() + 4 * out_off[:-1].astype(np.uint64)
    jobs["nan_rows"] = nan_rows.data_ptr() + 4 * np.arange(nd, dtype=np.uint64)
    jobs["k"], jobs["n"], jobs["nrows"], jobs["is_fp16"] = k, host["nlines"], host["nrows"], np.array(host["is16"], dtype=np.int32)
    stage = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).pin_memory() if nd else None
    jd = torch.empty(max(jobs.nbytes, 16) + 16, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    if nd:
        capi.check(capi.lib().svx_upload_pinned(jd.data_ptr(), stage.data_ptr(), stage.numel(), stream), "svx_upload_pinned")
        capi.check(capi.lib().svx_gather_doc_embedding(jd.data_ptr(), capi.hptr(jobs), nd, dim, stream), "svx_gather_doc_embedding")
    tensors = [out[int(out_off[i]):int(out_off[i + 1])].view(k, int(host["nlines"][i]), dim) for i in range(nd)]
    return tensors, nan_rows, (slab_d, tab_d, jd, stage, jobs, out, host)
