"""Batch driver for step 5.4 of the Speech-Vecalign pipeline — the GPU replacement of
``svecalign/seg_align/align.py`` (reference: a serial loop calling ``vecalign.align`` once per document
pair, align.py:206-230).

Same command line (positional ``metadata out_dir``; ``--src_lang --tgt_lang --seg_dir --concat_dir
--embed_dir --is_stopes_embed --fp16_embed -a --search_buffer_size -d --max_size_full_dp
--costs_sample_size --num_samps_for_norm --ign_indices_dir``, align.py:13-96), same file conventions
(``<dir>/<lang>/<audio stem>.txt|.embed``, ignore files ``<src>-<tgt>.{src,tgt}.txt``, output
``out_dir/<src>-<tgt>/<src stem>-<tgt stem>.txt``, align.py:117-179) and the same output format
(vecalign.py:174-184).  What changes is the execution: pairs are grouped into GPU batches bounded by
``--batch_gb`` of embeddings, every batch is one ``vecalign_batch`` call (one kernel launch per stage
for all its pairs), and with several processes (``torchrun`` or ``--rank/--n_shard``) the pairs are
length-balanced across GPUs; no collective is needed because every process writes its own files.
The batches are pipelined: a loader thread reads the files of batch b + 1 (row tables built by
svx_host_overlap_tables on several threads, embedding rows packed into one pinned slab in their on-disk
dtype) while the GPU gathers and aligns batch b (one upload, one gather launch, one launch chain) and
writer threads format the output files of batch b - 1.

Additions: ``--skip_existing`` (the reference's slow stages do this, preprocess/segment.py:105-128),
``--seed`` (the reference draws from the unseeded global RNG; here pair i draws from
RandomState(crc32(output name) ^ seed), so results do not depend on batching or sharding),
``--cost_mode exact|fast|tc``, ``--max_cost`` (writes step 6.1's filtered files from the same records,
postprocess/filter_by_cost.py:39-87).

    python -m speech_vecalign_b200.seg_align metadata.tsv out --src_lang en --tgt_lang de \\
        --seg_dir segments --concat_dir cat_segs --embed_dir embeds --is_stopes_embed -a 6
"""
import argparse
import logging
import os
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

from . import embedding_utils as eu
from .dp_utils import vecalign_batch
from .engine import host_threads, records_to_alignments
from .vecalign import load_ignore_index_file, make_alignment_types, print_alignments, width_over2_for

logger = logging.getLogger("seg_align")


def build_parser():
    p = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    p.add_argument("metadata", help="tab-separated pairs of audio paths, one pair per line")
    p.add_argument("out_dir", help="alignments are written to out_dir/<src_lang>-<tgt_lang>/")
    p.add_argument("--src_lang", required=True)
    p.add_argument("--tgt_lang", required=True)
    p.add_argument("--seg_dir", required=True, help="raw segment lists")
    p.add_argument("--concat_dir", required=True, help="concatenated-segment lists (embedding row keys)")
    p.add_argument("--embed_dir", required=True, help="embedding files")
    p.add_argument("--is_stopes_embed", action="store_true", help=".embed files written by stopes (.npy framed)")
    p.add_argument("--fp16_embed", action="store_true", help="raw fp16 dumps (SONAR / numpy)")
    p.add_argument("-a", "--alignment_max_size", type=int, default=6)
    p.add_argument("--search_buffer_size", type=int, default=5)
    p.add_argument("-d", "--del_percentile_frac", type=float, default=0.2)
    p.add_argument("--max_size_full_dp", type=int, default=300)
    p.add_argument("--costs_sample_size", type=int, default=20000)
    p.add_argument("--num_samps_for_norm", type=int, default=100)
    p.add_argument("--ign_indices_dir", default=None)
    # execution
    p.add_argument("--batch_gb", type=float, default=16.0, help="embedding bytes per GPU batch")
    p.add_argument("--skip_existing", action="store_true")
    p.add_argument("--host_gather", action="store_true", help="build the overlap tensors on the host (reference path) "
                                                             "instead of uploading the .embed rows and gathering on the GPU")
    p.add_argument("--max_cost", type=float, default=None,
                   help="also write the step-6.1 output (postprocess/filter_by_cost.py --max_cost) from the same records")
    p.add_argument("--filter_dir", default=None, help="where the filtered alignments go (default: <out_dir>_<max_cost>)")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--cost_mode", default="exact", choices=["exact", "fast", "tc"])
    p.add_argument("--rank", type=int, default=None, help="shard index (default: RANK env, else 0)")
    p.add_argument("--n_shard", type=int, default=None, help="number of shards (default: WORLD_SIZE env, else 1)")
    return p


def _existing(path, what):
    if path.exists():
        return True
    logger.warning("%s does not exist (%s); pair skipped", path, what)
    return False


def resolve_pairs(meta_lines, args):
    """metadata line -> dict of the files step 5.4 needs; pairs with a missing file are dropped with
    a warning, exactly the cases align.py:117-179 drops."""
    sl, tl = args.src_lang, args.tgt_lang
    out_dir = Path(args.out_dir) / f"{sl}-{tl}"
    ign_dir = Path(args.ign_indices_dir) / f"{sl}-{tl}" if args.ign_indices_dir else None
    jobs = []
    for line in meta_lines:
        line = line.strip()
        if not line:
            continue
        src_audio, tgt_audio = line.split("\t")[:2]
        s, t = Path(src_audio), Path(tgt_audio)
        item = {"out": out_dir / f"{s.stem}-{t.stem}.txt"}
        ok = True
        for key, root, suffix in (("seg", args.seg_dir, ".txt"), ("cat", args.concat_dir, ".txt"), ("emb", args.embed_dir, ".embed")):
            item["src_" + key] = (Path(root) / sl / s.name).with_suffix(suffix)
            item["tgt_" + key] = (Path(root) / tl / t.name).with_suffix(suffix)
            ok = ok and _existing(item["src_" + key], key) and _existing(item["tgt_" + key], key)
            if not ok:
                break
        if not ok:
            continue
        for side in ("src", "tgt"):
            f = ign_dir / f"{s.stem}-{t.stem}.{side}.txt" if ign_dir else None
            item[side + "_ign"] = f if f is not None and f.exists() else None
        jobs.append(item)
    return out_dir, jobs


def _count_lines(path):
    with open(path, "rb") as f:
        return sum(1 for _ in f)


def load_pair(item, k, args, on_device=False):
    """(vecs0, vecs1) as make_doc_embedding builds them (utils/embedding_utils.py:135-203).  on_device:
    the .embed rows are uploaded in their on-disk dtype (fp16 for SpeechLASER/SONAR dumps) and the (K, N, D)
    tensor is gathered on the GPU — the host->device bytes are the file, not the K-fold fp32 expansion."""
    out = []
    for side in ("src", "tgt"):
        lines = open(item[side + "_seg"], "rt", encoding="utf-8").readlines()
        ign = load_ignore_index_file(item[side + "_ign"]) if item[side + "_ign"] else None
        if on_device:
            key_to_row, rows = eu.read_in_embedding_rows(str(item[side + "_cat"]), str(item[side + "_emb"]),
                                                         args.is_stopes_embed, args.fp16_embed)
            out.append(eu.make_doc_embedding_device(key_to_row, rows, lines, k, ignore_indices=ign, overlap_segments=True))
        else:
            key_to_row, rows = eu.read_in_embeddings(str(item[side + "_cat"]), str(item[side + "_emb"]),
                                                     args.is_stopes_embed, args.fp16_embed)
            out.append(eu.make_doc_embedding(key_to_row, rows, lines, k, ignore_indices=ign, overlap_segments=True))
    return out[0], out[1]


def filter_by_cost(alignments, scores, max_cost):
    """Step 6.1 applied to fresh records (reference: postprocess/filter_by_cost.py:39-87, which re-parses the
    step-5.4 text file): insertions / deletions are dropped, then alignments whose cost exceeds max_cost.
    The cost compared and written is the one a reader of the "%.6f" file would see."""
    kept = []
    for (xs, ys), sc in zip(alignments, scores):
        cost = float("%.6f" % sc)
        if len(xs) == 0 or len(ys) == 0 or cost > max_cost:
            continue
        kept.append((list(xs), list(ys), cost))
    return kept


def write_filtered(path, kept):
    """filter_by_cost.py:74-80: '<src ids>:<tgt ids>:<cost>' with Python's float repr; nothing is written when
    every alignment was filtered out."""
    if not kept:
        logger.warning("Empty output. Will not write %s", path)
        return
    tmp = Path(str(path) + ".tmp")
    with open(tmp, "w") as f:
        for xs, ys, cost in kept:
            f.write(f"{xs}:{ys}:{cost}\n")
    os.replace(tmp, path)


def pair_seed(item, seed):
    return (zlib.crc32(item["out"].name.encode()) ^ (seed & 0xFFFFFFFF)) & 0xFFFFFFFF


def run(args):
    from .sharding import estimate_work, lpt_partition
    rank = args.rank if args.rank is not None else int(os.environ.get("RANK", "0"))
    nshard = args.n_shard if args.n_shard is not None else int(os.environ.get("WORLD_SIZE", "1"))
    assert 0 <= rank < nshard, f"invalid rank/n_shard {rank}/{nshard}"
    if "LOCAL_RANK" in os.environ:
        import torch
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))

    a = max(2, args.alignment_max_size)                           # vecalign.py:230-232
    k = a - 1
    types = make_alignment_types(a)
    w = width_over2_for(k, k, args.search_buffer_size)
    out_dir, jobs = resolve_pairs(open(args.metadata, "rt", encoding="utf-8"), args)
    out_dir.mkdir(parents=True, exist_ok=True)
    filt_dir = None
    if args.max_cost is not None:
        filt_dir = Path(args.filter_dir or (str(args.out_dir).rstrip("/") + f"_{args.max_cost}")) / f"{args.src_lang}-{args.tgt_lang}"
        filt_dir.mkdir(parents=True, exist_ok=True)
    if args.skip_existing:
        jobs = [j for j in jobs if not j["out"].exists()]
    sizes = [(_count_lines(j["src_seg"]), _count_lines(j["tgt_seg"])) for j in jobs]
    if nshard > 1 and jobs:
        mine = lpt_partition(estimate_work([s[0] for s in sizes], [s[1] for s in sizes], a, args.search_buffer_size), nshard)[rank]
        jobs, sizes = [jobs[i] for i in mine], [sizes[i] for i in mine]
    logger.info("shard %d/%d: %d document pairs", rank, nshard, len(jobs))

    budget = args.batch_gb * 2 ** 30
    batches, i = [], 0
    while i < len(jobs):
        lo, used = i, 0.0
        while i < len(jobs) and (i == lo or used + k * sum(sizes[i]) * eu.EMBED_DIM * 4 <= budget):
            used += k * sum(sizes[i]) * eu.EMBED_DIM * 4
            i += 1
        batches.append((lo, i))
    nthreads = host_threads(16)

    def load_host(lo, hi):
        """host half of a batch (files -> pinned slab + row tables), run one batch ahead on the loader thread"""
        docs = []
        for item, (ns, nt) in zip(jobs[lo:hi], sizes[lo:hi]):
            for side, n in (("src", ns), ("tgt", nt)):
                ign = load_ignore_index_file(item[side + "_ign"]) if item[side + "_ign"] else None
                docs.append({"seg": item[side + "_seg"], "cat": item[side + "_cat"], "emb": item[side + "_emb"], "nlines": n, "ignore": ign})
        return eu.prepare_documents_host(docs, k, args.is_stopes_embed, args.fp16_embed, nthreads)

    def write_batch(batch, res):
        bad = 0
        for item, r in zip(batch, res):
            if r["status"]:
                # the reference raises here (IndexError / 'traceback bug', dp_utils.py:123-124) and stops; a batch
                # driver keeps every other pair and leaves this one without an output file
                logger.error("traceback failed for %s (device status %d): no output written", item["out"].name, r["status"])
                bad += 1
                continue
            al, sc = records_to_alignments(r["recs"])
            tmp = item["out"].with_suffix(".txt.tmp")
            with open(tmp, "w") as f:
                print_alignments(al, scores=sc, ofile=f)
            os.replace(tmp, item["out"])
            if args.max_cost is not None:                     # step 6.1 straight from the records
                write_filtered(filt_dir / item["out"].name, filter_by_cost(al, sc, args.max_cost))
        return bad

    t_start = time.perf_counter()
    done, failed, writes = 0, 0, []
    loader = ThreadPoolExecutor(max_workers=1)
    writer = ThreadPoolExecutor(max_workers=2)
    pipelined = not args.host_gather
    fut = loader.submit(load_host, *batches[0]) if (pipelined and batches) else None
    for b, (lo, hi) in enumerate(batches):
        batch = jobs[lo:hi]
        if pipelined:
            host = fut.result()
            fut = loader.submit(load_host, *batches[b + 1]) if b + 1 < len(batches) else None
            tensors, nan_rows, keep = eu.gather_documents_device(host)
            pairs = [(tensors[2 * q], tensors[2 * q + 1]) for q in range(hi - lo)]
        else:
            pairs = [load_pair(item, k, args, on_device=False) for item in batch]
        res = vecalign_batch(pairs, types, args.del_percentile_frac, w, args.max_size_full_dp, args.costs_sample_size,
                             args.num_samps_for_norm, cost_mode=args.cost_mode, output="records",
                             seeds=[pair_seed(item, args.seed) for item in batch])
        if pipelined:
            nbad = int(nan_rows.sum().item())                 # one read-back per batch (the batch has been synchronised)
            if nbad:
                logger.error("loaded %d vector(s) with nan values; reset to zero", nbad)
            del tensors, pairs, keep
        writes.append(writer.submit(write_batch, batch, res))
        done += len(batch)
        logger.info("shard %d: %d/%d pairs aligned", rank, done, len(jobs))
    failed = sum(f.result() for f in writes)
    loader.shutdown()
    writer.shutdown()
    if failed:
        logger.error("shard %d: %d document pairs failed", rank, failed)
    dt = time.perf_counter() - t_start
    logger.info("shard %d: %d pairs, files to files in %.2f s (%.1f pairs/s)", rank, done - failed, dt, (done - failed) / max(dt, 1e-9))
    run.last_stats = {"pairs": done - failed, "seconds": dt}
    return done - failed


def main(argv=None):
    logging.basicConfig(level=os.environ.get("LOGLEVEL", "INFO"))
    args = build_parser().parse_args(argv)
    logger.info(args)
    return run(args)


if __name__ == "__main__":
    main()
