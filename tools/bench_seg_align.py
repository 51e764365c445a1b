#!/usr/bin/env python
"""tools/bench_seg_align.py — files-to-files throughput of the step-5.4 driver (speech_vecalign_b200.seg_align) on a
synthetic corpus laid out like the reference's data directories (segments/, cat_segs/, embeds/ per language,
metadata.tsv; BASELINE configs[3] lengths: 200-800 segments per side, -a 6, fp16 .embed files as SpeechLASER/SONAR
write them).

    python tools/bench_seg_align.py --pairs 512 --root /tmp/svx_corpus            # one GPU
    python -m torch.distributed.run --nproc-per-node 8 tools/bench_seg_align.py --pairs 512 --root /tmp/svx_corpus

The corpus is generated on the first run (rank 0, on the GPU) and reused.  Prints one JSON line: pairs, seconds of the
slowest rank (files read -> alignment files written), pairs/s, bytes of embedding files read.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_corpus(root, npairs, k, seed=1234):
    import torch
    from speech_vecalign_b200 import synth
    n0s, n1s = synth.batch_sizes(npairs, seed=seed)
    for sub in ("segments", "cat_segs", "embeds"):
        for lang in ("en", "de"):
            os.makedirs(os.path.join(root, sub, lang), exist_ok=True)
    meta = []
    total = 0
    for p in range(npairs):
        v0, v1 = synth.synth_pair_torch(int(n0s[p]), int(n1s[p]), k, seed=9_000_000 + p, device="cuda")
        for lang, v in (("en", v0), ("de", v1)):
            n = v.shape[1]
            t = np.cumsum(np.full(n + 1, 1.25)) - 1.25
            start, end = [f"{x:.2f}" for x in t[:-1]], [f"{x:.2f}" for x in t[1:]]
            name = f"doc{p:05d}_{lang}"
            with open(os.path.join(root, "segments", lang, name + ".txt"), "w") as f:
                f.write("".join(f"{s} {e}\n" for s, e in zip(start, end)))
            rows, keys = [], []
            for j in range(min(k, n)):                     # row (j, e): segments e-j .. e
                rows.append(v[j, j:])
                keys.extend(f"{start[e - j]} {end[e]}" for e in range(j, n))
            with open(os.path.join(root, "cat_segs", lang, name + ".txt"), "w") as f:
                f.write("\n".join(keys) + "\n")
            emb = torch.cat(rows).half().cpu().numpy()
            emb.tofile(os.path.join(root, "embeds", lang, name + ".embed"))
            total += emb.nbytes
        meta.append(f"/audio/en/doc{p:05d}_en.ogg\t/audio/de/doc{p:05d}_de.ogg")
    with open(os.path.join(root, "metadata.tsv"), "w") as f:
        f.write("\n".join(meta) + "\n")
    with open(os.path.join(root, "corpus.json"), "w") as f:
        json.dump({"pairs": npairs, "k": k, "embed_bytes": total}, f)
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=512)
    ap.add_argument("--root", default="/tmp/svx_corpus")
    ap.add_argument("-a", "--alignment_max_size", type=int, default=6)
    ap.add_argument("--batch_gb", type=float, default=4.0)
    ap.add_argument("--repeat", type=int, default=2, help="the first run also warms the page cache and the allocators")
    ap.add_argument("--host_gather", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from speech_vecalign_b200 import seg_align
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    k = args.alignment_max_size - 1
    info_path = os.path.join(args.root, "corpus.json")
    if rank == 0:
        info = json.load(open(info_path)) if os.path.exists(info_path) else None
        if not info or info["pairs"] != args.pairs or info["k"] != k:
            t0 = time.perf_counter()
            make_corpus(args.root, args.pairs, k)
            print(f"corpus of {args.pairs} pairs written in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
    if world > 1:
        dist.barrier()
    info = json.load(open(info_path))
    out_root = os.path.join(args.root, f"out_rank")
    argv = [os.path.join(args.root, "metadata.tsv"), out_root, "--src_lang", "en", "--tgt_lang", "de",
            "--seg_dir", os.path.join(args.root, "segments"), "--concat_dir", os.path.join(args.root, "cat_segs"),
            "--embed_dir", os.path.join(args.root, "embeds"), "--fp16_embed", "-a", str(args.alignment_max_size),
            "--batch_gb", str(args.batch_gb)] + (["--host_gather"] if args.host_gather else [])
    best = None
    for rep in range(args.repeat):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n = seg_align.main(argv)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt, float(n)], device="cuda", dtype=torch.float64)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt, n = float(tm[0]), int(t[1])
        best = dt if best is None else min(best, dt)
    if rank == 0:
        nout = len(os.listdir(os.path.join(out_root, "en-de")))
        print(json.dumps({"metric": "files -> files aligned doc pairs/sec (seg_align driver)", "pairs": n, "n_gpus": world,
                          "seconds": best, "value": n / best, "unit": "pairs/s", "embed_bytes_read": info["embed_bytes"],
                          "output_files": nout, "host_gather": bool(args.host_gather), "alignment_max_size": args.alignment_max_size}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
