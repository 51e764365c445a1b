#!/usr/bin/env python
"""tools/sass_excerpts.py - writes profiles/r02_sass_excerpts.md: per kernel of libsvx.so the tensor-core (UTCHMMA), TMA
(UTMALDG), TMEM (LDTM), mbarrier (SYNCS) and packed-FP32 (FFMA2) mnemonics found by `cuobjdump -sass`, with a short excerpt."""
import collections
import re
import subprocess
import sys

SO = "speech-vecalign_b200/libsvx.so"
WANT = [("k_margin_knn", "UTCHMMA|UTMALDG|UTCBAR|LDTM|SYNCS"), ("k_dense_costs_tc", "UTCHMMA|UTMALDG|UTCBAR|LDTM|SYNCS"),
        ("k_banded_costs_p2ILi4ELi32ELi4ELb1E", r"FFMA2|LDGSTS|SYNCS|LDS\.128"), ("k_banded_costs_p2ILi4ELi32ELi4ELb0E", r"FFMA2|LDGSTS|SYNCS|LDS\.128"),
        ("k_banded_costs_p2ILi5ELi16ELi4ELb1E", r"FFMA2|LDGSTS|SYNCS|LDS\.128"), ("k_banded_costs_blkILi1ELb1E", "FFMA2|FMUL|FADD|LDGSTS")]


def main(out_path="profiles/r02_sass_excerpts.md"):
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)
    out = ["# SASS evidence (cuobjdump -sass speech-vecalign_b200/libsvx.so, sm_100a) - round 2\n",
           "Regenerate: `python tools/sass_excerpts.py` after `make -C speech-vecalign_b200/csrc`.  Per kernel: the tensor-core / TMA /",
           "packed-FP32 mnemonics it contains (counts of static instructions) and a short excerpt of its inner loop.\n"]
    for name, pat in WANT:
        for f in funcs[1:]:
            head = f.split("\n", 1)[0]
            if name not in head:
                continue
            ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", f)
            c = collections.Counter(ins)
            out.append(f"## `{head.strip()}`\n")
            out.append("| mnemonic | static count |\n|---|---|")
            for k, v in sorted(c.items(), key=lambda kv: -kv[1]):
                if re.match(pat, k) or re.match("FFMA$|FMUL|FADD$|UTC|UTMA|LDTM|FFMA2", k):
                    out.append(f"| `{k}` | {v} |")
            out.append("\n```")
            for sub in pat.split("|"):              # a few lines per mnemonic of interest
                lines = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() for ln in f.split("\n") if re.search(sub, ln)]
                out.extend(ln.strip() for ln in lines[:3])
            out.append("```\n")
            break
    open(out_path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:])
