#!/usr/bin/env python
"""tools/ncu_summary.py — turns ncu exports into the summaries committed under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md> [--traffic profiles/traffic.json --workload cfg2]
        per-launch list of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
        (kernels of ONE step): time share per kernel, DRAM bytes; optionally the per-launcher DRAM traffic
        (sum over the launcher's kernels) that bench.py reports as roofline.traffic.
    python tools/ncu_summary.py full <report.ncu-rep> <out.md>
        key metrics of every kernel in an `ncu --set full` report.
"""
import collections
import csv
import json
import subprocess
import sys

LAUNCHER = [  # kernel-name prefix -> bench.py launcher name
    ("k_level_", "svx_level_prologue"), ("k_normalize", "svx_normalize_rows"), ("k_pairsum", "svx_downsample"),
    ("k_center", "svx_downsample"), ("k_sample_mean", "svx_sample_norms"), ("k_norms_gemv", "svx_sample_norms"),
    ("k_sort_samples", "svx_score_pairs"), ("k_score_pairs", "svx_score_pairs"), ("k_del_knob", "svx_del_knob"),
    ("k_dense_costs", "svx_dense_costs"), ("k_dense_dp", "svx_dense_dp"), ("k_upload", "svx_upload_pinned"),
    ("k_fill_f32", "svx_plan_upload"), ("k_margin", "svx_margin_scores"), ("k_gather", "svx_gather_doc_embedding"),
]


def short(name):
    n = name.replace("void ", "").replace("<unnamed>::", "")
    return n.split("(")[0]


def launcher_of(kname, level0):
    s = short(kname)
    for pre, nm in LAUNCHER:
        if s.startswith(pre):
            return nm
    if s.startswith("k_banded_costs"):
        return "svx_banded_costs_level0" if level0 else "svx_banded_costs_coarse"
    if s.startswith("k_banded_dp"):
        return "svx_banded_dp_level0" if level0 else "svx_banded_dp_coarse"
    return s


def launches(path, out, traffic=None, workload="cfg2"):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, mi, vi, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("ID")
    per = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > vi:
            per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    ls = list(per.values())
    # keep the last COMPLETE step of the capture: a step starts with the level-0 prologue (first prologue
    # kernel after a banded DP / upload / the start of the list) and ends with the level-0 banded DP
    starts = [i for i, d in enumerate(ls) if "k_level_" in d["k"] and (i == 0 or "k_banded_dp" in ls[i - 1]["k"] or "k_upload" in ls[i - 1]["k"])]
    ends = [i for i, d in enumerate(ls) if "k_banded_dp" in d["k"] and (i + 1 == len(ls) or "k_level_" in ls[i + 1]["k"])]
    if starts and ends and any(s_ < ends[-1] for s_ in starts):
        s_ = max(s_ for s_ in starts if s_ < ends[-1])
        ls = ls[s_:ends[-1] + 1]
    # the last banded costs / dp launch of a step is level 0
    last_cost = max(i for i, d in enumerate(ls) if "k_banded_costs" in d["k"])
    last_dp = max(i for i, d in enumerate(ls) if "k_banded_dp" in d["k"])
    tot = sum(d["gpu__time_duration.sum"] for d in ls)
    agg = collections.OrderedDict()
    with open(out, "w") as f:
        f.write(f"# ncu launch list (one step; cold-cache, serialised: compare SHARES) — source {path}\n\n")
        f.write("| # | kernel | time us | share | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|\n")
        for i, d in enumerate(ls):
            t = d["gpu__time_duration.sum"] / 1e3
            rd, wr = d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6
            f.write(f"| {i} | `{short(d['k'])}` | {t:.1f} | {100 * d['gpu__time_duration.sum'] / tot:.1f}% | {rd:.1f} | {wr:.1f} |\n")
            a = agg.setdefault(launcher_of(d["k"], i in (last_cost, last_dp)), [0.0, 0.0, 0])
            a[0] += t
            a[1] += rd + wr
            a[2] += 1
        f.write("\n| launcher (bench.py name) | launches | time us | share | DRAM MB |\n|---|---|---|---|---|\n")
        for nm, (t, b, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| {nm} | {n} | {t:.1f} | {100 * t * 1e3 / tot:.1f}% | {b:.1f} |\n")
    if traffic:
        try:
            tj = json.load(open(traffic))
        except Exception:
            tj = {}
        tj[workload] = {nm: int(b * 1e6) for nm, (t, b, n) in agg.items()}
        json.dump(tj, open(traffic, "w"), indent=1, sort_keys=True)


FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U = rows[0], rows[1]
    cols = [c for c in FULL if c in H]
    extra = [c for c in H if "tensor" in c and c not in cols][:6]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary — source {rep}\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[H.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for c in cols + extra:
                f.write(f"| {c} | {r[H.index(c)]} | {U[H.index(c)]} |\n")
            f.write("\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        tr = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
        wl = sys.argv[sys.argv.index("--workload") + 1] if "--workload" in sys.argv else "cfg2"
        launches(sys.argv[2], sys.argv[3], tr, wl)
    else:
        full(sys.argv[2], sys.argv[3])
